// Process-wide handle on the C-ABI context (include/bnpp_b200.h) used by the bn:: classes.
#ifndef BNPP_HOST_RUNTIME_HH
#define BNPP_HOST_RUNTIME_HH

#include "../../include/bnpp_b200.h"

namespace bn {
namespace gpu {

// Created on first use on device $BNPP_DEVICE (default 0).  There is no CPU fallback: if
// no CUDA device is usable the process prints the reason and exits with status 3.
bnpp_ctx *ctx();
// aborts with bnpp_last_error() when rc != 0 (CUDA errors -> non-zero exit)
void check(int rc, const char *what);
void shutdown();

}  // namespace gpu
}  // namespace bn

#endif
