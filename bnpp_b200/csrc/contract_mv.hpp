// Planner interface of the multi-valued elimination kernel (contract_mv.cu), used by contract_plan.
#pragma once
#include <vector>

#include "contract.hpp"

namespace bnpp {

struct MVAxis {
    uint32_t ext;               // > 1
    uint64_t s[kMaxK];          // operand strides (0 = the operand lacks the axis)
};

// union entries from which a multi-valued step takes this kernel (BNPP_MV_MIN_ENTRIES overrides; tests force 0)
uint64_t mv_min_entries();

// axes: the output's axes outermost first (dense row-major output).  BNPP_OK: `d` is a resolved contract_mv
// launch (table uploaded on ctx->stream); 1: not applicable, take another kernel; < 0: error.
int plan_mv(bnpp_ctx *ctx, LaunchDesc *d, int k, uint32_t cx, const uint64_t *sx, const uint64_t *op_bytes,
            const std::vector<MVAxis> &axes, uint64_t n_out, const ParamsHead &h);

// the TMA-staged variant (contract_mvt.cu): same contract; 1 when an operand's tile footprint is not a compact range
int plan_mvt(bnpp_ctx *ctx, LaunchDesc *d, int k, uint32_t cx, const uint64_t *sx, const uint64_t *op_bytes,
             const std::vector<MVAxis> &axes, uint64_t n_out, const ParamsHead &h);
bool mv_staged_enabled();
bool staged_async_enabled();    // ... and the other operands prefetched by cp.async into thread-private slots
bool staged_tma_enabled();      // transposed binary operands: tile by TMA bulk copies (contract_staged_bulk) or by LDGSTS

}  // namespace bnpp
