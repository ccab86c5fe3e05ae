// Interface between the VE plan (ve.cu) and the batched-evidence kernel (batched.cu).
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

struct bnpp_ctx;

namespace bnpp {

// One operand of a batched elimination step: an intermediate [entries][batch] or a resident
// CPT view read through a per-evidence-set base offset.
struct BatchedOperandDesc {
    const double *ptr;
    bool batched;
    const std::vector<uint32_t> *var, *card;
    const std::vector<int64_t> *stride;              // empty => dense
    const std::vector<std::pair<int64_t, int>> *obs; // (stride, evidence column) of observed axes
};

int contract_batched_step(bnpp_ctx *ctx, int k, const BatchedOperandDesc *ops, const std::vector<uint32_t> &out_var,
                          const std::vector<uint32_t> &out_card, int64_t elim, uint32_t nb, const uint8_t *ev_dev,
                          uint32_t n_obs, double *out_dev);

}  // namespace bnpp
