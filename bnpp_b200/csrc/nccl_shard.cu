// The cross-shard sum-out of wide-factor sharding: one NCCL all-reduce (fp64 sum) of the ranks'
// results over NVLink / NVSwitch.  See include/bnpp_b200_nccl.h.
#include <nccl.h>

#include "../../include/bnpp_b200_nccl.h"
#include "common.cuh"

extern "C" int bnpp_shard_allreduce_sum(bnpp_ctx *ctx, void *nccl_comm, double *buf_dev, uint64_t n)
{
    if (!ctx || !nccl_comm || !buf_dev) return BNPP_EINVAL;
    const ncclResult_t r = ncclAllReduce(buf_dev, buf_dev, n, ncclDouble, ncclSum, static_cast<ncclComm_t>(nccl_comm), ctx->stream);
    if (r != ncclSuccess) {
        ctx->last_error = std::string("ncclAllReduce: ") + ncclGetErrorString(r);
        return BNPP_ECUDA;
    }
    return BNPP_OK;
}

// One network sharded over the ranks of `nccl_comm` (wide-factor sharding): the plan was created with
// the shard variables among its observed ids; every rank runs it with its own values for them, then the ranks'
// result tables -- P(kept variables, evidence, shard variables = this rank's values) -- are summed over NVLink:
// the cross-shard sum-out.  A marginals plan is run unnormalised, summed, and normalised afterwards.
extern "C" int bnpp_ve_plan_run_sharded(bnpp_ctx *ctx, bnpp_ve_plan *plan, void *nccl_comm, const double *const *tables_dev,
                                        const uint32_t *obs_val, double *result_dev, double *z_dev)
{
    if (!ctx || !plan || !nccl_comm || !result_dev) return BNPP_EINVAL;
    uint64_t n = 0, total = 0;
    int rc = bnpp_ve_plan_result_size(plan, &n);
    if (rc != BNPP_OK) return rc;
    const bool is_mar = bnpp_mar_plan_layout(plan, 0, nullptr, nullptr, &total) == BNPP_OK;
    if (is_mar) {
        rc = bnpp_ve_plan_set_normalize(plan, 0);
        if (rc != BNPP_OK) return rc;
    }
    rc = bnpp_ve_plan_run(plan, tables_dev, obs_val, result_dev, z_dev);
    if (is_mar) bnpp_ve_plan_set_normalize(plan, 1);
    if (rc != BNPP_OK) return rc;
    ncclComm_t comm = static_cast<ncclComm_t>(nccl_comm);
    ncclResult_t r = ncclGroupStart();
    if (r == ncclSuccess) r = ncclAllReduce(result_dev, result_dev, n, ncclDouble, ncclSum, comm, ctx->stream);
    if (r == ncclSuccess && z_dev && !(z_dev >= result_dev && z_dev < result_dev + n))
        r = ncclAllReduce(z_dev, z_dev, 1, ncclDouble, ncclSum, comm, ctx->stream);
    const ncclResult_t r2 = ncclGroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) {
        ctx->last_error = std::string("ncclAllReduce: ") + ncclGetErrorString(r);
        return BNPP_ECUDA;
    }
    if (is_mar) return bnpp_mar_plan_normalize(plan, result_dev);
    return BNPP_OK;
}
