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
