/* bnpp_b200_nccl.h -- the one collective of the path, in its own library
 * (libbnpp_b200_nccl.so) so that the core library has no NCCL dependency.
 *
 * A factor too wide for one GPU is sharded by the variables of its widest clique that the
 * order eliminates last (bnpp_pick_shard_vars in bnpp_b200.h); every rank eliminates its own
 * slab with the ordinary plan, treating the shard variables as observed at its rank's
 * values.  Summing the shard variables out is then the ONLY cross-GPU step of the query:
 * an all-reduce (fp64 sum) of each rank's result over NVLink.  The reference has no
 * counterpart (single process, single thread).
 */
#ifndef BNPP_B200_NCCL_H_
#define BNPP_B200_NCCL_H_

#include "bnpp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* In-place sum over all ranks of buf_dev[0..n) (doubles), enqueued on the context's stream.
 * nccl_comm is an ncclComm_t created by the caller (one rank per GPU). */
int bnpp_shard_allreduce_sum(bnpp_ctx *ctx, void *nccl_comm, double *buf_dev, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
