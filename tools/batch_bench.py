#!/usr/bin/env python
"""Config 5: PR for a batch of evidence sets on the 500-variable network, sharded over ranks with
no collective.  Prints queries/s (device-timed, evidence resident; and end-to-end from pinned host
evidence with the result copied back).

    python tools/batch_bench.py [--sets 65536] [--iters 5]
    torchrun --nproc-per-node N tools/batch_bench.py ...
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bnpp_b200 import capi, model, sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--fused", type=int, default=1, help="1: the whole batch in one launch (K9, default); 0: one launch per bucket (K8)")
    ap.add_argument("--lanes", type=int, default=0, help="K9: force the lanes per evidence set (8, 16, 32, 128)")
    ap.add_argument("--ctas-per-sm", type=int, default=0, help="K9: cap on resident CTAs per SM")
    args = ap.parse_args()
    if not args.fused:
        os.environ["BNPP_FUSED"] = "0"
    if args.lanes:
        os.environ["BNPP_FUSED_G"] = str(args.lanes)
    if args.ctas_per_sm:
        os.environ["BNPP_FUSED_CTAS_PER_SM"] = str(args.ctas_per_sm)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = capi.Context(local)
    N, W, K, seed, nobs = 500, 6, 3, 11, 20
    _, bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))
    evs = synth.evidence_batch(N, nobs, args.sets, seed=5, fixed_ids=True)
    observed = sorted(evs[0])
    lo, hi = sharding.batch_slice(rank, world, args.sets)      # contiguous shard, no communication
    host = torch.tensor([[ev[v] for v in observed] for ev in evs[lo:hi]], dtype=torch.uint8).pin_memory()
    dev = host.cuda()
    s = ctx.torch_stream
    for _ in range(3):
        z = bn.partition_batch(observed, dev, "mf")
    ctx.sync()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(args.iters):
        z = bn.partition_batch(observed, dev, "mf")
    e1.record(s)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    launches = (ctx.launches - launches0) // args.iters
    plan0 = list(bn._plans.values())[0]
    plan_bytes = plan0.bytes      # 8 * (sum #operands + #out) per evidence set (CPT reads included)
    fused_info = plan0.fused_info(hi - lo)
    union_entries = plan0.union_entries
    bn.drop_plans()
    torch.cuda.synchronize()
    per_iter = []
    for _ in range(args.iters):
        t0 = time.perf_counter()
        bn.drop_plans()                 # nothing cached: ordering, planning, H2D of the evidence, launches, D2H of Z
        z = bn.partition_batch(observed, None, "mf", host_values=host)
        with torch.cuda.stream(s):
            zh = z.to("cpu", non_blocking=False)
        per_iter.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = sum(per_iter) / len(per_iter)
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = t.tolist()
    if rank == 0:
        out = {"metric": "VE PR queries/sec", "config": "config 5: %d evidence sets, 500-variable BN (W=6 K=3 seed=11), 20 observed ids fixed" % args.sets,
               "n_gpus": world, "value": args.sets / ms * 1e3, "unit": "queries/s", "ms_per_batch": ms,
               "e2e": {"value": args.sets / e2e_ms * 1e3, "ms_per_batch": e2e_ms, "h2d_bytes": host.numel() * world, "d2h_bytes": 8 * args.sets,
                       "ms_each": [round(x, 2) for x in per_iter]},
               "launches_per_batch": launches, "sample_Z": zh[:3].tolist(),
               "kernel": ("ve_fused: %d lanes per set, %d doubles of shared memory per set, %d steps in the launch" % fused_info)
               if fused_info[0] else "contract_batched: one launch per bucket",
               "entries_per_s": union_entries * args.sets / ms * 1e3,
               "algorithmic_GB_per_batch_per_rank": plan_bytes * (hi - lo) / 1e9,
               "GBs_per_rank": plan_bytes * (hi - lo) / ms / 1e6}
        print(json.dumps(out))
        if args.json:
            json.dump(out, open(args.json, "w"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
