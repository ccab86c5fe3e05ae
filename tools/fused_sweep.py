#!/usr/bin/env python
"""Config 5 through ve_fused under several (lanes per set, CTAs per SM) settings in ONE process -- model, evidence and
plan are built once, the knobs are environment variables the library reads at every run, so a setting costs a few
batches of GPU time instead of a Python start-up.  Also times the launch-per-bucket path.  JSON on stdout.

    python tools/fused_sweep.py [--sets 65536] [--iters 5] [--lanes 8,16,32] [--ctas 0,4,5,6,7]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bnpp_b200 import capi, model, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--lanes", default="8,16,32")
    ap.add_argument("--ctas", default="0,4,5,6,7")
    args = ap.parse_args()
    ctx = capi.Context(0)
    N, W, K, seed, nobs = 500, 6, 3, 11, 20
    _, bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))
    evs = synth.evidence_batch(N, nobs, args.sets, seed=5, fixed_ids=True)
    observed = sorted(evs[0])
    dev = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8).cuda()
    s = ctx.torch_stream
    order, _ = bn.order([v for v in range(N) if v not in evs[0]], evs[0], "mf")
    plan = bn.plan(observed, order)

    def timed():
        for _ in range(2):
            z = bn.partition_batch(observed, dev, "mf")
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(args.iters):
            z = bn.partition_batch(observed, dev, "mf")
        e1.record(s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters, z[:2].tolist()

    out = []
    plan.set_fused(False)
    for _ in range(2):
        bn.partition_batch(observed, dev, "mf")       # plain run, then the captured graph
    ms, z = timed()
    out.append({"path": "one launch per bucket", "ms_per_batch": ms, "queries_per_s": args.sets / ms * 1e3, "Z": z})
    plan.set_fused(True)
    for lanes in [int(x) for x in args.lanes.split(",")]:
        for ctas in [int(x) for x in args.ctas.split(",")]:
            os.environ["BNPP_FUSED_G"] = str(lanes)
            if ctas:
                os.environ["BNPP_FUSED_CTAS_PER_SM"] = str(ctas)
            else:
                os.environ.pop("BNPP_FUSED_CTAS_PER_SM", None)
            ms, z = timed()
            used = plan.fused_info(args.sets)[0]          # 0: this setting does not fuse (fell back to one launch per bucket)
            out.append({"path": "ve_fused" if used else "one launch per bucket (setting refused)", "lanes": lanes,
                        "ctas_per_sm_cap": ctas or None, "ms_per_batch": ms,
                        "queries_per_s": args.sets / ms * 1e3, "Z": z})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
