#!/usr/bin/env python
"""Per-launch breakdown of one VE PR query on the config-4 network: operands, union entries,
algorithmic bytes, CUDA-event ms, GB/s, kernel variant.   python tools/ve_profile.py [N W K seed]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnpp_b200 import capi, model, synth  # noqa: E402
import torch  # noqa: E402

args = [int(x) for x in sys.argv[1:5]] or [64, 40, 4, 5]
N, W, K, seed = args
ctx = capi.Context(0)
_, bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))
order, width = bn.order(list(range(N)), {}, "mf")
plan = bn.plan([], order)
res = torch.zeros(2, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr() + 8)
ctx.sync()
plan.set_profiling(True)
acc = None
R = 3
for _ in range(R):
    plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr() + 8)
    st = plan.step_stats()
    acc = st if acc is None else [dict(a, ms=a["ms"] + b["ms"]) for a, b in zip(acc, st)]
tot = sum(s["ms"] for s in acc) / R
print("width", width, "launches", len(acc), "total ms", tot, "Z", res[1].item())
for i, s in enumerate(acc):
    ms = s["ms"] / R
    if ms > 0.02:
        print("%3d k=%d entries=2^%-5.2f GB=%.3f ms=%.3f GB/s=%7.1f  %s" % (
            i, s["k"], __import__("math").log2(s["entries"]), s["bytes"] / 1e9, ms, s["bytes"] / ms / 1e6, s["kernel"]))
if len(sys.argv) > 5:
    json.dump(acc, open(sys.argv[5], "w"))
