#!/usr/bin/env python
"""Per-launch profile of PR (-mf) on a shipped network from oracle/_ref/models (present where the
reference was built).   python tools/real_profile.py Munin1 [Link ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bnpp_b200 import capi, model  # noqa: E402

ctx = capi.Context(0)
for name in sys.argv[1:] or ["Munin1"]:
    path = os.path.join(ROOT, "oracle", "_ref", "models", "bayesnets", name + ".uai")
    _, bn = model.from_uai_text(ctx, open(path).read())
    z, ms = bn.partition({}, "mf")
    bn.drop_plans()
    z, ms = bn.partition({}, "mf")
    reps = []
    for _ in range(5):
        t0 = time.perf_counter()
        z2, _ = bn.partition({}, "mf")
        reps.append((time.perf_counter() - t0) * 1e3)
    assert z2 == z
    launches0 = ctx.launches
    bn.partition({}, "mf")
    print(name, "Z", z, "e2e %.2f ms" % ms, bn.last_timing, "replay %.3f ms (min of 5), %d launches per query" % (min(reps), ctx.launches - launches0))
    order, width = bn.order(list(range(bn.nvars)), {}, "mf")
    plan = bn.plan([], order)
    res = torch.zeros(2, dtype=torch.float64, device="cuda")
    plan.set_profiling(True)
    plan.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr() + 8)
    st = plan.step_stats()
    tot = sum(s["ms"] for s in st)
    print("  width", width, "launches", len(st), "sum of launch ms %.3f" % tot, "union entries %.3e" % plan.union_entries,
          "GB %.3f" % (plan.bytes / 1e9))
    for s in sorted(st, key=lambda s: -s["ms"])[:6]:
        print("   %.3f ms  k=%d entries=%.3e GB/s=%.0f  %s" % (s["ms"], s["k"], s["entries"], s["bytes"] / s["ms"] / 1e6, s["kernel"]))
    bn.close()
