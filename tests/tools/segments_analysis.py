#!/usr/bin/env python
"""How launch-bound is each shipped network?  For PR (-mf) on every Bayesian network under oracle/_ref/models
(test infrastructure: the reference's model files), a DRY plan (no GPU) gives the launches per query, how many of them
are small (union table <= 2^14 entries: the K9 interpreter's range), and how many launches would remain if every run of
consecutive small steps became one launch (DESIGN.md gap 4).  Markdown table on stdout."""
import ctypes
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from bnpp_b200 import capi, model  # noqa: E402
from fused_interp import DryPlan, parse_uai  # noqa: E402

SMALL = 1 << 14
print("| network | variables | min-fill width | order ms (host, this container) | plan ms (host) | launches | small launches | share of union entries in small steps | "
      "K9 today | launches if small runs fused (elimination order) | launches with tasks (K10: one launch per dependency level + wide steps) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---|")
for path in sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "models", "bayesnets", "*.uai"))):
    cards, scopes, _ = parse_uai(open(path).read())
    n = len(cards)
    arr, keep = model._scopes(scopes, cards)          # marshalling is Python's cost, not the library's: outside the timers
    carr = capi._u32(cards)
    order, width = model.elim_order(cards, scopes, list(range(n)), "mf", _arr=arr, _cards=carr)
    t0 = time.perf_counter()
    order, width = model.elim_order(cards, scopes, list(range(n)), "mf", _arr=arr, _cards=carr)
    t1 = time.perf_counter()
    L = capi.lib()
    od, ob = capi._u32(order), capi._u32([])
    h = ctypes.c_void_p()
    t1b = time.perf_counter()
    L.bnpp_ve_plan_create(None, len(scopes), arr, 0, ctypes.cast(ob, capi.c_u32p), len(order), ctypes.cast(od, capi.c_u32p), ctypes.byref(h))
    t2 = time.perf_counter()
    L.bnpp_ve_plan_destroy(h)
    p = DryPlan(cards, scopes, [], order)
    vals = [ctypes.c_uint64() for _ in range(5)]
    L.bnpp_ve_plan_info(p.h, None, None, None, *[ctypes.byref(v) for v in vals])
    ns = vals[0].value
    ent = (ctypes.c_uint64 * max(1, ns))()
    L.bnpp_ve_plan_step_stats(p.h, ns, None, None, ctypes.cast(ent, capi.c_u64p), None)
    ent = list(ent[:ns])
    small = [e <= SMALL for e in ent]
    runs = sum(1 for i, s in enumerate(small) if s and (i == 0 or not small[i - 1]))
    after = (ns - sum(small)) + runs
    tot = sum(ent) or 1
    lanes = p.fused_info(1)[0]
    built = "-"
    if not lanes:
        q = DryPlan(cards, scopes, [], order)
        segs = q.segments(0)           # K10: tasks (subtrees of small steps) grouped by dependency level
        launches, groups, levels = q.launches()
        built = "%d (%d tasks in %d group launches, %d levels)" % (launches, len(segs), groups, levels)
        q.close()
    print("| %s | %d | %d | %.2f | %.2f | %d | %d | %.1f %% | %s | %d | %s |"
          % (os.path.basename(path)[:-4], n, width, (t1 - t0) * 1e3, (t2 - t1b) * 1e3, ns, sum(small),
             100.0 * sum(e for e, s in zip(ent, small) if s) / tot, ("one launch, %d lanes" % lanes) if lanes else "per bucket", 1 if lanes else after, built))
    p.close()
