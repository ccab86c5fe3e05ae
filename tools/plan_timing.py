#!/usr/bin/env python
"""Host-side cost of an end-to-end query: ordering and planning, per config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnpp_b200 import capi, model, synth
ctx = capi.Context(0)
for name, (N, W, K, seed), nobs in [("config4", (64, 40, 4, 5), 0), ("config5", (500, 6, 3, 11), 20), ("andes-like", (220, 12, 3, 7), 0)]:
    _, bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))
    ev = synth.evidence_batch(N, nobs, 1, seed=5)[0] if nobs else {}
    obs = sorted(ev)
    variables = [v for v in range(N) if v not in ev]
    for rep in range(3):
        t0 = time.perf_counter()
        order, w = bn.order(variables, ev, "mf")
        t1 = time.perf_counter()
        p = model.VEPlan(ctx, bn.cards, bn.scopes, obs, order, _arr=bn._scope_arr)
        t2 = time.perf_counter()
        p.close()
        t3 = time.perf_counter()
    print(name, "order %.3f ms  plan %.3f ms (%d launches)  destroy %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, p.n_launches, (t3 - t2) * 1e3))
