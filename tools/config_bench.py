#!/usr/bin/env python
"""Latency-bound BASELINE configs 1-3 on one GPU: asia PR/MAR (VE, -mf), grid3x3 PR/MAR, and the
40x40 Ising sum-product (sweeps, ms).  Prints one JSON object.   python tools/config_bench.py"""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bnpp_b200 import capi, model, synth  # noqa: E402

def run(ctx, reps=50):
    """-> dict of latencies / throughputs of BASELINE configs 1-3 on one GPU"""
    G = json.load(gzip.open(os.path.join(ROOT, "tests", "golden", "models.json.gz"), "rt"))
    out = {}

    def timed(fn, n=reps):
        fn()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(n):
            r = fn()
        ctx.sync()
        return (time.perf_counter() - t0) * 1e3 / n, r

    _, asia = model.from_uai_text(ctx, G["asia"]["uai"])
    ev = {0: 1, 2: 1}
    ms, (z, _) = timed(lambda: asia.partition(ev, "mf"))
    out["asia_pr_mf"] = {"ms": ms, "queries_per_s": 1e3 / ms, "Z": z}
    ms, mar = timed(lambda: asia.marginals(ev, None), max(2, reps // 3))
    out["asia_mar"] = {"ms": ms, "P(x7=0|e)": float(mar[7][0])}
    ms, mar = timed(lambda: asia.marginals_fast(ev, "mf"), max(2, reps // 3))
    out["asia_mar_bucket_tree"] = {"ms": ms, "P(x7=0|e)": float(mar[7][0])}
    asia.close()
    _, net = model.from_uai_text(ctx, G["network"]["uai"])
    ms, mar = timed(lambda: net.marginals_fast({}, "mf"), max(2, reps // 5))
    out["network120_mar_bucket_tree_mf"] = {"ms": ms, "P(x0=0)": float(mar[0][0])}
    net.close()
    _, grid = model.from_uai_text(ctx, G["grid3x3"]["uai"])
    ms, (z, _) = timed(lambda: grid.partition({0: 1, 4: 1, 5: 1}, "mf"))
    out["grid3x3_pr_mf"] = {"ms": ms, "Z": z}
    ms, mar = timed(lambda: grid.marginals({0: 1, 3: 1, 4: 1}, "mf"), max(2, reps // 3))
    out["grid3x3_mar_mf"] = {"ms": ms}
    grid.close()
    for J in (0.3, 0.5, 1.0):
        _, ising = model.from_uai_text(ctx, synth.ising_uai(40, 0.5, J, 7))
        t0 = time.perf_counter()
        fg, sweeps = ising.sum_product()
        mar = fg.marginals()
        ms_total = (time.perf_counter() - t0) * 1e3
        ms_update, _ = timed(lambda: (fg.reset(), fg.update(10000, 0.001)), 5)
        fg.close()
        ising.close()
        out["ising40_J%g" % J] = {"sweeps": sweeps, "ms_total_incl_setup": ms_total, "ms_update": ms_update,
                                 "us_per_sweep": ms_update * 1e3 / (sweeps + 1),
                                 "message_updates_per_s": 15680 * (sweeps + 1) / ms_update * 1e3, "P(x0=0)": float(mar[0][0])}
    return out


if __name__ == "__main__":
    print(json.dumps(run(capi.Context(0)), indent=1))
