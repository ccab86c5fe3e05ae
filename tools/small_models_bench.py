#!/usr/bin/env python
"""Configs 1-2 and the shipped toy networks: PR (min-fill) latency of ONE query, K9 (the whole plan
in one launch, `ve_fused`) against one launch per bucket.  `first` = first run of a fresh plan
(what a CLI invocation pays: program upload / launch resolution included), `replay` = a later run
of the same plan (CUDA graph for the launch-per-bucket path), both host wall time up to the
partition being back on the host.  JSON on stdout."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bnpp_b200 import capi, model  # noqa: E402

G = json.load(gzip.open(os.path.join(ROOT, "tests", "golden", "models.json.gz"), "rt"))
ctx = capi.Context(0)
out = []
for name in ["asia", "grid3x3", "child", "alarm", "win95pts", "hepar2", "hailfinder", "network"]:
    _, bn = model.from_uai_text(ctx, G[name]["uai"])
    bn.partition({}, "mf")          # context warm-up, ordering cache
    rec = {"network": name, "variables": bn.nvars}
    for fused in (1, 0):
        first, replay = [], []
        for rep in range(5):
            bn.drop_plans()
            order, _ = bn.order(list(range(bn.nvars)), {}, "mf")
            p = bn.plan([], order)
            p.set_fused(fused)
            ctx.sync()
            t0 = time.perf_counter()
            z, _ = bn.partition({}, "mf")
            first.append((time.perf_counter() - t0) * 1e3)
            for _ in range(3):
                bn.partition({}, "mf")
            t0 = time.perf_counter()
            for _ in range(20):
                z, _ = bn.partition({}, "mf")
            replay.append((time.perf_counter() - t0) * 1e3 / 20)
        key = "one_launch" if fused else "per_bucket"
        rec[key] = {"first_ms": min(first), "replay_ms": min(replay), "Z": z, "lanes": p.fused_info(1)[0] if fused else 0,
                    "launches": 1 if (fused and p.fused_info(1)[0]) else p.n_launches}
    out.append(rec)
    bn.close()
print(json.dumps(out))
