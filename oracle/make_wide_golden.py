#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- reference-computed PR for the WIDE networks of config 4 (tests/golden/wide.json).

The unmodified reference (oracle/_ref/ref_harness: BN::partition, code/model.cpp:275-294, min-fill) is run ONCE on
every problem a rank of `bench.py --gpus N` solves: the network of bnpp_b200.synth.WIDE[N] with its base evidence
(observed leaves) plus one assignment of the log2(N) shard variables.  Each such run is a width-27 query of ~2e9
union entries: 15-35 minutes and ~7 GiB in the reference.  The sum over the assignments is the PR of the unsharded
width-(27+log2 N) network (the reference itself cannot hold its 2^29..2^31-entry tables), and sums over the
assignments that agree with a coarser sharding give the partials of the strong-scaling series.

Any cutset is valid for this decomposition: P(e) = sum_x P(e, X=x).  The shard variables recorded here are the ones
bnpp_b200.sharding.pick_shard_vars chooses from the reference's own min-fill order (the harness's `order` command).

    nice -n 19 python oracle/make_wide_golden.py [--procs 5] [--only 1,2,4,8]

Results are appended to tests/golden/wide.json as they arrive (a rerun skips what is there).
"""
import argparse
import concurrent.futures
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from bnpp_b200 import sharding, synth  # noqa: E402  (pure-Python helpers: generator and cutset choice)
import oracle as orc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "wide.json")
TMP = "/tmp/bnpp_golden"


def evidset(ev):
    return "evidset %d %s" % (len(ev), " ".join("%d %d" % kv for kv in sorted(ev.items())))


def load():
    return json.load(open(OUT)) if os.path.exists(OUT) else {"networks": {}}


def save(g):
    tmp = OUT + ".tmp"
    with open(tmp, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    os.replace(tmp, OUT)


def one(job):
    path, ev = job
    rows = orc.RefHarness().run(["model " + path, evidset(ev), "opt mf", "pr", "order"], timeout=6 * 3600)
    pr = [r for r in rows if r[0] == "PR"][0]
    o = [r for r in rows if r[0] == "ORDER"][0]
    return {"assign": None, "pr": float(pr[1]), "ref_ms": float(pr[2]), "width": int(o[1]), "order": [int(x) for x in o[3:]]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=5)
    ap.add_argument("--only", default="1,2,4,8")
    args = ap.parse_args()
    os.makedirs(TMP, exist_ok=True)
    g = load()
    jobs = []
    for n in [int(x) for x in args.only.split(",")]:
        N, W, K, seed, ev = synth.wide_bn(n)
        text = synth.random_bn_uai(N, W, K, seed)
        path = os.path.join(TMP, "wide_%d.uai" % n)
        with open(path, "w") as f:
            f.write(text)
        rows = orc.RefHarness().run(["model " + path, evidset(ev), "opt mf", "order"])
        o = [r for r in rows if r[0] == "ORDER"][0]
        order, width = [int(x) for x in o[3:]], int(o[1])
        scopes, _ = synth.random_bn_scopes(N, W, K, seed)
        live = [[v for v in sc if v not in ev] for sc in scopes]
        gbits = n.bit_length() - 1
        shard_vars = sharding.pick_shard_vars(live, order, gbits)
        rec = g["networks"].setdefault(str(n), {})
        rec.update({"N": N, "W": W, "K": K, "seed": seed, "sha256": hashlib.sha256(text.encode()).hexdigest(),
                    "evidence": {str(k): v for k, v in ev.items()}, "width": width, "order": order,
                    "shard_vars": shard_vars})
        rec.setdefault("partials", {})
        for r in range(n):
            sev = sharding.shard_evidence(shard_vars, r)
            key = ",".join("%d=%d" % kv for kv in sorted(sev.items()))
            if key in rec["partials"]:
                continue
            full = dict(ev)
            full.update(sev)
            jobs.append((n, key, (path, full)))
    save(g)
    print("%d reference runs to do" % len(jobs), flush=True)
    with concurrent.futures.ThreadPoolExecutor(args.procs) as ex:
        futs = {ex.submit(one, j[2]): j for j in jobs}
        for fu in concurrent.futures.as_completed(futs):
            n, key, _ = futs[fu]
            res = fu.result()
            res["assign"] = key
            g = load()
            g["networks"][str(n)]["partials"][key] = res
            save(g)
            print("network %d [%s]: PR = %.17g  width %d  %.1f s" % (n, key, res["pr"], res["width"], res["ref_ms"] / 1e3), flush=True)
    g = load()
    for n, rec in g["networks"].items():
        if len(rec["partials"]) == int(n):
            rec["pr"] = sum(sorted(p["pr"] for p in rec["partials"].values()))
    save(g)


if __name__ == "__main__":
    main()
