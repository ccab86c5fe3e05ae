"""Host logic of launch-per-bucket plans without a GPU: the schedule a plan dumps (bnpp_ve_plan_describe) is executed
by the CPU emulator of tests/plan_emulator.py on one flat arena array and must reproduce the reference's PR / MAR
(golden fixtures) and the oracle's values on wide synthetic networks, where small tables are folded and the arena is
reused many times over."""
import math

import numpy as np

from bnpp_b200 import model, synth
from fused_interp import DryPlan, parse_uai
from plan_emulator import describe, run

REL = 1e-9


def _order(cards, scopes, variables, ev, flag):
    if not flag:
        return [v for v in variables if v not in ev]
    return model.elim_order(cards, scopes, variables, flag, observed=sorted(ev))[0]


def test_plan_schedule_partition_golden(golden_models):
    n = 0
    for name in ["asia", "child", "alarm", "insurance", "win95pts", "hepar2", "grid3x3"]:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        for case in m["pr"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            observed = sorted(ev)
            order = _order(cards, scopes, [v for v in range(len(cards)) if v not in ev], ev, case["flag"])
            p = DryPlan(cards, scopes, observed, order)
            res, z = run(describe(p.h), tables, [ev[v] for v in observed])
            assert math.isclose(res[0], case["pr"], rel_tol=REL), (name, case["flag"], res[0], case["pr"])
            assert z is not None and math.isclose(z, res[0], rel_tol=1e-15)
            p.close()
            n += 1
    assert n >= 40


def test_plan_schedule_marginals_golden(golden_models):
    for name in ["asia", "child", "alarm", "network"]:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        for case in m["mar"][:2]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            observed = sorted(ev)
            order = _order(cards, scopes, [v for v in range(len(cards)) if v not in ev], ev, "mf")
            p = DryPlan(cards, scopes, observed, order, marginals=True)
            off, size, total = p.layout
            res, _ = run(describe(p.h), tables, [ev[v] for v in observed])
            for v, want in enumerate(case["mar"]):
                seg = res[off[v]:off[v] + size[v]]
                got = np.array([1.0]) if size[v] == 1 else seg / seg.sum()
                assert np.allclose(got, want, rtol=REL, atol=1e-300), (name, v)
            p.close()


def test_plan_schedule_wide_network_with_folding_and_arena_reuse(golden_synth):
    """the config-4 generator at reduced width (tables up to 2^21 entries: small tables are folded ahead of the wide
    launches, arena slots are reused): PR against the compiled reference's value, and Z = 1 without evidence"""
    done = 0
    for rec in golden_synth["bn"]:
        if rec["N"] > 48:
            continue
        cards, scopes, tables = parse_uai(synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        ev = {int(k): v for k, v in rec["evidence"].items()}
        observed = sorted(ev)
        for case in rec["cases"]:
            if "pr" not in case and ev:
                continue
            order = _order(cards, scopes, [v for v in range(len(cards)) if v not in ev], ev, case["flag"])
            p = DryPlan(cards, scopes, observed, order)
            d = describe(p.h)
            widest = max(int(np.prod([c for _, c in st["scope"]])) for st in d["steps"])
            if widest > (1 << 21):
                p.close()
                continue
            res, _ = run(d, tables, [ev[v] for v in observed])
            want = case.get("pr", 1.0)
            assert math.isclose(res[0], want, rel_tol=REL), (rec["N"], case["flag"], res[0], want)
            # slots are reused (lifetimes end with a dependency level; slots are 32-double granules)
            assert d["arena"] <= sum((f["size"] + 31) // 32 * 32 for f in d["factors"] if f["src"] < 0)
            p.close()
            done += 1
    assert done >= 2


def test_plan_schedule_with_small_table_folding():
    """width 22 (2^23-entry union tables): ahead of a wide launch the plan multiplies the bucket's small CPTs into one
    table and copies raw input views into canonical axis order (fold_small, ve.cu) -- product-only steps that the
    emulator executes like any other; a normalised BN must come out at Z = 1, and P(x=0) + P(x=1) = 1 for a leaf"""
    cards, scopes, tables = parse_uai(synth.random_bn_uai(56, 30, 4, 2))
    n = len(cards)
    order, width = model.elim_order(cards, scopes, list(range(n)), "mf")
    assert 18 <= width <= 26
    p = DryPlan(cards, scopes, [], order)
    d = describe(p.h)
    folds = [st for st in d["steps"] if st["elim"] < 0 and st["out"] >= 0]
    assert folds, "no product-only step: the plan did not fold anything"
    res, _ = run(d, tables, [])
    assert math.isclose(res[0], 1.0, rel_tol=REL)
    assert d["arena"] < sum(f["size"] for f in d["factors"] if f["src"] < 0)
    p.close()
    leaf = n - 1
    order1, _ = model.elim_order(cards, scopes, [v for v in range(n) if v != leaf], "mf", observed=[leaf])
    p = DryPlan(cards, scopes, [leaf], order1)
    d = describe(p.h)
    z0, _ = run(d, tables, [0])
    z1, _ = run(d, tables, [1])
    assert math.isclose(z0[0] + z1[0], 1.0, rel_tol=REL)
    p.close()


def test_levelled_schedule_keeps_results_and_cuts_launches(golden_models):
    """K10 (tasks): the steps of a plan are re-ordered by dependency level -- the small tasks of a level first, then its
    wide steps -- and the arena is laid out again with level-wide lifetimes; the re-ordered schedule must give the
    reference's PR (emulator, flat arena), every small step must sit in exactly one task, and a run needs far fewer
    launches than one per bucket"""
    for name, at_most in [("insurance", 4), ("Water", 14), ("andes", 16), ("hailfinder", 1)]:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        case = [c for c in m["pr"] if c["flag"] == "mf"][-1]
        ev = {int(k): v for k, v in case["evidence"].items()}
        observed = sorted(ev)
        order = _order(cards, scopes, [v for v in range(len(cards)) if v not in ev], ev, "mf")
        p = DryPlan(cards, scopes, observed, order)
        segs = p.segments(0)
        d = describe(p.h)
        widest = max(int(np.prod([c for _, c in st["scope"]])) for st in d["steps"])
        if widest <= (1 << 22):
            res, _ = run(d, tables, [ev[v] for v in observed])
            assert math.isclose(res[0], case["pr"], rel_tol=REL), (name, res[0], case["pr"])
        assert d["steps"][-1]["out"] < 0, "the result step stays last"
        launches, groups, levels = p.launches()
        covered = sum(e - a for a, e, *_ in segs)
        assert all(x[1] <= y[0] for x, y in zip(segs, segs[1:])), "tasks are disjoint step ranges in schedule order"
        assert launches == len(d["steps"]) - covered + groups and groups <= levels
        assert launches <= at_most < len(d["steps"]), (name, launches, groups, levels, len(d["steps"]))
        p.close()


def test_plan_schedule_shrinks_buckets_with_many_factors():
    """a hub variable in 12 factors, eliminated first: the bucket exceeds BNPP_MAX_OPERANDS, the plan multiplies the
    smallest tables together first (shrink, ve.cu) -- same value as the oracle's one-by-one products"""
    import oracle as orc
    rng = np.random.default_rng(3)
    n = 12
    cards = [3] + [2] * (n - 1)
    scopes = [[0]] + [[0, i] for i in range(1, n)] + [[i, i + 1] for i in range(1, n - 1)]
    tables = [rng.uniform(0.1, 1.5, int(np.prod([cards[v] for v in sc]))) for sc in scopes]
    m = orc.OModel("MARKOV", cards, [orc.OFactor(sc, t) for sc, t in zip(scopes, tables)])
    for ev in ({}, {5: 1}):
        order = [v for v in range(n) if v not in ev]            # the hub first
        want = orc.partition(m, ev, order)
        observed = sorted(ev)
        p = DryPlan(cards, scopes, observed, order)
        p.set_fused(False)
        d = describe(p.h)
        assert max(len(st["ops"]) for st in d["steps"]) <= 6
        assert any(st["elim"] < 0 and st["out"] >= 0 for st in d["steps"]), "no product-only step: nothing was shrunk"
        res, _ = run(d, tables, [ev[v] for v in observed])
        assert math.isclose(res[0], want, rel_tol=1e-12), (ev, res[0], want)
        p.close()


def test_plan_schedule_random_networks():
    """random mixed-cardinality factor sets, random evidence, every heuristic: the dumped schedule on the flat arena
    against the oracle's bucket elimination -- with segments off and on (re-ordered steps, arena laid out again)"""
    import random
    import oracle as orc
    rng = random.Random(4242)
    checked = 0
    for trial in range(40):
        n = rng.randint(4, 14)
        cards = [rng.randint(2, 4) for _ in range(n)]
        scopes, tables = [], []
        for _ in range(rng.randint(n, 2 * n + 3)):
            sc = rng.sample(range(n), rng.randint(1, min(4, n)))
            scopes.append(sc)
            tables.append(np.array([rng.uniform(0.05, 2.0) for _ in range(int(np.prod([cards[v] for v in sc])))]))
        ev = {v: rng.randrange(cards[v]) for v in rng.sample(range(n), rng.randint(0, n // 3))}
        m = orc.OModel("MARKOV", cards, [orc.OFactor(sc, t) for sc, t in zip(scopes, tables)])
        observed = sorted(ev)
        variables = [v for v in range(n) if v not in ev]
        for flag in ("", "mf", "md", "wmf"):
            order = _order(cards, scopes, variables, ev, flag)
            want = orc.partition(m, ev, order)
            for segments in (False, True):
                p = DryPlan(cards, scopes, observed, order)
                if segments:
                    p.segments(0)
                d = describe(p.h)
                if not d["steps"] or d["steps"][-1]["out"] >= 0:      # nothing left to multiply: the result is the scalar 1
                    assert math.isclose(want, 1.0, rel_tol=1e-12) or not d["steps"]
                    p.close()
                    continue
                res, _ = run(d, tables, [ev[v] for v in observed])
                assert math.isclose(res[0], want, rel_tol=1e-12), (trial, flag, segments, res[0], want)
                p.close()
                checked += 1
    assert checked >= 250
