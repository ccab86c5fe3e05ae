"""K9 without a GPU: the step program a VE plan compiles itself into for the one-launch kernel
(`ve_fused`, bnpp_b200/csrc/fused.cu) is built on a DRY plan and executed by the CPU interpreter
of tests/fused_interp.py; the results must equal the reference's PR / MAR values (golden
fixtures dumped from the compiled reference) within 1e-9.  This pins the host half of K9 -- the
depth-first step order, the shared-memory arena layout, operand-offset tables, evidence folded
into CPT base offsets; the device half is pinned by tests/test_gpu_fused.py."""
import math

import numpy as np
import pytest

from bnpp_b200 import model, synth
from fused_interp import DryPlan, interpret, parse_uai

REL = 1e-9
SMALL = ["asia", "asia_positive", "cancer", "earthquake", "child", "alarm", "insurance", "grid3x3"]


def _order(cards, scopes, variables, ev, flag):
    if not flag:
        return [v for v in variables if v not in ev]
    return model.elim_order(cards, scopes, variables, flag, observed=sorted(ev))[0]


def test_fused_program_partition(golden_models):
    n = 0
    for name in SMALL:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        for case in m["pr"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            variables = [v for v in range(len(cards)) if v not in ev]
            order = _order(cards, scopes, variables, ev, case["flag"])
            if case["flag"]:
                assert order == case["order"]
            observed = sorted(ev)
            p = DryPlan(cards, scopes, observed, order)
            G, arena, n_steps = p.fused_info(1)
            if G == 0:
                p.close()
                continue
            assert G in (32, 128)
            prog, tab = p.program(1)
            res, z = interpret(prog, tab, n_steps, arena, tables, [ev[v] for v in observed], 1)
            assert math.isclose(res[0], case["pr"], rel_tol=REL), (name, case["flag"], res[0], case["pr"])
            assert z == res[0]
            p.close()
            n += 1
    assert n >= 40


def test_fused_program_marginals(golden_models):
    n = 0
    for name in ["asia", "child", "alarm", "grid3x3"]:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        for case in m["mar"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            variables = [v for v in range(len(cards)) if v not in ev]
            order = _order(cards, scopes, variables, ev, "mf")
            observed = sorted(ev)
            p = DryPlan(cards, scopes, observed, order, marginals=True)
            G, arena, n_steps = p.fused_info(1)
            if G == 0:
                p.close()
                continue
            off, size, total = p.layout
            prog, tab = p.program(1)
            res, _ = interpret(prog, tab, n_steps, arena, tables, [ev[v] for v in observed], total)
            for v, want in enumerate(case["mar"]):
                seg = res[off[v]:off[v] + size[v]]
                got = np.array([1.0]) if size[v] == 1 else seg / seg.sum()      # normalize_segments (elementwise.cu)
                assert len(got) == len(want), (name, v)
                assert np.allclose(got, want, rtol=REL, atol=1e-300), (name, v, got, want)
            p.close()
            n += 1
    assert n >= 4


def test_fused_program_batch(golden_synth):
    """config 5's network: one program for every evidence set, only the evidence values differ"""
    for rec in golden_synth["batch"]:
        if not rec["fixed_ids"]:
            continue
        cards, scopes, tables = parse_uai(synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        evs = synth.evidence_batch(rec["N"], rec["nobs"], rec["nsets"], seed=5, fixed_ids=True)
        observed = sorted(evs[0])
        variables = [v for v in range(len(cards)) if v not in evs[0]]
        order = _order(cards, scopes, variables, evs[0], "mf")
        p = DryPlan(cards, scopes, observed, order)
        G, arena, n_steps = p.fused_info(4096)
        assert G in (8, 16, 32), "config 5 must run fused"
        assert arena * 8 * (32 // G) <= 14080
        prog, tab = p.program(4096)
        for i in range(0, len(evs), 4 if rec["N"] > 100 else 1):
            res, _ = interpret(prog, tab, n_steps, arena, tables, [evs[i][v] for v in observed], 1)
            assert math.isclose(res[0], rec["pr"][i], rel_tol=REL), (rec["N"], i)
        # the same plan, one launch per bucket: not fused when asked
        p.set_fused(False)
        assert p.fused_info(4096)[0] == 0
        p.close()


def test_wide_plans_are_not_fused():
    """config 4 class networks keep one launch per bucket (their tables do not fit shared memory)"""
    scopes, _ = synth.random_bn_scopes(40, 24, 3, 3)
    cards = [2] * 40
    order = model.elim_order(cards, scopes, list(range(40)), "mf")[0]
    p = DryPlan(cards, scopes, [], order)
    assert p.fused_info(1)[0] == 0 and p.fused_info(1024)[0] == 0
    p.close()


def test_fused_program_random_mixed_cardinalities():
    """random factor sets over variables of cardinality 2..5, random evidence, every ordering heuristic:
    the interpreted program against the plain-C oracle's bucket elimination (oracle.partition); covers the
    non-pair entry path (cardinality > 2, CPT-only buckets, scalar factors, variables no factor mentions)"""
    import random
    import oracle as orc
    rng = random.Random(20261018)
    checked = 0
    for trial in range(40):
        n = rng.randint(3, 12)
        cards = [rng.randint(2, 5) for _ in range(n)]
        scopes, tables = [], []
        for _ in range(rng.randint(n, 2 * n)):
            w = rng.randint(1, min(4, n))
            sc = rng.sample(range(n), w)
            size = int(np.prod([cards[v] for v in sc]))
            scopes.append(sc)
            tables.append(np.array([rng.uniform(0.05, 2.0) for _ in range(size)]))
        ev = {v: rng.randrange(cards[v]) for v in rng.sample(range(n), rng.randint(0, n // 3))}
        m = orc.OModel("MARKOV", cards, [orc.OFactor(sc, t) for sc, t in zip(scopes, tables)])
        observed = sorted(ev)
        variables = [v for v in range(n) if v not in ev]
        for flag in ("", "mf", "md", "wmf"):
            order = _order(cards, scopes, variables, ev, flag)
            want = orc.partition(m, ev, order)
            p = DryPlan(cards, scopes, observed, order)
            G, arena, n_steps = p.fused_info(1)
            if G == 0:          # every factor fully observed: the result is a product of scalars, handled outside K9
                p.close()
                continue
            prog, tab = p.program(1)
            res, z = interpret(prog, tab, n_steps, arena, tables, [ev[v] for v in observed], 1)
            assert math.isclose(res[0], want, rel_tol=1e-12), (trial, flag, res[0], want)
            # a batch of the same plan runs the same program (when its arena leaves room for enough CTAs per SM)
            if p.fused_info(512)[0]:
                prog_b, tab_b = p.program(512)
                assert np.array_equal(prog, prog_b) and np.array_equal(tab, tab_b)
            p.close()
            checked += 1
    assert checked >= 120


def test_plan_with_large_variable_ids():
    """variable ids are arbitrary uint32 at the ABI: ids far beyond the dense rank table plan, order and
    evaluate like the same network with small ids"""
    import oracle as orc
    cards_small = [2, 3, 2, 2]
    scopes_small = [[0], [1, 0], [2, 1], [3, 2, 0]]
    rng = np.random.default_rng(7)
    tables = [rng.uniform(0.1, 1.0, int(np.prod([cards_small[v] for v in sc]))) for sc in scopes_small]
    want = orc.partition(orc.OModel("MARKOV", cards_small, [orc.OFactor(sc, t) for sc, t in zip(scopes_small, tables)]), {2: 1},
                         [0, 1, 3])
    base = (1 << 23) + 5                                  # beyond RankMap's dense range (ve.cu)
    ids = {v: base + 1000 * v for v in range(4)}
    scopes = [[ids[v] for v in sc] for sc in scopes_small]

    class Cards(dict):                                    # model._scopes indexes cards by variable id
        pass
    cards = Cards({ids[v]: cards_small[v] for v in range(4)})
    p = DryPlan(cards, scopes, [ids[2]], [ids[0], ids[1], ids[3]])
    G, arena, n_steps = p.fused_info(1)
    assert G > 0
    prog, tab = p.program(1)
    res, _ = interpret(prog, tab, n_steps, arena, tables, [1], 1)
    assert math.isclose(res[0], want, rel_tol=1e-12)
    p.close()


def test_fused_marginals_random_networks():
    """the all-marginals (bucket-tree) plan as a fused program on random mixed-cardinality networks with evidence,
    against the brute-force joint marginals of the oracle (code/model.cpp:69-101)"""
    import random
    import oracle as orc
    rng = random.Random(77)
    checked = 0
    for trial in range(25):
        n = rng.randint(3, 8)
        cards = [rng.randint(2, 4) for _ in range(n)]
        scopes, tables = [[v] for v in range(n)], []          # every variable is mentioned by a unary factor
        for _ in range(rng.randint(n - 1, 2 * n)):
            scopes.append(rng.sample(range(n), rng.randint(2, min(3, n))))
        for sc in scopes:
            tables.append(np.array([rng.uniform(0.05, 2.0) for _ in range(int(np.prod([cards[v] for v in sc])))]))
        ev = {v: rng.randrange(cards[v]) for v in rng.sample(range(n), rng.randint(0, 2))}
        m = orc.OModel("MARKOV", cards, [orc.OFactor(sc, t) for sc, t in zip(scopes, tables)])
        want = orc.joint_marginals(m, ev)
        observed = sorted(ev)
        variables = [v for v in range(n) if v not in ev]
        for flag in ("mf", "md"):
            order = _order(cards, scopes, variables, ev, flag)
            p = DryPlan(cards, scopes, observed, order, marginals=True)
            G, arena, n_steps = p.fused_info(1)
            assert G > 0
            off, size, total = p.layout
            prog, tab = p.program(1)
            res, _ = interpret(prog, tab, n_steps, arena, tables, [ev[v] for v in observed], total)
            for v in range(n):
                if v in ev:
                    assert size[v] == 1
                    continue
                seg = res[off[v]:off[v] + size[v]]
                assert np.allclose(seg / seg.sum(), want[v].values, rtol=1e-10, atol=0.0), (trial, flag, v)
            p.close()
            checked += 1
    assert checked == 50


def test_segment_programs_chain_through_the_global_arena(golden_models, golden_synth):
    """EXPERIMENTAL (DESIGN gap 4): a plan cut into fused segments of at most m steps -- each reads what earlier
    segments wrote to the plan's global arena and writes what later ones read -- evaluates to the reference's PR for
    every cut size; m = 1 is the degenerate case "every step its own launch"."""
    n = 0
    for name in ["asia", "child", "alarm", "hepar2", "win95pts"]:
        m = golden_models[name]
        cards, scopes, tables = parse_uai(m["uai"])
        for case in m["pr"]:
            if not case["flag"]:
                continue
            ev = {int(k): v for k, v in case["evidence"].items()}
            variables = [v for v in range(len(cards)) if v not in ev]
            order = _order(cards, scopes, variables, ev, case["flag"])
            observed = sorted(ev)
            p = DryPlan(cards, scopes, observed, order)
            for cut in (1, 2, 3, 7, 1000):
                segs = p.segments(cut)
                assert segs and segs[0][0] == 0 and segs[-1][1] == p.n_steps()
                assert all(a[1] == b[0] for a, b in zip(segs, segs[1:])), "every step in exactly one segment"
                glob, result, z = {}, np.full(1, np.nan), None
                for first, end, lanes, arena, prog, tab in segs:
                    assert end - first <= cut and lanes in (32, 128)
                    res, zz = interpret(prog, tab, end - first, arena, tables, [ev[v] for v in observed], 1, glob=glob, result=result)
                    z = zz if zz is not None else z
                assert math.isclose(result[0], case["pr"], rel_tol=REL), (name, case["flag"], cut, result[0])
                assert z == result[0]
                if cut == 1000:
                    assert len(segs) == 1 and not glob          # one segment = the whole plan: nothing goes through global memory
                n += 1
            p.close()
    assert n >= 100


def test_segments_of_a_mixed_plan():
    """a plan with wide steps keeps them as their own launches; the runs of small steps around them become segments"""
    scopes, _ = synth.random_bn_scopes(40, 24, 3, 3)            # min-fill width 15: buckets from 2^2 to 2^16 entries
    cards = [2] * 40
    order = model.elim_order(cards, scopes, list(range(40)), "mf")[0]
    p = DryPlan(cards, scopes, [], order)
    assert p.fused_info(1)[0] == 0
    segs = p.segments(0)
    steps = p.n_steps()
    covered = sum(end - first for first, end, *_ in segs)
    assert segs and 0 < covered < steps
    assert all(lanes == 128 for _, _, lanes, *_ in segs)
    p.close()
