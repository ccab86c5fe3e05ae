"""The multi-valued elimination kernel (contract_mv, csrc/contract_mv.cu): elimination of a variable with more than
two values -- `prod *= *pf` over a bucket then `prod.sum_out(var)`, code/model.cpp:414-418, code/factor.cpp:117-147,
182-212 -- through the C ABI, against the dumps of the compiled reference and the CPU oracle.  Bit-exact values."""
import math
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as orc  # noqa: E402

ZREL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(params=["staged", "gather"])
def mv_always(request):
    """every multi-valued step through the tiled kernels, whatever its size (the product's default starts at 2^15
    entries): once with the TMA-staged variant allowed (contract_mvt, the default), once gather only (contract_mv)"""
    from bnpp_b200 import capi
    old = capi.tuning_set("mv_min_entries", 0)
    capi.tuning_set("mv_staged", 1 if request.param == "staged" else 0)
    yield request.param
    capi.tuning_set("mv_min_entries", old)
    capi.tuning_set("mv_emax", 0)
    capi.tuning_set("mv_staged", 1)


def zclose(a, b):
    return a == b or math.isclose(a, b, rel_tol=ZREL, abs_tol=1e-300)


def rand_factor(ctx, rng, scope, cards):
    from bnpp_b200.factor import DeviceFactor
    n = int(np.prod([cards[v] for v in scope])) if scope else 1
    vals = np.random.default_rng(rng.randrange(1 << 30)).uniform(0.1, 1.0, n)
    of = orc.OFactor(scope, vals)
    return of, DeviceFactor.from_host(ctx, scope, [cards[v] for v in scope], vals, of.partition)


def test_reference_sum_out_dumps(ctx, golden_ops, mv_always):
    """Factor::sum_out of the reference's own product tables (code/factor.cpp:182-212), cards up to 7"""
    from bnpp_b200.factor import DeviceFactor
    seen = 0
    for c in golden_ops:
        cards = c["cards"]
        p = DeviceFactor.from_host(ctx, c["p"]["scope"], [cards[v] for v in c["p"]["scope"]], c["p"]["values"], c["p"]["partition"])
        s = p.sum_out(c["sum_var"])
        assert s.scope == c["s"]["scope"] and np.array_equal(s.values(), np.array(c["s"]["values"]))
        assert zclose(s.partition, c["s"]["partition"])
        seen += ("contract_mvt" if mv_always == "staged" else "contract_mv<") in ctx.last_launch()[0]
    assert seen >= 15, seen


@pytest.mark.parametrize("emax", [0, 64, 200, 4096])
def test_fused_step_random_multivalued(ctx, mv_always, emax):
    """k-ary product -> sum_out on random scopes, cards 2..9, random output axis orders, tile sizes from tiny to the
    largest: bit-exact against the oracle"""
    from bnpp_b200 import capi
    from bnpp_b200.factor import fused_product_sum_out
    capi.tuning_set("mv_emax", emax)
    rng = random.Random(100 + emax)
    used = 0
    for case in range(100):
        nvars = rng.randint(1, 7)
        cards = [rng.choice([2, 3, 3, 4, 5, 6, 7, 9]) for _ in range(nvars)]
        k = rng.randint(1, 6)
        ofs, dfs = [], []
        for _ in range(k):
            w = rng.randint(0, nvars)
            sc = rng.sample(range(nvars), w)
            o, d = rand_factor(ctx, rng, sc, cards)
            ofs.append(o)
            dfs.append(d)
        union = []
        for o in ofs:
            union += [v for v in o.scope if v not in union]
        if not union:
            continue
        elim = rng.choice(union)
        out_scope = [v for v in union if v != elim]
        rng.shuffle(out_scope)
        want = orc.product_sum_out(ofs, out_scope, elim, cards)
        got = fused_product_sum_out(ctx, dfs, out_scope, elim)
        name = ctx.last_launch()[0]
        assert np.array_equal(got.values(), want.values), (case, name)
        assert zclose(got.partition, want.partition), (case, name)
        used += ("contract_mvt" if mv_always == "staged" else "contract_mv<") in name
    assert used >= 15, used


@pytest.mark.parametrize("staged", [1, 0])
@pytest.mark.parametrize("card,elim_pos", [(3, "last"), (4, "last"), (5, "mid"), (7, "last"), (7, "first"), (8, "mid"), (21, "last")])
def test_wide_multivalued_step(ctx, card, elim_pos, staged):
    """a step of ~2^21 union entries at the product's default settings (the kernel the plan picks on Munin / Link /
    Barley sized buckets): canonical layout (eliminated variable innermost in both operands), eliminated variable in
    the middle and as the leading axis; against the oracle entry by entry"""
    from bnpp_b200 import capi
    from bnpp_b200.factor import fused_product_sum_out
    capi.tuning_set("mv_staged", staged)
    rng = random.Random(card * 7 + len(elim_pos))
    nv = 10
    cards = [5, 3, 4, 2, 6, 3, 2, 4, 5, 3]
    x = {"last": nv - 1, "mid": 3, "first": 0}[elim_pos]
    cards[x] = card
    sa = list(range(nv))
    sb = [v for v in sa if v not in (1, 4)] if x not in (1, 4) else [v for v in sa if v not in (2, 5)]
    oa, da = rand_factor(ctx, rng, sa, cards)
    ob, db = rand_factor(ctx, rng, sb, cards)
    out_scope = [v for v in sa if v != x]
    want = orc.product_sum_out([oa, ob], out_scope, x, cards)
    got = fused_product_sum_out(ctx, [da, db], out_scope, x)
    capi.tuning_set("mv_staged", 1)
    name = ctx.last_launch()[0]
    if elim_pos == "last":          # the eliminated variable is the operands' fastest axis: the tiled kernels
        assert ("contract_mvt" if staged else "contract_mv<") in name, name
    else:                           # a slow axis: one output entry per thread already reads whole rows (contract_generic)
        assert "contract_generic" in name or "contract_mv" in name, name
    assert np.array_equal(got.values(), want.values)
    assert zclose(got.partition, want.partition)


def test_multivalued_networks_pr(ctx, golden_models, mv_always):
    """PR on the shipped multi-valued networks with every multi-valued bucket through contract_mv, plan replayed
    (second run = CUDA graph with shared-memory launch nodes): the reference's values at 1e-9"""
    from bnpp_b200 import model
    for name in ["Water", "insurance", "hepar2", "alarm"]:
        if name not in golden_models:
            continue
        m = golden_models[name]
        bn = model.from_uai_text(ctx, m["uai"])[1]
        for case in m["pr"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            variables = [v for v in range(bn.nvars) if v not in ev]
            order, _ = bn.order(variables, ev, case["flag"] or None)
            p = bn.plan(sorted(ev), order)
            p.set_fused(False)
            for _ in range(3):
                z, _ = bn.partition(ev, case["flag"] or None)
                assert math.isclose(z, case["pr"], rel_tol=1e-9), (name, case["flag"], z, case["pr"])
        bn.close()
