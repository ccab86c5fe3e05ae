"""K9 on the device: a VE plan whose steps are all small runs as ONE launch (`ve_fused`,
bnpp_b200/csrc/fused.cu), the intermediates of an evidence set in shared memory.  Parity bar:
bit-identical to the launch-per-bucket path (same operand order, same summation order per
entry), which the other GPU tests pin to the reference at 1e-9 -- and directly against the
reference's golden PR / MAR values here."""
import math
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from bnpp_b200 import synth  # noqa: E402

REL = 1e-9
FUSED_MODELS = ["asia", "asia_positive", "cancer", "earthquake", "child", "alarm", "win95pts", "hailfinder", "hepar2",
                "grid3x3", "network"]


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def load(ctx, text):
    from bnpp_b200 import model
    return model.from_uai_text(ctx, text)[1]


def _pr_both(bn, ev, flag):
    """PR through the fused launch and through one launch per bucket -> (z_fused, z_buckets, launches_fused, lanes)"""
    ctx = bn.ctx
    variables = [v for v in range(bn.nvars) if v not in ev]
    order, _ = bn.order(variables, ev, flag)
    p = bn.plan(sorted(ev), order)
    p.set_fused(True)
    lanes = p.fused_info(1)[0]
    ctx.sync()
    l0 = ctx.launches
    zf, _ = bn.partition(ev, flag)
    lf = ctx.launches - l0
    p.set_fused(False)
    zb, _ = bn.partition(ev, flag)
    p.set_fused(True)
    return zf, zb, lf, lanes


def test_single_query_one_launch(ctx, golden_models):
    """configs 1-2 and the shipped toy networks: PR in one launch == one launch per bucket (bit for bit)
    == the reference (1e-9), binary and multi-valued variables, with and without evidence"""
    n = 0
    for name in FUSED_MODELS:
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        for case in m["pr"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            zf, zb, lf, lanes = _pr_both(bn, ev, case["flag"] or None)
            assert zf == zb, (name, case["flag"], zf, zb)
            assert math.isclose(zf, case["pr"], rel_tol=REL), (name, case["flag"], zf, case["pr"])
            if lanes == 0:          # the file order of the variables can make a wide plan: one launch per bucket
                assert not case["flag"], (name, case["flag"])
                continue
            assert lanes in (32, 128), (name, lanes)
            assert lf == 1, (name, lf)
            n += 1
        bn.close()
    assert n >= 50


def test_marginals_one_launch(ctx, golden_models):
    """the two-pass bucket-tree plan (all marginals) as one launch: equal to its launch-per-bucket run bit for bit"""
    for name in ["asia", "child", "alarm", "hepar2", "network"]:
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        for case in m["mar"][:2]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            ctx.sync()
            l0 = ctx.launches
            fused = bn.marginals_fast(ev, "mf")
            lf = ctx.launches - l0
            assert lf == 2, (name, lf)               # the plan + normalize_segments
            for p in bn._plans.values():
                p.set_fused(False)
            plain = bn.marginals_fast(ev, "mf")
            for p in bn._plans.values():
                p.set_fused(True)
            for v, (a, b, want) in enumerate(zip(fused, plain, case["mar"])):
                assert np.array_equal(a, b), (name, v, a, b)
                assert np.allclose(a, want, rtol=REL, atol=1e-300), (name, v)
        bn.close()


def test_result_table_one_launch(ctx, golden_models):
    """VE that keeps variables (query_ve's numerator, code/model.cpp:225-238): the result table and its
    partition from the fused launch"""
    for name in ["asia", "alarm", "child"]:
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        keep = [1, 3]
        order, _ = bn.order([v for v in range(bn.nvars) if v not in keep], {}, "mf")
        p = bn.plan([], order)
        assert p.fused_info(1)[0] > 0
        scope, cards, res_f = bn.variable_elimination({}, order)
        ctx.sync()
        res_f = res_f.cpu().numpy()
        p.set_fused(False)
        _, _, res_b = bn.variable_elimination({}, order)
        ctx.sync()
        res_b = res_b.cpu().numpy()
        assert scope == keep
        assert np.array_equal(res_f[:-1], res_b[:-1]), name
        want = [c["pr"] for c in m["pr"] if not c["evidence"]][0]       # sum of the joint marginal = P() of the reference
        assert math.isclose(res_f[-1], res_b[-1], rel_tol=1e-14) and math.isclose(res_f[-1], want, rel_tol=REL)
        bn.close()


def _batch(bn, observed, evs, fused, lanes=None):
    import torch
    vals = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    variables = [v for v in range(bn.nvars) if v not in set(observed)]
    p = bn.plan(observed, bn.order(variables, observed, "mf")[0])
    p.set_fused(fused)
    if lanes:
        os.environ["BNPP_FUSED_G"] = str(lanes)
    try:
        bn.ctx.sync()
        l0 = bn.ctx.launches
        z = bn.partition_batch(observed, vals, "mf")
        bn.ctx.sync()
        return z.cpu().numpy().copy(), bn.ctx.launches - l0
    finally:
        os.environ.pop("BNPP_FUSED_G", None)
        p.set_fused(True)


def test_batch_one_launch(ctx, golden_synth):
    """config 5: every evidence set of the batch in ONE launch; equal bit for bit to one launch per bucket
    and to the per-query path, equal to the reference's per-set PR (1e-9); every group width; a ragged tail"""
    for rec in golden_synth["batch"]:
        if not rec["fixed_ids"]:
            continue
        bn = load(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        nsets = 1027 if rec["N"] > 100 else 259            # not a multiple of the sets per CTA
        evs = synth.evidence_batch(rec["N"], rec["nobs"], nsets, seed=5, fixed_ids=True)
        observed = sorted(evs[0])
        zf, lf = _batch(bn, observed, evs, True)
        assert lf == 2, lf          # the evidence range check (sanitize_evidence) + ONE ve_fused launch
        zb, lb = _batch(bn, observed, evs, False)
        assert lb > 1
        assert np.array_equal(zf, zb), rec["N"]
        for i in range(rec["nsets"]):                      # the first sets are the golden ones (same generator, same seed)
            assert math.isclose(zf[i], rec["pr"][i], rel_tol=REL), (rec["N"], i)
        for lanes in (8, 16, 32, 128):
            zl, ll = _batch(bn, observed, evs, True, lanes)
            assert ll == 2           # evidence check + the fused launch
            assert np.array_equal(zl, zf), (rec["N"], lanes)
        rng = random.Random(3)
        for i in rng.sample(range(nsets), 5) + [0, nsets - 1]:
            zi, _ = bn.partition(evs[i], "mf")
            assert zi == zf[i], (rec["N"], i)
        bn.close()


def test_batch_large_one_launch(ctx):
    """16 384 sets on the 500-variable network: persistent CTAs loop over groups of sets"""
    N, nobs, nsets = 500, 20, 16384
    bn = load(ctx, synth.random_bn_uai(N, 6, 3, 11))
    evs = synth.evidence_batch(N, nobs, nsets, seed=5, fixed_ids=True)
    observed = sorted(evs[0])
    zf, lf = _batch(bn, observed, evs, True)
    zb, _ = _batch(bn, observed, evs, False)
    assert lf == 2               # evidence check + the fused launch
    assert np.array_equal(zf, zb)
    assert np.all(zf > 0) and np.all(zf < 1)
    bn.close()


def test_other_tables_other_evidence_same_program(ctx, golden_synth):
    """one fused plan serves other evidence values (base offsets) and other table addresses (the program is
    re-pointed, not rebuilt): a second copy of the model with its first CPT doubled gives exactly 2 Z"""
    import torch
    from bnpp_b200 import model
    from fused_interp import parse_uai
    rec = [r for r in golden_synth["batch"] if r["fixed_ids"] and r["N"] < 100][0]
    text = synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"])
    evs = synth.evidence_batch(rec["N"], rec["nobs"], rec["nsets"], seed=5, fixed_ids=True)
    cards, scopes, tables = parse_uai(text)
    bn1 = model.BN(ctx, cards, list(zip(scopes, tables)))
    bn2 = model.BN(ctx, cards, list(zip(scopes, [tables[0] * 2.0] + tables[1:])))
    observed = sorted(evs[0])
    order, _ = bn1.order([v for v in range(bn1.nvars) if v not in evs[0]], evs[0], "mf")
    p = bn1.plan(observed, order)
    assert p.fused_info(1)[0] > 0
    with torch.cuda.stream(ctx.torch_stream):
        res = torch.zeros(2, dtype=torch.float64, device="cuda")
    for i, ev in enumerate(evs[:8]):
        got = []
        for bn in (bn1, bn2, bn1):
            p.run(bn.table_ptrs, [ev[v] for v in observed], res.data_ptr(), res.data_ptr() + 8)
            ctx.sync()
            got.append(res.cpu().tolist())
        assert math.isclose(got[0][0], rec["pr"][i], rel_tol=REL), i
        assert got[0][0] == got[0][1] and got[2] == got[0]
        assert got[1][0] == 2.0 * got[0][0], i
    bn1.close()
    bn2.close()
