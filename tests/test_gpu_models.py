"""End-to-end parity through the C ABI: BN::partition / BN::marginals by device-resident
variable elimination (code/model.cpp:250-446) against the unmodified reference's results on
the same UAI models, evidence and ordering flags.  Tolerance 1e-9 relative (north_star);
elimination orders are asserted bit-exact on the way."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as orc  # noqa: E402
from bnpp_b200 import synth  # noqa: E402

REL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def load(ctx, text):
    from bnpp_b200 import model
    return model.from_uai_text(ctx, text)[1]


def test_partition_all_models(ctx, golden_models):
    """configs 1-2 and every small shipped network: PR under no heuristic / -mf / -wmf / -md"""
    n = 0
    for name, m in golden_models.items():
        bn = load(ctx, m["uai"])
        for case in m["pr"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            variables = [v for v in range(bn.nvars) if v not in ev]
            if case["flag"]:
                order, width = bn.order(variables, ev, case["flag"])
                assert order == case["order"] and width == case["width"], (name, case["flag"])
            z, _ = bn.partition(ev, case["flag"] or None)
            assert math.isclose(z, case["pr"], rel_tol=REL), (name, case["flag"], z, case["pr"])
            n += 1
        bn.close()
    assert n > 60


def test_marginals(ctx, golden_models):
    for name in ["asia", "asia_positive", "cancer", "earthquake", "child", "alarm", "grid3x3", "network"]:
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        for case in m["mar"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            got = bn.marginals(ev, case["flag"] or None)
            for v, (g, want) in enumerate(zip(got, case["mar"])):
                assert np.allclose(g, want, rtol=REL, atol=0.0), (name, case["flag"], v)
        bn.close()


def test_marginals_heuristic_with_evidence(ctx, golden_models):
    """`-mar -mf` with evidence crashes in the reference (SURVEY A.2 i); the result must equal the
    no-heuristic marginals, which the reference does compute"""
    m = golden_models["alarm"]
    bn = load(ctx, m["uai"])
    case = [c for c in m["mar"] if c["evidence"] and not c["flag"]][0]
    ev = {int(k): v for k, v in case["evidence"].items()}
    got = bn.marginals(ev, "mf")
    for g, want in zip(got, case["mar"]):
        assert np.allclose(g, want, rtol=REL, atol=0.0)
    bn.close()


def test_shipped_goldens(ctx, golden_models):
    """grid3x3.uai.PR / network.uai.PR as shipped by the reference (log10 Z, 6 digits)"""
    for name in ("grid3x3", "network"):
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        ev = orc.parse_evidence(m["shipped"]["PR_evid"]) if name == "grid3x3" else {}
        z, _ = bn.partition(ev, "mf")
        assert abs(math.log10(z) - float(m["shipped"]["PR"].split()[-1])) < 5e-4
        bn.close()


def test_synthetic_bn(ctx, golden_synth):
    """config 4 at reduced width: PR vs the reference, and Z = 1 without evidence"""
    for rec in golden_synth["bn"]:
        bn = load(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        ev = {int(k): v for k, v in rec["evidence"].items()}
        for case in rec["cases"]:
            z, _ = bn.partition(ev, case["flag"])
            if "pr" in case:
                assert math.isclose(z, case["pr"], rel_tol=REL), (rec["N"], case["flag"])
            if not ev:
                assert math.isclose(z, 1.0, rel_tol=REL)
        bn.close()


def test_evidence_batch_sample(ctx, golden_synth):
    """config 5: evidence sets over one plan (fixed observed ids) and over varying ids"""
    for rec in golden_synth["batch"]:
        bn = load(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        evs = synth.evidence_batch(rec["N"], rec["nobs"], rec["nsets"], seed=5, fixed_ids=rec["fixed_ids"])
        for i, ev in enumerate(evs):
            z, _ = bn.partition(ev, "mf")
            assert math.isclose(z, rec["pr"][i], rel_tol=REL), (rec["N"], i)
        if rec["fixed_ids"]:
            assert len(bn._plans) == 1          # one ordering, one plan, evidence values only move base pointers
        bn.close()


def test_wide_bn_normalised(ctx):
    """a BN without evidence sums to 1: width-22 instance of the config-4 generator (2^23-entry tables)"""
    from bnpp_b200 import model
    text = synth.random_bn_uai(56, 30, 4, 2)
    bn = load(ctx, text)
    order, width = bn.order(list(range(bn.nvars)), {}, "mf")
    assert 18 <= width <= 26
    z, _ = bn.partition({}, "mf")
    assert math.isclose(z, 1.0, rel_tol=REL)
    # one observed leaf: P(e) from the CPTs' own chain rule is not available in closed form, but
    # P(x=0) + P(x=1) must be 1
    leaf = bn.nvars - 1
    z0, _ = bn.partition({leaf: 0}, "mf")
    z1, _ = bn.partition({leaf: 1}, "mf")
    assert math.isclose(z0 + z1, 1.0, rel_tol=REL)
    bn.close()


def test_evidence_batch_one_launch_per_bucket(ctx, golden_synth):
    """K8: the whole batch through ONE plan run (batch = fastest axis of every intermediate);
    equals the reference's per-set PR and the per-query device path bit for bit"""
    import torch
    for rec in golden_synth["batch"]:
        if not rec["fixed_ids"]:
            continue
        bn = load(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))
        evs = synth.evidence_batch(rec["N"], rec["nobs"], rec["nsets"], seed=5, fixed_ids=True)
        observed = sorted(evs[0])
        vals = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        plan = bn.plan(observed, bn.order([v for v in range(bn.nvars) if v not in observed], observed, "mf")[0])
        plan.set_fused(False)               # this test is about the launch-per-bucket batch path (K8); K9: test_gpu_fused.py
        launches0 = ctx.launches
        z = bn.partition_batch(observed, vals, "mf")
        ctx.sync()
        per_batch = ctx.launches - launches0
        z = z.cpu().numpy()
        for i, ev in enumerate(evs):
            assert math.isclose(z[i], rec["pr"][i], rel_tol=REL), (rec["N"], i)
            zi, _ = bn.partition(ev, "mf")
            assert zi == z[i]                  # same arithmetic, same order: bit-identical to the per-query path
        # one launch per bucket for the whole batch (+1: the evidence matrix is transposed once)
        assert per_batch == bn.plan(observed, bn.order([v for v in range(bn.nvars) if v not in observed], observed, "mf")[0]).n_launches + 1
        bn.close()


def test_evidence_batch_large(ctx):
    """4096 evidence sets on the 500-variable network: batch result == per-query result on a sample,
    and sum over both values of one observed variable reproduces the marginal likelihood of the rest"""
    import random
    import torch
    N, nobs, nsets = 500, 20, 4096
    bn = load(ctx, synth.random_bn_uai(N, 6, 3, 11))
    evs = synth.evidence_batch(N, nobs, nsets, seed=5, fixed_ids=True)
    observed = sorted(evs[0])
    vals = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    z = bn.partition_batch(observed, vals, "mf")
    ctx.sync()
    z = z.cpu().numpy()
    rng = random.Random(1)
    for i in rng.sample(range(nsets), 12):
        zi, _ = bn.partition(evs[i], "mf")
        assert zi == z[i]
    assert np.all(z > 0) and np.all(z < 1)
    bn.close()


def test_marginals_bucket_tree(ctx, golden_models, golden_synth):
    """all marginals from one two-pass bucket-tree plan == the reference's N VE passes (1e-9),
    with and without evidence, several orderings, binary and multi-valued networks"""
    n = 0
    for name in ["asia", "asia_positive", "cancer", "earthquake", "child", "alarm", "grid3x3", "network"]:
        m = golden_models[name]
        bn = load(ctx, m["uai"])
        for case in m["mar"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            for h in ("mf", "md", "wmf", None):
                got = bn.marginals_fast(ev, h)
                for v, (g, want) in enumerate(zip(got, case["mar"])):
                    assert len(g) == len(want), (name, v)
                    assert np.allclose(g, want, rtol=REL, atol=1e-300), (name, h, v, g, want)
                n += 1
        bn.close()
    assert n >= 40
    # exact marginals of the 4x4 Ising grid (loopy graph, treewidth 4)
    rec = [r for r in golden_synth["ising"] if r["n"] == 4][0]
    bn = load(ctx, synth.ising_uai(4, rec["h"], rec["J"], rec["seed"]))
    got = bn.marginals_fast({}, "mf")
    assert np.allclose([g[0] for g in got], rec["exact_p0"], rtol=REL, atol=0.0)
    bn.close()


def test_marginals_bucket_tree_vs_passes_wide(ctx):
    """a width-15 network: the two-pass plan against one VE pass per variable (both on the device)"""
    bn = load(ctx, synth.random_bn_uai(40, 24, 3, 3))
    ev = {5: 1, 17: 0, 33: 1}
    fast = bn.marginals_fast(ev, "mf")
    slow = bn.marginals(ev, "mf")
    for v, (a, b) in enumerate(zip(fast, slow)):
        assert np.allclose(a, b, rtol=REL, atol=1e-300), v
    bn.close()


def test_evidence_batch_sliced(ctx):
    """a batch whose widest intermediate would exceed the slice budget is processed in slices of the batch"""
    import random
    import torch
    N, nobs, nsets = 48, 4, 16384
    bn = load(ctx, synth.random_bn_uai(N, 30, 3, 5))
    evs = synth.evidence_batch(N, nobs, nsets, seed=9, fixed_ids=True)
    observed = sorted(evs[0])
    vals = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for rep in range(3):                 # plain run, captured run, replayed run
        z = bn.partition_batch(observed, vals, "mf")
        ctx.sync()
        zh = z.cpu().numpy()
        rng = random.Random(rep)
        for i in rng.sample(range(nsets), 6) + [0, nsets - 1]:
            zi, _ = bn.partition(evs[i], "mf")
            assert zi == zh[i], (rep, i)
    p = [pl for pl in bn._plans.values()][0]
    assert p.max_step_entries // 2 * nsets > (1 << 27)      # widest intermediate x batch over the slice budget: >= 2 slices
    bn.close()


def test_shard_allreduce_single_rank(ctx):
    """the C-ABI collective (libbnpp_b200_nccl.so) on a one-rank communicator: identity, in place,
    on the context's stream; multi-rank sums are exercised by `bench.py --gpus N`"""
    import torch
    from bnpp_b200.nccl import ShardComm
    comm = ShardComm(ctx, 0, 1)
    with torch.cuda.stream(ctx.torch_stream):
        buf = torch.arange(5, dtype=torch.float64, device="cuda") + 0.25
    want = buf.clone()
    comm.allreduce_sum(buf.data_ptr(), 5)
    ctx.sync()
    assert torch.equal(buf, want)
    comm.close()
