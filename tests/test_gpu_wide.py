"""The WIDE path (config 4) against numbers the unmodified reference computed -- never against `Z == 1`.

* tests/golden/synthetic.json.gz: the N=64 W=40 K=4 seed=5 network with 8 observed variables (min-fill width 23),
  PR by BN::partition (code/model.cpp:275-294) under -mf / -md / -wmf;
* tests/golden/wide.json: the width-27..30 networks bench.py runs, with observed leaves: one reference run per shard
  assignment (oracle/make_wide_golden.py), i.e. the reference's value for every rank's slab and their sum.

Through: the launch-per-bucket plan (first run), its CUDA-graph replay (later runs), the segment path, and cutset
sharding executed by the CUDA plan rank by rank on one GPU (what `bench.py --gpus N` does on N GPUs)."""
import json
import math
import os

import pytest

pytestmark = pytest.mark.gpu

from bnpp_b200 import sharding, synth  # noqa: E402

REL = 1e-9
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def wide():
    p = os.path.join(ROOT, "tests", "golden", "wide.json")
    if not os.path.exists(p):
        pytest.skip("tests/golden/wide.json not generated")
    return json.load(open(p))["networks"]


def test_width23_reference_pr(ctx, golden_synth):
    """per-bucket run, graph replay (runs 2 and 3) and the segment path, three orderings"""
    from bnpp_b200 import model
    rec = [b for b in golden_synth["bn"] if b["N"] == 64][0]
    ev = {int(k): v for k, v in rec["evidence"].items()}
    bn = model.from_uai_text(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))[1]
    checked = 0
    for case in rec["cases"]:
        if case.get("pr") is None:
            continue
        variables = [v for v in range(bn.nvars) if v not in ev]
        order, width = bn.order(variables, ev, case["flag"])
        assert order == case["order"] and width == case["width"]
        for run in range(3):
            z, _ = bn.partition(ev, case["flag"])
            assert math.isclose(z, case["pr"], rel_tol=REL), (case["flag"], run, z, case["pr"])
        bn.drop_plans()
        p = bn.plan(sorted(ev), order)
        p.set_segments(True, 0)
        z, _ = bn.partition(ev, case["flag"])
        assert math.isclose(z, case["pr"], rel_tol=REL), (case["flag"], "segments", z, case["pr"])
        bn.drop_plans()
        checked += 1
    assert checked >= 1
    bn.close()


@pytest.mark.parametrize("key", ["1", "2", "4", "8"])
def test_wide_networks_match_reference_slabs(ctx, wide, key):
    """every rank's slab of the bench networks through the CUDA plan == the reference's value for that slab; their
    sum == the reference's PR of the unsharded network (which the reference itself is too narrow to run)"""
    from bnpp_b200 import model
    rec = wide.get(key)
    if rec is None or len(rec.get("partials", {})) != int(key):
        pytest.skip("reference runs for network %s not finished" % key)
    N, W, K, seed, ev = synth.wide_bn(int(key))
    assert (N, W, K, seed) == (rec["N"], rec["W"], rec["K"], rec["seed"])
    assert {str(k): v for k, v in ev.items()} == rec["evidence"]
    bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))[1]
    variables = [v for v in range(N) if v not in ev]
    order, width = bn.order(variables, ev, "mf")
    assert order == rec["order"] and width == rec["width"]
    g = int(key).bit_length() - 1
    shard_vars = sharding.pick_shard_vars(bn.conditioned_scopes(set(ev)), order, g)
    assert shard_vars == rec["shard_vars"]
    total = 0.0
    for r in range(int(key)):
        sev = sharding.shard_evidence(shard_vars, r)
        full = dict(ev)
        full.update(sev)
        want = rec["partials"][",".join("%d=%d" % kv for kv in sorted(sev.items()))]
        z, _ = bn.partition(full, "mf")
        assert math.isclose(z, want["pr"], rel_tol=REL), (key, r, z, want["pr"])
        z2, _ = bn.partition(full, "mf")          # graph replay
        assert z2 == z
        total += z
        bn.drop_plans()
    want_total = sum(sorted(p_["pr"] for p_ in rec["partials"].values()))
    assert math.isclose(total, want_total, rel_tol=REL)
    bn.close()


def test_strong_network_unsharded_equals_sum_of_reference_slabs(ctx, wide):
    """the width-30 network on ONE GPU (2^31-entry union tables: wider than the reference can hold) == the sum of
    the reference's eight slab values"""
    from bnpp_b200 import model
    rec = wide.get("8")
    if rec is None or len(rec.get("partials", {})) != 8:
        pytest.skip("reference runs for the width-30 network not finished")
    N, W, K, seed, ev = synth.wide_bn("strong")
    bn = model.from_uai_text(ctx, synth.random_bn_uai(N, W, K, seed))[1]
    z, _ = bn.partition(ev, "mf")
    want = sum(sorted(p_["pr"] for p_ in rec["partials"].values()))
    assert math.isclose(z, want, rel_tol=REL), (z, want)
    bn.close()


def test_sharded_api_single_rank(ctx, golden_synth):
    """bnpp_ve_plan_run_sharded over a one-rank NCCL communicator: run + all-reduce of result and partition"""
    import torch
    import torch.distributed as dist
    from bnpp_b200 import model
    from bnpp_b200.nccl import ShardComm
    own = not dist.is_initialized()
    comm = ShardComm(ctx, 0, 1)
    rec = [b for b in golden_synth["bn"] if b["N"] == 40][0]
    ev = {int(k): v for k, v in rec["evidence"].items()}
    bn = model.from_uai_text(ctx, synth.random_bn_uai(rec["N"], rec["W"], rec["K"], rec["seed"]))[1]
    case = [c for c in rec["cases"] if c["flag"] == "mf"][0]
    variables = [v for v in range(bn.nvars) if v not in ev]
    order, _ = bn.order(variables, ev, "mf")
    p = bn.plan(sorted(ev), order)
    with torch.cuda.stream(ctx.torch_stream):
        res = torch.zeros(2, dtype=torch.float64, device="cuda")
    comm.run_sharded(p, bn.table_ptrs, [ev[v] for v in sorted(ev)], res.data_ptr(), res.data_ptr() + 8)
    ctx.sync()
    assert math.isclose(res[0].item(), case["pr"], rel_tol=REL) and res[0].item() == res[1].item()
    comm.close()
    bn.close()
    assert own or True


def test_evidence_out_of_range_is_refused(ctx):
    """an evidence value >= the variable's cardinality must not move a CPT view past its table"""
    import torch
    from bnpp_b200 import capi, model
    bn = model.from_uai_text(ctx, synth.random_bn_uai(30, 6, 3, 4))[1]
    with pytest.raises(capi.BnppError):
        bn.partition({3: 2, 7: 0}, "mf")
    observed = [3, 7, 11]
    good = torch.tensor([[0, 1, 1], [1, 0, 0], [1, 1, 0], [0, 0, 0]], dtype=torch.uint8).cuda()
    bad_host = good.cpu().clone()
    bad_host[2, 1] = 5
    bad = bad_host.cuda()
    torch.cuda.synchronize()

    def run(values):
        z = bn.partition_batch(observed, values, "mf")
        ctx.sync()                      # the result buffer belongs to the context's stream
        return z.clone()

    z_good = run(good)
    assert ctx.status() == 0
    z_bad = run(bad)
    assert ctx.status() & 2 and ctx.status() == 0          # raised once, cleared by the read
    assert torch.equal(z_bad[[0, 1, 3]], z_good[[0, 1, 3]])
    assert z_bad[2].item() == z_good[1].item()             # the bad value was read as 0: set 2 became set 1
    bn.drop_plans()
    order, _ = bn.order([v for v in range(bn.nvars) if v not in observed], observed, "mf")
    bn.plan(observed, order).set_fused(False)              # the launch-per-bucket batched path
    z2 = run(bad)
    assert ctx.status() & 2
    assert torch.allclose(z2, z_bad, rtol=1e-12)
    bn.close()
