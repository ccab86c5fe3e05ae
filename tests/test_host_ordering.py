"""Host logic, CPU only: elimination orders through the C ABI (bnpp_elim_order /
bnpp_order_width) must be BIT-EXACT with the unmodified reference's Graph::ordering
(code/graph.cpp:41-237) on all shipped Bayesian networks x 3 heuristics, with and
without observed variables; and the library must export every symbol the header declares."""
import ctypes
import os
import re

from bnpp_b200 import capi, model, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "bnpp_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bnpp_[a-z0-9_]+)\s*\(", hdr)))
    L = capi.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(capi.EXPORTS) <= set(declared)
    assert L.bnpp_version() >= 100


def test_no_cpu_fallback():
    """without a device the product fails loudly instead of computing on the CPU"""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(capi.BnppError) as e:
        capi.Context(0)
    assert e.value.code == -2


def test_union_scope():
    assert capi.union_scope([3, 1], [2, 3], [1, 5, 0], [3, 4, 2]) == ([3, 1, 5, 0], [2, 3, 4, 2])
    assert capi.union_scope([], [], [2], [7]) == ([2], [7])


def test_orders_bit_exact(golden_orders):
    n = 0
    for name, rec in golden_orders.items():
        cards, scopes = rec["card"], rec["scopes"]
        for case in rec["cases"]:
            obs = set(case["observed"])
            variables = [v for v in range(len(cards)) if v not in obs]
            cond = [[v for v in sc if v not in obs] for sc in scopes]
            order, width = model.elim_order(cards, cond, variables, case["flag"])
            assert order == case["order"], (name, case["flag"], len(obs))
            assert width == case["width"]
            n += 1
        # the `width` prompt command: original order, then md / mf / wmf (code/bn.cpp:418-481)
        allv = list(range(len(cards)))
        assert model.order_width(cards, scopes, allv) == rec["widths"][0], name
    assert n >= 120


def test_orders_models_and_synthetic(golden_models, golden_synth):
    import oracle as orc
    for name, m in golden_models.items():
        mod = orc.parse_uai(m["uai"])
        cards = [int(c) for c in mod.card]
        scopes = [f.scope for f in mod.factors]
        for case in m["pr"]:
            if not case["flag"]:
                continue
            obs = {int(k) for k in case["evidence"]}
            variables = [v for v in range(len(cards)) if v not in obs]
            cond = [[v for v in sc if v not in obs] for sc in scopes]
            order, width = model.elim_order(cards, cond, variables, case["flag"])
            assert order == case["order"] and width == case["width"], (name, case["flag"])
    for rec in golden_synth["bn"]:
        scopes, _ = synth.random_bn_scopes(rec["N"], rec["W"], rec["K"], rec["seed"])
        obs = {int(k) for k in rec["evidence"]}
        cards = [2] * rec["N"]
        variables = [v for v in range(rec["N"]) if v not in obs]
        cond = [[v for v in sc if v not in obs] for sc in scopes]
        for case in rec["cases"]:
            order, width = model.elim_order(cards, cond, variables, case["flag"])
            assert order == case["order"] and width == case["width"]


def test_fast_orderer_equals_reference_containers(golden_orders):
    """the bit-matrix orderer and the std::unordered_set one (the reference's containers) agree on
    every shipped network, every heuristic, with and without observed variables handled in-library"""
    for name, rec in golden_orders.items():
        if len(rec["card"]) > 450:
            continue        # the container path is quadratic; the big nets are pinned by test_orders_bit_exact
        cards, scopes = rec["card"], rec["scopes"]
        for case in rec["cases"]:
            obs = case["observed"]
            allv = list(range(len(cards)))
            fast = model.elim_order(cards, scopes, allv, case["flag"], observed=obs)
            slow = model.elim_order(cards, scopes, allv, case["flag"], reference_containers=True, observed=obs)
            assert fast == slow == (case["order"], case["width"]), (name, case["flag"])


def test_header_is_plain_c(tmp_path):
    """the drop-in boundary is a C ABI: the header must compile as C99 on its own"""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "bnpp_b200.h"\nint main(void) { bnpp_scope s = {0, 0, 0}; return (int)bnpp_scope_size(&s) * 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])


def test_shard_vars_c_equals_python():
    """wide-factor shard variables: the C ABI (bnpp_pick_shard_vars) and the Python helper agree"""
    from bnpp_b200 import sharding
    for (N, W, K, seed) in [(64, 40, 4, 5), (68, 40, 4, 27), (72, 40, 4, 23), (76, 44, 4, 3), (30, 14, 3, 4)]:
        scopes, _ = synth.random_bn_scopes(N, W, K, seed)
        order, _w = model.elim_order([2] * N, scopes, list(range(N)), "mf")
        for g in (0, 1, 2, 3):
            assert sharding.pick_shard_vars(scopes, order, g) == sharding.pick_shard_vars_c(scopes, [2] * N, order, g)


def test_nccl_library_exports():
    so = os.path.join(ROOT, "bnpp_b200", "libbnpp_b200_nccl.so")
    assert os.path.exists(so)
    out = __import__("subprocess").run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    assert " T bnpp_shard_allreduce_sum" in out
