"""K10 (tasks): inside a plan that is not one launch altogether, the small steps are cut into tasks (subtrees of the
bucket tree, one CTA each, intermediates in shared memory) and all tasks of a dependency level run as ONE ve_tasks
launch.  The host half (the task programs, their traffic through the plan's global arena) is pinned on the CPU by
tests/test_fused_program_cpu.py::test_segment_programs_chain_through_the_global_arena; this is the device half:
results bit-identical to the plain launch-per-bucket run, far fewer launches, also when replayed as a CUDA graph."""
import math
import os

import pytest

pytestmark = pytest.mark.gpu

from bnpp_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def _pr(bn, ev, flag, segments, cut=0, fused=False):
    """PR on a FRESH plan (tasks can only be cut before a plan's first run) -> (Z, launches of the query)"""
    bn.drop_plans()
    variables = [v for v in range(bn.nvars) if v not in ev]
    order, _ = bn.order(variables, ev, flag)
    p = bn.plan(sorted(ev), order)
    p.set_fused(fused)
    p.set_segments(segments, cut)
    bn.ctx.sync()
    l0 = bn.ctx.launches
    z, _ = bn.partition(ev, flag)
    return z, bn.ctx.launches - l0


def test_segments_equal_per_bucket(ctx, golden_models):
    from bnpp_b200 import model
    for name in ["alarm", "insurance", "Water", "andes", "hepar2"]:
        m = golden_models[name]
        bn = model.from_uai_text(ctx, m["uai"])[1]
        for case in m["pr"]:
            if not case["flag"]:
                continue
            ev = {int(k): v for k, v in case["evidence"].items()}
            z0, l0 = _pr(bn, ev, case["flag"], False)
            for cut in (0, 1, 5):
                z1, l1 = _pr(bn, ev, case["flag"], True, cut)
                assert z1 == z0, (name, case["flag"], cut, z1, z0)
                if cut == 0:
                    assert l1 < l0, (name, l1, l0)
            assert math.isclose(z0, case["pr"], rel_tol=1e-9)
        bn.close()


def test_segments_on_the_wide_synthetic_network(ctx):
    """width-22 instance of the config-4 generator: small buckets in segments, wide ones as their own launches"""
    from bnpp_b200 import model
    bn = model.from_uai_text(ctx, synth.random_bn_uai(56, 30, 4, 2))[1]
    z0, l0 = _pr(bn, {}, "mf", False)
    z1, l1 = _pr(bn, {}, "mf", True)
    assert z1 == z0 and l1 < l0
    assert math.isclose(z0, 1.0, rel_tol=1e-9)
    bn.close()


def test_tasks_are_the_default_and_replay_as_a_graph(ctx, golden_models):
    """default settings: a plan with many small buckets (andes: 218 of 224) runs as a handful of launches; a plan with a
    few dozen buckets is cut into tasks only on request (Water, insurance).  The graph replays (runs 2..4) give the
    same bits as the first run and as one launch per bucket"""
    from bnpp_b200 import model
    for name, force in [("andes", False), ("Water", True), ("insurance", True)]:
        m = golden_models[name]
        bn = model.from_uai_text(ctx, m["uai"])[1]
        case = [c for c in m["pr"] if c["flag"] == "mf"][-1]
        ev = {int(k): v for k, v in case["evidence"].items()}
        variables = [v for v in range(bn.nvars) if v not in ev]
        order, _ = bn.order(variables, ev, "mf")
        p = bn.plan(sorted(ev), order)
        if force:
            p.set_segments(True, 0)
        zs = []
        for run in range(4):
            ctx.sync()
            l0 = ctx.launches
            z, _ = bn.partition(ev, "mf")
            zs.append(z)
            launches = ctx.launches - l0
        assert launches < p.n_launches / 2, (name, launches, p.n_launches)
        p.set_segments(False, 0)
        z_plain, _ = bn.partition(ev, "mf")
        assert all(z == z_plain for z in zs), (name, zs, z_plain)
        assert math.isclose(z_plain, case["pr"], rel_tol=1e-9)
        bn.close()
