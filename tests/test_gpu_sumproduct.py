"""K7 parity: batched loopy sum-product (code/graph.cpp:256-403) through the C ABI against
the unmodified reference's sweep counts and marginals (tests/golden)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle as orc  # noqa: E402
from bnpp_b200 import synth  # noqa: E402

REL = 1e-9   # north_star tolerance for marginals


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def run_bp(ctx, model, factors):
    from bnpp_b200.sumproduct import FactorGraph
    fg = FactorGraph(ctx, model.card, [(f.scope, f.values) for f in factors])
    sweeps = fg.update(10000, 0.001)
    mar = fg.marginals()
    fg.close()
    return sweeps, mar


def test_shipped_models(ctx, golden_models):
    for name in ["asia", "asia_positive", "cancer", "earthquake", "child", "alarm", "insurance", "grid3x3"]:
        m = golden_models[name]
        model = orc.parse_uai(m["uai"])
        for case in m["bp"]:
            ev = {int(k): v for k, v in case["evidence"].items()}
            factors = [orc.condition(f, ev, model.card) for f in model.factors] if case["cond"] else model.factors
            sweeps, mar = run_bp(ctx, model, factors)
            assert sweeps == case["sweeps"], (name, sweeps, case["sweeps"])
            for v in range(model.nvars):
                if case["cond"] and v in ev:
                    assert mar[v][0] == 1.0
                    continue
                assert np.allclose(mar[v], case["mar"][v], rtol=REL, atol=0.0), (name, v)


def test_ising(ctx, golden_synth):
    """config 3: 40x40 binary Ising grid, same sweep count and marginals as `bn -mar -sp`"""
    for rec in golden_synth["ising"]:
        model = orc.parse_uai(synth.ising_uai(rec["n"], rec["h"], rec["J"], rec["seed"]))
        sweeps, mar = run_bp(ctx, model, model.factors)
        assert sweeps == rec["sweeps"], (rec["n"], rec["J"], sweeps)
        p0 = np.array([m[0] for m in mar])
        assert np.allclose(p0, rec["p0"], rtol=REL, atol=0.0)


def test_sweep_error_semantics(ctx):
    """0/0 = NaN never raises maxerror, x/0 = inf forces another sweep (code/graph.cpp:349-356)"""
    from bnpp_b200.sumproduct import FactorGraph
    # a deterministic unary factor [1, 0]: its message has a zero entry, later sweeps compare 0 with 0
    fg = FactorGraph(ctx, [2, 2], [([0], [1.0, 0.0]), ([0, 1], [0.5, 0.5, 0.2, 0.8]), ([1], [0.3, 0.7])])
    o = orc.OFactorGraph([2, 2], [orc.OFactor([0], [1.0, 0.0]), orc.OFactor([0, 1], [0.5, 0.5, 0.2, 0.8]),
                                  orc.OFactor([1], [0.3, 0.7])])
    e1 = fg.sweep()
    assert e1 >= 1.0 - 1e-12          # uniform 0.5 -> 0 is a relative change of 1
    assert e1 == o.sweep()
    for _ in range(4):                # later sweeps compare the zero entry with itself: 0/0 = NaN is skipped
        e = fg.sweep()
        assert not np.isnan(e) and np.isclose(e, o.sweep(), rtol=1e-12, atol=1e-300)
    assert e == 0.0                   # converged, and the NaN comparisons never raised it
    fg.marginals()
    fg.close()


def test_one_launch_update_equals_sweep_by_sweep(ctx, golden_synth):
    """FactorGraph::update (code/graph.cpp:298-332) as ONE cooperative launch (grid barriers between the phases,
    convergence test on the device) == the host loop over bnpp_fg_sweep: same converging sweep, same messages"""
    from bnpp_b200.sumproduct import FactorGraph
    for rec in golden_synth["ising"]:
        model = orc.parse_uai(synth.ising_uai(rec["n"], rec["h"], rec["J"], rec["seed"]))
        facs = [(f.scope, f.values) for f in model.factors]
        a = FactorGraph(ctx, model.card, facs)
        l0 = ctx.launches
        sweeps = a.update(10000, 0.001)
        assert ctx.launches - l0 == 1, "update() must be one launch"
        b = FactorGraph(ctx, model.card, facs)
        it = 0
        while it < 10000 and not b.sweep() < 0.001:
            it += 1
        assert sweeps == it == rec["sweeps"]
        for ma, mb in zip(a.marginals(), b.marginals()):
            assert np.array_equal(ma, mb)
        # a capped run stops at max_sweeps, and reset() + update() repeats the first run exactly
        first = [m.copy() for m in a.marginals()]
        a.reset()
        assert a.update(3, 0.0) == 3
        a.reset()
        assert a.update(10000, 0.001) == rec["sweeps"]
        for m0, m1 in zip(first, a.marginals()):
            assert np.array_equal(m0, m1)
        a.close()
        b.close()
