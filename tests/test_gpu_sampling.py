"""BN::logical_sampling / BN::likelihood_weighting (code/model.cpp:540-690) with the samples drawn on
the GPU, one thread per sample.  The reference's draws are irreproducible (std::random_device), so parity is statistical:
the estimates must fall within a few standard errors of the EXACT P(evidence) the reference's VE computes (golden), a
run must repeat bit for bit under the same seed, and the bounded-variance rule must not depend on the batch size."""
import math

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", ["asia", "child", "alarm", "insurance"])
def test_sampling_estimates_the_exact_partition(ctx, golden_models, name):
    from bnpp_b200 import model
    m = golden_models[name]
    bn = model.from_uai_text(ctx, m["uai"])[1]
    case = [c for c in m["pr"] if c["evidence"]][0]
    ev = {int(k): v for k, v in case["evidence"].items()}
    exact = case["pr"]
    sp = model.Sampler(bn)
    M = 400_000
    p1 = sp.logical(ev, M, seed=7)
    assert sp.logical(ev, M, seed=7) == p1, "same seed, same samples"
    se = math.sqrt(exact * (1 - exact) / M)
    assert abs(p1 - exact) <= 5 * se + 1e-12, (name, p1, exact, se)
    assert sp.logical(ev, M, seed=8) != p1 or exact in (0.0, 1.0)
    # likelihood weighting: U = product of the CPT maxima (code/model.cpp:627-630), N* from delta = epsilon = 0.05
    import numpy as np
    hv = bn._host.numpy()
    U = float(np.prod([hv[o:o + n].max() for o, n in zip(bn._offs, bn._sizes)]))
    n_star = 4 * math.log(2 / 0.05) * 1.05 / 0.05 ** 2
    if exact / U > 1e-4:            # else the stopping rule needs more than ~6e7 samples
        e1, m1 = sp.likelihood(ev, U, n_star, seed=3, batch=1 << 14)
        e2, m2 = sp.likelihood(ev, U, n_star, seed=3, batch=777)
        assert (e1, m1) == (e2, m2), "the rule consumes samples in order: no dependence on the batch size"
        assert abs(e1 - exact) <= 0.05 * exact + 1e-12, (name, e1, exact, m1)      # the (epsilon, delta) guarantee, epsilon = 0.05
    sp.close()
    bn.close()
