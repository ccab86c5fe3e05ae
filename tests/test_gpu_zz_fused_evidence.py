"""K9 with heavy evidence: most variables of the 500-variable network observed, so resident CPTs are read
through base offsets built from up to four observed axes each (the odd and even lengths of the program's
observed-axis lists), and many buckets degenerate to scalars.  The batch in one launch against the CPU oracle (1e-9) and against single
queries through the same plan (bit for bit)."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from bnpp_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from bnpp_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("nobs", [60, 150, 330])
def test_heavy_evidence_one_launch_equals_per_bucket(ctx, nobs):
    import torch
    from bnpp_b200 import model
    N, nsets = 500, 301
    text = synth.random_bn_uai(N, 6, 3, 11)
    bn = model.from_uai_text(ctx, text)[1]
    evs = synth.evidence_batch(N, nobs, nsets, seed=17, fixed_ids=True)
    observed = sorted(evs[0])
    vals = torch.tensor([[ev[v] for v in observed] for ev in evs], dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    order, _ = bn.order([v for v in range(N) if v not in evs[0]], evs[0], "mf")
    p = bn.plan(observed, order)
    lanes = p.fused_info(nsets)[0]
    assert lanes > 0
    z = bn.partition_batch(observed, vals, "mf")
    ctx.sync()
    zf = z.cpu().numpy().copy()
    assert np.all(zf > 0)
    # against the CPU oracle (plain-C restatement of the reference's bucket elimination) on a few sets
    import oracle as orc
    m = orc.parse_uai(text)
    rng = random.Random(nobs)
    for i in rng.sample(range(nsets), 3):
        want = orc.partition(m, evs[i], order)
        assert abs(zf[i] - want) <= 1e-9 * abs(want), (i, zf[i], want)
    # single queries through the same plan (one launch with 32 or 128 lanes; one launch per bucket beyond 256 observed)
    for i in rng.sample(range(nsets), 4):
        zi, _ = bn.partition(evs[i], "mf")
        assert zi == zf[i], i
    if nobs == 60:
        p.set_fused(False)
        zb = bn.partition_batch(observed, vals, "mf")
        ctx.sync()
        assert np.array_equal(zb.cpu().numpy(), zf)
        p.set_fused(True)
    bn.close()
