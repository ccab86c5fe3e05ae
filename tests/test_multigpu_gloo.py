"""N > 1 host logic on CPU: two `gloo` ranks (world_size 2 and 4) partition the path exactly as
bench.py / tools/batch_bench.py do on GPUs -- wide-factor sharding by the last-eliminated variables
of the widest clique with ONE all-reduce of the partition, and evidence batches by contiguous
slices with no collective.  The per-rank arithmetic is done by the CPU oracle here (this is a test);
on the GPU box the same sharding feeds the CUDA plan."""
import math
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as orc
    from bnpp_b200 import model, sharding, synth
    N, W, K, seed = 30, 14, 3, 4
    m = orc.parse_uai(synth.random_bn_uai(N, W, K, seed))
    scopes = [f.scope for f in m.factors]
    cards = [2] * N
    # --- wide-factor sharding -------------------------------------------------------------------
    base_ev = {N - 1: 1, N - 3: 0}
    variables = [v for v in range(N) if v not in base_ev]
    order, width = model.elim_order(cards, scopes, variables, "mf", observed=sorted(base_ev))
    z_full = orc.partition(m, base_ev, order)
    g = world.bit_length() - 1
    shard = sharding.pick_shard_vars([[v for v in sc if v not in base_ev] for sc in scopes], order, g)
    assert len(shard) == g and sharding.shard_count(shard) == world
    ev = dict(base_ev)
    ev.update(sharding.shard_evidence(shard, rank))
    vars_r = [v for v in range(N) if v not in ev]
    order_r, width_r = model.elim_order(cards, scopes, vars_r, "mf", observed=sorted(ev))
    z = torch.tensor([orc.partition(m, ev, order_r)], dtype=torch.float64)
    dist.all_reduce(z)                      # the cross-shard sum-out of the shard variables
    ok_wide = math.isclose(z.item(), z_full, rel_tol=1e-12) and width_r <= width
    # --- evidence batch: contiguous slices, no collective on the path -----------------------------
    evs = synth.evidence_batch(N, 4, 10, seed=5, fixed_ids=True)
    lo, hi = sharding.batch_slice(rank, world, len(evs))
    observed = sorted(evs[0])
    o_b, _ = model.elim_order(cards, scopes, [v for v in range(N) if v not in observed], "mf", observed=observed)
    mine = [orc.partition(m, e, o_b) for e in evs[lo:hi]]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))      # host gather of the result vectors only
    if rank == 0:
        flat = [z for part in sorted(gathered) for z in part[2]]
        want = [orc.partition(m, e, o_b) for e in evs]
        cover = sorted((p[0], p[1]) for p in gathered)
        ok_batch = flat == want and cover[0][0] == 0 and cover[-1][1] == len(evs) and \
            all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
        q.put((ok_wide, ok_batch, z.item(), z_full))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharding_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    ok_wide, ok_batch, z, z_full = q.get(timeout=10)
    assert ok_wide, (z, z_full)
    assert ok_batch
