# small end-to-end exercise for compute-sanitizer: every kernel family once, small sizes
import sys, gzip, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from bnpp_b200 import capi, model, synth
from bnpp_b200.factor import DeviceFactor, fused_product_sum_out
ctx = capi.Context(0)
rng = np.random.default_rng(0)
def rf(scope, cards):
    n = int(np.prod([cards[v] for v in scope])) if scope else 1
    return DeviceFactor.from_host(ctx, scope, [cards[v] for v in scope], rng.uniform(0.1, 1, n))
cards = [2]*14
a, b, c = rf(list(range(14)), cards), rf([13, 5, 2, 0], cards), rf([1, 3], cards)
for elim in (0, 7, 13, None):
    out_scope = [v for v in range(14) if v != elim]
    fused_product_sum_out(ctx, [a, b, c], out_scope, elim).partition
cards2 = [3, 4, 5, 2, 3, 4]
d, e = rf([0, 1, 2, 3], cards2), rf([5, 4, 3, 1], cards2)
p = d.product(e); p.sum_out(1).partition; p.sum_out(3).partition; d.divide(e).partition; p.condition({0: 2, 5: 1}).partition
p.normalize().values(); p.max(); p.min()
G = json.load(gzip.open('/root/repo/tests/golden/models.json.gz', 'rt'))
_, bn = model.from_uai_text(ctx, G['alarm']['uai'])
print(bn.partition({3: 0, 17: 1, 30: 0}, 'mf')[0])
print([float(m[0]) for m in bn.marginals({}, 'mf')[:3]])
fg, sweeps = bn.sum_product(); print(sweeps, fg.marginals()[0]); fg.close()
_, wb = model.from_uai_text(ctx, synth.random_bn_uai(44, 26, 4, 2))
print(wb.partition({}, 'mf')[0])
evs = synth.evidence_batch(60, 6, 64, seed=5)
_, bb = model.from_uai_text(ctx, synth.random_bn_uai(60, 6, 3, 4))
obs = sorted(evs[0])
vals = torch.tensor([[ev[v] for v in obs] for ev in evs], dtype=torch.uint8, device='cuda')
torch.cuda.synchronize()
print(bb.partition_batch(obs, vals, 'mf')[:3].tolist())
ctx.sync()
print("done")
