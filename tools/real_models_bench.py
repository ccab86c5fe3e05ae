import gzip, json, sys, time
sys.path.insert(0, '/root/repo')
from bnpp_b200 import capi, model
G = json.load(gzip.open('/root/repo/tests/golden/models.json.gz', 'rt'))
ctx = capi.Context(0)
for name in ['Water', 'andes', 'hepar2', 'win95pts', 'insurance', 'alarm']:
    _, bn = model.from_uai_text(ctx, G[name]['uai'])
    for h in ['mf', 'md', 'wmf']:
        z, ms = bn.partition({}, h)
        t0 = time.perf_counter(); z, _ = bn.partition({}, h); ms2 = (time.perf_counter() - t0) * 1e3
        bn.drop_plans()
        t0 = time.perf_counter(); z, _ = bn.partition({}, h); ms3 = (time.perf_counter() - t0) * 1e3
        p = bn.plan([], bn.order(list(range(bn.nvars)), {}, h)[0])
        print(name, h, 'Z', z, 'first %.2f ms, cached plan %.2f ms, replan %.2f ms' % (ms, ms2, ms3), bn.last_timing, 'launches', p.n_launches, 'entries %.3e' % p.union_entries)
    if name == 'Water':
        p.set_profiling(True)
        import torch
        res = torch.zeros(2, dtype=torch.float64, device='cuda')
        p.run(bn.table_ptrs, [], res.data_ptr(), res.data_ptr() + 8)
        st = p.step_stats()
        for i, s in enumerate(st):
            if s['ms'] > 0.03: print('  ', i, s)
