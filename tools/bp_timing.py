import sys, time
sys.path.insert(0, '/root/repo')
from bnpp_b200 import capi, model, synth
from bnpp_b200.sumproduct import FactorGraph
import numpy as np
ctx = capi.Context(0)
for J in (0.3, 1.0, 1.0, 0.3):
    _, ising = model.from_uai_text(ctx, synth.ising_uai(40, 0.5, J, 7))
    hv = ising._host.numpy()
    t0 = time.perf_counter()
    facs = []
    for sc, p in zip(ising.scopes, ising.table_ptrs):
        o = (p - ising._dev.data_ptr()) // 8
        facs.append((sc, hv[o:o + 2 ** len(sc)]))
    t1 = time.perf_counter()
    fg = FactorGraph(ctx, ising.cards, facs)
    t2 = time.perf_counter()
    sweeps = fg.update(10000, 0.001)
    t3 = time.perf_counter()
    mar = fg.marginals()
    t4 = time.perf_counter()
    print(J, 'facs %.2f create %.2f update %.2f (%d sweeps) marg %.2f ms' % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, sweeps, (t4-t3)*1e3))
    fg.close()
