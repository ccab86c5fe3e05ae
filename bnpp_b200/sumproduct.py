"""Python mirror of `bn::FactorGraph` (reference code/graph.hh:39-55) over the C ABI."""
import ctypes

import numpy as np

from . import capi


class FactorGraph:
    """FactorGraph(variables, factors), code/graph.cpp:256-275.

    `cards[v]` is the cardinality of variable v; `factors` is a list of
    (scope ids, table values) in the reference's row-major, last-fastest layout."""

    def __init__(self, ctx, cards, factors):
        self.ctx = ctx
        self.cards = np.ascontiguousarray(cards, dtype=np.uint32)
        foff = [0]
        fscope = []
        toff = []
        tabs = []
        t = 0
        for scope, values in factors:
            values = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
            fscope += [int(v) for v in scope]
            foff.append(len(fscope))
            toff.append(t)
            tabs.append(values)
            t += values.size
        self._foff = np.array(foff, dtype=np.int32)
        self._fscope = np.array(fscope if fscope else [0], dtype=np.uint32)
        self._toff = np.array(toff if toff else [0], dtype=np.uint64)
        self._ftab = np.concatenate(tabs) if tabs else np.zeros(1)
        h = ctypes.c_void_p()
        ctx.check(ctx.L.bnpp_fg_create(
            ctx.h, len(self.cards), self.cards.ctypes.data_as(capi.c_u32p), len(factors),
            self._foff.ctypes.data_as(capi.c_i32p), self._fscope.ctypes.data_as(capi.c_u32p),
            self._toff.ctypes.data_as(capi.c_u64p), self._ftab.ctypes.data_as(capi.c_f64p), ctypes.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            self.ctx.L.bnpp_fg_destroy(self.h)
            self.h = None

    def sweep(self):
        e = ctypes.c_double()
        self.ctx.check(self.ctx.L.bnpp_fg_sweep(self.h, ctypes.byref(e)))
        return e.value

    def update(self, max_sweeps=10000, epsilon=0.001):
        """FactorGraph::update, code/graph.cpp:298-332 -> index of the converging sweep"""
        n = ctypes.c_uint32()
        self.ctx.check(self.ctx.L.bnpp_fg_update(self.h, max_sweeps, epsilon, ctypes.byref(n)))
        return n.value

    def reset(self):
        """messages back to the uniform start, code/graph.cpp:261-274"""
        self.ctx.L.bnpp_fg_reset.argtypes = [ctypes.c_void_p]
        self.ctx.check(self.ctx.L.bnpp_fg_reset(self.h))

    def marginals(self):
        """FactorGraph::marginal for every variable, code/graph.cpp:393-403 -> list of arrays"""
        out = np.zeros(int(self.cards.sum()))
        self.ctx.check(self.ctx.L.bnpp_fg_marginals(self.h, out.ctypes.data_as(capi.c_f64p)))
        off = [0] + [int(x) for x in np.cumsum(self.cards)]
        return [out[off[v]:off[v + 1]] for v in range(len(self.cards))]
