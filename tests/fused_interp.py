"""Test infrastructure: a CPU interpreter of the fused-VE step program (bnpp_b200/csrc/fused.hpp).

The program a plan compiles itself into (step order, shared-memory arena layout, operand-offset
tables, evidence folding) is host work; this executes it with numpy on the host exactly as a
group of lanes of `ve_fused` would -- same step sequence, same arena offsets, the same order of
multiplications and additions per entry -- so the compiler of the program is pinned against
the reference's PR / MAR values without a GPU.  Plans are created DRY (ctx == NULL).
"""
import ctypes

import numpy as np

from bnpp_b200 import capi, model

HEADER_WORDS, OPERAND_WORDS = 8, 4
TO_RESULT, WANT_Z, TO_GLOBAL = 1, 2, 8
TABLE, GLOBAL = 0xffffffff, 0xfffffffe        # what the high word of a dumped address says it points to


class DryPlan:
    """bnpp_ve_plan / bnpp_mar_plan created without a context"""

    def __init__(self, cards, scopes, observed, order, marginals=False):
        L = capi.lib()
        model._declare(L)
        L.bnpp_ve_plan_fused_program.argtypes = [ctypes.c_void_p, ctypes.c_uint32, capi.c_u32p, ctypes.c_uint64, capi.c_u64p,
                                                 capi.c_u32p, ctypes.c_uint64, capi.c_u64p]
        self.L = L
        arr, self._keep = model._scopes(scopes, cards)
        ov, od = capi._u32(observed), capi._u32(order)
        h = ctypes.c_void_p()
        if marginals:
            c = capi._u32(cards)
            rc = L.bnpp_mar_plan_create(None, len(cards), ctypes.cast(c, capi.c_u32p), len(scopes), arr, len(observed),
                                        ctypes.cast(ov, capi.c_u32p), len(order), ctypes.cast(od, capi.c_u32p), ctypes.byref(h))
        else:
            rc = L.bnpp_ve_plan_create(None, len(scopes), arr, len(observed), ctypes.cast(ov, capi.c_u32p), len(order),
                                       ctypes.cast(od, capi.c_u32p), ctypes.byref(h))
        assert rc == 0, rc
        self.h = h
        self.n_obs = len(observed)
        self.layout = None
        if marginals:
            n = len(cards)
            off, size = (ctypes.c_uint32 * max(1, n))(), (ctypes.c_uint32 * max(1, n))()
            total = ctypes.c_uint64()
            assert L.bnpp_mar_plan_layout(h, n, ctypes.cast(off, capi.c_u32p), ctypes.cast(size, capi.c_u32p), ctypes.byref(total)) == 0
            self.layout = (list(off[:n]), list(size[:n]), total.value)

    def fused_info(self, nb=1):
        g, a, n = ctypes.c_int32(), ctypes.c_uint32(), ctypes.c_uint32()
        assert self.L.bnpp_ve_plan_fused_info(self.h, nb, ctypes.byref(g), ctypes.byref(a), ctypes.byref(n)) == 0
        return g.value, a.value, n.value

    def set_fused(self, on):
        assert self.L.bnpp_ve_plan_set_fused(self.h, int(on)) == 0

    def program(self, nb=1):
        """-> (prog words, offset-table words) as numpy uint32 arrays; empty when the plan is not fused"""
        np_, nt = ctypes.c_uint64(), ctypes.c_uint64()
        assert self.L.bnpp_ve_plan_fused_program(self.h, nb, None, 0, ctypes.byref(np_), None, 0, ctypes.byref(nt)) == 0
        prog = np.zeros(max(1, np_.value), dtype=np.uint32)
        tab = np.zeros(max(1, nt.value), dtype=np.uint32)
        assert self.L.bnpp_ve_plan_fused_program(self.h, nb, prog.ctypes.data_as(capi.c_u32p), prog.size, ctypes.byref(np_),
                                                 tab.ctypes.data_as(capi.c_u32p), tab.size, ctypes.byref(nt)) == 0
        return prog[:np_.value], tab[:nt.value]

    def segments(self, max_steps):
        """EXPERIMENTAL: cut the plan into fused segments of at most max_steps steps -> list of
        (first step, end step, lanes, arena doubles, prog, tab)"""
        L = self.L
        L.bnpp_ve_plan_set_segments.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32]
        L.bnpp_ve_plan_segments.argtypes = [ctypes.c_void_p, ctypes.c_uint32, capi.c_u32p, capi.c_u32p, capi.c_u32p,
                                            ctypes.POINTER(ctypes.c_int32), capi.c_u32p]
        L.bnpp_ve_plan_segment_program.argtypes = [ctypes.c_void_p, ctypes.c_uint32, capi.c_u32p, ctypes.c_uint64, capi.c_u64p,
                                                   capi.c_u32p, ctypes.c_uint64, capi.c_u64p]
        assert L.bnpp_ve_plan_set_segments(self.h, 1, max_steps) == 0
        n = ctypes.c_uint32()
        assert L.bnpp_ve_plan_segments(self.h, 0, ctypes.byref(n), None, None, None, None) == 0
        cap = max(1, n.value)
        a, b, ar = (ctypes.c_uint32 * cap)(), (ctypes.c_uint32 * cap)(), (ctypes.c_uint32 * cap)()
        g = (ctypes.c_int32 * cap)()
        assert L.bnpp_ve_plan_segments(self.h, cap, ctypes.byref(n), a, b, g, ar) == 0
        out = []
        for i in range(n.value):
            np_, nt = ctypes.c_uint64(), ctypes.c_uint64()
            assert L.bnpp_ve_plan_segment_program(self.h, i, None, 0, ctypes.byref(np_), None, 0, ctypes.byref(nt)) == 0
            prog = np.zeros(max(1, np_.value), dtype=np.uint32)
            tab = np.zeros(max(1, nt.value), dtype=np.uint32)
            assert L.bnpp_ve_plan_segment_program(self.h, i, prog.ctypes.data_as(capi.c_u32p), prog.size, ctypes.byref(np_),
                                                  tab.ctypes.data_as(capi.c_u32p), tab.size, ctypes.byref(nt)) == 0
            out.append((a[i], b[i], g[i], ar[i], prog[:np_.value], tab[:nt.value]))
        return out

    def launches(self):
        """-> (kernel launches of a single-query run with tasks on, groups, dependency levels)"""
        L = self.L
        L.bnpp_ve_plan_launches.argtypes = [ctypes.c_void_p, capi.c_u64p, capi.c_u32p, capi.c_u32p]
        n, g, lv = ctypes.c_uint64(), ctypes.c_uint32(), ctypes.c_uint32()
        assert L.bnpp_ve_plan_launches(self.h, ctypes.byref(n), ctypes.byref(g), ctypes.byref(lv)) == 0
        return n.value, g.value, lv.value

    def n_steps(self):
        vals = [ctypes.c_uint64() for _ in range(5)]
        assert self.L.bnpp_ve_plan_info(self.h, None, None, None, *[ctypes.byref(v) for v in vals]) == 0
        return vals[0].value

    def close(self):
        if self.h:
            self.L.bnpp_ve_plan_destroy(self.h)
            self.h = None


def interpret(prog, tab, n_steps, arena_doubles, tables, ev, result_size, glob=None, result=None):
    """one evidence set.  tables: list of 1-D float64 arrays (the resident CPTs); ev: evidence values in the
    plan's observed order.  glob: {intermediate index: array} -- the plan's global arena, shared by the segments of
    a launch-per-bucket plan (read for operands made by earlier launches, written by kFusedToGlobal steps).
    -> (result array, partition or None)"""
    arena = np.full(max(1, arena_doubles), np.nan)
    if result is None:
        result = np.full(max(1, result_size), np.nan)
    glob = {} if glob is None else glob
    z = None
    pc = 0
    prog = [int(w) for w in prog]
    for _ in range(n_steps):
        n_out, cx, kf, out_off, tab_off, dst_lo, dst_hi = prog[pc:pc + 7]
        pc += HEADER_WORDS
        k, flags = kf & 0xff, kf >> 8
        ops = []
        for q in range(k):
            w0, w1, sx, w3 = prog[pc:pc + OPERAND_WORDS]
            pc += OPERAND_WORDS
            kind, nobs = w0 & 0xff, w0 >> 8
            if kind == 0:
                ops.append((arena, w1, sx))
            elif w3 == GLOBAL:
                assert nobs == 0
                ops.append((glob[w1], 0, sx))      # KeyError: read before any launch wrote it
            else:
                assert w3 == TABLE
                base = 0
                for j in range(0, nobs, 2):
                    s0, i0, s1, i1 = prog[pc:pc + 4]
                    pc += 4
                    base += s0 * ev[i0]
                    if j + 1 < nobs:
                        base += s1 * ev[i1]
                ops.append((tables[w1], base, sx))
        if not flags & (TO_RESULT | TO_GLOBAL):      # the output of a step never overlaps an arena operand the step still reads
            for mem, base, sx in ops:
                if mem is arena:
                    for q2, (m2, b2, s2) in enumerate(ops):
                        if m2 is mem and b2 == base:
                            idx = b2 + tab[tab_off + q2 * n_out:tab_off + (q2 + 1) * n_out].astype(np.int64)
                            touched = np.concatenate([idx + x * s2 for x in range(cx)])
                            assert not ((touched >= out_off) & (touched < out_off + n_out)).any(), "output aliases a live operand"
        out = np.empty(n_out)
        for o in range(n_out):
            acc = 0.0
            for x in range(cx):
                v = None
                for q, (mem, base, sx) in enumerate(ops):
                    t = mem[base + int(tab[tab_off + q * n_out + o]) + x * sx]
                    v = t if v is None else v * t
                acc = v if x == 0 else acc + v
            out[o] = acc
        assert not np.isnan(out).any(), "a step read an entry nobody wrote (stale or unallocated arena slot)"
        if flags & TO_RESULT:
            result[out_off:out_off + n_out] = out
            if flags & WANT_Z:
                z = float(out.sum())
        elif flags & TO_GLOBAL:
            assert dst_hi == GLOBAL
            glob[dst_lo] = out
        else:
            assert out_off + n_out <= arena_doubles
            arena[out_off:out_off + n_out] = out
    assert pc == len(prog)
    return result, z


def parse_uai(text):
    """-> (cards, scopes, tables) of a UAI model text (code/io.cpp:43-100 semantics)"""
    toks = []
    for line in text.splitlines():
        for t in line.split():
            if t.startswith("#"):
                break
            toks.append(t)
    it = iter(toks)
    next(it)
    n = int(next(it))
    cards = [int(next(it)) for _ in range(n)]
    m = int(next(it))
    scopes = []
    for _ in range(m):
        w = int(next(it))
        scopes.append([int(next(it)) for _ in range(w)])
    tables = []
    for _ in scopes:
        sz = int(next(it))
        tables.append(np.array([float(next(it)) for _ in range(sz)]))
    return cards, scopes, tables
