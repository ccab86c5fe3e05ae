"""Test infrastructure: a CPU emulator of a launch-per-bucket VE plan.

`bnpp_ve_plan_describe` dumps what every launch of a plan reads and writes and where each intermediate lives in the
plan's arena.  This executes that schedule with numpy on ONE flat arena array -- evidence as base offsets into the
resident tables, file-order strides for the input views, first-fit arena slots with their planned lifetimes -- so the
host logic of wide plans (bucket schedule, folding of small tables, operand-count shrinking, canonical axis orders,
arena reuse) is pinned against the oracle without a GPU.  A wrong lifetime shows up as a wrong number: a slot that
is overwritten while still needed corrupts a later step's operand."""
import ctypes

import numpy as np

from bnpp_b200 import capi


def describe(plan_handle):
    L = capi.lib()
    L.bnpp_ve_plan_describe.argtypes = [ctypes.c_void_p, capi.c_u64p, ctypes.c_uint64, capi.c_u64p]
    n = ctypes.c_uint64()
    assert L.bnpp_ve_plan_describe(plan_handle, None, 0, ctypes.byref(n)) == 0
    buf = (ctypes.c_uint64 * max(1, n.value))()
    assert L.bnpp_ve_plan_describe(plan_handle, buf, n.value, ctypes.byref(n)) == 0
    w = list(buf[:n.value])
    it = iter(w)
    nf, ns, arena, result_size = next(it), next(it), next(it), next(it)
    factors, steps = [], []
    for _ in range(nf):
        src, size, off, rank = next(it) - 1, next(it), next(it), next(it)
        axes = [(next(it), next(it), next(it)) for _ in range(rank)]
        nobs = next(it)
        obs = [(next(it), next(it)) for _ in range(nobs)]
        factors.append({"src": src, "size": size, "off": off, "axes": axes, "obs": obs})
    for _ in range(ns):
        elim, out, roff, want_z, k = next(it) - 1, next(it) - 1, next(it), next(it), next(it)
        ops = [next(it) for _ in range(k)]
        rank = next(it)
        scope = [(next(it), next(it)) for _ in range(rank)]
        steps.append({"elim": elim, "out": out, "roff": roff, "want_z": want_z, "ops": ops, "scope": scope})
    assert next(it, None) is None
    return {"factors": factors, "steps": steps, "arena": arena, "result_size": result_size}


def run(desc, tables, ev):
    """-> (result array, partition of the last want_z step or None)"""
    arena = np.full(max(1, desc["arena"]), np.nan)
    result = np.full(max(1, desc["result_size"]), np.nan)
    z = None

    def operand(fid):
        f = desc["factors"][fid]
        shape = [c for _, c, _ in f["axes"]]
        strides = [8 * s for _, _, s in f["axes"]]
        if f["src"] >= 0:
            base = sum(st * ev[col] for st, col in f["obs"])
            mem = tables[f["src"]][base:]
        else:
            mem = arena[f["off"]:f["off"] + f["size"]]
            assert not np.isnan(mem).any(), "operand slot holds no finished table (lifetime error)"
        if not shape:
            return [], np.asarray(mem[0])
        return [v for v, _, _ in f["axes"]], np.lib.stride_tricks.as_strided(mem, shape=shape, strides=strides)

    for st in desc["steps"]:
        out_vars = [v for v, _ in st["scope"]]
        union = out_vars + ([st["elim"]] if st["elim"] >= 0 else [])
        card = dict(st["scope"])
        acc = None
        for fid in st["ops"]:
            vars_, a = operand(fid)
            for v, c, _ in desc["factors"][fid]["axes"]:
                card[v] = c
            if vars_:
                order = sorted(range(len(vars_)), key=lambda i: union.index(vars_[i]))
                a = np.transpose(a, order)
                have = [vars_[i] for i in order]
                a = a.reshape([card[v] if v in have else 1 for v in union])
            else:
                a = a.reshape([1] * len(union))
            acc = a.astype(np.float64) if acc is None else acc * a
        acc = np.broadcast_to(acc, [card[v] for v in union])
        if st["elim"] >= 0:
            tot = acc[..., 0].copy()
            for x in range(1, card[st["elim"]]):
                tot = tot + acc[..., x]
            acc = tot
        flat = np.ascontiguousarray(acc).reshape(-1)
        if st["out"] < 0:
            result[st["roff"]:st["roff"] + flat.size] = flat
            if st["want_z"]:
                z = float(flat.sum())
        else:
            f = desc["factors"][st["out"]]
            assert f["size"] == flat.size and [v for v, _, _ in f["axes"]] == out_vars
            arena[f["off"]:f["off"] + flat.size] = flat
    return result, z
