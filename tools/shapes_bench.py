#!/usr/bin/env python
"""Times the fused product+sum-out kernel on the SURVEY §8d headline shapes
(F-elem / F-bcast / F-small x sum-out position x B order) and prints achieved GB/s of
ALGORITHMIC bytes 8*(#A + #B + #out) against the measured HBM peak.

    python tools/shapes_bench.py [--bits 28] [--iters 5] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bnpp_b200 import capi  # noqa: E402
from bnpp_b200.factor import DeviceFactor  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def shapes(nbits):
    allv = list(range(nbits))
    small = allv[:: max(1, nbits // 10)][:10]
    out = []
    for kind, a, b in (("elem", allv, allv), ("bcast", allv[:-1], allv[1:]), ("small", allv, small)):
        for k in (0, nbits // 2, nbits - 1):
            for rev in (False, True):
                bb = list(b)
                if kind == "small" and k not in bb:
                    bb = sorted(set(bb[:-1] + [k]))
                if k not in a and k not in bb:
                    continue
                if rev:
                    bb = bb[::-1]
                out.append((kind, k, rev, list(a), bb))
    return out


def _time_launch(ctx, fn, iters):
    import torch
    s_ = ctx.torch_stream
    torch.cuda.synchronize()
    times = []
    for it in range(iters + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s_)
        fn()
        e1.record(s_)
        e1.synchronize()
        if it >= 2:
            times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2]      # the median: one hiccup (a clock sample, a page fault) must not carry a row


def run_mv(ctx, cards_list, iters=5, verbose=True, peak=None, how=""):
    """multi-valued elimination (contract_mvt / contract_mv): A over 13 axes of 4 values + the eliminated variable
    innermost (the canonical VE layout), B the same minus two axes, >= 2^27 union entries; K = 1 and K = 2"""
    g = torch.Generator(device="cuda").manual_seed(1)
    if peak is None:
        peak, how = peak_gbs()
    rows = []
    for cx in cards_list:
        nax = 13
        cards = [4] * nax + [cx]
        sa = list(range(nax + 1))
        sb = [v for v in sa if v not in (3, 9)]
        A = DeviceFactor.empty(ctx, sa, [cards[v] for v in sa])
        B = DeviceFactor.empty(ctx, sb, [cards[v] for v in sb])
        A.buf[:-1] = torch.rand(A.size, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
        B.buf[:-1] = torch.rand(B.size, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
        out_scope = sa[:-1]
        out = DeviceFactor.empty(ctx, out_scope, [cards[v] for v in out_scope])
        for k_ops, ops in ((1, [(A.ptr, sa, A.cards, None)]), (2, [(A.ptr, sa, A.cards, None), (B.ptr, sb, B.cards, None)])):
            ms = _time_launch(ctx, lambda: ctx.product_sum_out(ops, out_scope, out.cards, nax, out.ptr, out.zptr), iters)
            nbytes = 8 * (A.size + (B.size if k_ops == 2 else 0) + out.size)
            row = {"kind": "mv", "card": cx, "k": k_ops, "entries": A.size, "ms": ms, "GBs": nbytes / ms / 1e6,
                   "frac": nbytes / ms / 1e6 / peak, "entries_per_s": A.size / ms * 1e3, "kernel": ctx.last_launch()[0]}
            rows.append(row)
            if verbose:
                print("mv  cx=%2d K=%d  %.3e entries  %8.3f ms  %8.1f GB/s  %5.1f%% of %s peak  %.3e entries/s  %s"
                      % (cx, k_ops, A.size, ms, row["GBs"], 100 * row["frac"], how, row["entries_per_s"], row["kernel"]), flush=True)
        del A, B, out
    return rows


def run_mixed(ctx, iters=5, verbose=True, peak=None, how=""):
    """the mixed-cardinality case of SURVEY 8d: cards cycle 2,3,4,5 over 16 axes (2.07e8 union entries)"""
    g = torch.Generator(device="cuda").manual_seed(1)
    if peak is None:
        peak, how = peak_gbs()
    rows = []
    cards = [2, 3, 4, 5] * 4
    allv = list(range(16))
    for k in (0, 1, 6, 15):
        sa, sb = allv, [v for v in allv if v != (k + 3) % 16]      # B lacks one axis, keeps the eliminated one
        A = DeviceFactor.empty(ctx, sa, [cards[v] for v in sa])
        B = DeviceFactor.empty(ctx, sb, [cards[v] for v in sb])
        A.buf[:-1] = torch.rand(A.size, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
        B.buf[:-1] = torch.rand(B.size, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
        out_scope = [v for v in allv if v != k]
        out = DeviceFactor.empty(ctx, out_scope, [cards[v] for v in out_scope])
        ops = [(A.ptr, sa, A.cards, None), (B.ptr, sb, B.cards, None)]
        ms = _time_launch(ctx, lambda: ctx.product_sum_out(ops, out_scope, out.cards, k, out.ptr, out.zptr), iters)
        nbytes = 8 * (A.size + B.size + out.size)
        row = {"kind": "mixed", "k": k, "card": cards[k], "entries": A.size, "ms": ms, "GBs": nbytes / ms / 1e6,
               "frac": nbytes / ms / 1e6 / peak, "entries_per_s": A.size / ms * 1e3, "kernel": ctx.last_launch()[0]}
        rows.append(row)
        if verbose:
            print("mixed  k=%2d card=%d  %8.3f ms  %8.1f GB/s  %5.1f%% of %s peak  %.3e entries/s  %s"
                  % (k, cards[k], ms, row["GBs"], 100 * row["frac"], how, row["entries_per_s"], row["kernel"]), flush=True)
        del A, B, out
    return rows


def run_binary(ctx, bits=28, iters=5, only=None, no_reverse=False, verbose=True, peak=None, how=""):
    """F-elem / F-bcast / F-small x sum-out position x B order, binary variables"""
    g = torch.Generator(device="cuda").manual_seed(1)
    if peak is None:
        peak, how = peak_gbs()
    cache = {}

    def table(scope, which):
        key = (len(scope), which)     # A and B must never alias: the kernel would read one table only
        if key not in cache:
            f = DeviceFactor.empty(ctx, list(range(key[0])), [2] * key[0])
            f.buf[:-1] = torch.rand(f.size, generator=g, device="cuda", dtype=torch.float64) * 0.9 + 0.1
            cache[key] = f
        return cache[key]

    rows = []
    for kind, k, rev, sa, sb in shapes(bits):
        if only and kind != only:
            continue
        if rev and no_reverse:
            continue
        A, B = table(sa, 'A'), table(sb, 'B')
        union = sa + [v for v in sb if v not in sa]
        out_scope = [v for v in union if v != k]
        out = DeviceFactor.empty(ctx, out_scope, [2] * len(out_scope))
        ops = [(A.ptr, sa, [2] * len(sa), None), (B.ptr, sb, [2] * len(sb), None)]
        ms = _time_launch(ctx, lambda: ctx.product_sum_out(ops, out_scope, [2] * len(out_scope), k, out.ptr, out.zptr), iters)
        nbytes = 8 * (A.size + B.size + out.size)
        gbs = nbytes / ms / 1e6
        entries = 2 ** len(union)
        row = {"kind": kind, "k": k, "reverse_b": rev, "union_bits": len(union), "ms": ms, "GBs": gbs,
               "frac": gbs / peak, "entries_per_s": entries / ms * 1e3, "kernel": ctx.last_launch()[0]}
        rows.append(row)
        if verbose:
            print("%-6s k=%2d rev=%d  %8.3f ms  %8.1f GB/s  %5.1f%% of %s peak  %.3e entries/s  %s"
                  % (kind, k, rev, ms, gbs, 100 * gbs / peak, how, row["entries_per_s"], row["kernel"]), flush=True)
        del out
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=28)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="kind filter")
    ap.add_argument("--no-reverse", action="store_true")
    ap.add_argument("--mixed", action="store_true", help="the mixed-cardinality case of SURVEY 8d: cards cycle 2,3,4,5 over 16 axes (2.07e8 union entries)")
    ap.add_argument("--mv", default=None, help="multi-valued elimination: comma-separated cardinalities of the eliminated variable, e.g. 3,4,5,7,8")
    args = ap.parse_args()
    ctx = capi.Context(0)
    peak, how = peak_gbs()
    if args.mv:
        rows = run_mv(ctx, [int(c) for c in args.mv.split(",")], args.iters, peak=peak, how=how)
    elif args.mixed:
        rows = run_mixed(ctx, args.iters, peak=peak, how=how)
    else:
        rows = run_binary(ctx, args.bits, args.iters, args.only, args.no_reverse, peak=peak, how=how)
    if args.json:
        json.dump({"peak_gbs": peak, "peak_kind": how, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
