#!/usr/bin/env python
"""Turns ncu outputs into the small, committed summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r1_launches.csv  > profiles/r1_launches.md
  python tools/ncu_summary.py full     gpurun_out/r1_canon.ncu-rep > profiles/r1_canon_full.md
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    im = hdr.index("Metric Name") if "Metric Name" in hdr else None
    iu = hdr.index("Metric Unit") if "Metric Unit" in hdr else None
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if im is not None and r[im] != "gpu__time_duration.sum":
            continue                     # a launch list may carry other metrics (DRAM bytes) beside the durations
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        if iu is not None and r[iu] in ("us", "usecond"):
            v *= 1e3
        elif iu is not None and r[iu] in ("ms", "msecond"):
            v *= 1e6
        name = r[ik].split("(bnpp::")[0].replace("void ", "").replace("bnpp::", "")
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, v in tot.most_common():
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, cnt[k], v / 1e3, 100 * v / s))
    print("\ntotal %.1f us over %d launches (ncu-serialised, cold-cache: compare shares, not absolutes)" % (s / 1e3, sum(cnt.values())))


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
]


def full(path, min_us=50.0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(w, lab, hdr.index(w)) for w, lab in WANT if w in hdr]
    ik = hdr.index("Kernel Name")
    print("| # | kernel | " + " | ".join("%s (%s)" % (lab, units[i]) if units[i] else lab for w, lab, i in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for n, r in enumerate(data):
        t = float(r[hdr.index("gpu__time_duration.sum")])
        t_us = t * (1e3 if units[hdr.index("gpu__time_duration.sum")] == "ms" else 1.0)
        if t_us < min_us:
            continue
        name = r[ik].replace("void ", "").split("(")[0]
        vals = []
        for w, lab, i in cols:
            try:
                vals.append("%.4g" % float(r[i].replace(",", "")))
            except ValueError:
                vals.append(r[i])
        print("| %d | `%s` | " % (n, name) + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
