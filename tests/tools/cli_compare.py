#!/usr/bin/env python
"""Test infrastructure (it executes the reference under oracle/_ref).
`bn <model> -pr -mf` on every shipped Bayesian network: this build vs the unmodified reference
(oracle/_ref), partition line and the tools' own `Executed in` times.  Markdown table on stdout."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")


def run(exe, path, timeout):
    try:
        p = subprocess.run([exe, path, "-pr", "-mf"], capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return None, None
    z = p.stdout.splitlines()[0].replace(">> Partition = ", "") if p.stdout else "?"
    m = re.search(r"Executed in ([-+0-9.e]+)ms", p.stdout)
    return z, float(m.group(1)) if m else None


print("| network | Z (reference) | Z (this build) | reference ms | this build ms | speed-up |")
print("|---|---|---|---:|---:|---:|")
for path in sorted(glob.glob(os.path.join(REF, "models", "bayesnets", "*.uai"))):
    zr, tr = run(os.path.join(REF, "bn"), path, 200)
    zo, to = run(os.path.join(ROOT, "bnpp_b200", "bin", "bn"), path, 300)
    sp = "%.0fx" % (tr / to) if tr and to else "-"
    print("| %s | %s | %s | %s | %s | %s |" % (os.path.basename(path), zr if zr else "> 200 s", zo,
                                              "%.2f" % tr if tr else "-", "%.2f" % to if to else "-", sp), flush=True)
