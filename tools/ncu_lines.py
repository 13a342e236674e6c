#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py REPORT KERNEL [min_instr_per_warp]
"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern,
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, h = None, None
agg = {}
warps = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 4 and r[0] == "Line No":
        h = r
        iI, iS = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) < len(h) or not r[0]:
        continue
    try:
        n = int(r[iI]); smp = int(r[iS])
    except ValueError:
        continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += n; a[1] += smp
grid = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", kern, "--launch-skip", "0", "--launch-count", "1",
                       "--metrics", "launch__grid_size,launch__block_size"], capture_output=True, text=True).stdout
g = list(csv.reader(grid.splitlines()))
d = dict(zip(g[0], g[2]))
warps = int(d["launch__grid_size"]) * int(d["launch__block_size"]) // 32
tot = sum(a[0] for a in agg.values())
smp = sum(a[1] for a in agg.values())
print(f"{kern}: {tot / warps:.0f} instr/warp, {smp} samples, {warps} warps")
for (f, ln), a in sorted(agg.items()):
    if a[0] / warps >= thresh:
        print(f"{a[0] / warps:7.1f} {100.0 * a[1] / max(smp, 1):5.1f}%  {f}:{ln:<5d} {a[2].strip()[:130]}")
