#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions, stall samples and shared-memory wavefronts of an .ncu-rep (needs -lineinfo and
--import-source on).  Usage: ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
key = 1 if (len(sys.argv) > 3 and sys.argv[3] == "smp") else 0      # sort by executed instructions (default) or by samples
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, ""])
fname, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        col = {}
        for k, v in zip(hdr, r):
            col.setdefault(k, v)          # the first "Source" column is the CUDA line
        try:
            ex = float(col.get("Instructions Executed") or 0)
            smp = float(col.get("# Samples") or 0)
            wf = float(col.get("L1 Wavefronts Shared") or 0)
            wfi = float(col.get("L1 Wavefronts Shared Ideal") or 0)
        except ValueError:
            continue
        if not col["Line No"].strip():
            continue                      # per-file subtotal rows
        a = agg[(fname, col["Line No"])]
        a[0] += ex; a[1] += smp; a[2] += wf; a[3] += wfi; a[4] = col["Source"].strip()[:86]
tot = sum(a[0] for a in agg.values()) or 1.0
tots = sum(a[1] for a in agg.values()) or 1.0
print(f"total executed warp instructions {tot:.4g}, samples {tots:.0f}")
print("file:line              ex%    smp%   smem wavefronts (ideal)   source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
    print(f"{f + ':' + ln:22s} {100 * a[0] / tot:5.2f}  {100 * a[1] / tots:5.2f}  {a[2]:12.4g} ({a[3]:10.4g})  {a[4]}")

# optional: sums over line ranges of one file:  ncu_lines.py rep N key file:lo-hi[:name] ...
if len(sys.argv) > 4:
    print("-- ranges")
    for spec in sys.argv[4:]:
        parts = spec.split(":")
        f, rng = parts[0], parts[1]
        name = parts[2] if len(parts) > 2 else rng
        lo, hi = (int(x) for x in rng.split("-"))
        ex = sum(a[0] for (ff, ln), a in agg.items() if ff == f and lo <= int(ln) <= hi)
        sm = sum(a[1] for (ff, ln), a in agg.items() if ff == f and lo <= int(ln) <= hi)
        print(f"  {name:28s} ex {100 * ex / tot:5.2f}%  smp {100 * sm / tots:5.2f}%")
    others = defaultdict(lambda: [0.0, 0.0])
    for (ff, ln), a in agg.items():
        others[ff][0] += a[0]; others[ff][1] += a[1]
    for ff, a in sorted(others.items(), key=lambda kv: -kv[1][0]):
        print(f"  file {ff:30s} ex {100 * a[0] / tot:5.2f}%  smp {100 * a[1] / tots:5.2f}%")
