#!/usr/bin/env python
"""Turns the raw `ncu --metrics gpu__time_duration.sum --csv` log of a bench.py run into the per-step launch list kept under profiles/.
usage: launch_list_summary.py launches.csv step_index > profiles/...csv     (a step = the launches from one aabb_kernel to the next)"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
step_index = int(sys.argv[2]) if len(sys.argv) > 2 else -2
launch = [(int(r[0]), r[4].split("(")[0].replace("void ", ""), float(r[14].replace(",", "")) / 1000.0) for r in rows]
ours = [l for l in launch if "tnsb::" in l[1]]
starts = [i for i, l in enumerate(ours) if "aabb_kernel" in l[1]]
lo = starts[step_index]
hi = starts[step_index + 1] if step_index + 1 < len(starts) and step_index != -1 else len(ours)
step = ours[lo:hi]
tot = sum(l[2] for l in step)
agg = OrderedDict()
for _, k, us in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
print(f"# {len(step)} launches of our kernels in this step, {tot / 1000.0:.3f} ms in total under ncu (per-launch times are cold-cache and serialised:")
print("# compare SHARES, not absolutes)")
print("kernel,launches,duration_us,share_pct")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{us:.1f},{100.0 * us / tot:.1f}")
print("# per launch, in order")
print("id,kernel,duration_us")
for i, k, us in step:
    print(f"{i},{k},{us:.1f}")
