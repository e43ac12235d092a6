#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS hot spots) into a small text file suitable for profiles/."""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, v = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
for name in want:
    for i, n in enumerate(h):
        if n == name:
            print(f"{name} = {v[i]} {units[i]}")
print("-- warp stall reasons (per issue active)")
for i, n in enumerate(h):
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", n)
    if m and float(v[i] or 0) > 0.05:
        print(f"  {m.group(1):24s} {float(v[i]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh = rows[1]
data = rows[2:]
isrc, ie, iss = hh.index("Source"), hh.index("Instructions Executed"), hh.index("# Samples")
tot = sum(float(r[ie] or 0) for r in data)
tots = sum(float(r[iss] or 0) for r in data)
ops = Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[isrc])
    if m:
        ops[m.group(2)] += float(r[ie] or 0)
print(f"-- executed warp instructions by opcode (total {tot:.4g})")
for op, c in ops.most_common(22):
    print(f"  {op:10s} {100 * c / tot:5.1f}%")
print("-- top sampled SASS instructions")
for r in sorted(data, key=lambda r: -float(r[iss] or 0))[:14]:
    print(f"  {100 * float(r[iss] or 0) / tots:5.2f}% smp {100 * float(r[ie] or 0) / tot:5.2f}% ex  {r[isrc].strip()[:80]}")
