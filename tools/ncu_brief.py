#!/usr/bin/env python
"""Print the handful of ncu raw metrics we look at first.  usage: tools/ncu_brief.py report.ncu-rep [row]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, v = r[0], r[2 + row]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size"]
for i, n in enumerate(h):
    if n in want:
        print(f"{n:70s} {r[1][i]:14s} {v[i]}")
st = [(float(v[i]), n) for i, n in enumerate(h) if "issue_stalled" in n and n.endswith("per_issue_active.ratio") and v[i]]
for x, n in sorted(st, reverse=True)[:8]:
    print(f"  stall {n.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {x:.3f}")
