#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: key metrics of every captured launch + hottest source lines.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
print(f"# ncu summary of {rep.split('/')[-1]}")
for r in rows[2:]:
    d = dict(zip(h, r))
    u = dict(zip(h, units))
    print()
    for k in KEYS:
        if k in d and d[k] != "":
            print(f"{k:70s} {d[k]} {u.get(k, '')}")
    stalls = sorted(((float(v), k) for k, v in d.items()
                     if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", k) and v), reverse=True)
    print("warp stall cycles per issued instruction:",
          ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in stalls[:8]))
print()
print("# hottest source lines (executed warp instructions, share of stall samples)")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_lines.py"), rep, "25"])
