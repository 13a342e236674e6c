#!/usr/bin/env python
"""Brief per-kernel table from an .ncu-rep (duration, instructions, issue activity, occupancy, DRAM bytes, bank conflicts)."""
import csv
import subprocess
import sys

M = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
     "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
     "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
     "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for row in rows[2:]:
    d = dict(zip(h, row))
    name = d["Kernel Name"].split("(")[0]
    print(name)
    for m in M:
        print(f"   {m:75s} {d.get(m, '?')}")
