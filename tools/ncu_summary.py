#!/usr/bin/env python
"""Condense an .ncu-rep into the handful of counters DESIGN.md / bench.py cite: `python tools/ncu_summary.py REP OUT.txt`."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "sm__cycles_elapsed.max"]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none, condensed from {rep.split('/')[-1]} by tools/ncu_summary.py\n")
    names = [r[hdr.index("Kernel Name")].split("(")[0] for r in data]
    f.write("metric | unit | " + " | ".join(f"launch {i} ({n})" for i, n in enumerate(names)) + "\n")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            f.write(f"{k} | {units[i]} | " + " | ".join(r[i] for r in data) + "\n")
print(open(out).read())
