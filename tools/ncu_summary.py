#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into the few numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, io

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name"), " grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
    for k in keys:
        if k in d:
            print(f"  {k:95s} {d[k]:>18s} {u[k]}")
