#!/usr/bin/env python3
"""Print the handful of ncu metrics we track (dev tool): python tools/ncu_keys.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__average_cache_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum",
        "gcc__xbar2gcc_sectors.sum", "gcc__xbar2gcc_sectors.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "smsp__inst_executed_op_branch.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
d = dict(zip(hdr, zip(units, vals)))
for k in keys:
    if k in d:
        print(f"{k:90s} {d[k][1]:>20s} {d[k][0]}")
