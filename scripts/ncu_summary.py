#!/usr/bin/env python
"""Condense an `ncu --page details --csv` export into one block per kernel launch (the metrics the
roofline discussion in DESIGN.md uses); with the `--page raw --csv` export of the same report as third argument
also the pipe and DRAM counters that the details page does not print (FP64 / tensor / LSU pipes, DRAM bytes, stall
ratios per issued instruction).  Usage: ncu_summary.py details.csv [kernel-substring] [raw.csv]"""
import csv
import sys
from collections import OrderedDict

KEEP = ["Duration", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput",
        "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy",
        "Grid Size", "Block Size", "Executed Instructions", "Warp Cycles Per Issued Instruction",
        "Eligible Warps Per Scheduler", "No Eligible", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block",
        "Local Load", "Mem Busy", "Max Bandwidth"]


RAW_KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.per_cycle_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    h = rows[0]
    idx = {n: h.index(n) for n in ("ID", "Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value")}
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    launches = OrderedDict()
    for r in rows[1:]:
        if len(r) <= idx["Metric Value"]:
            continue
        k = (r[idx["ID"]], r[idx["Kernel Name"]])
        launches.setdefault(k, []).append(r)
    raw = {}
    if len(sys.argv) > 3:
        rr = list(csv.reader(open(sys.argv[3], errors="ignore")))
        rh, ru = rr[0], rr[1]
        for r in rr[2:]:
            raw[r[rh.index("ID")]] = {k: (v, u) for k, v, u in zip(rh, r, ru)}
    for (i, name), rs in launches.items():
        if want not in name:
            continue
        print(f"== launch {i}: {name[:110]}")
        for k in RAW_KEEP:
            if i in raw and k in raw[i]:
                v, u = raw[i][k]
                print(f"   {k:80s} {v:>16s} {u}")
        for r in rs:
            m = r[idx["Metric Name"]]
            if m in KEEP or (r[idx["Section Name"]].startswith("Warp State") and "Stall" in m):
                print(f"   {m:45s} {r[idx['Metric Value']]:>16s} {r[idx['Metric Unit']]}")


if __name__ == "__main__":
    main()
