#!/usr/bin/env python
"""Condense an `ncu --page details --csv` export into one block per kernel launch (the metrics the
roofline discussion in DESIGN.md uses).  Usage: ncu_summary.py details.csv [kernel-substring]"""
import csv
import sys
from collections import OrderedDict

KEEP = ["Duration", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Compute (SM) Throughput",
        "Executed Ipc Active", "Issue Slots Busy", "Registers Per Thread", "Achieved Occupancy", "Theoretical Occupancy",
        "Grid Size", "Block Size", "Executed Instructions", "Warp Cycles Per Issued Instruction",
        "Eligible Warps Per Scheduler", "No Eligible", "Dynamic Shared Memory Per Block", "Static Shared Memory Per Block",
        "Local Load", "Mem Busy", "Max Bandwidth"]


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
    for (i, name), rs in launches.items():
        if want not in name:
            continue
        print(f"== launch {i}: {name[:110]}")
        for r in rs:
            m = r[idx["Metric Name"]]
            if m in KEEP or (r[idx["Section Name"]].startswith("Warp State") and "Stall" in m):
                print(f"   {m:45s} {r[idx['Metric Value']]:>16s} {r[idx['Metric Unit']]}")


if __name__ == "__main__":
    main()
