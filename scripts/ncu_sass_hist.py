#!/usr/bin/env python
"""Opcode histogram (weighted by executed warp instructions) and stall-sample hot spots per kernel from an
`ncu --page source --csv` export.  Usage: ncu_sass_hist.py source.csv kernel-substring [launch-index]"""
import collections
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
want = sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
sel = [b for b in blocks if want in b["name"]]
b = sel[which]
ops, samples = collections.Counter(), []
tot = 0
for r in b["rows"]:
    sass, smp, ex = r[1].strip(), int(r[2] or 0), int(r[5] or 0)
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ops[op.split(".")[0]] += ex
    tot += ex
    samples.append((smp, ex, sass))
print(b["name"][:100], "executed warp instructions", tot)
for op, n in ops.most_common(25):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot:5.1f}%")
print("top stall-sample instructions:")
for smp, ex, sass in sorted(samples, reverse=True)[:18]:
    print(f"  {smp:7d} samples  {ex:10d} exec  {sass[:90]}")
