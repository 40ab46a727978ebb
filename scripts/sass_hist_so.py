#!/usr/bin/env python
"""Static SASS opcode histogram of the shipped library (cuobjdump -sass attrici_b200/libattrici_b200.so): the mnemonics
that show which hardware paths the kernels use - UBLKCP (cp.async.bulk, the TMA engine's bulk copies), SYNCS (mbarrier),
HMMA (mma.sync TF32), LDGSTS (cp.async), DFMA / MUFU - per kernel for the kernels named on the command line.

    python scripts/sass_hist_so.py [kernel-substring ...] > profiles/rN_sass_hist.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "UTCMMA", "UTCHMMA", "LDTM", "STTM", "SYNCS", "HMMA", "DMMA", "LDGSTS", "DFMA", "FFMA", "MUFU",
         "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "NANOSLEEP", "FENCE", "MEMBAR"]


def main():
    wanted = sys.argv[1:]
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "attrici_b200", "libattrici_b200.so")], capture_output=True,
                         text=True, check=True).stdout
    total, per = collections.Counter(), collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"attrici::\(anonymous namespace\)::", "", name)
            cur = next((name for w in wanted if w in name), None)
            if cur:
                per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            total[op] += 1
            if cur:
                per[cur][op] += 1
    print("# static SASS instruction counts, sm_100a, whole library:", sum(total.values()), "instructions")
    print("  " + "  ".join(f"{k} {total[k]}" for k in WATCH if total[k]))
    print("  absent:", " ".join(k for k in WATCH if not total[k]))
    for name, c in per.items():
        n = sum(c.values())
        print(f"\n{name[:150]}\n  {n} instructions: " + "  ".join(f"{k} {c[k]}" for k in WATCH if c[k]))


if __name__ == "__main__":
    main()
