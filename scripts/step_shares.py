#!/usr/bin/env python
"""Per-kernel share of one device-resident step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv`): the last group of launches that starts with the min/max pass of
the full 10 000-cell batch and ends before the next basis / min/max launch.

    python scripts/step_shares.py launches.csv [cells]
"""
import collections
import csv
import re
import sys


def short_name(full):
    m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", full)
    return m.group(1) if m else full


def main():
    path = sys.argv[1]
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(short_name(r[ki]), r[gi], float(r[vi].replace(",", "")) / 1000.0) for r in rows[1:]]
    want_x = ((cells + 127) // 128 + 3) // 4
    starts = [i for i, (k, g, _) in enumerate(seq) if k.startswith("prep_minmax") and g.startswith(f"({want_x},")]
    i0 = starts[-1]
    i1 = i0 + 1
    while i1 < len(seq) and not (seq[i1][0].startswith("prep_minmax") or seq[i1][0].startswith("build_basis")):
        i1 += 1
    tot = collections.OrderedDict()
    for k, _, v in seq[i0:i1]:
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    print(f"# one device-resident step ({cells} tas cells x 43 464 days) from {path.split('/')[-1]}")
    print("# ncu gpu__time_duration.sum, serialised launches, cold caches: the SHARE is what compares with bench.py's segments")
    print(f"# {'kernel':48s} launches   total us   share")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:48s} {n:5d}  {t:10.1f}  {100 * t / total:5.1f} %")
    print(f"  {'sum':48s}        {total:10.1f}")


if __name__ == "__main__":
    main()
