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
    # a step starts with the first pass over the series: the one-pass scaling kernel (Normal family on a gap-free axis;
    # 256 cells per CTA) or the min / max pass (512 cells per CTA)
    first = {"prep_onepass_kernel": (cells + 255) // 256, "prep_minmax": ((cells + 127) // 128 + 3) // 4}

    def is_start(k, g):
        return any(k.startswith(n) and g.startswith(f"({x},") for n, x in first.items())

    starts = [i for i, (k, g, _) in enumerate(seq) if is_start(k, g)]
    i0 = starts[-1]
    i1 = i0 + 1
    while i1 < len(seq) and not (is_start(seq[i1][0], seq[i1][1]) or seq[i1][0].startswith("build_basis")):
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
