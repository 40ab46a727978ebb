#!/usr/bin/env python
"""Condense an `ncu --page source --csv` export (SASS view with warp-stall sampling) into: stall reasons, opcode mix by
executed instructions and by samples, and the hottest instructions.  Usage: ncu_source_hot.py source.csv [min_exec]
(min_exec splits the instructions into a hot loop region, executed at least that often, and the rest)."""
import collections
import csv
import sys

csv.field_size_limit(10**9)


def fresh():
    return ({"hot": dict(ex=0, samples=0, static=0, stalls=collections.Counter(), ops=collections.Counter(), ops_s=collections.Counter()),
             "rest": dict(ex=0, samples=0, static=0, stalls=collections.Counter(), ops=collections.Counter(), ops_s=collections.Counter())}, [])


def main():
    path = sys.argv[1]
    split = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    hdr = None
    regions, lines = fresh()
    for row in csv.reader(open(path, errors="ignore")):
        if row and row[0] == "Kernel Name":
            report(regions, lines, top)
            regions, lines = fresh()
            print("kernel:", row[1][:120])
            continue
        if row and row[0] == "Address":
            hdr = row
            continue
        if not hdr or len(row) < len(hdr) - 2:
            continue
        try:
            s, ex = int(row[4]), int(row[5])
        except ValueError:
            continue
        r = regions["hot" if ex >= split else "rest"]
        r["ex"] += ex
        r["samples"] += s
        r["static"] += 1
        t = row[1].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        op = ".".join(op.split(".")[:2])
        r["ops"][op] += ex
        r["ops_s"][op] += s
        lines.append((s, ex, row[1].strip()))
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try:
                    r["stalls"][h[6:]] += int(row[i])
                except ValueError:
                    pass
    report(regions, lines, top)


def report(regions, lines, top):
    if not lines:
        return
    for name, r in regions.items():
        if not r["ex"]:
            continue
        S = max(1, sum(r["stalls"].values()))
        print(f"== {name}: executed {r['ex']}, samples {r['samples']}, static instructions {r['static']}")
        print("   stalls:", ", ".join(f"{a}={100 * b / S:.0f}%" for a, b in r["stalls"].most_common(9)))
        print("   by executed:", ", ".join(f"{a}={100 * b / r['ex']:.1f}%" for a, b in r["ops"].most_common(24)))
        print("   by samples :", ", ".join(f"{a}={100 * b / max(1, r['samples']):.1f}%" for a, b in r["ops_s"].most_common(14)))
    lines.sort(reverse=True)
    print("== hottest instructions (samples, executed, SASS)")
    for s, ex, t in lines[:top]:
        print(f"   {s:6d} {ex:9d}  {t}")


if __name__ == "__main__":
    main()
