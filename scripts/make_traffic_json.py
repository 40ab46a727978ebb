#!/usr/bin/env python
"""profiles/rN_traffic.json from `ncu --page raw --csv` exports: DRAM bytes and duration of the first launch of
each kernel (the file bench.py reads for roofline.traffic).

    python scripts/make_traffic_json.py OUT.json SOURCE_NOTE raw1.csv [raw2.csv ...]
"""
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def short_name(full):
    m = re.search(r"(\w+_kernel(?:<[^>]*>)?)", full)
    return m.group(1) if m else full


def main():
    out, note, files = sys.argv[1], sys.argv[2], sys.argv[3:]
    kernels = {}
    for path in files:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {n: hdr.index(n) for n in ("Kernel Name", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                         "gpu__time_duration.sum")}
        for r in rows[2:]:
            name = short_name(r[col["Kernel Name"]])
            if name in kernels:
                continue

            def val(n):
                return float(r[col[n]].replace(",", "")) * SCALE[units[col[n]]]

            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            kernels[name] = {"grid": r[col["Grid Size"]], "dram_bytes_read": rd, "dram_bytes_write": wr,
                             "duration_s_under_ncu": val("gpu__time_duration.sum"), "dram_bytes": rd + wr}
    json.dump({"source": note, "cells": 10000, "days": 43464, "kernels": kernels}, open(out, "w"), indent=1)
    for k, v in kernels.items():
        print(f"{k:48s} {v['dram_bytes'] / 1e9:7.3f} GB  {v['duration_s_under_ncu'] * 1e6:8.1f} us")


if __name__ == "__main__":
    main()
