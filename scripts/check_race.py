#!/usr/bin/env python
"""Determinism check of the scale -> fit path on the GPU: the same big batch detrended several times must give
bit-identical coefficients (a handshake race in a streaming kernel shows up as a few cells that differ between
runs).  Usage: check_race.py [cells] [runs] [save.npy | compare.npy]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.solver import CellBatchSolver  # noqa: E402


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 67420
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    path = sys.argv[3] if len(sys.argv) > 3 else None
    T = synthetic.DAYS_1901_2019
    dev = torch.device("cuda:0")
    y = synthetic.tas_tile_torch(C, T, device=dev, nan_fraction=0.001)
    gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
    s = CellBatchSolver("tas", device=dev)
    first = None
    for r in range(runs):
        out = s.detrend_device(y, days, gmt, keep_scaled=False)
        p = out["fit"].params.cpu().numpy().copy()
        sc = out["scaling"].cpu().numpy().copy() if "scaling" in out else np.zeros((C, 2))
        chk = float(torch.nan_to_num(out["cfact"]).sum().item())
        rep = int(out["replaced"].sum().item())
        del out
        if first is None:
            first = (p, chk, rep, sc)
            print(f"run 0: params checksum {p.sum():.17g} cfact checksum {chk:.17g} replaced {rep}")
            continue
        bad = np.nonzero((p != first[0]).any(axis=1))[0]
        print(f"run {r}: {bad.size} cells differ from run 0; cfact checksum equal {chk == first[1]}; replaced equal {rep == first[2]}")
        if bad.size:
            print("   cells", bad[:24], "cell % 256", (bad % 256)[:24], "max |d|", np.abs(p - first[0]).max())
            print("   by warp of the CTA (cell % 256 // 64):", np.bincount((bad % 256) // 64, minlength=4),
                  "by cell % 256 // 32:", np.bincount((bad % 256) // 32, minlength=8))
            print("   CTAs touched:", np.unique(bad // 256).size, "of", (C + 255) // 256, "| scaling differs in",
                  int((sc != first[3]).any(axis=1).sum()), "cells")
            ctas, counts = np.unique(bad // 256, return_counts=True)
            print("   bad cells per touched CTA: min", counts.min(), "median", int(np.median(counts)), "max", counts.max())
    if path and os.path.exists(path):
        q = np.load(path)
        bad = np.nonzero((q != first[0]).any(axis=1))[0]
        print(f"against {path}: {bad.size} cells differ", bad[:24], (bad % 256)[:24], np.abs(q - first[0]).max() if bad.size else 0.0)
    elif path:
        np.save(path, first[0])
        print("saved", path)


if __name__ == "__main__":
    main()
