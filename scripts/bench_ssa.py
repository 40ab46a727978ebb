#!/usr/bin/env python
"""Timing of the SSA predictor smoothing (csrc/ssa.cu) on the reference's fixture series (4 493 values, window 365)
and on a daily series, beside the oracle's O(n W) NumPy version on the host."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from attrici_b200.preprocessing import ssa_first_component  # noqa: E402
from oracle import ssa_oracle  # noqa: E402  (checker / CPU timing only)


def timed(f, reps=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "ssa_component.npz"))
    x = torch.as_tensor(g["raw_gmt10"].astype(np.float64), device="cuda")
    _, info = ssa_first_component(x, 365, return_info=True)
    gpu = timed(lambda: ssa_first_component(x, 365))
    t0 = time.perf_counter()
    ssa_oracle.ssa_first_component(g["raw_gmt10"], 365)
    cpu = time.perf_counter() - t0
    print(f"fixture series n=4493 W=365: GPU {gpu * 1e3:.2f} ms ({info['iterations']} iterations), NumPy O(nW) oracle "
          f"{cpu * 1e3:.0f} ms; the reference's (W,W,K) tensor path took 16.6 s in the build container")
    n = 43464
    t = np.arange(n)
    xd = torch.as_tensor(287.0 + 1.2 * (t / n) ** 2 + 0.4 * np.sin(2 * np.pi * t / 365.25), device="cuda")
    print(f"daily series n={n} W=365: GPU {timed(lambda: ssa_first_component(xd, 365)) * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
