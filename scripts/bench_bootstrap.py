#!/usr/bin/env python
"""Throughput of the block bootstrap (attrici_b200.bootstrap.bootstrap_cells) at BASELINE length:
100 tas cells x 100 samples x 43 464 days = 10 000 refits, beside one refit by the oracle on the host."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.bootstrap import bootstrap_cells  # noqa: E402
from attrici_b200.solver import cell_seeds  # noqa: E402


def main():
    T = synthetic.DAYS_1901_2019
    Cn = int(os.environ.get("CELLS", 100))
    N = int(os.environ.get("SAMPLES", 100))
    y = synthetic.tas_cells(Cn, T, seed=5)
    gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
    years = (np.datetime64("1901-01-01") + days.astype("timedelta64[D]")).astype("datetime64[Y]").astype(int) + 1970
    lat, lon = synthetic.tile_coordinates(10, (Cn + 9) // 10)
    seeds = cell_seeds(0, lat[:Cn], lon[:Cn])
    yd = torch.as_tensor(y, device="cuda")
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = bootstrap_cells("tas", yd, days, gmt, years, seeds, N)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"run {it}: {Cn} cells x {N} samples x {T} days: {dt * 1e3:.1f} ms = {Cn * N / dt:,.0f} refits/s; "
              f"refit passes mean {out['refit_passes'].double().mean().item():.2f}, "
              f"std of the fitted mean, median over days and cells: {out['bootstrap_std'].median().item():.4f} K")


if __name__ == "__main__":
    main()
