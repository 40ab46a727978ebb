#!/usr/bin/env python
"""Segment timings of the other variables (configs[2], [3]) on synthetic tiles: scale / fit / quantile map.
Not a bench line (see bench.py); numbers go to profiles/ as context."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.solver import CellBatchSolver, cell_seeds  # noqa: E402

C = int(os.environ.get("CELLS", 1024))
T = int(os.environ.get("DAYS", 43464))
dev = "cuda:0"
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
for var in os.environ.get("VARS", "pr,hurs,sfcWind,tasrange,rsds").split(","):
    # generate a few distinct cells on the host and tile them (generation cost, not GPU cost, limits C)
    base = synthetic.GENERATORS[var](min(C, 64), T)
    y = torch.from_numpy(np.tile(base, (1, (C + base.shape[1] - 1) // base.shape[1]))[:, :C].copy()).to(dev)
    s = CellBatchSolver(var, modes=4, device=dev)
    s.set_predictor(days, gmt, C)
    lat, lon = synthetic.tile_coordinates(1, C)
    seeds = cell_seeds(0, lat, lon) if var == "pr" else None
    best = [1e9] * 3
    for it in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        ys, sc = s.scale(y)
        ev[1].record()
        fit = s.fit(ys)
        ev[2].record()
        s.quantile_map(y, ys, fit.params, sc, seeds)
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        best = [min(a, b) for a, b in zip(best, t)]
    tot = sum(best)
    st = fit.status.cpu().numpy()
    print(f"{var:9s} C={C} T={T} scale {best[0]:8.2f} ms  fit {best[1]:9.2f} ms  qmap {best[2]:9.2f} ms  -> {C / tot * 1e3:9.0f} cells/s  "
          f"passes mean {fit.n_pass.float().mean().item():.1f} max {fit.n_pass.max().item()}  converged {np.mean(st == 0):.3f}", flush=True)
