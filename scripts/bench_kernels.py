#!/usr/bin/env python
"""Segment timings of the tas tile (scale / fit / quantile map) for kernel tuning; env toggles are read by the
library (ATTRICI_QMAP_V, ATTRICI_NO_FOLD, ...).  Not a bench line: see bench.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.solver import CellBatchSolver  # noqa: E402

C = int(os.environ.get("CELLS", 10000))
T = 43464
dev = "cuda:0"
y = synthetic.tas_tile_torch(C, T, seed=1234, device=dev, nan_fraction=0.001)
gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
s = CellBatchSolver("tas", modes=4, device=dev)
s.set_predictor(days, gmt, C)
cfact = torch.empty((T, C), dtype=torch.float64, device=dev)
keep = os.environ.get("KEEP_SCALED", "0") == "1"
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
best = [1e9] * 3
for it in range(6):
    ev[0].record()
    ys, sc = s.scale(y, None, keep_scaled=keep)
    ev[1].record()
    fit = s.fit(ys)
    ev[2].record()
    s.quantile_map(y, ys, fit.params, sc, out=cfact)
    ev[3].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    best = [min(a, b) for a, b in zip(best, t)]
tot = sum(best)
print(f"{os.environ.get('LABEL', '')} C={C} scale {best[0]:.3f} ms  fit {best[1]:.3f} ms  qmap {best[2]:.3f} ms  total {tot:.3f} ms "
      f"-> {C / tot * 1e3:.0f} cells/s   qmap {13.0 * T * C / best[2] / 1e6:.0f} GB/s  scale {8.0 * T * C / best[0] / 1e6:.0f} GB/s "
      f"passes {fit.n_pass.float().mean().item():.2f}")
