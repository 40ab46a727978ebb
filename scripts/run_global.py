#!/usr/bin/env python
"""BASELINE configs[4]: a global 0.5 degree land run (67 420 cells x 43 464 days by default) of tas and pr, cells
sharded over the ranks by the reference's partition rule, per-cell results gathered, and the merged NetCDF file
written from the shards (``io.write_merged_output_sharded``).  Prints one JSON line per variable with the wall time
of each stage; the synthetic input is generated per rank and not timed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_global.py \
        [--cells 67420 --days 43464 --variables tas,pr --out DIR --write-days N --no-write --keep]

Single process: ``python scripts/run_global.py --cells 8000``.  The land cells are drawn as a fixed pseudo-random
26 % of the 360 x 720 grid (the share of land in the ISIMIP mask), lat-major like attrici/detrend.py:672.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from attrici_b200 import io as aio  # noqa: E402
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.detrend import Config, gather_cells, get_task_indices  # noqa: E402
from attrici_b200.solver import CellBatchSolver, cell_seeds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=67420)
ap.add_argument("--days", type=int, default=synthetic.DAYS_1901_2019)
ap.add_argument("--variables", default="tas,pr")
ap.add_argument("--out", default="/tmp/attrici_global")   # not gpurun_out/: the full file is ~100 GB per variable
ap.add_argument("--no-write", action="store_true")
ap.add_argument("--write-days", type=int, default=0, help="write only the first N days (0 = all)")
ap.add_argument("--keep", action="store_true", help="keep the files (default: removed once their size is recorded)")
ap.add_argument("--slab-cells", type=int, default=2048)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
host_group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    host_group = dist.new_group(backend="gloo")      # host tensors of the final gather to NetCDF


def wall():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter()


# land mask: a fixed subset of the global 0.5 degree grid, stacked lat-major / lon-minor
glat, glon = synthetic.tile_coordinates(360, 720, lat0=-89.75, lon0=-179.75)
pick = np.sort(np.random.default_rng(2021).choice(glat.size, size=args.cells, replace=False))
lat, lon = glat[pick], glon[pick]
T, C = args.days, args.cells
days, gmt = synthetic.time_axis(T), synthetic.smoothed_gmt(T)
idx = get_task_indices(C, rank, world)
counts = [len(get_task_indices(C, r, world)) for r in range(world)]
make = {"tas": synthetic.tas_cells, "pr": synthetic.pr_cells}
os.makedirs(args.out, exist_ok=True)

for variable in args.variables.split(","):
    # every rank generates and solves only its shard (what detrend_cells does with a full [T, C] array)
    t0 = time.perf_counter()
    y_local = torch.from_numpy(make[variable](len(idx), T, first_cell=int(idx[0]) if len(idx) else 0)).pin_memory()
    t_synth = time.perf_counter() - t0
    cfg = Config(variable, slab_cells=args.slab_cells, device=f"cuda:{local_rank}")
    solver = CellBatchSolver(variable, device=cfg.device)
    seeds = cell_seeds(cfg.seed, lat[idx], lon[idx]) if variable == "pr" else None
    n, P = len(idx), solver.n_params
    out = {"cfact": torch.empty((T, n), dtype=torch.float64, pin_memory=True),      # pinned outputs allocated (and
           "replaced": torch.empty((T, n), dtype=torch.uint8, pin_memory=True),     # the scratch sized by a warm-up
           "params": torch.empty((n, P), dtype=torch.float64, pin_memory=True),     # slab) outside the timed region
           "logp": torch.empty(n, dtype=torch.float64, pin_memory=True),
           "scaling": torch.empty((n, 2), dtype=torch.float64, pin_memory=True),
           "status": torch.empty(n, dtype=torch.int32, pin_memory=True),
           "n_pass": torch.empty(n, dtype=torch.int32, pin_memory=True)}
    w = min(cfg.slab_cells, n)
    solver.detrend_host(y_local[:, :w], days, gmt, seeds=None if seeds is None else seeds[:w], slab_cells=w)
    a = wall()
    res = solver.detrend_host(y_local, days, gmt, seeds=seeds, slab_cells=cfg.slab_cells, out=out)
    b = wall()
    per_cell = {k: res[k].numpy() for k in ("params", "logp", "scaling", "status", "n_pass")}
    if world > 1:
        per_cell = gather_cells(per_cell, counts, dist, cfg.device)
    c = wall()
    path = os.path.join(args.out, f"{variable}_cfact_merged.nc")
    if not args.no_write:
        nw = args.write_days or T
        series = {"cfact": res["cfact"][:nw], "replaced": res["replaced"][:nw]}
        if world > 1:
            aio.write_merged_output_sharded(path, series, counts, lat, lon, days[:nw], dist, group=host_group,
                                            scalars={"logp": per_cell["logp"]}, shared={"gmt_scaled": gmt[:nw]})
        else:
            aio.write_merged_output(path, series, lat, lon, days[:nw], scalars={"logp": per_cell["logp"]},
                                    shared={"gmt_scaled": gmt[:nw]})
        if rank == 0:
            aio.write_merged_trace(os.path.join(args.out, f"{variable}_trace_merged.nc"), variable,
                                   per_cell["params"], per_cell["logp"], per_cell["scaling"], lat, lon)
    d = wall()
    if rank == 0:
        status = per_cell["status"]
        print(json.dumps({
            "variable": variable, "cells": C, "days": T, "n_gpus": world,
            "detrend_s": b - a, "cells_per_s": C / (b - a), "gather_s": c - b, "write_s": None if args.no_write else d - c,
            "file_bytes": os.path.getsize(path) if not args.no_write else None,
            "days_written": None if args.no_write else (args.write_days or T),
            "converged_fraction": float((status == 0).mean()), "passes_mean": float(per_cell["n_pass"].mean()),
            "synthetic_input_s_rank0": t_synth,
            "note": "detrend_s = H2D + scale + fit + quantile map + D2H, pipelined over cell slabs (not separable)",
        }), flush=True)
        if not args.no_write and not args.keep:
            os.remove(path)
    del res, y_local

if world > 1:
    dist.destroy_process_group()
