#!/usr/bin/env python
"""BASELINE configs[4]: a global 0.5 degree land run (67 420 cells x 43 464 days by default) of tas and pr, cells
sharded over the ranks by the reference's partition rule, per-cell results gathered, and the merged NetCDF file
written from the shards (``io.write_merged_output_sharded``).  Prints one JSON line per variable with the wall time
of each stage; the synthetic input is generated per rank and not timed.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_global.py \
        [--cells 67420 --days 43464 --variables tas,pr --out DIR --write-days N --write-budget-s S --no-write --keep]

Single process: ``python scripts/run_global.py --cells 8000``.  The land cells are drawn as a fixed pseudo-random
26 % of the 360 x 720 grid (the share of land in the ISIMIP mask), lat-major like attrici/detrend.py:672.

Stages reported per variable (wall seconds, max over ranks through the barriers):
  detrend_s   H2D + scale + fit + quantile map + D2H through ``CellBatchSolver.detrend_host`` (pipelined over cell slabs)
  device      CUDA-event times of scale / fit / quantile map for one slab resident in HBM (what the pipeline overlaps
              with the copies), and the PCIe volumes of the run
  gather_s    final gather of the per-cell results (coefficients, logp, scaling, status) to every rank
  write_s     final host gather of the counterfactual series + streaming of the merged file on rank 0
  check       --check N: N cells recomputed on rank 0 by a single-GPU device-resident run, and 2 of them by the oracle
"""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from attrici_b200 import io as aio  # noqa: E402
from attrici_b200 import synthetic  # noqa: E402
from attrici_b200.detrend import Config, bind_to_gpu_cpus, gather_cells, get_task_indices  # noqa: E402
from attrici_b200.solver import CellBatchSolver, cell_seeds  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=67420)
ap.add_argument("--days", type=int, default=synthetic.DAYS_1901_2019)
ap.add_argument("--variables", default="tas,pr")
ap.add_argument("--out", default="/tmp/attrici_global")   # not gpurun_out/: the full file is tens of GB per variable
ap.add_argument("--no-write", action="store_true")
ap.add_argument("--write-days", type=int, default=0, help="write only the first N days (0 = all)")
ap.add_argument("--write-variables", default="tas", help="variables whose merged series file is written")
ap.add_argument("--keep", action="store_true", help="keep the files (default: removed once their size is recorded)")
ap.add_argument("--slab-cells", type=int, default=2048)
ap.add_argument("--cfact-dtype", default="float64", choices=["float64", "float32"])
ap.add_argument("--file-dtype", default="float32", choices=["float64", "float32"],
                help="dtype of cfact in the merged file (float32 = the dtype of ISIMIP input / output files)")
ap.add_argument("--check", type=int, default=6, help="cells re-solved on one GPU (and 2 of them by the CPU oracle)")
ap.add_argument("--json", default=None, help="also append the JSON lines to this file (rank 0)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
bind_to_gpu_cpus(local_rank)
host_group = None
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    host_group = dist.new_group(backend="gloo")      # host tensors of the final gather to NetCDF


def wall():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter()


def emit(line):
    text = json.dumps(line)
    print(text, flush=True)
    if args.json:
        with open(args.json, "a") as f:
            f.write(text + "\n")


# land mask: a fixed subset of the global 0.5 degree grid, stacked lat-major / lon-minor
# (latitudes -56 .. 84 like the ISIMIP land mask: 280 x 720 grid points, a third of them land)
glat, glon = synthetic.tile_coordinates(280, 720, lat0=-55.75, lon0=-179.75)
pick = np.sort(np.random.default_rng(2021).choice(glat.size, size=args.cells, replace=False))
lat, lon = glat[pick], glon[pick]
T, C = args.days, args.cells
days, gmt = synthetic.time_axis(T), synthetic.ssa_gmt(T)
idx = get_task_indices(C, rank, world)
counts = [len(get_task_indices(C, r, world)) for r in range(world)]
make = {"tas": synthetic.tas_cells, "pr": synthetic.pr_cells}
cdtype = torch.float32 if args.cfact_dtype == "float32" else torch.float64
os.makedirs(args.out, exist_ok=True)
if rank == 0:
    du = shutil.disk_usage(args.out)
    emit({"stage": "setup", "n_gpus": world, "cells": C, "days": T, "cells_per_rank": counts, "host_cpus": os.cpu_count(),
          "disk_free_gb": round(du.free / 1e9, 1), "cfact_dtype": args.cfact_dtype})

for variable in args.variables.split(","):
    # every rank generates and solves only its shard (what detrend_cells does with a full [T, C] array)
    t0 = time.perf_counter()
    n = len(idx)
    first = int(idx[0]) if n else 0
    if variable == "tas":   # generated on the device (the NumPy generator needs ~2 ms per cell), moved to pinned memory
        y_dev = synthetic.tas_tile_torch(n, T, seed=1234 + first, device=f"cuda:{local_rank}", nan_fraction=0.001)
        y_local = torch.empty((T, n), dtype=torch.float32, pin_memory=True)
        y_local.copy_(y_dev)
        del y_dev
        torch.cuda.empty_cache()
    else:                   # distinct cells from the host generator, tiled
        base = make[variable](min(n, 256), T, first_cell=first)
        reps = (n + base.shape[1] - 1) // base.shape[1]
        y_local = torch.from_numpy(np.tile(base, (1, reps))[:, :n].copy()).pin_memory()
    t_synth = time.perf_counter() - t0
    cfg = Config(variable, slab_cells=args.slab_cells, device=f"cuda:{local_rank}")
    solver = CellBatchSolver(variable, device=cfg.device)
    seeds = cell_seeds(cfg.seed, lat[idx], lon[idx]) if variable == "pr" else None
    P = solver.n_params
    host_kw = dict(cfact_dtype=cdtype, replaced="sparse", slab_major=True)
    # device-resident segment times of one slab (what the pipeline overlaps with the copies)
    w = min(cfg.slab_cells, n)
    yd = y_local[:, :w].to(cfg.device)
    sd = None if seeds is None else seeds[:w]
    seg = None
    for _ in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        solver.set_predictor(days, gmt, w)
        ev[0].record()
        ys, sc = solver.scale(yd, None, keep_scaled=variable != "tas")
        ev[1].record()
        fit = solver.fit(ys)
        ev[2].record()
        solver.quantile_map(yd, ys, fit.params, sc, sd)
        ev[3].record()
        torch.cuda.synchronize()
        seg = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
    del yd, ys, fit
    torch.cuda.empty_cache()
    # warm-up of the host pipeline: allocates the pinned outputs and the device scratch outside the timed region
    res = solver.detrend_host(y_local, days, gmt, seeds=seeds, slab_cells=cfg.slab_cells, **host_kw)
    a = wall()
    res = solver.detrend_host(y_local, days, gmt, seeds=seeds, slab_cells=cfg.slab_cells, out=res, **host_kw)
    b = wall()
    per_cell = {k: res[k].numpy() for k in ("params", "logp", "scaling", "status", "n_pass")}
    n_replaced_local = int(res["replaced_sparse"].shape[0])
    if world > 1:
        per_cell = gather_cells(per_cell, counts, dist, cfg.device)
    c = wall()
    path = os.path.join(args.out, f"{variable}_cfact_merged.nc")
    wrote = (not args.no_write) and variable in args.write_variables.split(",")
    if wrote:
        nw = args.write_days or T
        fdt = torch.float32 if args.file_dtype == "float32" else torch.float64
        series = {"cfact": aio.ColumnBlocks([s[:nw] for s in res["cfact_slabs"]], dtype=fdt),
                  "replaced": aio.SparseReplaced(res["replaced_sparse"].numpy(), nw, n)}
        kw = dict(scalars={"logp": per_cell["logp"]}, shared={"gmt_scaled": gmt[:nw]}, replaced_as_flags=False)
        if world > 1:
            aio.write_merged_output_sharded(path, series, counts, lat, lon, days[:nw], dist, group=host_group, **kw)
        else:
            aio.write_merged_output(path, series, lat, lon, days[:nw], **kw)
        if rank == 0:
            aio.write_merged_trace(os.path.join(args.out, f"{variable}_trace_merged.nc"), variable,
                                   per_cell["params"], per_cell["logp"], per_cell["scaling"], lat, lon)
    d = wall()
    # ---- spot checks on rank 0: N cells of rank 0's shard re-solved device-resident on one GPU, 2 by the CPU oracle
    check = None
    if rank == 0 and args.check > 0 and n > 0:
        from oracle import attrici_oracle as oracle

        sel = np.unique(np.linspace(0, n - 1, args.check).astype(int))
        ysel = y_local[:, sel].contiguous()
        s1 = CellBatchSolver(variable, device=cfg.device)
        one = s1.detrend_device(ysel.to(cfg.device), days, gmt, seeds=None if seeds is None else seeds[sel],
                                keep_scaled=variable != "tas")
        cf1 = one["cfact"].cpu().numpy()
        blocks = aio.ColumnBlocks(res["cfact_slabs"])
        cols = blocks.columns(sel).numpy().astype(np.float64)
        fin = np.isfinite(cf1) & np.isfinite(cols)
        rel = np.abs(cols[fin] - cf1[fin]) / np.maximum(np.abs(cf1[fin]), 1e-30)
        check = {"cells": sel.tolist(), "single_gpu_max_rel_cfact": float(rel.max()),
                 "single_gpu_max_abs_logp": float(np.abs(one["fit"].logp.cpu().numpy() - per_cell["logp"][idx[sel]]).max()),
                 "nan_pattern_equal": bool(np.array_equal(np.isnan(cf1), np.isnan(cols)))}
        orc = []
        for j in sel[:2]:
            kw = {} if seeds is None else {"seed": cfg.seed, "lat": float(lat[idx[j]]), "lon": float(lon[idx[j]])}
            ref = oracle.detrend_cell(variable, y_local[:, j].numpy(), gmt, days, 4, **kw)
            col = blocks.columns([j]).numpy()[:, 0].astype(np.float64)
            ok = np.isfinite(ref["cfact"]) & np.isfinite(col) & (np.abs(ref["cfact"]) > 0)
            orc.append({"cell": int(j), "logp_rel": float(abs(per_cell["logp"][idx[j]] - ref["logp"]) / abs(ref["logp"])),
                        "cfact_max_rel": float((np.abs(col[ok] - ref["cfact"][ok]) / np.abs(ref["cfact"][ok])).max()),
                        "zero_pattern_equal": bool(np.array_equal(col == 0, ref["cfact"] == 0))})
        check["oracle"] = orc
    if rank == 0:
        status = per_cell["status"]
        esize = 4 if cdtype == torch.float32 else 8
        emit({
            "variable": variable, "cells": C, "days": T, "n_gpus": world,
            "detrend_s": b - a, "cells_per_s": C / (b - a), "gather_s": c - b, "write_s": (d - c) if wrote else None,
            "file_bytes": os.path.getsize(path) if wrote else None,
            "write_gb_per_s": (os.path.getsize(path) / 1e9 / (d - c)) if wrote else None,
            "days_written": (args.write_days or T) if wrote else None, "file_cfact_dtype": args.file_dtype if wrote else None,
            "device": {"slab_cells": w, "scale_ms": seg[0], "fit_ms": seg[1], "quantile_map_ms": seg[2],
                       "device_cells_per_s_per_gpu": w / (sum(seg) / 1e3),
                       "h2d_gb_per_rank": 4.0 * T * counts[0] / 1e9, "d2h_gb_per_rank": esize * T * counts[0] / 1e9,
                       "pcie_gbs_per_rank_achieved": (4.0 + esize) * T * counts[0] / 1e9 / (b - a)},
            "outputs": f"{args.cfact_dtype} cfact, sparse replaced list ({n_replaced_local} entries on rank 0), slab-major",
            "converged_fraction": float((status == 0).mean()), "passes_mean": float(per_cell["n_pass"].mean()),
            "synthetic_input_s_rank0": t_synth, "check": check,
            "note": "detrend_s = H2D + scale + fit + quantile map + D2H, pipelined over cell slabs (not separable); "
                    "device = the kernels of one slab alone",
        })
        if wrote and not args.keep:
            os.remove(path)
    del res, y_local, solver
    torch.cuda.empty_cache()

if world > 1:
    dist.destroy_process_group()
