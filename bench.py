#!/usr/bin/env python
"""Headline benchmark: gridcell-fits/sec (43 464-day series, modes=4) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # the CUDA path (this repo)
    python bench.py --impl reference [...]                       # the reference's CPU algorithm

A "step" is one pass of the hot path (scale -> MAP fit -> factual/counterfactual parameters -> quantile
map -> rescale/replacement) over one batch: the 100x100-cell tas tile of BASELINE.json configs[1]
(C = 10 000 cells, T = 43 464 days, synthetic, float32).  With N > 1 every rank (one process per GPU, torchrun)
detrends its own tile of the same size - weak scaling - and the step ends with the path's one collective,
an all-gather of the per-cell coefficients.

``value``  : cells / second with the tile resident in HBM (device-timed with CUDA events, max over ranks).
``e2e``    : the same through the reference-facing host API (``CellBatchSolver.detrend_host`` ->
             ``attrici_detrend_host``): float32 series in pinned host memory, float64 counterfactuals and
             flags back in pinned host memory, copies inside the timed region.
``roofline``: the dominant segment of the step (scaling / quantile map: HBM streaming kernels), achieved =
             its algorithmic bytes / its CUDA-event time inside the timed steps; ``segments`` has all three.
``cpu_baseline`` / ``--impl reference``: the oracle's port of the reference algorithm (SciPy L-BFGS-B with
             finite-difference gradients on the reference objective, scipy.stats quantile map) on the
             host cores, one cell per process, on a bounded sample of the same tile.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_DAYS = 43464
TILE_CELLS = 10000
MODES = 4
METRIC = "gridcell-fits/sec (43k-day series, modes=4)"
UNIT = "cells/s"


def bench_config(cells=TILE_CELLS):
    """``config`` of the JSON line: identical for the CUDA arm and the reference arm (what was sampled / measured
    per arm is in ``cpu_baseline.sample``, ``e2e`` and ``fit_stats``)."""
    return {"workload": "tas Gaussian 100x100-cell tile, T=43464, modes=4, SSA-smoothed GMT predictor (BASELINE configs[1])",
            "cells_per_gpu": cells, "days": T_DAYS, "modes": MODES,
            "l2": "inputs larger than L2 (1.74 GB float32 series per tile; streamed twice per step: scale + sums, "
                  "quantile map)",
            "report_variables": ["cfact", "replaced"]}


def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/r1_traffic.json, same tile)."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    if d.get("cells") != TILE_CELLS or d.get("days") != T_DAYS:
        return None
    for name, v in d["kernels"].items():
        if name.startswith(kernel_prefix):
            return v["dram_bytes"]
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    The sampler is started before the warm-up steps (nvidia-smi needs a few hundred ms to deliver its first line);
    every line carries its timestamp and ``stop`` keeps those inside [t_begin, t_end] of the timed region.  If the
    region is shorter than the sampling period, the samples taken since the warm-up began are used and the result
    says so."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        self.t_start = time.time()

    def begin_region(self):
        self.t_begin = time.time()

    def end_region(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        import datetime

        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[2]), float(f[3]), [n for n, v in zip(names, f[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        t0 = self.t_begin if self.t_begin is not None else self.t_start
        t1 = self.t_end if self.t_end is not None else time.time()
        inside = [r for r in rows if t0 - 0.01 <= r[0] <= t1 + 0.01]
        note = None
        if not inside:
            inside = [r for r in rows if r[0] <= t1 + 0.01] or rows
            note = "timed region shorter than the sampling period: samples since the warm-up steps began"
        sm = [r[1] for r in inside]
        smax = [r[2] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        busy = [x for x in sm if x > 0.5 * max(smax)] or sm
        res = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": reasons, "samples": len(sm)}
        if note:
            res["note"] = note
        return res


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port), one cell per process
# ------------------------------------------------------------------------------------------------
_CPU_GMT = {}


def cpu_kind():
    """"reference": the unmodified reference from baseline/_ref (or /root/reference) through oracle/reference_runner;
    "port": the oracle's restatement of the same algorithm, when the reference is not installed."""
    from oracle import reference_runner

    return "reference" if reference_runner.available() else "port"


CPU_KIND_TEXT = {
    "reference": "the UNMODIFIED reference (baseline/_ref: create_variable -> ModelScipy.fit = SciPy L-BFGS-B with "
                 "finite differences -> estimate_distribution x2 -> quantile_mapping -> rescale), xarray replaced by a "
                 "1-D container stand-in",
    "port": "oracle port of the reference algorithm (SciPy L-BFGS-B + finite differences on the reference objective, "
            "scipy.stats quantile map); the reference itself is not installed here",
}


def _cpu_cell(args):
    cell, n_days = args
    from attrici_b200 import synthetic

    y = synthetic.tas_cells(1, n_days, seed=1234, first_cell=cell)[:, 0]
    if n_days not in _CPU_GMT:
        from oracle.ssa_oracle import calc_gmt_by_ssa

        _CPU_GMT[n_days] = synthetic.ssa_gmt(n_days, calc_gmt_by_ssa=calc_gmt_by_ssa)
    gmt, days = _CPU_GMT[n_days], synthetic.time_axis(n_days)
    t0 = time.perf_counter()
    if cpu_kind() == "reference":
        from oracle import reference_runner

        out = reference_runner.detrend_cell("tas", y, gmt, days, MODES)
        nfev = out["nfev"]
    else:
        from oracle import attrici_oracle as oracle

        out = oracle.detrend_cell("tas", y, gmt, days, MODES, solver="lbfgsb")
        nfev = out["trace"]["nfev"]
    return time.perf_counter() - t0, int(nfev)


def _cpu_init():
    from attrici_b200 import synthetic  # noqa: F401
    from oracle import attrici_oracle, reference_runner  # noqa: F401

    if reference_runner.available():
        reference_runner._modules()


def cpu_pool(cores):
    """Worker pool with one BLAS thread per process, as in the reference's HPC recipe (USAGE.md:128-174);
    the variables must be in the environment the workers are spawned with."""
    import multiprocessing as mp

    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"
    try:
        pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init)
        pool.map(abs, range(cores))  # workers up and imported before anything is timed
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return pool


def cpu_sample(pool, n_cells, first_cell=0):
    """Detrend ``n_cells`` synthetic tas cells with the reference algorithm; returns (wall s, per-cell results)."""
    t0 = time.perf_counter()
    res = pool.map(_cpu_cell, [(first_cell + i, T_DAYS) for i in range(n_cells)], chunksize=1)
    return time.perf_counter() - t0, res


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def reference_cells_per_step(steps, cores, rounds=16):
    """Cells per step of the reference arm: one per core, fewer once ``steps`` rounds of the pool would exceed
    ``rounds`` (~11 s each)."""
    budget_cells = rounds * cores
    return cores if steps * cores <= budget_cells else max(1, budget_cells // steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = min(host_cores(), 64)
    # One cell takes ~11 s on one core.  A step is one cell per core while the whole run stays within ~3 minutes
    # (16 rounds of the pool); for more steps than that the per-step sample shrinks, and the steps' samples are fed
    # to the pool back to back (no barrier between steps: a step smaller than the pool would leave cores idle), so
    # that the K steps are timed as one region and all cores stay busy.
    n_cells = reference_cells_per_step(args.steps, cores)
    kind = cpu_kind()
    with cpu_pool(cores) as pool:  # worker start-up and imports are the warm-up; the CPU path has no other
        total, _ = cpu_sample(pool, n_cells * args.steps)
    value = n_cells * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(args.cells),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{n_cells} cells of the tile per step ({n_cells * args.steps} in all, fed to the pool "
                                   f"back to back), one process per core; {CPU_KIND_TEXT[kind]}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist

    from attrici_b200 import _lib, synthetic
    from attrici_b200.solver import CellBatchSolver

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    from attrici_b200.detrend import bind_to_gpu_cpus

    bound_cpus = bind_to_gpu_cpus(local_rank)  # pinned host buffers on the GPU's NUMA node
    if world > 1:
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print its version on stdout - the contract is
        # one JSON line there; an explicit INFO / TRACE request of the caller is left alone
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # whatever level is asked for: not on stdout
        dist.init_process_group("nccl", device_id=torch.device(device))

    C, T = args.cells, T_DAYS
    y = synthetic.tas_tile_torch(C, T, seed=1234 + rank, device=device, nan_fraction=0.001)
    gmt, days = synthetic.ssa_gmt(T), synthetic.time_axis(T)   # SSA-smoothed GMT (library kernels), configs[1]
    solver = CellBatchSolver("tas", modes=MODES, device=device, profile=os.environ.get("ATTRICI_NO_FUSED") == "1")
    solver.set_predictor(days, gmt, C)
    P = solver.n_params
    gathered = torch.empty((world * C, P), dtype=torch.float64, device=device) if world > 1 else None

    # pre-allocated outputs: nothing but the library's kernels runs between the events of a step
    cfact = torch.empty((T, C), dtype=torch.float64, device=device)
    seg_names = ("scale", "fit", "quantile_map")

    def step(ev=None):
        """scale -> fit -> quantile map, outputs left in HBM.  The scaled series is an intermediate the
        Normal family does not materialise (report_variables = cfact, replaced)."""
        if ev:
            ev[0].record()
        ys, scaling = solver.scale(y, None, keep_scaled=False)
        if ev:
            ev[1].record()
        fit = solver.fit(ys)
        if ev:
            ev[2].record()
        _, replaced, _ = solver.quantile_map(y, ys, fit.params, scaling, out=cfact)
        if ev:
            ev[3].record()
        if world > 1:  # the path's one collective: final gather of the per-cell coefficients
            dist.all_gather_into_tensor(gathered, fit.params)
        return fit

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        fit = step()
    barrier()
    n_pass = fit.n_pass.cpu().numpy()
    status = fit.status.cpu().numpy()
    _lib.fit_profile(reset=True)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    barrier()
    sampler.begin_region()
    ev0.record()
    for k in range(args.steps):
        step(seg_ev[k])
    ev1.record()
    barrier()
    sampler.end_region()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    seg_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in seg_ev) / args.steps for i in range(3)]
    launches = _lib.launch_count() - launches0
    pass_ms, pass_launches, cell_days = _lib.fit_profile(reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * C * args.steps / (ms / 1e3)

    # end to end through the host API: pinned host buffers, copies inside the timed region
    del cfact, fit
    y_host = torch.empty((T, C), dtype=torch.float32, pin_memory=True)
    y_host.copy_(y)
    del y
    torch.cuda.empty_cache()
    solver_h = CellBatchSolver("tas", modes=MODES, device=device)
    res = solver_h.detrend_host(y_host, days, gmt, slab_cells=args.slab)  # warm-up, allocates pinned outputs
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = solver_h.detrend_host(y_host, days, gmt, slab_cells=args.slab, out=res)
        _ = float(res["logp"][0])  # result read on the host
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * C * e2e_steps / e2e_s
    h2d = y_host.numel() * 4 + T * 12
    d2h = sum(int(v.numel() * v.element_size()) for k, v in res.items()
              if isinstance(v, torch.Tensor) and not k.startswith("_") and k != "replaced_sparse")
    if res.get("replaced_sparse") is not None:
        d2h += int(res["replaced_sparse"].numel() * 4)

    if rank == 0:
        peak, peak_src = measured_peaks()
        # algorithmic bytes per cell-day of each segment (DESIGN.md "Kernels and rooflines"):
        #   scale: the series is read twice (min/max, then scaling + per-phase sums)          8 B
        #   quantile map: float32 observation in, float64 counterfactual + uint8 flag out     13 B
        #   fit: per pass 3 float64 sums per phase instead of the series (1461 x 24 B per cell and pass)
        seg_bytes = {"scale": 8.0 * T * C, "quantile_map": 13.0 * T * C,
                     "fit": 24.0 * min(T, 1461) * C * float(n_pass.mean() + 1)}
        segments = {n: {"ms": seg_ms[i], "share_of_step": seg_ms[i] / (ms / args.steps),
                        "algorithmic_gbs": seg_bytes[n] / 1e9 / (seg_ms[i] / 1e3)}
                    for i, n in enumerate(seg_names)}
        dom = max(("scale", "quantile_map"), key=lambda n: segments[n]["ms"])
        dom_kernel = {"scale": "prep_minmax_kernel + prep_scale_fold_kernel (2 launches)",
                      "quantile_map": "qmap_fold_lean_kernel"}[dom]
        achieved = segments[dom]["algorithmic_gbs"]
        traffic = None
        if C == TILE_CELLS:
            traffic = (ncu_traffic("qmap_fold_lean_kernel") if dom == "quantile_map" else
                       sum(filter(None, [ncu_traffic("prep_minmax_kernel"), ncu_traffic("prep_scale_fold_kernel")])) or None)
        cpu = None
        if not args.no_cpu_baseline:
            cores = min(host_cores(), 64)
            with cpu_pool(cores) as pool:
                wall, r = cpu_sample(pool, cores)
            kind = cpu_kind()
            cpu = {"value": cores / wall, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{cores} cells of the same tile, one process per core; {CPU_KIND_TEXT[kind]}; "
                             f"{int(np.mean([x[1] for x in r]))} objective evaluations per cell on average"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 scaling (bit-exact), f64 sums/likelihood/solve/quantile map",
            "data": "synthetic",
            "config": {"workload": "tas Gaussian 100x100-cell tile, T=43464, modes=4 (BASELINE configs[1])",
                       "cells_per_gpu": C, "days": T, "modes": MODES,
                       "l2": "inputs larger than L2 (1.74 GB float32 series per tile, streamed three times per step)",
                       "report_variables": ["cfact", "replaced"],
                       "passes_per_cell_mean": float(n_pass.mean()), "passes_per_cell_max": int(n_pass.max()),
                       "converged_fraction": float((status == 0).mean())},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "slab_cells": args.slab, "cpus_bound_per_rank": bound_cpus},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dom_kernel, "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read + write, profiles/r1_traffic.json)",
                         "algorithmic_bytes": seg_bytes[dom], "peak_source": peak_src,
                         "launches": args.steps, "avg_launch_ms": segments[dom]["ms"],
                         "share_of_step": segments[dom]["share_of_step"]},
            "segments": segments,
            "fit_passes": {"launches": int(pass_launches), "avg_launch_ms": pass_ms / max(pass_launches, 1),
                           "share_of_step": pass_ms / ms},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cells", type=int, default=TILE_CELLS, help="cells per GPU (default: the 100x100 tile)")
    ap.add_argument("--slab", type=int, default=2048, help="cells per device slab of the host pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
