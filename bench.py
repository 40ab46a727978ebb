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
    return {"workload": WORKLOADS["tas"],
            "cells_per_gpu": cells, "days": T_DAYS, "modes": MODES,
            "l2": "inputs larger than L2 (1.74 GB float32 series per tile; streamed twice per step: min/max + per-phase "
                  "sums in one pass, then the quantile map)",
            "report_variables": ["cfact", "replaced"]}


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
# measured FP64 FMA peak of this pool's B200 (scripts/micro/fp64_peak.cu / dmma_peak.cu, profiles/r2_dmma_peak.txt:
# 57.4 DFMA / clk / SM scalar, 63.6 through DMMA - the same pipe): the roofline of the fit kernel
FP64_PEAK_GFMAS = 16700.0
# FP64 operations of one (cell, phase pair) of one likelihood pass of the fused fit, counted on the mathematics
# (DESIGN.md section 5): predictors 27, two exp 2 x 14, residual sums 2 x 8, nll 2 x 3, gradient weights 2 x 6 + 6,
# gradient moments 27, u-space map 8
FIT_FP64_PER_PAIR = 124.0
# Normal-family variables: scale + sums in one pass, fused fit on the per-phase sums, y_scaled never materialised
NORMAL_VARIABLES = ("tas", "rsds")
WORKLOADS = {
    "tas": "tas Gaussian 100x100-cell tile, T=43464, modes=4, SSA-smoothed GMT predictor (BASELINE configs[1])",
    "pr": "pr Bernoulli-Gamma (wet/dry mask + gamma fit) on the same tile (BASELINE configs[2])",
    "sfcWind": "sfcWind Weibull on the same tile (BASELINE configs[3])",
    "hurs": "hurs Beta on the same tile (BASELINE configs[3])",
    "tasrange": "tasrange Gamma on the same tile",
    "rsds": "rsds Gaussian with the <= 0 rule on the same tile (BASELINE configs[3])",
}


def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch of a kernel from the committed ncu capture (profiles/r2_traffic.json, same tile)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        with open(p) as f:
            d = json.load(f)
        if d.get("cells") != TILE_CELLS or d.get("days") != T_DAYS:
            continue
        for k, v in d["kernels"].items():
            if k.startswith(kernel_prefix):
                return v["dram_bytes"], name
    return None, None


def family_tile(variable, C, T, device, seed):
    """Synthetic tile of a non-tas variable in HBM: distinct cells generated on the host (the generators are NumPy
    loops), tiled to C columns."""
    import torch

    from attrici_b200 import synthetic

    base = synthetic.GENERATORS[variable](min(C, 64), T, seed=seed)
    reps = (C + base.shape[1] - 1) // base.shape[1]
    return torch.from_numpy(np.tile(base, (1, reps))[:, :C].copy()).to(device)


def measure_device(solver, y, seeds, steps, warmup, keep_scaled, barrier, after_step=None):
    """W warm-up + K timed steps of scale -> fit -> quantile map with everything resident in HBM.
    Returns (ms total, per-segment ms per step, last fit)."""
    import torch

    T, C = y.shape
    cfact = torch.empty((T, C), dtype=torch.float64, device=y.device)

    def step(ev=None):
        if ev:
            ev[0].record()
        ys, scaling = solver.scale(y, None, keep_scaled=keep_scaled)
        if ev:
            ev[1].record()
        fit = solver.fit(ys)
        if ev:
            ev[2].record()
        solver.quantile_map(y, ys, fit.params, scaling, seeds, out=cfact)
        if ev:
            ev[3].record()
        if after_step:
            after_step(fit)
        return fit

    for _ in range(warmup):
        fit = step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seg_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
    barrier()
    t_begin = time.time()
    ev0.record()
    for k in range(steps):
        fit = step(seg_ev[k])
    ev1.record()
    barrier()
    t_end = time.time()
    ms = ev0.elapsed_time(ev1)
    seg_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in seg_ev) / steps for i in range(3)]
    del cfact
    return ms, seg_ms, fit, (t_begin, t_end)


def segment_rooflines(variable, seg_ms, n_pass_mean, T, C, peak_hbm):
    """Per-segment roofline statements.  Algorithmic bytes per cell-day: Normal family on the
    one-pass path - scale 4 B (the series once), quantile map 13 B (4 in, 8 + 1 out); families that keep y_scaled -
    scale 12 B (read twice, y_scaled written), every likelihood pass 4 B, quantile map 17 B.  The Normal fit runs on
    35 KB of per-phase sums per cell and is FP64-bound: FIT_FP64_PER_PAIR operations per (cell, pair, pass)."""
    normal = variable in NORMAL_VARIABLES
    cd = float(T) * C
    out = {}
    sb = (4.0 if normal else 12.0) * cd
    qb = (13.0 if normal else 17.0) * cd
    out["scale"] = {"ms": seg_ms[0], "bound": "hbm", "algorithmic_bytes": sb, "achieved": sb / 1e9 / (seg_ms[0] / 1e3),
                    "peak": peak_hbm, "unit": "GB/s"}
    out["quantile_map"] = {"ms": seg_ms[2], "bound": "hbm", "algorithmic_bytes": qb,
                           "achieved": qb / 1e9 / (seg_ms[2] / 1e3), "peak": peak_hbm, "unit": "GB/s"}
    if normal:
        ops = FIT_FP64_PER_PAIR * 731 * (n_pass_mean + 1.0) * C   # + the evaluation of the returned point
        out["fit"] = {"ms": seg_ms[1], "bound": "fp64", "algorithmic_fma": ops, "achieved": ops / 1e9 / (seg_ms[1] / 1e3),
                      "peak": FP64_PEAK_GFMAS, "unit": "GFMA/s (FP64)"}
    else:
        fb = 4.0 * cd * (n_pass_mean + 1.0)
        out["fit"] = {"ms": seg_ms[1], "bound": "hbm", "algorithmic_bytes": fb, "achieved": fb / 1e9 / (seg_ms[1] / 1e3),
                      "peak": peak_hbm, "unit": "GB/s"}
    for v in out.values():
        v["frac"] = v["achieved"] / v["peak"]
    return out


SEG_KERNEL = {
    ("tas", "scale"): "prep_onepass_kernel", ("tas", "fit"): "fit_fused_kernel", ("tas", "quantile_map"): "qmap_fold_bulk_kernel",
}
LIMITER = {
    "pr": "quantile map (qmap_fold_pr_kernel): FP64 series of the incomplete gamma function on the wet days, which are "
          "compacted so that all 32 lanes of a warp work (the day-ordered kernel had 9.8 of 32 active); 2 CTAs x 8 warps "
          "per SM at 128 registers, IPC 1.8, no DRAM traffic to speak of (profiles/r2_final_pr_ncu.txt); not memory-bound",
    "hurs": "quantile map: incomplete-beta inverses (Halley iterations of continued fractions), FP64 pipe ~30 % at 17 % "
            "occupancy (profiles/r2_qmap_pr_hurs_ncu.txt, day-ordered qmap_kernel); not memory-bound",
}


def run_workload(variable, device, C, T, days, gmt, steps, warmup, barrier):
    """Device-resident throughput of one of the other workloads (configs[2], [3]) on the same tile shape."""
    import torch

    from attrici_b200 import synthetic
    from attrici_b200.solver import CellBatchSolver, cell_seeds

    y = family_tile(variable, C, T, device, seed=4321)
    solver = CellBatchSolver(variable, modes=MODES, device=device)
    solver.set_predictor(days, gmt, C)
    seeds = None
    if variable == "pr":
        lat, lon = synthetic.tile_coordinates(100, (C + 99) // 100)
        seeds = cell_seeds(0, lat[:C], lon[:C])
    # the Normal family (rsds) takes the path of the tas line: sums in one pass, y_scaled never materialised
    normal = variable in NORMAL_VARIABLES
    ms, seg_ms, fit, _ = measure_device(solver, y, seeds, steps, warmup, keep_scaled=variable not in NORMAL_VARIABLES, barrier=barrier)
    n_pass = fit.n_pass.float().mean().item()
    peak, _ = measured_peaks()
    segs = segment_rooflines(variable, seg_ms, n_pass, T, C, peak)
    dom = max(segs, key=lambda k: segs[k]["ms"])
    res = {"workload": WORKLOADS[variable], "value": C * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
           "steps": steps, "warmup": warmup, "cells": C, "passes_per_cell_mean": n_pass,
           "converged_fraction": float((fit.status == 0).float().mean().item()),
           "segments": {k: {"ms": v["ms"], "frac": v["frac"], "bound": v["bound"]} for k, v in segs.items()},
           "roofline": dict(segs[dom], segment=dom)}
    if variable in LIMITER:
        res["limiter"] = LIMITER[variable]
    del y, solver
    torch.cuda.empty_cache()
    return res


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from attrici_b200 import _lib, synthetic
    from attrici_b200.solver import CellBatchSolver, cell_seeds

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

    variable = args.workload
    C, T = args.cells, T_DAYS
    gmt, days = synthetic.ssa_gmt(T), synthetic.time_axis(T)   # SSA-smoothed GMT (library kernels), configs[1]
    normal = variable == "tas"
    if normal:
        y = synthetic.tas_tile_torch(C, T, seed=1234 + rank, device=device, nan_fraction=0.001)
        seeds = None
    else:
        y = family_tile(variable, C, T, device, seed=4321 + rank)
        lat, lon = synthetic.tile_coordinates(100, (C + 99) // 100)
        seeds = cell_seeds(0, lat[:C], lon[:C]) if variable == "pr" else None
    solver = CellBatchSolver(variable, modes=MODES, device=device, profile=os.environ.get("ATTRICI_NO_FUSED") == "1")
    solver.set_predictor(days, gmt, C)
    P = solver.n_params
    gathered = torch.empty((world * C, P), dtype=torch.float64, device=device) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def after_step(fit):
        if world > 1:  # the path's one collective: final gather of the per-cell coefficients
            dist.all_gather_into_tensor(gathered, fit.params)

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = [0]

    def counted_barrier():
        barrier()
        launches0.append(_lib.launch_count())

    # the scaled series is an intermediate the Normal family does not materialise (report_variables = cfact, replaced)
    ms, seg_ms, fit, (t_begin, t_end) = measure_device(solver, y, seeds, args.steps, max(args.warmup, 3), keep_scaled=variable not in NORMAL_VARIABLES,
                                                       barrier=counted_barrier, after_step=after_step)
    sampler.t_begin, sampler.t_end = t_begin, t_end
    clocks = sampler.stop()
    # launches between the barrier before the timed steps and the one after them (the second and third barrier)
    launches = launches0[3] - launches0[2]
    n_pass = fit.n_pass.cpu().numpy()
    status = fit.status.cpu().numpy()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * C * args.steps / (ms / 1e3)

    # ---- end to end through the host API: pinned host buffers, copies inside the timed region ----
    del fit
    y_host = torch.empty((T, C), dtype=torch.float32, pin_memory=True)
    y_host.copy_(y)
    del y
    torch.cuda.empty_cache()
    solver_h = CellBatchSolver(variable, modes=MODES, device=device)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_run(**kw):
        res = solver_h.detrend_host(y_host, days, gmt, seeds=seeds, slab_cells=args.slab, **kw)  # warm-up, allocates pinned outputs
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = solver_h.detrend_host(y_host, days, gmt, seeds=seeds, slab_cells=args.slab, out=res, **kw)
            _ = float(res["logp"][0])  # result read on the host
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        d2h = sum(int(v.numel() * v.element_size()) for k, v in res.items()
                  if isinstance(v, torch.Tensor) and not k.startswith("_") and k != "replaced_sparse")
        if res.get("replaced_sparse") is not None:
            d2h += int(res["replaced_sparse"].numel() * 4)
        return world * C * e2e_steps / dt, d2h

    h2d = y_host.numel() * 4 + T * 12 + (4 * C if seeds is not None else 0)
    # the reference's outputs: float64 counterfactual + the replacement flags (as the list of their non-zero entries),
    # slabs of cells contiguous in host memory
    e2e_value, d2h = e2e_run(replaced="sparse", slab_major=True)
    variants = {}
    if rank == 0 or world > 1:
        v, b = e2e_run(cfact_dtype=torch.float32, replaced="sparse", slab_major=True)
        variants["float32_cfact"] = {"value": v, "unit": UNIT, "d2h_bytes_per_step": b,
                                     "note": "opt-in: the FP64 result rounded once to the dtype of the input files"}
        v, b = e2e_run(replaced="dense", slab_major=False)
        variants["dense_time_major"] = {"value": v, "unit": UNIT, "d2h_bytes_per_step": b,
                                        "note": "round-1 layout: [T x C] float64 + uint8 plane, pitched copies"}

    if rank == 0:
        peak, peak_src = measured_peaks()
        segs = segment_rooflines(variable, seg_ms, float(n_pass.mean()), T, C, peak)
        step_ms = ms / args.steps
        for v in segs.values():
            v["share_of_step"] = v["ms"] / step_ms
        dom = max(segs, key=lambda k: segs[k]["ms"])
        roof = dict(segs[dom])
        roof["segment"] = dom
        roof["kernel"] = SEG_KERNEL.get((variable, dom), dom)
        roof["launches"] = args.steps
        roof["avg_launch_ms"] = segs[dom]["ms"]
        roof["peak_source"] = (peak_src if roof["bound"] == "hbm" else
                               "measured FP64 FMA rate of this pool's B200 (scripts/micro/fp64_peak.cu, profiles/r2_dmma_peak.txt)")
        traffic, traffic_src = (ncu_traffic(roof["kernel"]) if C == TILE_CELLS and normal else (None, None))
        roof["traffic"] = traffic
        roof["traffic_unit"] = f"bytes per launch (ncu dram__bytes_read + write, profiles/{traffic_src})" if traffic_src else None
        # the step as a whole against the HBM roofline: bytes that must move (series in once, counterfactual + flags out)
        must = 13.0 * T * C
        moved = sum(v.get("algorithmic_bytes", 0.0) for v in segs.values()) + (24.0 * 1464 * C * 2 if normal else 0.0)
        whole = {"must_move_bytes": must, "frac_of_hbm_peak_must_move": must / 1e9 / (step_ms / 1e3) / peak,
                 "moved_bytes_model": moved, "frac_of_hbm_peak_moved": moved / 1e9 / (step_ms / 1e3) / peak}
        cpu = None
        if not args.no_cpu_baseline and normal:
            cores = min(host_cores(), 64)
            with cpu_pool(cores) as pool:
                wall, r = cpu_sample(pool, cores)
            kind = cpu_kind()
            cpu = {"value": cores / wall, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"{cores} cells of the same tile, one process per core; {CPU_KIND_TEXT[kind]}; "
                             f"{int(np.mean([x[1] for x in r]))} objective evaluations per cell on average"}
        workloads = None
        if world == 1 and normal and not args.no_workloads:
            del solver
            torch.cuda.empty_cache()
            workloads = {}
            for wv in ("pr", "sfcWind", "hurs", "rsds"):
                workloads[wv] = run_workload(wv, device, C, T, days, gmt, steps=3, warmup=3, barrier=barrier)
        cfg = bench_config(C)   # identical for both arms; per-arm measurements live outside of it
        cfg["workload"] = WORKLOADS[variable]
        fit_stats = {"passes_per_cell_mean": float(n_pass.mean()), "passes_per_cell_max": int(n_pass.max()),
                     "converged_fraction": float((status == 0).mean())}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 (sums, likelihood, solve, quantile map; f32 bit-exact scaling, TF32 tensor-core Fisher moments)",
            "data": "synthetic",
            "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "slab_cells": args.slab, "cpus_bound_per_rank": bound_cpus,
                    "outputs": "float64 cfact, replaced as sparse (day, cell, code) list, slab-major host layout",
                    "variants": variants},
            "gpu_launches": int(launches),
            "roofline": roof,
            "segments": segs,
            "whole_step": whole,
            "fit_stats": fit_stats,
            "workloads": workloads,
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
    ap.add_argument("--workload", default="tas", choices=sorted(WORKLOADS),
                    help="variable of the tile (default tas = BASELINE configs[1]; the default single-GPU tas run also "
                         "reports pr / sfcWind / hurs / rsds under \"workloads\")")
    ap.add_argument("--no-workloads", action="store_true", help="skip the other workloads in the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
