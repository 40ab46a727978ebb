"""Batched driver: what replaces the per-cell loop of ``attrici.detrend.detrend`` for ``solver="cuda"``.

The reference walks over grid cells one at a time (attrici/detrend.py:760-788) and lets independent OS
processes take contiguous index ranges (``get_task_indices``, :119-169; USAGE.md:128-137).  Here the
cells of a rank's range are processed as one batch on its GPU, and ranks (one process per GPU) take the
ranges the same partition function assigns.  Nothing is exchanged during the fit; the only collective is
a final gather of per-cell results.
"""

from dataclasses import dataclass

import numpy as np

from . import _lib
from .variables import check_bounds, check_units, create_variable


@dataclass
class Config:
    """The fields of ``attrici.detrend.Config`` (attrici/detrend.py:46-116) the hot path reads, with the
    reference's names and defaults, plus the knobs of the batched solver."""

    variable: str
    modes: int = 4
    seed: int = 0
    task_id: int = 0
    task_count: int = 1
    solver: str = "cuda"
    slab_cells: int = 2048       # cells per device slab of the host pipeline
    use_fp64: bool = False       # FP64 day arithmetic in the likelihood passes
    device: str = None           # default: cuda:<LOCAL_RANK>
    validate: bool = False       # bounds checks of the input and of cfact on the host (see detrend_cells)
    fit_only: bool = False       # stop after the fit (attrici/detrend.py:353-354): no cfact / replaced
    # variables of the per-cell output dataset to return (attrici/detrend.py:447-467; the reference's default is
    # ("all",)): beyond cfact / replaced, "y_scaled", "cfact_scaled", the distribution fields ("mu", "sigma", ...)
    # and their counterfactual twins ("mu_cfact", ...) come back in res["series"]
    report_variables: tuple = ("cfact", "replaced")
    window_size: int = None      # rolling-window model (Normal family): detrend_cells(..., time=...) with datetime64 days


def get_task_indices(indices_len, task_id, task_count):
    """Cell indices of one task: same rule as attrici/detrend.py:119-169 (every task gets
    ``indices_len // task_count + 1`` cells until they run out, the first task that would overshoot gets
    the remainder, later tasks are empty; an exact division is split evenly)."""
    if task_id < 0 or task_id >= task_count:
        raise ValueError("Task ID must be between 0 and task count")
    if indices_len % task_count == 0:
        calls = np.full(task_count, indices_len // task_count, dtype=np.int64)
    else:
        calls = np.full(task_count, indices_len // task_count + 1, dtype=np.int64)
        over = np.flatnonzero(np.cumsum(calls) > indices_len)
        calls[over] = 0
        calls[over[0]] = indices_len - calls.sum()
    cum = np.cumsum(calls)
    start = 0 if task_id == 0 else int(cum[task_id - 1])
    end = int(cum[task_id]) - 1
    return np.arange(start, end + 1)


def scale_predictor(gmt, n_time=None, days=None, gmt_days=None, full_extrapolation=False):
    """Predictor on the observation axis, min-max scaled to [0, 1] (attrici/detrend.py:687-707).

    Default branch (:698-707): the GMT series is stretched over the normalised observation times
    ``(t - t.min()) / (t.max() - t.min())`` — whatever dates the GMT file carries — with ``np.interp``.  ``days``:
    integer day offsets of the observation axis (any origin, gaps allowed); ``n_time`` alone means the gap-free
    axis ``0 .. n_time-1``.
    ``full_extrapolation`` (:688-696): the GMT values sit on a subset of the observation dates (``gmt_days``, same
    origin as ``days``, e.g. every 10th day); linear interpolation in between and linear extrapolation of the first /
    last segment beyond its ends — the arithmetic of ``scipy.interpolate.interp1d(kind="linear",
    fill_value="extrapolate")`` that ``xarray.DataArray.interp`` runs on the dates as float64 nanoseconds since
    the first GMT date."""
    gmt = np.asarray(gmt, dtype=np.float64)
    if days is None:
        if n_time is None:
            raise ValueError("give n_time or days")
        days = np.arange(n_time, dtype=np.int64)
    days = np.asarray(days)
    if not np.issubdtype(days.dtype, np.integer):
        raise TypeError("days must be integer day offsets")
    days = days.astype(np.int64)
    if n_time is not None and len(days) != n_time:
        raise ValueError("days must have n_time entries")
    if full_extrapolation:
        if gmt_days is None:
            raise ValueError("full_extrapolation needs the day offsets of the GMT values (gmt_days)")
        gd = np.asarray(gmt_days).astype(np.int64)
        if gd.shape != gmt.shape or gd.size < 2 or (np.diff(gd) <= 0).any():
            raise ValueError("gmt_days must be increasing, one entry per GMT value, at least two")
        ns = 86400.0e9                                   # xarray works on nanoseconds since the first GMT date
        x = (gd - gd[0]).astype(np.float64) * ns
        x_new = (days - gd[0]).astype(np.float64) * ns
        hi = np.clip(np.searchsorted(x, x_new), 1, len(x) - 1)
        lo = hi - 1
        span = x[hi] - x[lo]                             # SciPy 1.18 `_call_linear`: two weights, no slope
        g = ((x_new - x[lo]) / span) * gmt[hi] + ((x[hi] - x_new) / span) * gmt[lo]
    else:
        # timedelta64 / timedelta64 in the reference: a float64 quotient of integer tick counts, which for whole days
        # equals the quotient of the day counts
        t_scaled = (days - days.min()).astype(np.float64) / np.float64(days.max() - days.min())
        g = np.interp(t_scaled, np.linspace(0, 1, len(gmt)), gmt)
    return (g - g.min()) / (g.max() - g.min())


def stack_latlon(lat, lon):
    """Coordinates of the stacked cells, ``obs_data.stack(latlon=("lat", "lon"))`` (attrici/detrend.py:672): lat-major,
    lon-minor; cell ``i * len(lon) + j`` is ``(lat[i], lon[j])``, i.e. column ``i * len(lon) + j`` of the ``[T, C]``
    view of a ``(time, lat, lon)`` array."""
    la, lo = np.meshgrid(np.asarray(lat), np.asarray(lon), indexing="ij")
    return la.ravel(), lo.ravel()


def select_cells(lat, lon, cells=None, mask=None, mask_lat=None, mask_lon=None):
    """Indices of the stacked cells a run works on (attrici/detrend.py:674-685): ``cells`` - a list of
    ``(lat, lon)`` tuples, taken in the order given, ``KeyError`` if one is not a cell of the input
    (``obs_data.sel(latlon=[...])``) - and then ``mask`` - a ``(lat, lon)`` array on ``mask_lat`` x ``mask_lon``
    whose cells equal to 1 are kept, in the mask's stacked order (``mask.where(mask == 1).dropna("latlon")``; a
    masked-in cell that is not among the (selected) input cells is a ``KeyError`` as well).  ``lat`` / ``lon``: the
    coordinates of the stacked cells (``stack_latlon``).  Coordinates are compared exactly, as xarray's index does."""
    lat, lon = np.asarray(lat), np.asarray(lon)
    index = {(a, b): i for i, (a, b) in enumerate(zip(lat.tolist(), lon.tolist()))}
    chosen = np.arange(lat.size)
    if cells:
        try:
            chosen = np.array([index[(float(a), float(b))] for a, b in cells], dtype=np.int64)
        except KeyError as e:
            raise KeyError(f"Not all cells could be found in input data: {e.args[0]}") from None
    if mask is not None:
        mask = np.asarray(mask)
        mla, mlo = stack_latlon(mask_lat, mask_lon)
        if mask.shape != (np.asarray(mask_lat).size, np.asarray(mask_lon).size):
            raise ValueError("mask must be a (lat, lon) array on mask_lat x mask_lon")
        keep = np.flatnonzero(mask.ravel() == 1)
        available = {(lat[i].item(), lon[i].item()): i for i in chosen.tolist()}
        try:
            chosen = np.array([available[(mla[k].item(), mlo[k].item())] for k in keep], dtype=np.int64)
        except KeyError as e:
            raise KeyError(f"masked-in cell {e.args[0]} is not a cell of the input data") from None
    return chosen


def fit_window_indices(time, start_date=None, stop_date=None):
    """``[i0, i1)`` of the days the model is fitted on: ``subset_times = time[(time >= start) & (time <= stop)]``
    (attrici/detrend.py:709-722), both ends inclusive, ``None`` = first / last day of the series.  ``time``:
    increasing datetime64 axis; the dates anything ``np.datetime64`` accepts (``"1901-01-01"``, ``datetime.date``)."""
    t = np.asarray(time)
    if not np.issubdtype(t.dtype, np.datetime64):
        raise TypeError("time must be a datetime64 axis")
    start = t[0] if start_date is None else np.datetime64(start_date)
    stop = t[-1] if stop_date is None else np.datetime64(stop_date)
    keep = np.flatnonzero((t >= start) & (t <= stop))
    if keep.size == 0:
        raise ValueError(f"no day of the series lies in [{start}, {stop}]")
    if keep[-1] - keep[0] + 1 != keep.size:
        raise ValueError("time axis must be increasing")
    return int(keep[0]), int(keep[-1]) + 1


def bind_to_gpu_cpus(device_index):
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that the pinned
    host buffers allocated afterwards are first-touched on that node and the PCIe copies do not cross the
    socket interconnect.  One process per GPU makes this the natural placement; a no-op when NVML or the
    affinity call is unavailable or the mask is empty.  Returns the number of CPUs bound to (0 = unchanged)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                index = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:  # noqa: BLE001 - placement is an optimisation, never a failure
        pass
    return 0


def validate_bounds(spec, values, skip_inf=False):
    """``Variable.validate`` (attrici/variables.py:317-320 and the other classes): ``check_bounds`` with the
    variable's bounds, for a whole ``[T, C]`` batch."""
    check_bounds(values, spec.lower_bound, spec.upper_bound, skip_inf=skip_inf)


def _dist_info():
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_rank(), dist.get_world_size()
    except ImportError:  # pragma: no cover
        pass
    return None, 0, 1


def gather_cells(local, counts, dist=None, device=None):
    """Final gather of per-cell arrays to every rank: the one collective of the path.

    ``local``: dict name -> array whose first axis is this rank's cells;
    ``counts``: cells per rank.  Uses ``all_gather`` on tensors padded to the largest shard (shards are
    uneven under the reference's partition rule).  Meant for the per-cell results (coefficients, logp, scaling,
    status: MBs); a per-day series does not belong here - ``io.write_merged_output_sharded`` gathers those a block of
    days at a time - and an array above 256 MB per rank is refused."""
    import torch

    if dist is None:
        dist, _, world = _dist_info()
    if dist is None:
        return {k: np.asarray(v) for k, v in local.items()}
    world = dist.get_world_size()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device(device) if (on_gpu and device) else torch.device("cuda" if on_gpu else "cpu")
    cmax = int(max(counts))
    out = {}
    for name, arr in local.items():
        a = np.ascontiguousarray(arr)
        if a.nbytes > (1 << 28):
            raise ValueError(f"gather_cells is for per-cell results; {name!r} holds {a.nbytes >> 20} MB on this rank "
                             "(per-day series go through io.write_merged_output_sharded)")
        pad = np.zeros((cmax,) + a.shape[1:], dtype=a.dtype)
        pad[: a.shape[0]] = a
        t = torch.from_numpy(pad).to(dev)
        bucket = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bucket, t)
        out[name] = np.concatenate([b.cpu().numpy()[: int(n)] for b, n in zip(bucket, counts)], axis=0)
    return out


def report_series(spec, report_variables):
    """Names of the extra per-day series a ``report_variables`` list asks for (everything of the reference's output
    dataset, attrici/detrend.py:447-467, that is not an input, ``cfact``, ``replaced`` or ``logp``); unknown names
    raise ``KeyError`` like the reference's ``ds[config.report_variables]``."""
    fields = list(spec.parameters) + [f"{k}_cfact" for k in spec.parameters]
    known = ["y", "gmt_scaled", "y_scaled", "cfact_scaled", "cfact", "logp", "replaced"] + fields
    if "all" in report_variables:
        return ["y_scaled", "cfact_scaled"] + fields
    for name in report_variables:
        if name not in known:
            raise KeyError(name)
    return [k for k in ["y_scaled", "cfact_scaled"] + fields if k in report_variables]


def passthrough_series(report_variables, y_local, gmt_scaled):
    """The inputs a ``report_variables`` list names (``y``, ``gmt_scaled``: attrici/detrend.py:449-451) as they go into
    ``res["series"]`` next to the computed ones: ``y`` [T, n_local] (the caller's array, not a copy), ``gmt_scaled`` [T].
    ``logp`` is the per-cell ``res["logp"]`` whatever the list says."""
    out = {}
    want_all = "all" in report_variables
    if want_all or "y" in report_variables:
        out["y"] = y_local.numpy() if hasattr(y_local, "numpy") else np.asarray(y_local)
    if want_all or "gmt_scaled" in report_variables:
        out["gmt_scaled"] = np.asarray(gmt_scaled, dtype=np.float64)
    return out


def _slabs_on_device(solver, config, y_local, days, gmt, fit_window, seeds, trace_local, want_map, extras=()):
    """The two modes of the reference's cell function that are not the whole path: ``fit_only`` (scale and fit,
    attrici/detrend.py:353-354) and a given trace (no fit; the quantile map runs on the stored coefficients and
    rescales with the stored scaling constants, :319-324, :335) — and the runs that report more than ``cfact`` /
    ``replaced`` (``extras``: names from ``report_series``).  Slab by slab through the solver's device calls."""
    import torch

    T, n = y_local.shape
    P = solver.n_params
    pin = torch.cuda.is_available()
    res = {"cfact": torch.empty((T, n), dtype=torch.float64, pin_memory=pin) if want_map else None,
           "replaced": torch.empty((T, n), dtype=torch.uint8, pin_memory=pin) if want_map else None,
           "params": torch.empty((n, P), dtype=torch.float64), "logp": torch.empty(n, dtype=torch.float64),
           "scaling": torch.empty((n, 2), dtype=torch.float64), "status": torch.zeros(n, dtype=torch.int32),
           "n_pass": torch.zeros(n, dtype=torch.int32)}
    series = {k: torch.empty((T, n), dtype=torch.float32 if k == "y_scaled" else torch.float64, pin_memory=pin)
              for k in extras}
    slab = max(1, int(config.slab_cells))
    for a in range(0, n, slab):
        b = min(n, a + slab)
        yd = y_local[:, a:b].contiguous().to(solver.device, non_blocking=True)
        if not getattr(solver, "_basis_ready", False) or (solver.T, solver.C) != (T, b - a):
            solver.set_predictor(days, gmt, b - a)           # shared basis: once per slab shape
        ys, scaling = solver.scale(yd, fit_window)
        if trace_local is None:
            fit = solver.fit(ys)
            params, logp = fit.params, fit.logp
            res["status"][a:b], res["n_pass"][a:b] = fit.status.cpu(), fit.n_pass.cpu()
        else:
            params = torch.from_numpy(np.ascontiguousarray(trace_local["params"][a:b], dtype=np.float64)).to(solver.device)
            logp = torch.from_numpy(np.ascontiguousarray(trace_local["logp"][a:b], dtype=np.float64))
            if trace_local.get("scaling") is not None:   # variable.scaling taken from the trace (:319-324)
                scaling = torch.from_numpy(np.ascontiguousarray(trace_local["scaling"][a:b], dtype=np.float64)).to(solver.device)
        res["params"][a:b], res["logp"][a:b], res["scaling"][a:b] = params.cpu(), logp.cpu(), scaling.cpu()
        if want_map:
            cfact, replaced, cs = solver.quantile_map(yd, ys, params, scaling, None if seeds is None else seeds[a:b],
                                                      want_scaled="cfact_scaled" in series)
            res["cfact"][:, a:b].copy_(cfact)
            res["replaced"][:, a:b].copy_(replaced)
            if cs is not None:
                series["cfact_scaled"][:, a:b].copy_(cs)
        for k in series:
            if k == "y_scaled":
                series[k][:, a:b].copy_(ys)
            elif k != "cfact_scaled":                        # distribution fields: estimate_distribution x2, :362-370
                name, cf = (k[: -len("_cfact")], True) if k.endswith("_cfact") else (k, False)
                series[k][:, a:b].copy_(solver.eval_parameter(params, name, counterfactual=cf))
    if torch.device(solver.device).type == "cuda":
        torch.cuda.synchronize(solver.device)
    res["series"] = series
    return res


def _detrend_cells_window(config, y, gmt_scaled, time, lat, lon, fit_window, gather, units):
    """``detrend_cells`` for the rolling-window model (``config.window_size``; Normal family): slabs of cells through
    ``window.WindowBatchSolver`` with everything of a slab resident on the device.  ``params`` is [C, 3, 366] in the order
    of ``window.TRACE_KEYS``; ``time``: datetime64 days of the axis (the model is indexed by day of the year)."""
    import os

    import torch

    from .window import WindowBatchSolver

    if time is None or not np.issubdtype(np.asarray(time).dtype, np.datetime64):
        raise TypeError("the rolling-window model needs time= as datetime64 days")
    g = np.asarray(gmt_scaled, dtype=np.float64)
    spec = create_variable(config.variable)
    check_units(spec, units)
    y = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.asarray(y))
    T, C = y.shape
    dist, rank, world = _dist_info()
    task_id, task_count = (rank, world) if world > 1 else (config.task_id, config.task_count)
    idx = get_task_indices(C, task_id, task_count)
    device = config.device or f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    res = {"indices": idx}
    n_local = len(idx)
    keys = ("params", "logp", "status", "n_pass", "scaling")
    local = {"params": np.zeros((n_local, 3, 366)), "logp": np.zeros(n_local), "status": np.zeros(n_local, np.int32),
             "n_pass": np.zeros(n_local, np.int32), "scaling": np.zeros((n_local, 2))}
    cfact = np.empty((T, n_local), dtype=np.float64)
    replaced = np.empty((T, n_local), dtype=np.uint8)
    if n_local:
        solver = WindowBatchSolver(spec, config.window_size, device=device)
        c0 = int(idx[0])
        for a in range(0, n_local, config.slab_cells):
            b = min(n_local, a + config.slab_cells)
            out = solver.detrend_device(y[:, c0 + a:c0 + b].to(device), time, g, fit_window)
            local["params"][a:b] = out["params"].cpu().numpy()
            local["logp"][a:b] = out["logp"].cpu().numpy()
            local["status"][a:b] = out["status"].cpu().numpy()
            local["n_pass"][a:b] = out["n_iter"].cpu().numpy()
            local["scaling"][a:b] = out["scaling"].cpu().numpy()
            cfact[:, a:b] = out["cfact"].cpu().numpy()
            replaced[:, a:b] = out["replaced"].cpu().numpy()
    res["cfact"], res["replaced"] = cfact, replaced
    if gather and world > 1:
        counts = [len(get_task_indices(C, r, world)) for r in range(world)]
        flat = dict(local, params=local["params"].reshape(n_local, -1))
        got = gather_cells(flat, counts, dist, device)
        got["params"] = got["params"].reshape(-1, 3, 366)
        local = got
    res.update({k: local[k] for k in keys})
    return res


def detrend_cells(config, y, gmt_scaled, days, lat, lon, fit_window=None, gather=True, out=None, units=None,
                  trace=None, time=None):
    """Detrend the cells of ``y`` ([T, C] float32 host array, time-major, cells stacked lat-major as
    attrici/detrend.py:672 does) that the reference's partition assigns to this rank.

    Returns a dict: ``indices`` (this rank's cell indices), ``cfact`` [T, n_local] float64 and
    ``replaced`` [T, n_local] uint8 for those cells, and per-cell ``params`` [C', P], ``logp``, ``status``,
    ``n_pass``, ``scaling`` — gathered over all ranks when ``gather`` and torch.distributed is
    initialised (then C' = C), else local.

    ``units``: the ``units`` attribute of the input variable, checked like ``check_units``
    (attrici/variables.py:134-156).  ``config.validate`` adds the reference's bounds checks of the input
    (``<Variable>.__init__`` -> ``validate``, e.g. variables.py:317-320) and of ``cfact`` (detrend.py:411) for this
    rank's cells — host reductions over the whole batch, which cost more than the device work, hence opt-in.
    ``trace``: ``{"params": [C, P], "logp": [C], "scaling": [C, 2] (optional)}`` for all ``C`` cells, e.g. from
    ``io.read_merged_trace`` — the ``--trace-file`` mode (detrend.py:741-751): no fit, the quantile map runs on
    these coefficients.  ``config.fit_only`` stops after the fit (``cfact`` / ``replaced`` are None).
    ``fit_window``: ``[i0, i1)`` on the time axis (``fit_window_indices`` for start / stop dates).
    ``config.report_variables`` beyond ``cfact`` / ``replaced`` adds ``res["series"]``: name -> [T, n_local]
    (``y_scaled`` float32, ``cfact_scaled``, distribution fields and ``<field>_cfact`` float64; detrend.py:447-467;
    ``y`` as given and ``gmt_scaled`` [T] are passed through; ``logp`` is always ``res["logp"]``).
    """
    import os

    import torch

    from .solver import CellBatchSolver, cell_seeds

    if config.solver != "cuda":
        raise ValueError(f"attrici_b200 only provides solver='cuda', got {config.solver!r}")
    if (config.modes is None) == (config.window_size is None):
        raise ValueError("Exactly one of `modes` and `window_size` must be set")      # detrend.py:646-647
    if config.window_size is not None and config.window_size % 2 == 0:
        raise ValueError("Window size must be an odd number")                          # detrend.py:648-649
    if config.window_size is not None:
        return _detrend_cells_window(config, y, gmt_scaled, time, lat, lon, fit_window, gather, units)
    g = np.asarray(gmt_scaled, dtype=np.float64)
    if g.ndim != 1:
        raise ValueError("GMT data must have dimension 'time'")                         # detrend.py:654-655
    if np.any(np.isnan(g)) or np.any(np.isinf(g)):
        raise ValueError("GMT data must not contain NaN or infinite values")           # detrend.py:656-657
    spec = create_variable(config.variable)
    check_units(spec, units)
    extras = [] if config.fit_only else report_series(spec, config.report_variables)
    y = np.asarray(y) if not isinstance(y, torch.Tensor) else y
    if y.ndim != 2:
        raise ValueError("y must be [T, C]")
    T, C = y.shape
    lat, lon = np.asarray(lat), np.asarray(lon)
    if lat.shape != (C,) or lon.shape != (C,):
        raise ValueError("lat / lon must have one entry per cell")
    dist, rank, world = _dist_info()
    if world > 1 and config.task_count > 1:
        # ranks of one process group already are the tasks; a SLURM array on top needs one group per array task
        raise ValueError("task_id / task_count cannot be combined with an initialised torch.distributed process group "
                         f"(world size {world}): the ranks take the reference's task ranges themselves")
    task_id, task_count = (rank, world) if world > 1 else (config.task_id, config.task_count)
    idx = get_task_indices(C, task_id, task_count)
    device = config.device or f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    if world > 1:
        bind_to_gpu_cpus(torch.device(device).index or 0)
    n_local = len(idx)
    P = spec.n_params(config.modes)
    res = {"indices": idx}
    if n_local:
        solver = CellBatchSolver(spec, modes=config.modes, device=device, use_fp64=config.use_fp64)
        sl = slice(int(idx[0]), int(idx[-1]) + 1)  # contiguous by construction
        seeds = cell_seeds(config.seed, lat[sl], lon[sl]) if spec.family == _lib.BERNOULLI_GAMMA else None
        y_local = y[:, sl]
        if isinstance(y_local, np.ndarray):
            y_local = torch.from_numpy(y_local)
        if config.validate:
            validate_bounds(spec, y_local, skip_inf=True)
        if trace is not None or config.fit_only or extras:
            if trace is not None:
                tp = np.asarray(trace["params"])
                if tp.shape != (C, P):
                    raise ValueError(f"trace params must be [{C}, {P}], got {tp.shape}")
            t_local = None if trace is None else {
                "params": tp[sl], "logp": np.asarray(trace.get("logp", np.full(C, np.nan)))[sl],
                "scaling": None if trace.get("scaling") is None else np.asarray(trace["scaling"])[sl]}
            o = _slabs_on_device(solver, config, y_local, days, gmt_scaled, fit_window, seeds, t_local,
                                 want_map=not config.fit_only, extras=extras)
            res["series"] = {k: v.numpy() for k, v in o.pop("series").items()}
        else:
            o = solver.detrend_host(y_local, days, gmt_scaled, fit_window, seeds, config.slab_cells, out=out)
        local = {k: (o[k].numpy() if o[k] is not None else None) for k in o if not k.startswith("_")}
        passed = {} if config.fit_only else passthrough_series(config.report_variables, y_local, gmt_scaled)
        if passed:
            res.setdefault("series", {}).update(passed)
        if config.validate and local["cfact"] is not None:
            validate_bounds(spec, local["cfact"])
    else:
        local = {"cfact": None if config.fit_only else np.empty((T, 0)),
                 "replaced": None if config.fit_only else np.empty((T, 0), dtype=np.uint8),
                 "params": np.empty((0, P)), "logp": np.empty(0), "scaling": np.empty((0, 2)),
                 "status": np.empty(0, dtype=np.int32), "n_pass": np.empty(0, dtype=np.int32)}
        if extras:
            res["series"] = {k: np.empty((T, 0), dtype=np.float32 if k == "y_scaled" else np.float64) for k in extras}
    res["cfact"], res["replaced"] = local["cfact"], local["replaced"]
    per_cell = {k: local[k] for k in ("params", "logp", "scaling", "status", "n_pass")}
    if gather and world > 1:
        counts = [len(get_task_indices(C, r, world)) for r in range(world)]
        per_cell = gather_cells(per_cell, counts, dist, device)
    res.update(per_cell)
    return res
