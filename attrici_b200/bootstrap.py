"""Block bootstrap of the fitted trend on the GPU (the bootstrap section of ``fit_and_detrend_cell``,
attrici/detrend.py:479-612, for a batch of cells).

Per cell the reference reshuffles the factual quantiles of the observations in blocks of one calendar year,
maps them back through the factual distribution of the target day, refits the model and collects the fitted
expectation; it reports mean / standard deviation / percentiles over ``bootstrap_sample_count`` such samples.
Here the samples of a batch of cells become extra cells of one wide series and are refitted by the same kernels
as the original fit (``CellBatchSolver``); ``libattrici_b200`` adds the year-block choices (NumPy's legacy
``np.random.randint`` stream per cell, seeded as detrend.py:285), the resampling and the statistics
(``csrc/bootstrap.cu``).  No CPU fallback.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .solver import CellBatchSolver
from .variables import VariableSpec, create_variable

QUANTILES = (0, 0.01, 0.025, 0.05, 0.25, 0.5, 0.75, 0.95, 0.975, 0.99, 1)   # detrend.py:575
MAX_SAMPLES = 256


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def year_blocks(years):
    """First day of each calendar-year block (+ T) and the block of each day (detrend.py:491-498).

    ``years``: calendar year of every day of the (gap-free, ordered) time axis, e.g. ``time.dt.year``."""
    years = np.asarray(years)
    if years.ndim != 1 or years.size == 0 or np.any(np.diff(years) < 0):
        raise ValueError("years must be a non-decreasing 1-D array (one entry per day)")
    change = np.flatnonzero(np.diff(years) != 0) + 1
    start = np.concatenate(([0], change, [years.size])).astype(np.int32)
    if start.size - 1 > 32767:
        raise ValueError("too many year blocks")
    block_of_day = (np.searchsorted(start, np.arange(years.size), side="right") - 1).astype(np.int16)
    return start, block_of_day


def bootstrap_cells(variable, y, days, gmt, years, seeds, n_samples, modes=4, device="cuda:0", max_columns=8192,
                    return_samples=False):
    """Bootstrap statistics of the fitted expectation for every cell of ``y`` [T, C] (float32 array / tensor).

    ``seeds``: per-cell NumPy legacy seeds (``attrici_b200.solver.cell_seeds``: detrend.py:285).
    Returns a dict of CUDA tensors shaped like the reference's bootstrap dataset (detrend.py:576-612):
    ``expected_values, bootstrap_mean, bootstrap_std`` [T, C], ``bootstrap_quantiles`` [11, T, C],
    ``bootstrap_replaced`` [T, C] (int32), plus ``quantile`` (the levels) and the original ``fit``.
    The whole series is the fit window, as the reference demands (detrend.py:480-486)."""
    spec = create_variable(variable) if isinstance(variable, str) else variable
    n_samples = int(n_samples)
    if not 1 <= n_samples <= MAX_SAMPLES:
        raise ValueError(f"n_samples must be in 1..{MAX_SAMPLES}")
    lib = _lib.load()
    dev = torch.device(device)
    yt = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
    yt = yt.to(device=dev, dtype=torch.float32).contiguous()
    T, Cn = yt.shape
    start_np, bod_np = year_blocks(years)
    if start_np[-1] != T:
        raise ValueError("years must have one entry per day of y")
    n_blocks = int(start_np.size - 1)
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    if seeds.shape != (Cn,):
        raise ValueError("seeds must have one entry per cell")

    # the original fit (detrend.py:326-370)
    solver = CellBatchSolver(spec, modes=modes, device=dev)
    solver.set_predictor(days, gmt, Cn)
    ys, scaling = solver.scale(yt)
    fit = solver.fit(ys)
    expected_scaled = solver.eval_parameter(fit.params, _lib.WHICH_EXPECTATION)   # Distribution.expectation()

    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        seeds_d = torch.as_tensor(seeds.view(np.int32)).to(dev)
        choices = torch.empty((Cn, n_samples * n_blocks), dtype=torch.int16, device=dev)
        # pr: the quantile map drew np.random.rand(T) from the cell's generator first (variables.py:458): 2 T words
        skip_words = 2 * T if spec.family == _lib.BERNOULLI_GAMMA else 0
        _lib.check(lib.attrici_bootstrap_block_choices(_ptr(seeds_d), Cn, n_samples * n_blocks, n_blocks, skip_words,
                                                       _ptr(choices), stream))
        scores = torch.empty((T, Cn), dtype=torch.float64, device=dev)
        ws = solver._ws
        _lib.check(lib.attrici_bootstrap_scores(_ptr(ws), ws.numel(), C.byref(solver.var), _ptr(ys), ys.stride(0), T, Cn,
                                                _ptr(fit.params), _ptr(scores), Cn, stream))
        start_d = torch.as_tensor(start_np).to(dev)
        bod_d = torch.as_tensor(bod_np).to(dev)
        replaced = torch.zeros((T, Cn), dtype=torch.int32, device=dev)
        E = torch.empty((T, n_samples * Cn), dtype=torch.float64, device=dev)

        # samples as extra cells: R samples of all cells per refit batch, column r * C + c
        R = max(1, min(n_samples, int(max_columns) // Cn))
        # the samples are already in scaled units: same family, no scaling, no masks
        sample_spec = VariableSpec("bootstrap-sample", spec.family, _lib.SCALE_NONE)
        folded = spec.family == _lib.NORMAL   # the Normal fit runs on per-phase sums and needs no scaled copy
        refit = {}
        y_new = torch.empty((T, R * Cn), dtype=torch.float32, device=dev)
        n_pass = []
        for r0 in range(0, n_samples, R):
            r = min(R, n_samples - r0)
            cols = r * Cn
            if cols not in refit:
                refit[cols] = CellBatchSolver(sample_spec, modes=modes, device=dev)
                refit[cols].set_predictor(days, gmt, cols)
            sv = refit[cols]
            yb = y_new[:, :cols]
            _lib.check(lib.attrici_bootstrap_resample(
                _ptr(ws), ws.numel(), C.byref(solver.var), _ptr(ys), ys.stride(0), T, Cn, _ptr(fit.params), _ptr(scores),
                Cn, _ptr(choices), n_samples, n_blocks, r0, r, _ptr(start_d), _ptr(bod_d), _ptr(yb), yb.stride(0),
                _ptr(replaced), Cn, stream))
            ysb, _ = sv.scale(yb, keep_scaled=not folded)
            fb = sv.fit(ysb)
            n_pass.append(fb.n_pass)
            sv.eval_parameter(fb.params, _lib.WHICH_EXPECTATION, out=E[:, r0 * Cn:r0 * Cn + cols])
        levels = torch.as_tensor(np.asarray(QUANTILES, dtype=np.float64)).to(dev)
        mean = torch.empty((T, Cn), dtype=torch.float64, device=dev)
        std = torch.empty((T, Cn), dtype=torch.float64, device=dev)
        quant = torch.empty((len(QUANTILES), T, Cn), dtype=torch.float64, device=dev)
        # expected_values = Variable.rescale(expectation of the original fit) (detrend.py:579-581): the statistics
        # kernel over a single "sample" (its mean is the rescaled value itself; std / quant are overwritten below)
        expected = torch.empty((T, Cn), dtype=torch.float64, device=dev)
        _lib.check(lib.attrici_bootstrap_stats(C.byref(solver.var), _ptr(expected_scaled), expected_scaled.stride(0), T, Cn,
                                               1, _ptr(scaling), _ptr(levels), len(QUANTILES), _ptr(expected), _ptr(std),
                                               _ptr(quant), stream))
        _lib.check(lib.attrici_bootstrap_stats(C.byref(solver.var), _ptr(E), E.stride(0), T, Cn, n_samples, _ptr(scaling),
                                               _ptr(levels), len(QUANTILES), _ptr(mean), _ptr(std), _ptr(quant), stream))
    out = {"expected_values": expected, "bootstrap_mean": mean, "bootstrap_std": std, "bootstrap_quantiles": quant,
           "bootstrap_replaced": replaced, "quantile": np.asarray(QUANTILES), "fit": fit, "scaling": scaling,
           "refit_passes": torch.cat(n_pass)}
    if return_samples:
        out["samples_scaled"] = E.view(T, n_samples, Cn)   # fitted mu of every sample, scaled units
        out["choices"] = choices.view(Cn, n_samples, n_blocks)
    return out
