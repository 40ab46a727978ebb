"""Rolling-window model (``--window-size``) of the Normal family: host side.

The reference's second model family (attrici/estimation/model_pymc5.py:366-547, data gathered by
``collect_windows``, attrici/util.py:107-172; ``modes`` XOR ``window_size``, attrici/detrend.py:646-649): every day of
the year 1..366 has its own intercept (and, for the GMT-dependent ``mu``, trend) with N(0, 1) priors, fitted on the
``window_size`` days around every day of the fit series with that day of the year.  The integers - day of the year of
every day, the calendar-year runs of the fit series - are computed here with NumPy datetime64 arithmetic (bit-equal to
pandas ``time.dt.dayofyear``, tested) and handed to the kernels of ``csrc/window.cu``; the device never derives a date.

Parity: the reference implements this model for PyMC5 only (``ModelScipy`` raises ``NotImplementedError``), has no test
or fixture for it and PyMC is not in this image - the kernels are checked against ``oracle/window_oracle.py``, a
restatement of ``collect_windows`` and of the PyMC5 log posterior ("parity unpinned").
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from .solver import CellBatchSolver, _ptr

N_DOY = 366
TRACE_KEYS = ("weights_mu_longterm_intercept", "weights_mu_longterm_trend", "weights_sigma_longterm_intercept")


def dayofyear(time):
    """``time.dt.dayofyear`` (1..366) of a datetime64 axis as int16, from integer day / year arithmetic."""
    d = np.asarray(time).astype("datetime64[D]")
    return ((d - d.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.int64) + 1).astype(np.int16)


def window_runs(doy):
    """The fit series as runs of consecutive days whose day of the year increases by one (calendar years, or the pieces
    of them a series with gaps or a mid-year start has): arrays (start index, first day of the year, length), int32.
    With them day k of the year has the centres ``start + (k - first)`` of the runs that contain k, which is
    ``data.time.groupby("time.dayofyear").groups[k]`` (util.py:126), and day 366 the day after each day 365 (:132)."""
    doy = np.asarray(doy, dtype=np.int64)
    if doy.ndim != 1 or doy.size == 0:
        raise ValueError("doy must be a non-empty 1-D array")
    brk = np.flatnonzero(np.diff(doy) != 1) + 1
    start = np.concatenate([[0], brk])
    length = np.diff(np.concatenate([start, [doy.size]]))
    return start.astype(np.int32), doy[start].astype(np.int32), length.astype(np.int32)


class WindowBatchSolver:
    """Scale -> rolling-window fit -> quantile map for a batch of cells of one Normal-family variable."""

    def __init__(self, variable, window_size, device="cuda:0"):
        if window_size is None or int(window_size) % 2 == 0:
            raise ValueError("Window size must be an odd number")                      # detrend.py:648-649
        self.window_size = int(window_size)
        self.base = CellBatchSolver(variable, modes=1, device=device)
        if self.base.spec.family != _lib.NORMAL:
            raise NotImplementedError("Rolling window is implemented for the Normal family (tas, ps, rlds, rsds, tasskew)")
        self.device = self.base.device
        self.lib = self.base.lib

    def fit(self, y_scaled, time, gmt, fit_window=None):
        """MAP weights of the 366 days of the year per cell.  ``y_scaled`` [T, C] float32 CUDA, ``time`` datetime64 [T],
        ``gmt`` [T].  Returns (params [C, 3, 366] in the order of ``TRACE_KEYS``, logp [C], status [C], n_iter [C])."""
        T, C_ = y_scaled.shape
        i0, i1 = (0, T) if fit_window is None else fit_window
        doy = dayofyear(time)
        start, first, length = window_runs(doy[i0:i1])
        dev = self.device
        g = torch.as_tensor(np.ascontiguousarray(gmt, dtype=np.float64)).to(dev)
        rs, rf, rl = (torch.as_tensor(a).to(dev) for a in (start, first, length))
        nbytes = C.c_size_t(0)
        _lib.check(self.lib.attrici_window_workspace_bytes(C_, C.byref(nbytes)))
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        params = torch.empty((C_, 3, N_DOY), dtype=torch.float64, device=dev)
        logp = torch.empty(C_, dtype=torch.float64, device=dev)
        status = torch.empty(C_, dtype=torch.int32, device=dev)
        n_iter = torch.empty(C_, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.attrici_window_fit(_ptr(ws), ws.numel(), C.byref(self.base.var), _ptr(y_scaled),
                                                   y_scaled.stride(0), T, C_, i0, i1, _ptr(g), _ptr(rs), _ptr(rf), _ptr(rl),
                                                   int(start.size), self.window_size, _ptr(params), _ptr(logp), _ptr(status),
                                                   _ptr(n_iter), self.base._stream()))
            torch.cuda.current_stream(dev).synchronize()   # ws is released on return
        return params, logp, status, n_iter

    def quantile_map(self, y, y_scaled, time, gmt, params, scaling, want_scaled=False):
        """``estimate_distribution`` by day of the year (model_pymc5.py:435-447, 524-538) for the factual and the
        counterfactual (gmt = 0) predictor, quantile map, rescale and replacement (as ``CellBatchSolver.quantile_map``)."""
        T, C_ = y.shape
        dev = self.device
        doy = torch.as_tensor(dayofyear(time)).to(dev)
        g = torch.as_tensor(np.ascontiguousarray(gmt, dtype=np.float64)).to(dev)
        cfact = torch.empty((T, C_), dtype=torch.float64, device=dev)
        replaced = torch.empty((T, C_), dtype=torch.uint8, device=dev)
        cs = torch.empty((T, C_), dtype=torch.float64, device=dev) if want_scaled else None
        with torch.cuda.device(dev):
            _lib.check(self.lib.attrici_window_quantile_map(C.byref(self.base.var), _ptr(y), _ptr(y_scaled), y.stride(0), T, C_,
                                                            _ptr(g), _ptr(doy), _ptr(params), _ptr(scaling), _ptr(cfact),
                                                            _ptr(replaced), _ptr(cs), cfact.stride(0), self.base._stream()))
            torch.cuda.current_stream(dev).synchronize()
        return cfact, replaced, cs

    def detrend_device(self, y, time, gmt, fit_window=None):
        """Whole path for a [T, C] float32 CUDA batch: the Variable class' scaling (``CellBatchSolver.scale``), the
        rolling-window fit, the quantile map."""
        T, C_ = y.shape
        days = (np.asarray(time).astype("datetime64[D]") - np.asarray(time).astype("datetime64[D]")[0]).astype(np.int32)
        self.base.set_predictor(days, gmt, C_)
        ys, scaling = self.base.scale(y, fit_window, keep_scaled=True)
        params, logp, status, n_iter = self.fit(ys, time, gmt, fit_window)
        cfact, replaced, _ = self.quantile_map(y, ys, time, gmt, params, scaling)
        return {"y_scaled": ys, "scaling": scaling, "params": params, "logp": logp, "status": status, "n_iter": n_iter,
                "cfact": cfact, "replaced": replaced}


def trace_of(params_cell, logp):
    """One cell's result as the reference's PyMC5 trace dict (``weights_*`` entries of length 366 and ``logp``)."""
    p = np.asarray(params_cell)
    out = {k: p[i].copy() for i, k in enumerate(TRACE_KEYS)}
    out["logp"] = float(logp)
    return out
