"""Batched CUDA solver: the host side of the hot path.

``CellBatchSolver`` drives ``libattrici_b200.so`` for a batch of grid cells laid out time-major
(``[T, C]``, the layout ``obs_data.stack(latlon=("lat", "lon"))`` gives, attrici/detrend.py:672): it is
what replaces, for ``solver="cuda"``, the body of ``fit_and_detrend_cell`` between ``create_variable``
and ``estimate_logp`` (attrici/detrend.py:313-416) for all cells at once.  torch is used for device
memory and streams only; every computation is a kernel of the C-ABI library.  No CPU fallback.
"""

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .variables import create_variable


def cell_seeds(seed, lat, lon):
    """Per-cell NumPy legacy seed of attrici/detrend.py:285:
    ``(seed + int(lat * 1e9) + int(lon * 1e6)) % 2**32`` (Python ``int()`` truncates toward zero)."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    out = np.empty(lat.shape, dtype=np.uint32)
    flat_lat, flat_lon, flat_out = lat.ravel(), lon.ravel(), out.ravel()
    for i in range(flat_lat.size):
        flat_out[i] = (int(seed) + int(flat_lat[i] * 1e9) + int(flat_lon[i] * 1e6)) % 2**32
    return out


@dataclass
class FitResult:
    """Device-resident result of ``CellBatchSolver.fit`` (SciPy trace schema per cell:
    ``{"logp": float, "params": ndarray[P]}``, attrici/estimation/model_scipy.py:524-527)."""

    params: torch.Tensor   # [C, P] float64, reference theta layout
    logp: torch.Tensor     # [C] float64
    status: torch.Tensor   # [C] int32, _lib.STATUS_*
    n_pass: torch.Tensor   # [C] int32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class CellBatchSolver:
    """Scale -> fit -> quantile-map for a batch of cells of one variable on one GPU."""

    def __init__(self, variable, modes=4, device="cuda:0", use_fp64=False, zero_init=False, max_pass=60,
                 tol_pred=1e-11, profile=False, exact_scaled_sums=False):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.AttriciCudaError("attrici_b200 needs a CUDA device (no CPU fallback)")
        self.spec = create_variable(variable) if isinstance(variable, str) else variable
        self.modes = int(modes)
        self.var = self.spec.descriptor(self.modes)
        self.device = torch.device(device)
        self.n_params = self.spec.n_params(self.modes)
        self.options = _lib.default_fit_options()
        self.options.use_fp64 = int(use_fp64)
        self.options.zero_init = int(zero_init)
        self.options.max_pass = int(max_pass)
        self.options.tol_pred = float(tol_pred)
        self.options.profile = int(profile)
        # Normal family without a materialised y_scaled: True keeps the two-pass sums of the float32-rounded scaled
        # values the reference fits; the default gathers sums of the unrounded ones in one pass over the series
        # (logp differs by ~2e-9 relative; attrici_prep_scale_mask_ex)
        self.prep_flags = _lib.PREP_EXACT_SUMS if exact_scaled_sums else 0
        self._ws = None
        self._ws_key = None
        self.T = self.C = None

    # -- plumbing -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, T, C_):
        key = (T, C_)
        if self._ws_key != key:
            n = _lib.workspace_bytes(self.spec.family, T, C_)
            self._ws = None  # release before allocating the next one
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._ws_key = key
            self._basis_ready = False
        return self._ws

    def _check_series(self, y, name):
        if not (isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == torch.float32 and y.dim() == 2):
            raise TypeError(f"{name} must be a 2-D float32 CUDA tensor [T, C]")
        if y.shape[1] > 1 and y.stride(1) != 1:
            raise ValueError(f"{name} must be time-major with unit stride along cells")
        return y.shape[0], y.shape[1], y.stride(0)

    # -- steps ----------------------------------------------------------------------------------
    def set_predictor(self, days, gmt, n_cells):
        """Shared Fourier/GMT basis (attrici/util.py:85-104, model_scipy.py:180-197, 280-289).

        ``days``: integer day offsets of the time axis, ``gmt``: scaled predictor on that axis
        (attrici/detrend.py:698-707)."""
        days = torch.as_tensor(np.ascontiguousarray(days, dtype=np.int32)).to(self.device)
        gmt = torch.as_tensor(np.ascontiguousarray(gmt, dtype=np.float64)).to(self.device)
        if days.numel() != gmt.numel():
            raise ValueError("days and gmt must have the same length")
        self.T, self.C = int(days.numel()), int(n_cells)
        ws = self._workspace(self.T, self.C)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_build_basis(_ptr(ws), ws.numel(), self.spec.family, _ptr(days), _ptr(gmt),
                                                    self.T, self.C, self._stream()))
        self._days, self._gmt = days, gmt
        self._basis_ready = True

    def _need_basis(self, T, C_):
        if not getattr(self, "_basis_ready", False) or (self.T, self.C) != (T, C_):
            raise RuntimeError("call set_predictor(days, gmt, n_cells) for this batch shape first")

    def scale(self, y, fit_window=None, keep_scaled=True):
        """``create_variable`` arithmetic (variables.py:92-111, 363-403, 574-591, 619-623, 722-732, 770-780)
        and Inf -> NaN (detrend.py:311).  Returns (y_scaled [T, C] float32, scaling [C, 2] float64).

        ``keep_scaled=False`` (Normal family only) does not materialise ``y_scaled`` (returned as None): on a
        gap-free daily axis the fit runs on per-phase sums gathered here and the quantile map recomputes the
        scaled values bit-identically from ``y``."""
        T, C_, ld = self._check_series(y, "y")
        self._need_basis(T, C_)
        i0, i1 = (0, T) if fit_window is None else fit_window
        # the C ABI has one leading dimension for y and y_scaled: the scaled series takes the layout of its input
        ys = (torch.empty_strided((T, C_), (ld, 1), dtype=torch.float32, device=self.device)
              if keep_scaled else None)
        scaling = torch.empty((C_, 2), dtype=torch.float64, device=self.device)
        ws = self._ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_prep_scale_mask_ex(_ptr(ws), ws.numel(), C.byref(self.var), _ptr(y), _ptr(ys),
                                                           ld, T, C_, i0, i1, _ptr(scaling), self.prep_flags,
                                                           self._stream()))
        self._fit_window = (i0, i1)
        return ys, scaling

    def nll_grad(self, y_scaled, theta, fit_window=None, use_fp64=True, want_fisher=False):
        """Negative log posterior, gradient and Fisher information at ``theta`` [C, P] (known-answer
        entry; the reference objective is model_scipy.py:311-331 summed as in :520).  ``scale`` must have
        run on this batch (it records the per-cell phase origins)."""
        T, C_, ld = self._check_series(y_scaled, "y_scaled")
        self._need_basis(T, C_)
        i0, i1 = self._fit_window if fit_window is None else fit_window
        theta = torch.as_tensor(theta, dtype=torch.float64, device=self.device).contiguous()
        if theta.shape != (C_, self.n_params):
            raise ValueError(f"theta must be [{C_}, {self.n_params}]")
        nll = torch.empty(C_, dtype=torch.float64, device=self.device)
        grad = torch.empty((C_, self.n_params), dtype=torch.float64, device=self.device)
        fisher = (torch.empty((C_, self.n_params, self.n_params), dtype=torch.float64, device=self.device)
                  if want_fisher else None)
        ws = self._ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_nll_grad(_ptr(ws), ws.numel(), C.byref(self.var), _ptr(y_scaled), ld, T, C_,
                                                 i0, i1, int(use_fp64), _ptr(theta), _ptr(nll), _ptr(grad),
                                                 _ptr(fisher), self._stream()))
        return nll, grad, fisher

    def fit(self, y_scaled, fit_window=None):
        """Batched MAP fit (replaces ModelScipy.fit, model_scipy.py:505-527).  ``y_scaled=None``: fit from the
        per-phase sums of the last ``scale`` call (Normal family, gap-free daily axis)."""
        if y_scaled is None:
            T, C_, ld = self.T, self.C, self.C
        else:
            T, C_, ld = self._check_series(y_scaled, "y_scaled")
        self._need_basis(T, C_)
        i0, i1 = self._fit_window if fit_window is None else fit_window
        params = torch.empty((C_, self.n_params), dtype=torch.float64, device=self.device)
        logp = torch.empty(C_, dtype=torch.float64, device=self.device)
        status = torch.empty(C_, dtype=torch.int32, device=self.device)
        n_pass = torch.empty(C_, dtype=torch.int32, device=self.device)
        ws = self._ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_fit_batch(_ptr(ws), ws.numel(), C.byref(self.var), _ptr(y_scaled), ld, T, C_,
                                                  i0, i1, C.byref(self.options), _ptr(params), _ptr(logp),
                                                  _ptr(status), _ptr(n_pass), self._stream()))
        return FitResult(params, logp, status, n_pass)

    def quantile_map(self, y, y_scaled, params, scaling, seeds=None, want_replaced=True, want_scaled=False,
                     out=None):
        """``estimate_distribution`` x2 + ``quantile_mapping`` + ``rescale`` + replacement
        (model_scipy.py:547-570, variables.py:287-303 / 422-480, detrend.py:392-411).
        Returns (cfact [T, C] float64, replaced [T, C] uint8 codes or None, cfact_scaled or None)."""
        T, C_, ld = self._check_series(y, "y")
        if y_scaled is not None:
            T2, C2, ld2 = self._check_series(y_scaled, "y_scaled")
            if (T2, C2, ld2) != (T, C_, ld):
                raise ValueError("y and y_scaled must have the same shape and stride")
        self._need_basis(T, C_)
        cfact = out if out is not None else torch.empty((T, C_), dtype=torch.float64, device=self.device)
        replaced = torch.empty((T, C_), dtype=torch.uint8, device=self.device) if want_replaced else None
        if cfact.dtype == torch.float32:
            # opt-in output type (``out=`` a float32 tensor): the FP64 result rounded once; Normal variables without a
            # post rule on a gap-free daily axis (attrici_quantile_map_f32)
            if not want_replaced or want_scaled:
                raise ValueError("a float32 counterfactual comes with `replaced` and without cfact_scaled")
            ws = self._ws
            with torch.cuda.device(self.device):
                _lib.check(self.lib.attrici_quantile_map_f32(_ptr(ws), ws.numel(), C.byref(self.var), _ptr(y), _ptr(y_scaled),
                                                             ld, T, C_, _ptr(params), _ptr(scaling), _ptr(cfact),
                                                             _ptr(replaced), cfact.stride(0), self._stream()))
            return cfact, replaced, None
        cs = torch.empty((T, C_), dtype=torch.float64, device=self.device) if want_scaled else None
        if seeds is not None:
            seeds = torch.as_tensor(np.ascontiguousarray(seeds, dtype=np.uint32).view(np.int32)).to(self.device)
        ws = self._ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_quantile_map(_ptr(ws), ws.numel(), C.byref(self.var), _ptr(y), _ptr(y_scaled),
                                                     ld, T, C_, _ptr(params), _ptr(scaling), _ptr(seeds), _ptr(cfact),
                                                     _ptr(replaced), _ptr(cs), cfact.stride(0), self._stream()))
        return cfact, replaced, cs

    def eval_parameter(self, params, which, counterfactual=False, out=None):
        """Per-day distribution parameter (``report_variables``, detrend.py:462-467): ``which`` is a
        name of ``self.spec.parameters`` or its index.  ``out``: optional [T, C] float64 CUDA view (unit stride
        along cells) to write into."""
        if isinstance(which, str):
            which = _lib.WHICH_EXPECTATION if which == "expectation" else self.spec.parameters.index(which)
        self._need_basis(self.T, self.C)
        if out is None:
            out = torch.empty((self.T, self.C), dtype=torch.float64, device=self.device)
        elif (out.dtype != torch.float64 or not out.is_cuda or tuple(out.shape) != (self.T, self.C)
              or (self.C > 1 and out.stride(1) != 1)):
            raise ValueError("out must be a [T, C] float64 CUDA tensor with unit stride along cells")
        ws = self._ws
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_eval_parameter(_ptr(ws), ws.numel(), C.byref(self.var), self.T, self.C,
                                                       _ptr(params), int(which), int(bool(counterfactual)),
                                                       _ptr(out), out.stride(0), self._stream()))
        return out

    def mt19937_uniform(self, seeds, n):
        """``np.random.seed(s); np.random.rand(n)`` per cell, as [n, C] float64."""
        seeds = torch.as_tensor(np.ascontiguousarray(seeds, dtype=np.uint32).view(np.int32)).to(self.device)
        out = torch.empty((n, seeds.numel()), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.attrici_mt19937_uniform(_ptr(seeds), seeds.numel(), n, _ptr(out), seeds.numel(),
                                                        self._stream()))
        return out

    # -- whole path -----------------------------------------------------------------------------
    def detrend_device(self, y, days, gmt, fit_window=None, seeds=None, want_replaced=True, keep_scaled=True):
        """Whole hot path with inputs and outputs resident in HBM.  ``keep_scaled=False`` (Normal family on a
        gap-free daily axis) skips the ``y_scaled`` intermediate, see ``scale``."""
        T, C_, _ = self._check_series(y, "y")
        if not getattr(self, "_basis_ready", False) or (self.T, self.C) != (T, C_):
            self.set_predictor(days, gmt, C_)
        if not keep_scaled:
            d = np.asarray(days)
            if self.spec.family != _lib.NORMAL or not np.array_equal(d - d[0], np.arange(len(d))):
                keep_scaled = True
        ys, scaling = self.scale(y, fit_window, keep_scaled)
        fit = self.fit(ys)
        cfact, replaced, _ = self.quantile_map(y, ys, fit.params, scaling, seeds, want_replaced)
        return {"y_scaled": ys, "scaling": scaling, "fit": fit, "cfact": cfact, "replaced": replaced}

    def host_slabs(self, n_cells, slab_cells=2048):
        """Cell ranges ``[(cell0, n_cells), ...]`` of the slabs ``detrend_host`` cuts a batch into
        (``attrici_host_slab_range``; a short first slab shortens the fill of the copy / compute pipeline)."""
        slab = int(min(slab_cells, n_cells))
        c0, nc = C.c_int(0), C.c_int(0)
        n = self.lib.attrici_host_slab_range(n_cells, slab, -1, None, None)
        if n < 0:
            _lib.check(n)
        out = []
        for k in range(n):
            self.lib.attrici_host_slab_range(n_cells, slab, k, C.byref(c0), C.byref(nc))
            out.append((c0.value, nc.value))
        return out

    def detrend_host(self, y_host, days, gmt, fit_window=None, seeds=None, slab_cells=2048, out=None,
                     want_replaced=True, cfact_dtype=torch.float64, replaced="dense", slab_major=False,
                     sparse_capacity=None):
        """Whole hot path with HOST buffers (pinned recommended): slabs of cells, copies and kernels
        overlapped on the library's two internal streams (``attrici_detrend_host_ex``).

        ``y_host``: [T, C] float32 numpy array or CPU tensor.  Returns a dict of CPU tensors (pinned):
        cfact [T, C] (``cfact_dtype``: float64 like the reference's output, or float32 = the FP64 result rounded once
        to the dtype of the input files), replaced, params [C, P], logp [C], scaling [C, 2], status [C], n_pass [C].

        What crosses PCIe bounds this call, so the result layout is selectable:
        ``replaced="dense"``: [T, C] uint8 codes; ``"sparse"``: ``replaced_sparse`` [n, 3] int32 rows (day, cell, code)
        of the non-zero entries (unordered; ``io.dense_replaced`` rebuilds the plane); ``None`` / want_replaced=False: none.
        ``slab_major=True``: ``cfact`` (and a dense ``replaced``) come back as one flat pinned tensor holding the cell
        slabs one after the other, each [T, n_k] contiguous (one linear copy per slab instead of a T-row pitched one);
        ``cfact_slabs`` / ``replaced_slabs`` are the per-slab [T, n_k] views and ``slabs`` their (cell0, n_k)."""
        yt = y_host if isinstance(y_host, torch.Tensor) else torch.from_numpy(np.asarray(y_host))
        if yt.dtype != torch.float32 or yt.dim() != 2 or yt.stride(1) != 1 or yt.is_cuda:
            raise TypeError("y_host must be a 2-D float32 host array [T, C] with unit stride along cells")
        if cfact_dtype not in (torch.float64, torch.float32):
            raise TypeError("cfact_dtype must be torch.float64 or torch.float32")
        if not want_replaced:
            replaced = None
        if replaced not in ("dense", "sparse", None):
            raise ValueError("replaced must be 'dense', 'sparse' or None")
        T, C_ = yt.shape
        i0, i1 = (0, T) if fit_window is None else fit_window
        slab = int(min(slab_cells, C_))
        need = _lib.detrend_host_scratch_bytes(self.spec.family, T, slab)
        if getattr(self, "_scratch", None) is None or self._scratch.numel() < need:
            self._scratch = None
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        pin = torch.cuda.is_available()
        if sparse_capacity is None:
            sparse_capacity = max(4096, (T * C_) // 16)
        key = (T, C_, cfact_dtype, replaced, bool(slab_major))
        if out is not None and "_key" not in out:
            out["_key"] = key      # buffers allocated by the caller (time-major, dense): used as they are
        if out is None or out.get("_key") != key:
            shape = (T * C_,) if slab_major else (T, C_)
            out = {
                "_key": key,
                "cfact": torch.empty(shape, dtype=cfact_dtype, pin_memory=pin),
                "replaced": torch.empty(shape, dtype=torch.uint8, pin_memory=pin) if replaced == "dense" else None,
                "params": torch.empty((C_, self.n_params), dtype=torch.float64, pin_memory=pin),
                "logp": torch.empty(C_, dtype=torch.float64, pin_memory=pin),
                "scaling": torch.empty((C_, 2), dtype=torch.float64, pin_memory=pin),
                "status": torch.empty(C_, dtype=torch.int32, pin_memory=pin),
                "n_pass": torch.empty(C_, dtype=torch.int32, pin_memory=pin),
            }
            if replaced == "sparse":
                out["_sparse_buf"] = torch.empty((int(sparse_capacity), 3), dtype=torch.int32, pin_memory=pin)
            if slab_major:
                out["slabs"] = self.host_slabs(C_, slab)
                out["cfact_slabs"] = [out["cfact"][T * c0:T * (c0 + n)].view(T, n) for c0, n in out["slabs"]]
                if replaced == "dense":
                    out["replaced_slabs"] = [out["replaced"][T * c0:T * (c0 + n)].view(T, n) for c0, n in out["slabs"]]
        days = np.ascontiguousarray(days, dtype=np.int32)
        gmt = np.ascontiguousarray(gmt, dtype=np.float64)
        if len(days) != T or len(gmt) != T:
            raise ValueError("days / gmt must have length T")
        seeds_np = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)

        def hp(t):
            return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

        ho = _lib.HostOptions()
        _lib.check(self.lib.attrici_default_host_options(C.byref(ho)))
        ho.prep_flags = self.prep_flags
        ho.cfact_dtype = _lib.F32 if cfact_dtype == torch.float32 else _lib.F64
        ho.replaced_mode = {"dense": _lib.FLAGS_DENSE, "sparse": _lib.FLAGS_SPARSE, None: _lib.FLAGS_OFF}[replaced]
        ho.slab_major = int(bool(slab_major))
        sparse_buf = out.get("_sparse_buf")
        ho.sparse_capacity = 0 if sparse_buf is None else sparse_buf.shape[0]
        n_sparse = C.c_int64(0)
        with torch.cuda.device(self.device):
            # the library works on its own streams: whatever the torch stream still has queued on memory the caching
            # allocator may hand out again (the scratch above, outputs of an earlier call) must have finished
            torch.cuda.current_stream(self.device).synchronize()
            _lib.check(self.lib.attrici_detrend_host_ex(
                C.byref(self.var), C.byref(self.options), C.byref(ho), hp(yt), yt.stride(0), T, C_, i0, i1,
                C.c_void_p(days.ctypes.data), C.c_void_p(gmt.ctypes.data),
                C.c_void_p(seeds_np.ctypes.data) if seeds_np is not None else C.c_void_p(0),
                hp(out["cfact"]), 0 if slab_major else out["cfact"].stride(0), hp(out["replaced"]), hp(sparse_buf),
                C.byref(n_sparse), hp(out["params"]), hp(out["logp"]),
                hp(out["scaling"]), hp(out["status"]), hp(out["n_pass"]), _ptr(self._scratch),
                self._scratch.numel(), slab))
        if sparse_buf is not None:
            out["replaced_sparse"] = sparse_buf[:n_sparse.value]
        return out
