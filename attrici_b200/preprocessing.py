"""SSA smoothing of the GMT predictor on the GPU, mirroring ``attrici.preprocessing`` of the reference.

``calc_gmt_by_ssa(gmt, times, window_size=365, subset=10)`` has the reference's signature and return value
(attrici/preprocessing.py:13-40): the first reconstructed SSA component of ``gmt[::subset]`` and ``times[::subset]``.
The reference builds all ``window_size`` elementary matrices of the vendored pyts class (a (W, W, K) tensor) and
keeps row 0; ``libattrici_b200`` computes that component alone (``csrc/ssa.cu``: Gram matrix, leading eigenvector
by power iteration, projection, diagonal averaging, all float64).  There is no CPU fallback.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import AttriciCudaError


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def ssa_first_component(x, window_size, device=None, max_iter=200000, tol=1e-13, return_info=False):
    """Row 0 of ``SingularSpectrumAnalysis(window_size).transform(x[None, :])`` (groups=None).

    ``x``: 1-D array-like or tensor (any float dtype; computed in float64).  Returns a float64 numpy array
    (a CUDA tensor if ``x`` is a CUDA tensor).  Raises ``ValueError`` for a window outside ``[2, len(x)]`` like
    pyts (singularspectrumanalysis.py:253-262) and ``AttriciCudaError`` if the eigenvector iteration does not
    reach ``tol`` within ``max_iter`` steps (two leading eigenvalues that coincide to ~1e-5 relative)."""
    if not isinstance(window_size, (int, np.integer)):
        raise TypeError("'window_size' must be an integer.")
    if not torch.cuda.is_available():
        raise AttriciCudaError("attrici_b200 needs a CUDA device; there is no CPU fallback")
    lib = _lib.load()
    on_device = isinstance(x, torch.Tensor) and x.is_cuda
    dev = x.device if on_device else torch.device(device if device is not None else "cuda")
    xt = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))
    xt = xt.reshape(-1).to(device=dev, dtype=torch.float64).contiguous()
    n = int(xt.numel())
    W = int(window_size)
    if not 2 <= W <= n:
        raise ValueError("If 'window_size' is an integer, it must be greater than or equal to 2 and lower than or "
                         f"equal to n_timestamps (got {W}).")
    need = C.c_size_t(0)
    _lib.check(lib.attrici_ssa_workspace_bytes(n, W, C.byref(need)))
    with torch.cuda.device(dev):
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        info = torch.zeros(2, dtype=torch.int32, device=dev)
        lam = torch.zeros(1, dtype=torch.float64, device=dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.attrici_ssa_first_component(_ptr(xt), n, W, _ptr(out), _ptr(ws), ws.numel(), int(max_iter),
                                                   float(tol), _ptr(info), _ptr(lam), stream))
        iters, ok = (int(v) for v in info.cpu())
    if not ok:
        raise AttriciCudaError(f"SSA eigenvector iteration did not reach tol={tol} in {iters} steps")
    res = out if on_device else out.cpu().numpy()
    if return_info:
        return res, {"iterations": iters, "eigenvalue": float(lam.item())}
    return res


def calc_gmt_by_ssa(gmt, times, window_size=365, subset=10, device=None):
    """Smoothed GMT by SSA, ``(ssa[0, :], times[::subset])`` as attrici/preprocessing.py:13-40."""
    g = np.squeeze(np.asarray(gmt)[::subset])
    return ssa_first_component(g, window_size, device=device), np.asarray(times)[::subset]
