"""CPU oracle for the SSA smoothing of the GMT predictor (TEST INFRASTRUCTURE ONLY).

Restates, in O(n W) NumPy, what ``calc_gmt_by_ssa`` (attrici/preprocessing.py:13-40) takes from the vendored
pyts class (attrici/vendored/singularspectrumanalysis.py:161-207): the first row of
``SingularSpectrumAnalysis(window_size).transform(gmt[::subset])`` with ``groups=None``, i.e. the first
elementary component

    X[i, k] = x[i + k]                      (W x K trajectory matrix, K = n - W + 1;  :165-169)
    v       = leading eigenvector of X X^T  (:170-172)
    E       = v v^T X = v (X^T v)^T         (:174, _outer_dot :51-57, first slice only)
    out[t]  = mean of E over the anti-diagonal i + k = t   (_diagonal_averaging :60-76; the W >= K
                                                            branch :178-182 transposes E, same means)

The reference materialises all W elementary matrices (a (W, W, K) tensor: 4.4 GB and minutes for the 1901-2023
daily series at subset=10); only the first is ever used.  Pinned by ``tests/golden/ssa_component.npz`` (outputs
of the unmodified reference, made by ``tests/golden/make_golden_ssa.py``) and by the reference's own fixture
pair ``20crv3-era5_gmt_raw.nc -> 20CRv3-ERA5_germany_ssa_gmt.nc`` (reference tests/test_ssa.py:27-44).

Only ``tests/`` import this module; the product (``attrici_b200.preprocessing``) never does.
"""

import numpy as np


def ssa_first_component(x, window_size):
    """First SSA elementary component of the 1-D series ``x`` (float64 arithmetic)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    W = int(window_size)
    if not 2 <= W <= n:
        raise ValueError(f"window_size must be between 2 and n_timestamps (got {W}, n = {n})")
    K = n - W + 1
    X = np.lib.stride_tricks.sliding_window_view(x, K)          # X[i, k] = x[i + k], shape (W, K)
    G = X @ X.T
    w, V = np.linalg.eigh(G)
    v = V[:, -1]                                                # eigenvalues ascending: last = largest
    u = X.T @ v                                                 # (K,)
    out = np.zeros(n)
    cnt = np.zeros(n)
    for i in range(W):
        out[i:i + K] += v[i] * u
        cnt[i:i + K] += 1.0
    return out / cnt


def calc_gmt_by_ssa(gmt, times, window_size=365, subset=10):
    """attrici/preprocessing.py:13-40 with the O(n W) component above."""
    g = np.squeeze(np.asarray(gmt)[::subset])
    return ssa_first_component(g, window_size), np.asarray(times)[::subset]
