"""Rolling-window model (``--window-size``) on the GPU against oracle/window_oracle.py (parity unpinned: the reference runs
this model through PyMC5 only and ships no test or fixture for it; the oracle's integers are pinned to pandas in
tests/test_window_oracle.py).  attrici/util.py:107-172, attrici/estimation/model_pymc5.py:366-547."""
import numpy as np
import pytest
import torch

from attrici_b200 import synthetic
from attrici_b200.detrend import Config, detrend_cells
from attrici_b200.window import TRACE_KEYS, WindowBatchSolver
from oracle import window_oracle as wo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def case(start, T, C, seed, nan_block=True):
    time = (np.datetime64(start) + np.arange(T)).astype("datetime64[D]")
    y = synthetic.tas_cells(C, T, seed=seed)
    if nan_block:
        y[300:340, 1] = np.nan
        y[::7, 2] = np.nan
    return time, y, synthetic.smoothed_gmt(T)


@pytest.mark.parametrize("start,T,window,fit_window", [("1990-03-05", 2600, 31, None), ("1901-01-01", 3300, 15, (200, 3100)),
                                                       ("1999-12-30", 1500, 5, None)])
def test_window_model_against_oracle(start, T, window, fit_window):
    C = 5
    time, y, gmt = case(start, T, C, seed=3)
    s = WindowBatchSolver("tas", window, device=DEV)
    out = s.detrend_device(torch.from_numpy(y).to(DEV), time, gmt, fit_window)
    assert (out["status"].cpu().numpy() == 0).all()
    ys = out["y_scaled"].cpu().numpy()
    params, logp = out["params"].cpu().numpy(), out["logp"].cpu().numpy()
    cfact, replaced = out["cfact"].cpu().numpy(), out["replaced"].cpu().numpy()
    scaling = out["scaling"].cpu().numpy()
    i0, i1 = fit_window or (0, T)
    for c in range(C):
        ref_p, ref_logp = wo.fit(ys[i0:i1, c], gmt[i0:i1], time[i0:i1], window)
        np.testing.assert_allclose(logp[c], ref_logp, rtol=1e-9)
        np.testing.assert_allclose(params[c], ref_p, rtol=0, atol=2e-6)
        # the quantile map at the kernel's own weights
        cf, rep, _ = wo.detrend_cell(y[:, c], ys[:, c], scaling[c], gmt, time, params[c])
        assert np.array_equal(rep, replaced[:, c])
        np.testing.assert_allclose(cfact[:, c], cf, rtol=1e-10, equal_nan=True)
    # sigma does not depend on the predictor: the map is a shift by the change of mu on unreplaced days
    k = wo.dayofyear(time) - 1
    shift = (cfact[:, 0] - y[:, 0].astype(np.float64)) / scaling[0, 1]
    ok = (replaced[:, 0] == 0) & (np.abs((ys[:, 0] - (params[0, 0, k] + params[0, 1, k] * gmt)) / np.exp(params[0, 2, k])) < 3.5)
    np.testing.assert_allclose(shift[ok], (-params[0, 1, k] * gmt)[ok], rtol=0, atol=2e-6)


def test_window_sums_slide_exactly():
    """sliding window sums == the multiset sums of collect_windows, for a window longer than the shortest run"""
    T, C, window = 1200, 3, 41
    time, y, gmt = case("2003-11-20", T, C, seed=9)
    s = WindowBatchSolver("tas", window, device=DEV)
    out = s.detrend_device(torch.from_numpy(y).to(DEV), time, gmt)
    ys = out["y_scaled"].cpu().numpy()
    ref_p, ref_logp = wo.fit(ys[:, 2], gmt, time, window)
    np.testing.assert_allclose(out["logp"].cpu().numpy()[2], ref_logp, rtol=1e-9)


def test_window_model_through_the_plugin_and_the_driver():
    """``ModelCuda(window_size=...)`` (the reference's Model interface, PyMC5 trace schema) and
    ``detrend_cells(Config(window_size=...))`` give the batch solver's results; families other than Normal and an even
    window are refused like the reference does (detrend.py:646-649, model_scipy.py:58-59)."""
    from attrici_b200.estimation import ModelCuda

    T, C, window = 2000, 4, 31
    time, y, gmt = case("1950-01-01", T, C, seed=5, nan_block=False)
    lat, lon = synthetic.tile_coordinates(2, 2)
    res = detrend_cells(Config("tas", modes=None, window_size=window, slab_cells=3), y, gmt, None, lat, lon, time=time)
    direct = WindowBatchSolver("tas", window, device=DEV).detrend_device(torch.from_numpy(y).to(DEV), time, gmt)
    assert np.array_equal(res["params"], direct["params"].cpu().numpy())
    assert np.array_equal(res["cfact"], direct["cfact"].cpu().numpy(), equal_nan=True)

    class Normal:
        def __init__(self, mu, sigma):
            self.mu, self.sigma = mu, sigma

    class Arr:
        def __init__(self, values, time):
            self.values, self.time = values, self
            self._t = time

        def __array__(self, dtype=None, copy=None):
            return self._t

    class TimeCoord:
        def __init__(self, t):
            self.values = t

    class Series:
        def __init__(self, values, t):
            self.values, self.time = values, TimeCoord(t)

    ys = direct["y_scaled"].cpu().numpy()[:, 0]
    m = ModelCuda(Normal, {"mu": object(), "sigma": object()}, Series(ys, time), Series(gmt, time), window_size=window)
    trace = m.fit()
    assert set(trace) == set(TRACE_KEYS) | {"logp"} and all(trace[k].shape == (366,) for k in TRACE_KEYS)
    np.testing.assert_allclose(trace["logp"], direct["logp"].cpu().numpy()[0], rtol=1e-12)
    d = m.estimate_distribution(trace, Series(gmt, time))
    k = wo.dayofyear(time) - 1
    np.testing.assert_allclose(d.mu, trace[TRACE_KEYS[0]][k] + trace[TRACE_KEYS[1]][k] * gmt, rtol=1e-15)
    with pytest.raises(ValueError, match="odd"):
        detrend_cells(Config("tas", modes=None, window_size=30), y, gmt, None, lat, lon, time=time)
    with pytest.raises(NotImplementedError):
        WindowBatchSolver("pr", 31, device=DEV)
    with pytest.raises(ValueError, match="Exactly one"):
        detrend_cells(Config("tas", modes=4, window_size=31), y, gmt, None, lat, lon, time=time)
