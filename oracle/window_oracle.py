"""CPU oracle of the rolling-window model (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product).

Restates, with NumPy / SciPy, for the Normal family:
  * ``collect_windows`` (attrici/util.py:107-172): day-of-year groups, day 366 = the day after each day 365 (:126-133),
    series mirrored by half a window at both ends (:139-150), centred rolling windows (:153-157), windows of a day of the
    year joined over the years (:160-169) - here as index multisets ``windows(time, window_size)[k]``;
  * the PyMC5 rolling-window log posterior (attrici/estimation/model_pymc5.py:366-547, 610-617): per day of the year
    ``mu = a_k + b_k * predictor``, ``sigma = exp(c_k)``, Normal likelihood of the windowed days, N(0, 1) priors on all
    weights (PRIOR_INTERCEPT_MU / SIGMA, :104-105, used for intercepts and trends alike, :403-414);
  * ``estimate`` by ``time.dt.dayofyear`` (:435-447, :524-538) and the Normal quantile map of the main oracle.

PARITY UNPINNED: the reference runs this model through PyMC5 only (its SciPy solver raises NotImplementedError,
model_scipy.py:58-59), has no test and no fixture for it, and PyMC is not installed in this image.  What is pinned: the
day-of-year integers and the window membership, against pandas (``tests/test_window_oracle.py``), which is what xarray's
``dt.dayofyear`` / ``rolling().construct()`` compute with.  ``logp`` here is the full log posterior at the MAP (PyMC's
``logp`` Deterministic holds the prior part only).
"""

import numpy as np
from scipy import optimize, special

LN_SQRT_2PI = 0.9189385332046727


def dayofyear(time):
    d = np.asarray(time).astype("datetime64[D]")
    return (d - d.astype("datetime64[Y]").astype("datetime64[D]")).astype(np.int64) + 1


def windows(time, window_size):
    """index multiset (into the fit series) of every day of the year: dict k -> int array, k = 1..366"""
    n = len(time)
    half = window_size // 2
    doy = dayofyear(time)
    groups = {k: np.flatnonzero(doy == k) for k in range(1, 367)}
    groups[366] = groups[365] + 1                                    # util.py:132
    # extended series: indices of the original days it holds (util.py:139-150)
    ext = np.concatenate([np.arange(half, 0, -1), np.arange(n), np.arange(n - 2, n - half - 3, -1)])
    out = {}
    for k in range(1, 367):
        idx = []
        for c in groups[k]:                                          # window of centre c = ext[c .. c + 2 half]
            idx.append(ext[c:c + window_size])
        out[k] = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
    return out


def sums(y_scaled, gmt, time, window_size):
    """[366, 6] {n, sum g, sum g^2, sum y, sum y^2, sum y g} over the valid days of every window set"""
    y = np.asarray(y_scaled, dtype=np.float64)
    g = np.asarray(gmt, dtype=np.float64)
    w = windows(time, window_size)
    s = np.zeros((366, 6))
    for k in range(1, 367):
        yy, gg = y[w[k]], g[w[k]]
        ok = ~np.isnan(yy)
        yy, gg = yy[ok], gg[ok]
        s[k - 1] = [ok.sum(), gg.sum(), (gg * gg).sum(), yy.sum(), (yy * yy).sum(), (yy * gg).sum()]
    return s


def objective(theta, y, g):
    a, b, c = theta
    z = (y - a - b * g) * np.exp(-c)
    return np.sum(0.5 * z * z + c + LN_SQRT_2PI) + 0.5 * (a * a + b * b + c * c) + 3 * LN_SQRT_2PI


def fit(y_scaled, gmt, time, window_size):
    """MAP per day of the year from the windowed days themselves (not from sums): params [3, 366], logp"""
    y = np.asarray(y_scaled, dtype=np.float64)
    g = np.asarray(gmt, dtype=np.float64)
    w = windows(time, window_size)
    params = np.zeros((3, 366))
    total = 0.0
    for k in range(1, 367):
        yy, gg = y[w[k]], g[w[k]]
        ok = ~np.isnan(yy)
        yy, gg = yy[ok], gg[ok]
        if yy.size == 0:
            total += 3 * LN_SQRT_2PI
            continue
        A = np.stack([np.ones_like(gg), gg], axis=1)
        ab, *_ = np.linalg.lstsq(A, yy, rcond=None)
        r = yy - A @ ab
        x0 = np.array([ab[0], ab[1], 0.5 * np.log(max(np.mean(r * r), 1e-30))])
        res = optimize.minimize(objective, x0, args=(yy, gg), method="BFGS", options={"gtol": 1e-10, "maxiter": 500})
        x = res.x
        for _ in range(20):                                          # polish: exact Newton steps
            a, b, c = x
            e2 = np.exp(-2 * c)
            rr = yy - a - b * gg
            grad = np.array([-e2 * rr.sum() + a, -e2 * (rr * gg).sum() + b, yy.size - e2 * (rr * rr).sum() + c])
            H = np.array([[e2 * yy.size + 1, e2 * gg.sum(), 2 * e2 * rr.sum()],
                          [e2 * gg.sum(), e2 * (gg * gg).sum() + 1, 2 * e2 * (rr * gg).sum()],
                          [2 * e2 * rr.sum(), 2 * e2 * (rr * gg).sum(), 2 * e2 * (rr * rr).sum() + 1]])
            step = np.linalg.solve(H, -grad)
            if objective(x + step, yy, gg) <= objective(x, yy, gg):
                x = x + step
            if np.abs(step).max() < 1e-14:
                break
        params[:, k - 1] = x
        total += objective(x, yy, gg)
    return params, -total


def detrend_cell(y, y_scaled, scaling, gmt_full, time_full, params):
    """Normal quantile map by day of the year: cfact, replaced codes (0 / 1 NaN / 2 +Inf / 3 -Inf), cfact_scaled"""
    doy = dayofyear(time_full) - 1
    a, b, c = params[0][doy], params[1][doy], params[2][doy]
    g = np.asarray(gmt_full, dtype=np.float64)
    ys = np.asarray(y_scaled, dtype=np.float64)
    sigma = np.exp(c)
    with np.errstate(invalid="ignore", divide="ignore"):
        q = special.ndtr((ys - (a + b * g)) / sigma)
        cs = a + sigma * special.ndtri(q)
    datamin, scale = scaling
    cf = cs * scale + datamin
    rep = np.zeros(len(ys), dtype=np.uint8)
    yo = np.asarray(y, dtype=np.float64)
    rep[np.isnan(cf)] = 1
    rep[np.isposinf(cf)] = 2
    rep[np.isneginf(cf)] = 3
    cf = np.where(rep > 0, yo, cf)
    return cf, rep, cs
