"""Rolling-window model, CPU side: the day-of-year integers and the window membership of oracle/window_oracle.py and of
attrici_b200/window.py against pandas (what xarray's ``dt.dayofyear`` / ``rolling().construct()`` of the reference's
``collect_windows``, attrici/util.py:107-172, compute with)."""
import numpy as np
import pandas as pd
import pytest

from attrici_b200 import window as win
from oracle import window_oracle as wo


def axis(start, n):
    return (np.datetime64(start) + np.arange(n)).astype("datetime64[D]")


@pytest.mark.parametrize("start,n", [("1901-01-01", 4000), ("1999-03-17", 1300), ("2003-12-30", 800), ("1896-02-27", 1800)])
def test_dayofyear_bit_equal_to_pandas(start, n):
    t = axis(start, n)
    want = pd.DatetimeIndex(t).dayofyear.values
    assert np.array_equal(win.dayofyear(t), want) and win.dayofyear(t).dtype == np.int16
    assert np.array_equal(wo.dayofyear(t), want)


def brute_force_windows(time, window_size):
    """collect_windows with pandas: the mirrored extension, centred rolling windows of the INDEX series, day-of-year
    groups with the 366 rule - written against util.py:126-169 line by line"""
    n = len(time)
    half = window_size // 2
    doy = pd.DatetimeIndex(time).dayofyear.values
    groups = {k: list(np.flatnonzero(doy == k)) for k in range(1, 367)}
    groups[366] = [i + 1 for i in groups[365]]
    data = np.arange(n)
    ext = np.concatenate([data[list(reversed(range(1, half + 1)))], data, data[list(reversed(range(n - half - 2, n - 1)))]])
    roll = [w.values for w in pd.Series(ext).rolling(window_size, center=True)][half:len(ext) - half]
    return {k: np.concatenate([roll[i] for i in groups[k]]).astype(np.int64) if groups[k] else np.zeros(0, np.int64)
            for k in range(1, 367)}


@pytest.mark.parametrize("start,n,w", [("1901-01-01", 1500, 31), ("1999-03-17", 900, 5), ("1903-07-02", 1200, 15),
                                        ("2000-01-01", 366 + 365, 31)])
def test_window_membership_against_pandas(start, n, w):
    t = axis(start, n)
    got, want = wo.windows(t, w), brute_force_windows(t, w)
    for k in range(1, 367):
        assert np.array_equal(got[k], want[k]), k
    # every window is complete and every index lies inside the series
    assert all(len(got[k]) % w == 0 and (len(got[k]) == 0 or (got[k].min() >= 0 and got[k].max() < n)) for k in got)


@pytest.mark.parametrize("start,n", [("1901-01-01", 3000), ("1999-03-17", 1300), ("2003-12-30", 800)])
def test_runs_reproduce_the_dayofyear_groups(start, n):
    """centres from (start, first day of the year, length) runs == groupby("time.dayofyear").groups, incl. day 366"""
    t = axis(start, n)
    doy = win.dayofyear(t)
    s, f, ln = win.window_runs(doy)
    assert s[0] == 0 and ln.sum() == n and np.array_equal(s[1:], np.cumsum(ln)[:-1])
    for k in range(1, 366):
        centres = [int(s[r] + (k - f[r])) for r in range(len(s)) if f[r] <= k <= min(f[r] + ln[r] - 1, 365)]
        assert centres == list(np.flatnonzero(doy == k)), k
    c366 = [int(s[r] + (365 - f[r]) + 1) for r in range(len(s)) if f[r] <= 365 <= f[r] + ln[r] - 1]
    assert c366 == [i + 1 for i in np.flatnonzero(doy == 365)]


def test_oracle_sums_and_fit_are_consistent():
    """the MAP found on the windowed days is the stationary point of the objective written on the six sums"""
    rng = np.random.default_rng(0)
    t = axis("1950-01-01", 2200)
    g = np.linspace(0, 1, len(t))
    y = (0.4 + 0.1 * np.cos(2 * np.pi * np.arange(len(t)) / 365.25) + 0.2 * g + 0.05 * rng.standard_normal(len(t))).astype(np.float32)
    y[100:130] = np.nan
    S = wo.sums(y, g, t, 31)
    params, logp = wo.fit(y, g, t, 31)
    for k in (0, 59, 200, 364, 365):
        n, sg, sgg, sy, syy, syg = S[k]
        a, b, c = params[:, k]
        e2 = np.exp(-2 * c)
        r0, r1 = sy - a * n - b * sg, syg - a * sg - b * sgg
        Q = syy - a * sy - b * syg - a * r0 - b * r1
        grad = np.array([-e2 * r0 + a, -e2 * r1 + b, n - e2 * Q + c])
        assert np.abs(grad).max() < 1e-7 * max(n, 1.0), (k, grad)
    assert np.isfinite(logp)
