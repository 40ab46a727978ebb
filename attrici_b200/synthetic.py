"""Seeded synthetic inputs of the BASELINE.json configurations.

There is no network for ISIMIP data, so ``bench.py`` and the parity tests draw daily series with the
statistical structure the reference models: a Fourier annual cycle whose coefficients depend linearly on
a smoothed GMT predictor scaled to [0, 1], cast to float32 like ISIMIP files.  ``numpy`` generators are
used by the tests (the oracle consumes the same arrays); ``*_torch`` variants build large tiles directly
in HBM for the bench.
"""

import numpy as np

DAYS_1901_2019 = 43464  # 1901-01-01 ... 2019-12-31, proleptic Gregorian


def time_axis(n_days=DAYS_1901_2019):
    """Integer day offsets of a gap-free daily axis."""
    return np.arange(n_days, dtype=np.int32)


def smoothed_gmt(n_days=DAYS_1901_2019):
    """A smooth warming curve standing in for SSA-smoothed GMT (attrici/preprocessing.py:13-40), stretched
    to the daily axis and min-max scaled to [0, 1] as attrici/detrend.py:698-707 does."""
    x = np.arange(n_days, dtype=np.float64) / (n_days - 1)
    g = 1.0 / (1.0 + np.exp(-(x - 0.72) / 0.10)) + 0.06 * np.sin(2 * np.pi * x * 1.7) + 0.1 * x
    return (g - g.min()) / (g.max() - g.min())


def raw_gmt(n_days=DAYS_1901_2019, seed=7):
    """Daily raw global-mean temperature in K: 14 degC + a 1.2 K sigmoid-like ramp + ENSO-like
    AR(1) noise + a small annual cycle - the kind of series ``attrici ssa`` reads (attrici/ssa.py:12-59)."""
    rng = np.random.default_rng(seed)
    x = np.arange(n_days, dtype=np.float64) / max(n_days - 1, 1)
    ramp = 1.2 / (1.0 + np.exp(-(x - 0.72) / 0.10))
    eps = rng.standard_normal(n_days) * 0.02
    ar = np.empty(n_days)
    acc = 0.0
    for i in range(n_days):                      # AR(1), e-folding time ~ 200 days
        acc = 0.995 * acc + eps[i]
        ar[i] = acc
    return 287.15 + ramp + ar + 0.1 * np.cos(2 * np.pi * np.arange(n_days) / 365.25)


def ssa_gmt(n_days=DAYS_1901_2019, seed=7, window_size=365, subset=10, calc_gmt_by_ssa=None):
    """The predictor of BASELINE configs[1]: ``raw_gmt`` smoothed by SSA (``calc_gmt_by_ssa``, window 365 on every
    10th day: attrici/preprocessing.py:13-40), stretched to the daily axis and scaled to [0, 1]
    (attrici/detrend.py:698-707).  ``calc_gmt_by_ssa``: the smoother, default the library's SSA kernels
    (``attrici_b200.preprocessing``); the CPU arm of bench.py and the tests pass the oracle's restatement."""
    from .detrend import scale_predictor

    if calc_gmt_by_ssa is None:
        from .preprocessing import calc_gmt_by_ssa
    raw = raw_gmt(n_days, seed)
    w = min(window_size, len(raw[::subset]))
    smooth, _ = calc_gmt_by_ssa(raw, np.arange(n_days), window_size=w, subset=subset)
    return scale_predictor(np.asarray(smooth), n_time=n_days)


def _tau(n_days):
    return np.arange(n_days, dtype=np.float64) / 365.25


def tile_coordinates(n_lat, n_lon, lat0=0.25, lon0=0.25, step=0.5):
    """lat-major, lon-minor stacking of a regular tile (attrici/detrend.py:672)."""
    lat = lat0 + step * np.arange(n_lat)
    lon = lon0 + step * np.arange(n_lon)
    la, lo = np.meshgrid(lat, lon, indexing="ij")
    return la.ravel(), lo.ravel()


def tas_cells(n_cells, n_days=DAYS_1901_2019, seed=1234, nan_fraction=0.0, first_cell=0):
    """Config 2: Gaussian tas, [n_days, n_cells] float32 (time-major)."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1, s1, c2 = np.cos(2 * np.pi * tau), np.sin(2 * np.pi * tau), np.cos(4 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        mu0 = rng.uniform(250, 300)
        a1, b1 = rng.normal(0, 8, 2)
        a2 = rng.normal(0, 2)
        s = rng.normal(1.5, 1.0)
        sig0, phi = rng.uniform(1.5, 4.0), rng.uniform(0, 2 * np.pi)
        sigma = sig0 * np.exp(0.3 * np.cos(2 * np.pi * tau + phi))
        col = mu0 + a1 * c1 + b1 * s1 + a2 * c2 + s * gmt + sigma * rng.standard_normal(n_days)
        if nan_fraction > 0 and rng.uniform() < nan_fraction:
            start = int(rng.integers(1, n_days - 400))
            col[start : start + 365] = np.nan
        y[:, c] = col.astype(np.float32)
    return y


def pr_cells(n_cells, n_days=DAYS_1901_2019, seed=4321, first_cell=0):
    """Config 3: Bernoulli-Gamma pr in kg m-2 s-1, [n_days, n_cells] float32."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1 = np.cos(2 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        a, b, cc = rng.normal(0, 0.7), rng.normal(0, 0.4), rng.normal(0, 0.3)
        p_dry = 1.0 / (1.0 + np.exp(-(a + b * c1 + cc * gmt)))
        k = rng.uniform(0.6, 1.2)
        al, be, ga = rng.normal(0.0, 0.3), rng.normal(0, 0.3), rng.normal(0.1, 0.2)
        mean = np.exp(al + be * c1 + ga * gmt)
        wet = rng.uniform(size=n_days) >= p_dry
        amount = rng.gamma(k, mean / k) * 3e-5 + 1.1574e-6 * 1.5
        y[:, c] = np.where(wet, amount, 0.0).astype(np.float32)
    return y


def wind_cells(n_cells, n_days=DAYS_1901_2019, seed=777, first_cell=0):
    """Config 4: Weibull sfcWind in m s-1."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1, s1 = np.cos(2 * np.pi * tau), np.sin(2 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        shape = 2.0 + rng.normal(0, 0.3)
        scale = np.exp(rng.normal(1.5, 0.2) + 0.2 * rng.normal() * c1 + 0.15 * rng.normal() * s1 + 0.1 * rng.normal() * gmt)
        y[:, c] = (scale * rng.weibull(max(shape, 1.0), n_days)).astype(np.float32)
    return y


def hurs_cells(n_cells, n_days=DAYS_1901_2019, seed=888, first_cell=0):
    """Config 4: Beta hurs in percent (values are clipped into (0, 100), a few land on the masks)."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1, s1 = np.cos(2 * np.pi * tau), np.sin(2 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        eta = rng.normal(0.8, 0.4) + 0.5 * rng.normal() * c1 + 0.3 * rng.normal() * s1 - 0.2 * abs(rng.normal()) * gmt
        mu = 1.0 / (1.0 + np.exp(-eta))
        phi = rng.uniform(8, 30)
        v = rng.beta(mu * phi, (1 - mu) * phi) * 100.0
        y[:, c] = np.clip(v, 0.0, 100.0).astype(np.float32)
    return y


def tasrange_cells(n_cells, n_days=DAYS_1901_2019, seed=999, first_cell=0):
    """Config 4: Gamma tasrange in K."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1, s1 = np.cos(2 * np.pi * tau), np.sin(2 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        mean = np.exp(rng.normal(2.0, 0.2) + 0.3 * rng.normal() * c1 + 0.2 * rng.normal() * s1 + 0.05 * rng.normal() * gmt)
        k = rng.uniform(4, 12)
        y[:, c] = rng.gamma(k, mean / k).astype(np.float32)
    return y


def rsds_cells(n_cells, n_days=DAYS_1901_2019, seed=555, first_cell=0):
    """Config 4: Normal rsds with a polar-night-like winter floor near 0 (exercises ``<= 0 -> NaN -> replaced``)."""
    tau, gmt = _tau(n_days), smoothed_gmt(n_days)
    y = np.empty((n_days, n_cells), dtype=np.float32)
    c1 = np.cos(2 * np.pi * tau)
    for c in range(n_cells):
        rng = np.random.default_rng(seed + first_cell + c)
        base = rng.uniform(90, 140)
        col = base - (base - 3.0) * c1 * rng.uniform(0.9, 1.0) + 4.0 * gmt * rng.normal() + rng.uniform(4, 9) * rng.standard_normal(n_days)
        y[:, c] = np.maximum(col, 0.5).astype(np.float32)
    return y


GENERATORS = {"tas": tas_cells, "pr": pr_cells, "sfcWind": wind_cells, "hurs": hurs_cells, "tasrange": tasrange_cells,
              "rsds": rsds_cells}


def tas_tile_torch(n_cells, n_days=DAYS_1901_2019, seed=1234, device="cuda:0", nan_fraction=0.001):
    """Config 2 built directly in HBM: same model as ``tas_cells`` (different random stream), [T, C] float32."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dev = torch.device(device)
    tau = torch.arange(n_days, dtype=torch.float64, device=dev) / 365.25
    gmt = torch.as_tensor(smoothed_gmt(n_days), device=dev)

    def u(lo, hi):
        return lo + (hi - lo) * torch.rand(n_cells, generator=g, device=dev, dtype=torch.float32)

    def n(m, s):
        return m + s * torch.randn(n_cells, generator=g, device=dev, dtype=torch.float32)

    mu0, a1, b1, a2, s = u(250, 300), n(0, 8), n(0, 8), n(0, 2), n(1.5, 1.0)
    sig0, phi = u(1.5, 4.0), u(0, 2 * np.pi)
    y = torch.empty((n_days, n_cells), dtype=torch.float32, device=dev)
    step = 4096
    for t0 in range(0, n_days, step):
        t1 = min(n_days, t0 + step)
        ph = (2 * np.pi * tau[t0:t1]).to(torch.float32)[:, None]
        gm = gmt[t0:t1].to(torch.float32)[:, None]
        sigma = sig0[None, :] * torch.exp(0.3 * torch.cos(ph + phi[None, :]))
        eps = torch.randn((t1 - t0, n_cells), generator=g, device=dev, dtype=torch.float32)
        y[t0:t1] = (mu0[None, :] + a1[None, :] * torch.cos(ph) + b1[None, :] * torch.sin(ph)
                    + a2[None, :] * torch.cos(2 * ph) + s[None, :] * gm + sigma * eps)
    n_bad = int(round(nan_fraction * n_cells))
    if n_bad:
        cols = torch.randperm(n_cells, generator=g, device=dev)[:n_bad]
        starts = torch.randint(1, n_days - 400, (n_bad,), generator=g, device=dev)
        for c, st in zip(cols.tolist(), starts.tolist()):
            y[st : st + 365, c] = float("nan")
    return y
