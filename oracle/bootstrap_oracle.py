"""CPU oracle for the block bootstrap of the fitted trend (TEST INFRASTRUCTURE ONLY).

NumPy restatement of the bootstrap section of ``fit_and_detrend_cell`` (attrici/detrend.py:479-630) for one cell:
year blocks (:491-498), ``get_random_block_indices`` (:500-531), the resampling of the factual quantiles and
their inversion through the factual distribution of the target day (:533-557), the refit (:559-572) and the
statistics of the refitted expectations (:574-612).  The random block choices come from NumPy's legacy global
generator exactly as in the reference (seeded per cell at detrend.py:285).

``random_block_indices`` is pinned by ``tests/golden/bootstrap_blocks.npz``, which
``tests/golden/make_golden_bootstrap.py`` produces by executing the reference's own nested function (its source
text, unmodified).  The refit uses the oracle's exact MAP solver (``attrici_oracle.fit_newton``), whose optimum is
what the reference's L-BFGS-B approximates.

Only ``tests/`` import this module; the product never does.
"""

import numpy as np
from scipy import stats

from . import attrici_oracle as ao

QUANTILES = [0, 0.01, 0.025, 0.05, 0.25, 0.5, 0.75, 0.95, 0.975, 0.99, 1]   # detrend.py:575


def year_blocks(years):
    """detrend.py:491-498: one index array per calendar year, in year order."""
    years = np.asarray(years)
    return [np.flatnonzero(years == y) for y in np.unique(years)]


def random_block_indices(blocks, target_length):
    """detrend.py:500-531 (draws one ``np.random.randint`` from the global legacy generator)."""
    block = np.random.randint(0, len(blocks))
    idx = blocks[block]
    if len(idx) < target_length:
        missing = target_length - len(idx)
        if block >= len(blocks) - 1:
            idx = np.concatenate((blocks[block - 1][-missing:], idx))
        else:
            idx = np.concatenate((idx, blocks[block + 1][:missing]))
    return idx[:target_length]


def resample_indices(blocks):
    """detrend.py:541-546: source day of every target day for one bootstrap sample."""
    return np.concatenate([random_block_indices(blocks, len(b)) for b in blocks])


def bootstrap_cell(variable, y, gmt_scaled, days, years, n_samples, modes=4, seed=0, lat=0.0, lon=0.0,
                   return_samples=False):
    """detrend.py:479-612 for one cell (full time series, as the reference demands at :480-486)."""
    family = ao.VARIABLES[variable][0]
    np.random.seed(ao.cell_seed(seed, lat, lon))
    if family == "bernoulli_gamma":
        np.random.rand(len(y))   # consumed by Pr.quantile_mapping before the bootstrap starts (variables.py:458)
    ys, scaling, _ = ao.scale_variable(variable, y)
    prob = ao.build_problem(family, ys, gmt_scaled, days, modes)
    trace = ao.fit_newton(prob)
    ref = ao.estimate_params(family, trace["params"], gmt_scaled, days, modes)
    # Distribution.expectation(), attrici/distributions.py:89-90, 157-158, 186-187, 215-216
    expectation = {"normal": lambda p: p["mu"], "gamma": lambda p: p["mu"], "beta": lambda p: p["mu"],
                   "weibull": lambda p: stats.weibull_min.mean(p["alpha"], scale=p["beta"]),
                   "bernoulli_gamma": lambda p: (1 - p["p"]) * p["mu"]}[family]
    blocks = year_blocks(years)
    ys64 = np.asarray(ys, dtype=np.float64)
    with np.errstate(all="ignore"):
        q0 = ao.cdf(family, ref, ys64)
    replaced = np.zeros(len(ys64))
    samples, series = [], []
    for _ in range(n_samples):
        src = resample_indices(blocks)
        with np.errstate(all="ignore"):
            y_new = ao.invcdf(family, ref, q0[src])
        invalid = np.isnan(y_new) | np.isinf(y_new)
        replaced[invalid] += 1
        y_new[invalid] = ys64[invalid]
        tr = ao.fit_newton(ao.build_problem(family, y_new, gmt_scaled, days, modes))
        p = ao.estimate_params(family, tr["params"], gmt_scaled, days, modes)
        samples.append(ao.rescale(variable, expectation(p), scaling))
        if return_samples:
            series.append(y_new)
    samples = np.asarray(samples)
    out = {
        "expected_values": ao.rescale(variable, expectation(ref), scaling),
        "bootstrap_std": np.std(samples, axis=0),
        "bootstrap_mean": np.mean(samples, axis=0),
        "bootstrap_quantiles": np.asarray([np.percentile(samples, q * 100, axis=0) for q in QUANTILES]),
        "bootstrap_replaced": replaced,
    }
    if return_samples:
        out["samples"] = samples
        out["series"] = np.asarray(series)
    return out
