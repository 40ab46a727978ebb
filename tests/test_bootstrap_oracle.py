"""Block bootstrap: the oracle's restatement of detrend.py:479-612 against index vectors produced by the
reference's own get_random_block_indices (tests/golden/make_golden_bootstrap.py executes its source text)."""
import os

import numpy as np

from attrici_b200 import synthetic
from oracle import bootstrap_oracle as bo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_resampling_indices_match_the_reference_function():
    g = np.load(os.path.join(GOLDEN, "bootstrap_blocks.npz"))
    assert int(g["n_cases"]) == 3
    for i in range(int(g["n_cases"])):
        blocks = bo.year_blocks(g[f"case{i}_years"])
        np.random.seed(int(g[f"case{i}_seed"]))
        want = g[f"case{i}_indices"]
        got = np.asarray([bo.resample_indices(blocks) for _ in range(want.shape[0])])
        assert np.array_equal(got, want), f"case {i}"


def test_host_side_block_table_matches_the_oracle_blocks():
    # attrici_b200.bootstrap.year_blocks is host logic (no GPU): same blocks as detrend.py:491-498
    from attrici_b200.bootstrap import year_blocks

    g = np.load(os.path.join(GOLDEN, "bootstrap_blocks.npz"))
    years = g["case2_years"]           # partial first and last year
    start, block_of_day = year_blocks(years)
    blocks = bo.year_blocks(years)
    assert len(start) - 1 == len(blocks)
    for b, idx in enumerate(blocks):
        assert start[b] == idx[0] and start[b + 1] == idx[-1] + 1
        assert np.all(block_of_day[idx] == b)


def test_oracle_bootstrap_of_one_short_cell():
    T, N = 3 * 365 + 1, 3
    y = synthetic.tas_cells(1, T, seed=3)[:, 0]
    gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
    years = (np.datetime64("1901-01-01") + days.astype("timedelta64[D]")).astype("datetime64[Y]").astype(int) + 1970
    out = bo.bootstrap_cell("tas", y, gmt, days, years, N, seed=0, lat=50.75, lon=9.25, return_samples=True)
    assert out["samples"].shape == (N, T) and out["bootstrap_quantiles"].shape == (11, T)
    # quantile levels 0 / 0.5 / 1 are min / median / max of the samples (detrend.py:575)
    np.testing.assert_array_equal(out["bootstrap_quantiles"][0], out["samples"].min(axis=0))
    np.testing.assert_array_equal(out["bootstrap_quantiles"][-1], out["samples"].max(axis=0))
    np.testing.assert_allclose(out["bootstrap_quantiles"][5], np.median(out["samples"], axis=0), rtol=1e-15)
    assert out["bootstrap_replaced"].sum() == 0
    # a resampled series is a permutation of year blocks of standardised residuals: same length, finite
    assert np.isfinite(out["series"]).all()
