"""SSA predictor smoothing: the O(n W) oracle against outputs of the unmodified reference (CPU).

Golden vectors: tests/golden/ssa_component.npz (made by tests/golden/make_golden_ssa.py, which imports
/root/reference's attrici.preprocessing.calc_gmt_by_ssa) and tests/golden/ssa_gmt.npz (the reference's own
fixture 20CRv3-ERA5_germany_ssa_gmt.nc, target of reference tests/test_ssa.py:27-44)."""
import os

import numpy as np
import pytest

from oracle import ssa_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ssa_component.npz"))


def test_oracle_matches_reference_on_its_own_fixture(gold):
    out = ssa_oracle.ssa_first_component(gold["raw_gmt10"], 365)
    np.testing.assert_allclose(out, gold["ref_f64"], rtol=1e-11)
    # as shipped the reference runs this in float32 (the file dtype): its own test tolerance
    np.testing.assert_allclose(out, gold["ref_f32"], rtol=1e-5, atol=1e-6)
    fixture = np.load(os.path.join(GOLDEN, "ssa_gmt.npz"))
    np.testing.assert_allclose(out, fixture["tas"], rtol=1e-5, atol=1e-6)   # reference tests/test_ssa.py:44


def test_oracle_matches_reference_on_synthetic_cases(gold):
    for i in range(int(gold["n_cases"])):
        x, w, ref = gold[f"case{i}_x"], int(gold[f"case{i}_w"]), gold[f"case{i}_out"]
        out = ssa_oracle.ssa_first_component(x, w)
        np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-11, err_msg=f"case {i} (n={x.size}, W={w})")


def test_subset_and_shapes_like_the_reference_test():
    # reference tests/test_ssa.py:10-24
    times = np.arange(0, 365 * 10)
    rng = np.random.default_rng(0)
    gmt = rng.normal(np.linspace(15, 18, len(times)), 2)
    ssa, ssa_times = ssa_oracle.calc_gmt_by_ssa(gmt, times, window_size=365, subset=10)
    assert ssa.shape == (len(times[::10]),) and ssa_times.shape == (len(times[::10]),)


def test_window_bounds():
    with pytest.raises(ValueError):
        ssa_oracle.ssa_first_component(np.arange(10.0), 11)
    with pytest.raises(ValueError):
        ssa_oracle.ssa_first_component(np.arange(10.0), 1)
