"""SSA predictor smoothing on the GPU (csrc/ssa.cu through the C ABI) against the reference's outputs and the oracle."""
import os
import time

import numpy as np
import pytest
import torch

from oracle import ssa_oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "ssa_component.npz"))


def test_reference_fixture_full_length(gold):
    """20crv3-era5_gmt_raw.nc -> SSA (window 365, every 10th day): the reference's own golden pair."""
    from attrici_b200.preprocessing import ssa_first_component

    out, info = ssa_first_component(gold["raw_gmt10"], 365, return_info=True)
    np.testing.assert_allclose(out, gold["ref_f64"], rtol=1e-10)                      # reference, float64 input
    np.testing.assert_allclose(out, gold["ref_f32"], rtol=1e-5, atol=1e-6)            # reference as shipped (float32)
    fixture = np.load(os.path.join(GOLDEN, "ssa_gmt.npz"))
    np.testing.assert_allclose(out, fixture["tas"], rtol=1e-5, atol=1e-6)             # reference tests/test_ssa.py:44
    assert info["iterations"] < 100   # the mean dominates the spectrum


def test_synthetic_cases_against_reference_outputs(gold):
    from attrici_b200.preprocessing import ssa_first_component

    for i in range(int(gold["n_cases"])):
        x, w, ref = gold[f"case{i}_x"], int(gold[f"case{i}_w"]), gold[f"case{i}_out"]
        out = ssa_first_component(x, w)
        # case 6 is white noise: the two leading eigenvalues are a few per cent apart, which amplifies the
        # rounding differences between eigh and the iteration
        tol = 1e-7 if i == 6 else 1e-9
        np.testing.assert_allclose(out, ref, rtol=tol, atol=tol * np.abs(ref).max(), err_msg=f"case {i}")


def test_calc_gmt_by_ssa_signature_and_subset():
    """Reference tests/test_ssa.py:10-24 (shapes) plus values against the oracle."""
    from attrici_b200.preprocessing import calc_gmt_by_ssa

    times = np.arange(0, 365 * 10)
    rng = np.random.default_rng(1)
    gmt = rng.normal(np.linspace(15, 18, len(times)), 2)
    ssa, ssa_times = calc_gmt_by_ssa(gmt, times, window_size=365, subset=10)
    assert ssa.shape == (len(times[::10]),) and ssa_times.shape == (len(times[::10]),)
    ref, _ = ssa_oracle.calc_gmt_by_ssa(gmt, times, window_size=365, subset=10)
    np.testing.assert_allclose(ssa, ref, rtol=1e-10)


def test_daily_series_without_subsetting():
    """43 464 daily values, window 365 (the reference would need a 365 x 365 x 43 100 tensor = 46 GB)."""
    from attrici_b200.preprocessing import ssa_first_component

    n = 43464
    t = np.arange(n)
    rng = np.random.default_rng(2)
    x = 287.0 + 1.2 / (1 + np.exp(-(t - 30000) / 4000.0)) + 0.4 * np.sin(2 * np.pi * t / 365.25) + rng.normal(0, 0.15, n)
    xd = torch.as_tensor(x, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ssa_first_component(xd, 365)
    torch.cuda.synchronize()
    gpu_s = time.perf_counter() - t0
    ref = ssa_oracle.ssa_first_component(x, 365)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-10)
    assert out.is_cuda and gpu_s < 5.0


def test_errors_like_the_reference():
    from attrici_b200.preprocessing import ssa_first_component

    with pytest.raises(ValueError):
        ssa_first_component(np.arange(10.0), 11)
    with pytest.raises(ValueError):
        ssa_first_component(np.arange(10.0), 1)
    with pytest.raises(TypeError):
        ssa_first_component(np.arange(10.0), 0.5)
