"""Block bootstrap on the GPU (csrc/bootstrap.cu + attrici_b200/bootstrap.py) against NumPy's generator, the
reference's own resampling function (golden indices) and the oracle's restatement of detrend.py:479-612."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import helpers
from attrici_b200 import synthetic
from oracle import bootstrap_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _choices(seeds, n_draws, n_blocks, skip=0):
    from attrici_b200 import _lib

    lib = _lib.load()
    sd = torch.as_tensor(np.asarray(seeds, dtype=np.uint32).view(np.int32)).to(DEV)
    out = torch.empty((len(seeds), n_draws), dtype=torch.int16, device=DEV)
    _lib.check(lib.attrici_bootstrap_block_choices(_ptr(sd), len(seeds), n_draws, n_blocks, skip, _ptr(out),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.cpu().numpy()


@pytest.mark.parametrize("n_blocks", [1, 2, 3, 119, 128, 129, 1000])
def test_block_choices_are_numpy_legacy_randint(n_blocks):
    seeds = [0, 1, 12345, 2**32 - 1, 50750009250]   # the last one wraps like detrend.py:285
    seeds = [s % 2**32 for s in seeds]
    got = _choices(seeds, 1500, n_blocks)
    for i, s in enumerate(seeds):
        np.random.seed(s)
        want = np.array([np.random.randint(0, n_blocks) for _ in range(1500)])
        assert np.array_equal(got[i], want), (n_blocks, s)


def test_block_choices_after_skipped_draws():
    np.random.seed(77)
    np.random.rand(1001)   # 2002 generator words (what pr's quantile map consumes before its bootstrap)
    want = np.array([np.random.randint(0, 119) for _ in range(300)])
    assert np.array_equal(_choices([77], 300, 119, skip=2002)[0], want)


def test_resampled_days_match_the_reference_function():
    """Source day of every target day, against index vectors produced by the reference's own
    get_random_block_indices (tests/golden/make_golden_bootstrap.py): leap years, partial first / last year."""
    from attrici_b200 import _lib
    from attrici_b200.bootstrap import year_blocks
    from attrici_b200.solver import CellBatchSolver

    lib = _lib.load()
    g = np.load(os.path.join(GOLDEN, "bootstrap_blocks.npz"))
    for i in range(int(g["n_cases"])):
        years, seed, want = g[f"case{i}_years"], int(g[f"case{i}_seed"]), g[f"case{i}_indices"]
        n, T = want.shape
        start, bod = year_blocks(years)
        nb = len(start) - 1
        s = CellBatchSolver("tas", device=DEV)
        s.set_predictor(np.arange(T, dtype=np.int32), np.linspace(0, 1, T), 1)
        # mu = 0, sigma = 1 (all coefficients zero) and score(t) = t: the sample IS the source-day index
        params = torch.zeros((1, s.n_params), dtype=torch.float64, device=DEV)
        scores = torch.arange(T, dtype=torch.float64, device=DEV).reshape(T, 1).contiguous()
        ys = torch.full((T, 1), -1.0, dtype=torch.float32, device=DEV)
        choices = torch.as_tensor(_choices([seed], n * nb, nb)).to(DEV)
        y_new = torch.empty((T, n), dtype=torch.float32, device=DEV)
        rc = torch.zeros((T, 1), dtype=torch.int32, device=DEV)
        start_d, bod_d = torch.as_tensor(start).to(DEV), torch.as_tensor(bod).to(DEV)
        _lib.check(lib.attrici_bootstrap_resample(
            _ptr(s._ws), s._ws.numel(), C.byref(s.var), _ptr(ys), 1, T, 1, _ptr(params), _ptr(scores), 1, _ptr(choices), n,
            nb, 0, n, _ptr(start_d), _ptr(bod_d), _ptr(y_new), n, _ptr(rc), 1,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        got = y_new.cpu().numpy().T.astype(np.int64)
        assert np.array_equal(got, want), f"case {i}"
        assert int(rc.sum()) == 0


def test_bootstrap_statistics_against_oracle():
    from attrici_b200.bootstrap import bootstrap_cells
    from attrici_b200.solver import cell_seeds

    T, Cn, N = 6 * 365 + 2, 3, 8
    y = synthetic.tas_cells(Cn, T, seed=41)
    y[400:430, 1] = np.nan          # missing block: resampled onto other days, replaced where it lands on NaN days
    y[1000, 2] += 200.0             # z >> 8: its factual quantile is 1, every day that draws it is replaced
    gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
    years = (np.datetime64("1901-01-01") + days.astype("timedelta64[D]")).astype("datetime64[Y]").astype(int) + 1970
    lat, lon = np.array([50.75, 50.75, 51.25]), np.array([9.25, 9.75, 9.25])
    seeds = cell_seeds(0, lat, lon)
    out = bootstrap_cells("tas", y, days, gmt, years, seeds, N, device=DEV, return_samples=True)
    for j in range(Cn):
        ref = bootstrap_oracle.bootstrap_cell("tas", y[:, j], gmt, days, years, N, modes=helpers.MODES, seed=0,
                                              lat=lat[j], lon=lon[j], return_samples=True)
        np.testing.assert_array_equal(out["bootstrap_replaced"][:, j].cpu().numpy(), ref["bootstrap_replaced"])
        # six years barely identify the GMT slope (a flat direction of the posterior): equal optima to ~1e-9 in logp
        # still leave the fitted mean a few 1e-5 K apart, the same 1e-5-relative bar as for cfact elsewhere
        np.testing.assert_allclose(out["expected_values"][:, j].cpu().numpy(), ref["expected_values"], rtol=2e-6)
        # every sample's fitted expectation, then the statistics (the refit converges to the MAP the oracle's
        # Newton solver finds; the resampled series is stored in float32 here, float64 in the reference)
        sc = out["scaling"][j].cpu().numpy()
        samples = out["samples_scaled"][:, :, j].cpu().numpy().T * sc[1] + sc[0]
        np.testing.assert_allclose(samples, ref["samples"], rtol=2e-6)
        np.testing.assert_allclose(out["bootstrap_mean"][:, j].cpu().numpy(), ref["bootstrap_mean"], rtol=2e-6)
        np.testing.assert_allclose(out["bootstrap_quantiles"][:, :, j].cpu().numpy(), ref["bootstrap_quantiles"], rtol=2e-6)
        np.testing.assert_allclose(out["bootstrap_std"][:, j].cpu().numpy(), ref["bootstrap_std"], rtol=1e-3, atol=1e-6)
    assert int(out["bootstrap_replaced"][:, 2].sum()) > 0 and int(out["bootstrap_replaced"][:, 0].sum()) == 0


@pytest.mark.parametrize("variable", ["tasrange", "sfcWind", "hurs", "pr"])
def test_bootstrap_other_families_against_oracle(variable):
    """Gamma / Weibull / Beta: quantiles inverted per sample-day by the quantile map's own functions, the samples
    refitted by the day-ordered / phase-ordered kernels of those families."""
    from attrici_b200.bootstrap import bootstrap_cells
    from attrici_b200.solver import cell_seeds

    T, Cn, N = 4 * 365 + 1, 2, 4
    make = {"tasrange": synthetic.tasrange_cells, "sfcWind": synthetic.wind_cells, "hurs": synthetic.hurs_cells,
            "pr": synthetic.pr_cells}[variable]
    y = make(Cn, T)
    gmt, days = synthetic.smoothed_gmt(T), synthetic.time_axis(T)
    years = (np.datetime64("1901-01-01") + days.astype("timedelta64[D]")).astype("datetime64[Y]").astype(int) + 1970
    lat, lon = np.array([50.75, 51.25]), np.array([9.25, 9.75])
    out = bootstrap_cells(variable, y, days, gmt, years, cell_seeds(0, lat, lon), N, device=DEV, return_samples=True)
    for j in range(Cn):
        ref = bootstrap_oracle.bootstrap_cell(variable, y[:, j], gmt, days, years, N, modes=helpers.MODES, seed=0,
                                              lat=lat[j], lon=lon[j], return_samples=True)
        np.testing.assert_array_equal(out["bootstrap_replaced"][:, j].cpu().numpy(), ref["bootstrap_replaced"])
        # four years of pr (two likelihood terms, float32 day arithmetic, a third of the days dry) pin the optimum
        # less sharply than the other variables
        tol = 1e-3 if variable == "pr" else 1e-4
        np.testing.assert_allclose(out["expected_values"][:, j].cpu().numpy(), ref["expected_values"], rtol=tol)
        np.testing.assert_allclose(out["bootstrap_mean"][:, j].cpu().numpy(), ref["bootstrap_mean"], rtol=tol)
        np.testing.assert_allclose(out["bootstrap_quantiles"][:, :, j].cpu().numpy(), ref["bootstrap_quantiles"], rtol=tol)
        np.testing.assert_allclose(out["bootstrap_std"][:, j].cpu().numpy(), ref["bootstrap_std"], rtol=200 * tol,
                                   atol=1e-6 * float(np.abs(ref["bootstrap_std"]).max()))
