"""``ModelCuda`` behind the reference's REAL plugin call (no GPU: the batched solver is stubbed).

``attrici.variables.create_variable(name, data).create_model(ModelCuda, predictor, modes=4)`` - the call
``fit_and_detrend_cell`` makes (attrici/detrend.py:313, 326-331; variables.py:323-334, 406-419, 594-605, 631-642,
677-688, 735-746, 783-794) - is run with the unmodified reference classes (``baseline/_ref`` or ``/root/reference``,
xarray replaced by the 1-D stand-in of ``oracle/standins``) for every variable: distribution class identity, the
ordered parameter dict (theta layout, incl. Weibull's ``[alpha | beta]`` and ``wind`` / ``sfcWind``), what the Variable
hands over as ``observed`` (scaled, masked days removed - all days for pr), and the trace / distribution objects that
come back.  Also: the reference-side patch of ``integration/`` applies to the reference and adds the dispatch branch
and the CLI choice."""

import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import reference_runner

pytestmark = pytest.mark.skipif(not reference_runner.available(), reason="reference not installed (baseline/install_ref.sh)")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIABLES = ["tas", "ps", "rlds", "rsds", "tasskew", "pr", "tasrange", "hurs", "sfcWind", "wind"]
EXPECT = {  # variable -> (distribution class name, ordered parameters, dependent flags)
    "tas": ("Normal", ("mu", "sigma"), (True, False)), "ps": ("Normal", ("mu", "sigma"), (True, False)),
    "rlds": ("Normal", ("mu", "sigma"), (True, False)), "rsds": ("Normal", ("mu", "sigma"), (True, False)),
    "tasskew": ("Normal", ("mu", "sigma"), (True, False)),
    "pr": ("BernoulliGamma", ("p", "mu", "nu"), (True, True, False)),
    "tasrange": ("Gamma", ("mu", "nu"), (True, False)), "hurs": ("Beta", ("mu", "phi"), (True, False)),
    "sfcWind": ("Weibull", ("alpha", "beta"), (False, True)), "wind": ("Weibull", ("alpha", "beta"), (False, True)),
}


def _series(variable, n, rng):
    if variable in ("tas",):
        y = rng.normal(285, 6, n)
    elif variable == "ps":
        y = rng.normal(98000, 900, n)
    elif variable in ("rlds", "rsds"):
        y = rng.normal(250, 40, n)
    elif variable == "tasskew":
        y = np.clip(rng.normal(0.5, 0.15, n), 0, 1)
        y[[3, 17]] = [0.00005, 0.99995]                     # masked by the thresholds (variables.py:619-623)
    elif variable == "pr":
        y = np.where(rng.random(n) < 0.4, 0.0, rng.gamma(0.8, 4e-5, n))
    elif variable == "tasrange":
        y = rng.gamma(6, 1.5, n)
        y[5] = 0.005                                        # <= 0.01 -> masked (variables.py:722-732)
    elif variable == "hurs":
        y = np.clip(rng.beta(8, 3, n) * 100, 0, 100)
        y[[2, 9]] = [0.005, 99.995]
    else:
        y = rng.weibull(2.0, n) * 5
        y[7] = 0.001
    y = y.astype(np.float32)
    y[11] = np.nan
    return y


class _StubSolver:
    """Stands in for CellBatchSolver: remembers what it was given, returns results that encode the inputs."""

    last = None

    def __init__(self, spec, modes=4, device="cpu", use_fp64=False, zero_init=False):
        self.device, self.spec, self.modes = torch.device("cpu"), spec, modes
        self.n_params = spec.n_params(modes)
        type(self).last = self

    def set_predictor(self, days, gmt, n_cells):
        self.days, self.gmt = np.asarray(days).copy(), np.asarray(gmt).copy()

    def scale(self, y):
        self.y = y.numpy().copy()
        return y, torch.tensor([[0.0, 1.0]], dtype=torch.float64)

    def fit(self, ys):
        from attrici_b200.solver import FitResult

        return FitResult(torch.arange(self.n_params, dtype=torch.float64)[None, :], torch.tensor([-3.25], dtype=torch.float64),
                         torch.tensor([0], dtype=torch.int32), torch.tensor([5], dtype=torch.int32))

    def eval_parameter(self, params, which):
        return (100.0 * (which + 1) + torch.arange(len(self.days), dtype=torch.float64))[:, None]


@pytest.mark.parametrize("variable", VARIABLES)
def test_real_create_model_with_model_cuda(variable, monkeypatch):
    m = reference_runner._modules()
    xr, create_variable = m["xr"], m["create_variable"]
    import attrici.distributions as ref_distributions

    from attrici_b200 import _lib, estimation
    from attrici_b200.variables import create_variable as our_spec

    monkeypatch.setattr(estimation, "CellBatchSolver", _StubSolver)
    n = 400
    rng = np.random.default_rng(5)
    t = reference_runner.time_axis(np.arange(n))
    y = _series(variable, n, rng)
    data = xr.DataArray(y.copy(), coords={"time": t}, attrs={"units": reference_runner.UNITS[variable]})
    predictor = xr.DataArray(np.linspace(0.0, 1.0, n), coords={"time": t})
    window = t[20:380]                                        # a fit window inside the series (detrend.py:709-722)
    with np.errstate(all="ignore"):
        var = create_variable(variable, data)                 # the reference's class, unmodified
        model = var.create_model(estimation.ModelCuda, predictor.sel(time=window), modes=4, window_size=None)
    name, order, dependent = EXPECT[variable]
    assert isinstance(model, estimation.ModelCuda)
    assert model._distribution_class is getattr(ref_distributions, name)              # class identity, as ModelScipy
    spec = our_spec(variable)
    assert model._family == spec.family and tuple(spec.parameters) == order
    assert model._solver.n_params == spec.n_params(4) == (45 if name == "BernoulliGamma" else 27)
    assert model._solver.spec.scaling == _lib.SCALE_NONE                              # already scaled by the Variable
    # what the Variable handed over: y_scaled on the fit window, masked / missing days dropped (pr keeps all days)
    ys = np.asarray(var.y_scaled.values)[20:380]
    keep = np.ones(360, dtype=bool) if variable == "pr" else ~np.isnan(ys)
    np.testing.assert_array_equal(model._y, ys[keep].astype(np.float32))
    np.testing.assert_array_equal(model._days, (np.arange(20, 380)[keep] - np.arange(20, 380)[keep][0]).astype(np.int32))
    np.testing.assert_array_equal(model._gmt, np.linspace(0.0, 1.0, n)[20:380][keep])
    trace = model.fit_cached({"data": 1}, cache_dir=None, timeout=None)               # detrend.py:335-350
    assert set(trace) == {"logp", "params"} and trace["params"].shape == (spec.n_params(4),)
    assert model.estimate_logp(trace) == -3.25                                        # detrend.py:416
    dist = model.estimate_distribution(trace, predictor=predictor)                    # detrend.py:362-364
    assert type(dist) is getattr(ref_distributions, name)
    for i, field in enumerate(order):
        np.testing.assert_array_equal(getattr(dist, field), 100.0 * (i + 1) + np.arange(n))
    assert _StubSolver.last.days.tolist() == list(range(n))                           # phase origin = first day of the full axis
    # the reference's own quantile_mapping runs on the distributions ModelCuda returns
    cf = model.estimate_distribution(trace, predictor=predictor)
    with np.errstate(all="ignore"):
        out = var.quantile_mapping(dist, cf)
    assert np.asarray(out).shape == (n,)


def test_distribution_identity_is_required_for_reference_classes(monkeypatch):
    m = reference_runner._modules()
    import attrici.distributions as ref_distributions

    from attrici_b200 import estimation

    monkeypatch.setattr(estimation, "CellBatchSolver", _StubSolver)
    Fake = type("Normal", (), {"__module__": "attrici.estimation.model_scipy"})      # same name, not the class
    with pytest.raises(ValueError, match="not supported"):
        estimation._family_of(Fake)
    Sub = type("Normal", (ref_distributions.Normal,), {"__module__": "attrici.distributions"})
    with pytest.raises(ValueError, match="not supported"):                            # `is`, not issubclass
        estimation._family_of(Sub)
    assert estimation._family_of(ref_distributions.Weibull) == estimation._lib.WEIBULL
    assert m is not None


def test_reference_side_patch_applies(tmp_path):
    """integration/attrici_cuda_solver.patch against the unmodified reference: dispatch branch (detrend.py:724-739),
    batched branch in front of the cell loop (:760-788), CLI choice (commands/detrend.py:265-271)."""
    patch = os.path.join(ROOT, "integration", "attrici_cuda_solver.patch")
    if shutil.which("patch") is None:
        pytest.skip("no patch(1)")
    dst = tmp_path / "ref"
    shutil.copytree(os.path.join(reference_runner.reference_root(), "attrici"), dst / "attrici")
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=dst, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    det = (dst / "attrici" / "detrend.py").read_text()
    cli = (dst / "attrici" / "commands" / "detrend.py").read_text()
    assert 'elif config.solver == "cuda":' in det and "from attrici_b200.estimation import ModelCuda" in det
    assert det.index('elif config.solver == "cuda":') < det.index('raise ValueError(f"Unknown solver {config.solver}")')
    assert "cuda_detrend.detrend_cells(" in det and det.index("cuda_detrend.detrend_cells(") < det.index(
        "for i, latlon in enumerate(indices):")
    assert 'choices=["pymc5", "scipy", "pymc3", "cuda"]' in cli
    for f in ("attrici/detrend.py", "attrici/commands/detrend.py"):
        subprocess.run([sys.executable, "-m", "py_compile", str(dst / f)], check=True)
    # every attrici_b200 name the patch uses exists with the arguments it passes
    import inspect

    from attrici_b200 import detrend as d, io as aio

    for kw in ("modes", "window_size", "seed", "fit_only", "report_variables"):
        assert kw in d.Config.__dataclass_fields__
    sig = inspect.signature(d.detrend_cells).parameters
    assert all(k in sig for k in ("fit_window", "units", "trace", "gather"))
    sig = inspect.signature(aio.write_cell_files).parameters
    assert list(sig)[:7] == ["output_dir", "variable", "lat", "lon", "time", "series", "logp"]
    assert all(k in sig for k in ("params", "scaling", "shared", "overwrite", "attrs"))
    assert callable(d.fit_window_indices)
