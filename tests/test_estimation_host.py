"""Host side of ``ModelCuda`` (attrici_b200/estimation.py) without a GPU: the checks made before any CUDA call -
the reference's ``Model`` constructor contract (attrici/estimation/model.py:51-116, model_scipy.py:58-59, 379-503) -
and the data flow of ``fit`` / ``estimate_logp`` / ``estimate_distribution`` / ``fit_cached`` with the batched solver
replaced by a stub."""

from dataclasses import dataclass

import numpy as np
import pytest
import torch

from attrici_b200 import estimation
from attrici_b200.estimation import ModelCuda, _day_offsets
from attrici_b200.solver import FitResult


@dataclass
class Normal:  # stands for attrici.distributions.Normal (dataclass with these fields, distributions.py:64-90)
    mu: np.ndarray
    sigma: np.ndarray


@dataclass
class BernoulliGamma:  # distributions.py:93-129
    p: np.ndarray
    mu: np.ndarray
    nu: np.ndarray


@dataclass
class Par:  # attrici.estimation.model.Parameter: link + dependent flag
    link: object
    dependent: bool


class Arr:  # the part of a 1-D xarray.DataArray the constructor reads
    def __init__(self, values, time):
        self.values = values
        self.time = type("T", (), {"values": time})()


TIME = np.datetime64("1901-01-01") + np.arange(50).astype("timedelta64[D]")
NORMAL_PARS = {"mu": Par(None, True), "sigma": Par(np.exp, False)}


def test_day_offsets():
    assert _day_offsets(Arr(None, TIME)).tolist() == list(range(50))
    gapped = np.delete(TIME, [3, 4])
    assert _day_offsets(Arr(None, gapped))[:5].tolist() == [0, 1, 2, 5, 6] and _day_offsets(Arr(None, gapped)).dtype == np.int32
    assert _day_offsets(np.array([10, 11, 15])).tolist() == [0, 1, 5]                # integer days, any origin
    with pytest.raises(ValueError, match="whole-day"):
        _day_offsets(Arr(None, TIME + np.timedelta64(6, "h") * (np.arange(50) % 2)))
    with pytest.raises(TypeError):
        _day_offsets(np.linspace(0, 1, 5))


def test_constructor_rejects_what_the_reference_rejects():
    obs, pred = Arr(np.zeros(50, np.float32), TIME), Arr(np.linspace(0, 1, 50), TIME)
    with pytest.raises(ValueError, match="Exactly one of"):                           # detrend.py:646-647
        ModelCuda(Normal, NORMAL_PARS, obs, pred, modes=4, window_size=31)
    with pytest.raises(ValueError, match="Exactly one of"):
        ModelCuda(Normal, NORMAL_PARS, obs, pred)
    with pytest.raises(ValueError, match="odd"):                                      # detrend.py:648-649
        ModelCuda(Normal, NORMAL_PARS, obs, pred, window_size=30)
    with pytest.raises(ValueError, match="not supported"):                            # unknown distribution class
        ModelCuda(dict, NORMAL_PARS, obs, pred, modes=4)
    with pytest.raises(ValueError, match="in this order"):                            # theta layout = dict order
        ModelCuda(Normal, {"sigma": Par(np.exp, False), "mu": Par(None, True)}, obs, pred, modes=4)
    with pytest.raises(ValueError, match="dependent"):                                # not a model of the reference
        ModelCuda(Normal, {"mu": Par(None, False), "sigma": Par(np.exp, False)}, obs, pred, modes=4)


class _Stub:
    """Stands in for CellBatchSolver: one cell, results that encode their inputs."""

    def __init__(self, spec, modes=4, device="cpu", use_fp64=False, zero_init=False):
        self.device, self.spec, self.modes, self.calls = torch.device("cpu"), spec, modes, []
        self.n_params = 27 if len(spec.parameters) == 2 else 45

    def set_predictor(self, days, gmt, n_cells):
        self.calls.append(("predictor", np.asarray(days).copy(), np.asarray(gmt).copy(), n_cells))
        self._T = len(days)

    def scale(self, y):
        self.calls.append(("scale", tuple(y.shape), y.dtype))
        return y, torch.tensor([[0.0, 1.0]], dtype=torch.float64)

    def fit(self, ys):
        return FitResult(torch.arange(self.n_params, dtype=torch.float64)[None, :], torch.tensor([-12.5], dtype=torch.float64),
                         torch.tensor([0], dtype=torch.int32), torch.tensor([6], dtype=torch.int32))

    def eval_parameter(self, params, which):
        return (params[0, 0] + 10.0 * (which + 1) + torch.arange(self._T, dtype=torch.float64))[:, None]


def test_fit_and_estimate_flow_with_stubbed_solver(monkeypatch):
    monkeypatch.setattr(estimation, "CellBatchSolver", _Stub)
    y = np.random.default_rng(0).random(50)
    obs, pred = Arr(y, TIME), Arr(np.linspace(0, 1, 50), TIME)
    m = ModelCuda(Normal, NORMAL_PARS, obs, pred, modes=4)
    assert m._solver.spec.scaling == estimation._lib.SCALE_NONE                       # already scaled by the Variable class
    trace = m.fit()
    assert set(trace) == {"logp", "params"} and trace["logp"] == -12.5               # model_scipy.py:524-527
    assert isinstance(trace["params"], np.ndarray) and trace["params"].shape == (27,)
    assert m.estimate_logp(trace) == -12.5
    cached = m.fit_cached({"data": 1}, cache_dir=None, timeout=5)                     # model.py:118-188: same trace
    assert cached["logp"] == trace["logp"] and np.array_equal(cached["params"], trace["params"])
    kinds = [c[0] for c in m._solver.calls]
    assert kinds[:2] == ["predictor", "scale"] and m._solver.calls[1][1:] == ((50, 1), torch.float32)
    # estimate_distribution: predictor of the FULL series (phase origin = its first day), distribution class kept
    full_time = np.datetime64("1900-12-22") + np.arange(70).astype("timedelta64[D]")
    d = m.estimate_distribution(trace, Arr(np.linspace(0, 2, 70), full_time))
    assert isinstance(d, Normal) and d.mu.shape == (70,) and d.sigma.shape == (70,)
    np.testing.assert_array_equal(d.mu, 10.0 + np.arange(70))
    np.testing.assert_array_equal(d.sigma, 20.0 + np.arange(70))
    last = m._solver.calls[-1]
    assert last[0] == "predictor" and last[1].tolist() == list(range(70)) and last[3] == 1
    # Bernoulli-Gamma: three fields in the dataclass order p, mu, nu
    pars = {"p": Par(None, True), "mu": Par(np.exp, True), "nu": Par(np.exp, False)}
    m = ModelCuda(BernoulliGamma, pars, obs, pred, modes=4)
    d = m.estimate_distribution({"params": np.zeros(45), "logp": 0.0}, pred)
    assert isinstance(d, BernoulliGamma) and d.nu[0] == 30.0
    with pytest.raises(ValueError, match="1-D and aligned"):
        ModelCuda(Normal, NORMAL_PARS, Arr(y[:40], TIME[:40]), pred, modes=4)


def test_fit_cached_cache_and_timeout(monkeypatch, tmp_path):
    """attrici/estimation/model.py:118-188: joblib cache keyed by the defining inputs, ``None`` on timeout."""
    import time

    fits = []

    class Counting(_Stub):
        def fit(self, ys):
            fits.append(1)
            return super().fit(ys)

    monkeypatch.setattr(estimation, "CellBatchSolver", Counting)
    obs, pred = Arr(np.zeros(50), TIME), Arr(np.linspace(0, 1, 50), TIME)
    m = ModelCuda(Normal, NORMAL_PARS, obs, pred, modes=4)
    key = {"data": np.arange(5), "modes": 4, "seed": 0, "solver": "cuda"}
    a = m.fit_cached(key, cache_dir=str(tmp_path), timeout=30)
    b = m.fit_cached(key, cache_dir=str(tmp_path), timeout=30)                        # served from the cache
    assert len(fits) == 1 and a["logp"] == b["logp"] and np.array_equal(a["params"], b["params"])
    m.fit_cached({**key, "seed": 1}, cache_dir=str(tmp_path))                         # other inputs: a new fit
    assert len(fits) == 2
    m.fit_cached(key)                                                                  # no cache directory: always fits
    assert len(fits) == 3

    class Slow(_Stub):
        def fit(self, ys):
            time.sleep(1.0)
            return super().fit(ys)

    monkeypatch.setattr(estimation, "CellBatchSolver", Slow)
    m = ModelCuda(Normal, NORMAL_PARS, obs, pred, modes=4)
    t0 = time.perf_counter()
    assert m.fit_cached(key, timeout=0.1) is None                                     # model.py:185-187
    assert time.perf_counter() - t0 < 0.9
    assert m.fit_cached(key, timeout=10)["logp"] == -12.5


def test_fit_cached_passes_errors_through(monkeypatch):
    class Failing(_Stub):
        def fit(self, ys):
            raise RuntimeError("boom")

    monkeypatch.setattr(estimation, "CellBatchSolver", Failing)
    m = ModelCuda(Normal, NORMAL_PARS, Arr(np.zeros(50), TIME), Arr(np.linspace(0, 1, 50), TIME), modes=4)
    with pytest.raises(RuntimeError, match="boom"):
        m.fit_cached({"data": 1}, timeout=5)
    with pytest.raises(RuntimeError, match="boom"):
        m.fit_cached({"data": 1})
