"""``ModelCuda``: the CUDA solver behind the reference's solver plugin interface.

Mirrors ``attrici.estimation.model.Model`` (attrici/estimation/model.py:51-116) and the constructor
keywords / trace schema of ``ModelScipy`` (attrici/estimation/model_scipy.py:379-570), so that
``Variable.create_model(ModelCuda, predictor, modes=...)`` works in the reference's per-cell loop
(attrici/detrend.py:326-416) and ``write_trace`` / ``--trace-file`` keep their format.  One cell per
instance is the reference's interface; production runs use ``attrici_b200.detrend.detrend_cells``,
which batches cells (the per-cell loop of attrici/detrend.py:760-788 is what this package replaces).

There is no CPU fallback: constructing a model without the CUDA library or a GPU raises.
"""

import numpy as np
import torch

from . import _lib
from .solver import CellBatchSolver
from .variables import FAMILY_PARAMETERS, VariableSpec

_FAMILY_OF_DISTRIBUTION = {
    "Normal": _lib.NORMAL,
    "Gamma": _lib.GAMMA,
    "Beta": _lib.BETA,
    "Weibull": _lib.WEIBULL,
    "BernoulliGamma": _lib.BERNOULLI_GAMMA,
}
# which parameters depend on the predictor (attrici/variables.py create_model tables)
_DEPENDENT = {
    _lib.NORMAL: (True, False),
    _lib.GAMMA: (True, False),
    _lib.BETA: (True, False),
    _lib.WEIBULL: (False, True),
    _lib.BERNOULLI_GAMMA: (True, True, False),
}


def _family_of(distribution):
    """Family of a distribution class.  ``ModelScipy`` compares the class with the classes of
    ``attrici.distributions`` by identity (model_scipy.py:423, 448, 461, 474, 487); so does this for a class that
    comes from that module (a subclass or a same-named class of another module of the reference is rejected like
    there).  Classes from elsewhere - the dataclasses of a host application that mirrors the reference - are matched
    by name."""
    import sys

    name = getattr(distribution, "__name__", None)
    if name not in _FAMILY_OF_DISTRIBUTION:
        raise ValueError(f"Distribution {distribution} not supported")
    module = getattr(distribution, "__module__", "")
    if module.startswith("attrici."):
        ref = sys.modules.get("attrici.distributions")
        if module != "attrici.distributions" or ref is None or getattr(ref, name, None) is not distribution:
            raise ValueError(f"Distribution {distribution} not supported")
    return _FAMILY_OF_DISTRIBUTION[name]


def _values(x):
    return np.asarray(getattr(x, "values", x))


def _day_offsets(time):
    """Integer day offsets of a datetime64 (or already integer) time vector."""
    t = _values(getattr(time, "time", time))
    if np.issubdtype(t.dtype, np.datetime64):
        d = (t - t.min()) / np.timedelta64(1, "D")
        days = np.rint(d).astype(np.int64)
        if not np.array_equal(days.astype(np.float64), d):
            raise ValueError("time axis must be at whole-day spacing")
    elif np.issubdtype(t.dtype, np.integer):
        days = t.astype(np.int64) - int(t.min())
    else:
        raise TypeError("time coordinate must be datetime64 or integer days")
    if days.max(initial=0) >= 2**31:
        raise ValueError("time axis too long")
    return days.astype(np.int32)


class ModelCuda:
    """Single-cell view of the batched CUDA solver (``solver="cuda"``)."""

    def __init__(self, distribution, parameters, observed, predictor, modes=None, window_size=None,
                 device="cuda:0", use_fp64=False, zero_init=False):
        if (modes is None) == (window_size is None):
            raise ValueError("Exactly one of `modes` and `window_size` must be set")   # detrend.py:646-647
        family = _family_of(distribution)
        self._window_size = None
        if window_size is not None:
            # the reference's second model family (model_pymc5.py:366-547); built for the Normal family, the others
            # raise like ModelScipy does for every family (model_scipy.py:58-59)
            if family != _lib.NORMAL:
                raise NotImplementedError("Rolling window is implemented for the Normal family only in the CUDA solver")
            if int(window_size) % 2 == 0:
                raise ValueError("Window size must be an odd number")                  # detrend.py:648-649
            self._window_size = int(window_size)
        name = distribution.__name__
        expect = FAMILY_PARAMETERS[family]
        if tuple(parameters) != expect:
            raise ValueError(f"{name} expects parameters {expect} in this order, got {tuple(parameters)}")
        for (pname, par), dep in zip(parameters.items(), _DEPENDENT[family]):
            if bool(getattr(par, "dependent", dep)) != dep:
                raise ValueError(f"parameter {pname} of {name}: dependent={not dep} is not a model of the reference")
        self._distribution_class = distribution
        self._family = family
        self._modes = int(modes) if modes is not None else 1
        self._time = _values(getattr(predictor, "time", np.zeros(0)))
        y = _values(observed).astype(np.float32)
        gmt = _values(predictor).astype(np.float64)
        if y.shape != gmt.shape or y.ndim != 1:
            raise ValueError("observed and predictor must be 1-D and aligned")
        self._days = _day_offsets(predictor)
        self._y = y
        self._gmt = gmt
        # the series handed over is already scaled and masked by the Variable class: identity scaling
        spec = VariableSpec("cell", family, _lib.SCALE_NONE)
        self._solver = CellBatchSolver(spec, modes=self._modes, device=device, use_fp64=use_fp64, zero_init=zero_init)

    def _window_solver(self):
        from .window import WindowBatchSolver

        if not np.issubdtype(self._time.dtype, np.datetime64):
            raise TypeError("the rolling-window model needs a datetime64 time coordinate (day of the year)")
        return WindowBatchSolver(self._solver.spec, self._window_size, device=self._solver.device)

    def fit(self, **kwargs):
        """Returns the SciPy trace schema ``{"logp": float, "params": ndarray[P]}`` (model_scipy.py:524-527); with
        ``window_size`` the PyMC5 trace schema of the rolling-window model (``weights_<par>_longterm_intercept`` /
        ``_trend`` of length 366 and ``logp``, model_pymc5.py:403-414, 686-689)."""
        if self._window_size is not None:
            from .window import trace_of

            w = self._window_solver()
            ys = torch.from_numpy(np.ascontiguousarray(self._y.reshape(-1, 1))).to(w.device)
            params, logp, status, n_iter = w.fit(ys, self._time, self._gmt)
            self._status, self._n_pass = int(status.item()), int(n_iter.item())
            return trace_of(params[0].cpu().numpy(), logp.item())
        s = self._solver
        s.set_predictor(self._days, self._gmt, 1)
        y = torch.from_numpy(np.ascontiguousarray(self._y.reshape(-1, 1))).to(s.device)
        ys, _ = s.scale(y)
        fit = s.fit(ys)
        self._status = int(fit.status.item())
        self._n_pass = int(fit.n_pass.item())
        return {"logp": float(fit.logp.item()), "params": fit.params[0].cpu().numpy()}

    def estimate_logp(self, trace, **kwargs):
        return trace["logp"]

    def estimate_distribution(self, trace, predictor, **kwargs):
        """Per-day distribution parameters on ``predictor``'s time axis (phase origin = its first day, as
        model_scipy.py:565-567 re-sets the predictor data) wrapped in the distribution class given to the
        constructor."""
        if self._window_size is not None:
            # estimate() of the rolling-window parameter classes (model_pymc5.py:435-447, 524-538): weights by day of year
            from .window import TRACE_KEYS, dayofyear

            k = dayofyear(_values(predictor.time)).astype(np.int64) - 1
            g = _values(predictor).astype(np.float64)
            a, b, c = (np.asarray(trace[key], dtype=np.float64) for key in TRACE_KEYS)
            return self._distribution_class(mu=a[k] + b[k] * g, sigma=np.exp(c[k]))
        s = self._solver
        gmt = _values(predictor).astype(np.float64)
        s.set_predictor(_day_offsets(predictor), gmt, 1)
        params = torch.as_tensor(np.asarray(trace["params"], dtype=np.float64)[None, :]).to(s.device)
        out = {}
        for i, pname in enumerate(FAMILY_PARAMETERS[self._family]):
            out[pname] = s.eval_parameter(params, i)[:, 0].cpu().numpy()
        return self._distribution_class(**out)

    def fit_cached(self, defining_data_inputs=None, cache_dir=None, timeout=None, **kwargs):
        """``Model.fit_cached`` (attrici/estimation/model.py:118-188), which is what ``fit_and_detrend_cell`` calls
        (attrici/detrend.py:335-350): with a ``cache_dir`` the trace is stored in / taken from a joblib cache keyed
        by ``defining_data_inputs``; with a ``timeout`` (seconds) ``None`` is returned if the fit takes longer (the
        reference then skips the cell).  A fit takes milliseconds here, so both matter for interface parity only."""
        def cached_fit():
            if cache_dir is None:
                return self.fit(**kwargs)
            from joblib import Memory

            def _fit(inputs):  # `inputs` only defines for which data the cached trace is valid
                return self.fit(**kwargs)

            return Memory(cache_dir, verbose=0).cache(_fit)(defining_data_inputs)

        if timeout is None:
            return cached_fit()
        import threading

        # a daemon thread (like func_timeout's): a fit that never returns must not keep the process alive at exit
        box = {}

        def run():
            try:
                box["result"] = cached_fit()
            except BaseException as e:  # noqa: BLE001 - re-raised in the caller's thread
                box["error"] = e

        worker = threading.Thread(target=run, daemon=True)
        worker.start()
        worker.join(timeout)
        if worker.is_alive():
            return None
        if "error" in box:
            raise box["error"]
        return box["result"]
