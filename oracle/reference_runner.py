"""Runs the UNMODIFIED reference on one grid cell (TEST INFRASTRUCTURE ONLY).

Used by ``bench.py --impl reference`` / its ``cpu_baseline`` leg (``kind: "reference"``) and by the tests that
cross-check the oracle port against the reference itself.  Nothing in ``attrici_b200`` imports this module.

The reference package is taken from ``baseline/_ref`` (installed by ``baseline/install_ref.sh``, git-ignored, travels
to the GPU box) or, in the build container, from ``/root/reference``; ``xarray`` / ``func_timeout`` / ``tomlkit``,
which this image does not have, come from the import stand-ins of ``oracle/standins`` (1-D containers: the hot-path
modules use nothing else of them).  ``attrici/detrend.py`` itself needs the real xarray / netCDF4, so the glue of
``fit_and_detrend_cell`` (attrici/detrend.py:285, 311-416) is restated around the reference's own objects:

    create_variable -> Variable.create_model(ModelScipy) -> fit -> estimate_distribution x2
    -> quantile_mapping -> rescale -> NaN / Inf replacement -> estimate_logp

For pr / tasrange / hurs / sfcWind the reference's ``PredictorDependentParam.estimate`` is wrapped at run time to
apply its link (SURVEY.md section 8c defect 1; the file is untouched) - without it the shipped SciPy solver
returns theta = 0 and ``cfact == y`` for those variables.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINK_CORRECTED = {"pr", "tasrange", "hurs", "sfcWind", "wind"}
UNITS = {"tas": "K", "ps": "Pa", "rlds": "W m-2", "rsds": "W m-2", "tasskew": "1", "pr": "kg m-2 s-1",
         "tasrange": "K", "hurs": "%", "sfcWind": "m s-1", "wind": "m s-1"}


def reference_root():
    """Directory that holds the reference's ``attrici`` package, or None."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(p, "attrici", "estimation", "model_scipy.py")):
            return p
    return None


def available():
    return reference_root() is not None


_imported = {}


def _modules():
    if not _imported:
        ref = reference_root()
        if ref is None:
            raise ImportError("the reference is not installed (baseline/install_ref.sh)")
        for p in (ref, os.path.join(ROOT, "oracle", "standins")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import xarray as xr  # the stand-in
        from attrici.estimation import model_scipy
        from attrici.variables import create_variable

        _imported.update(xr=xr, model_scipy=model_scipy, create_variable=create_variable)
    return _imported


def _apply_link_correction(model_scipy):
    cls = model_scipy.AttriciGLMScipy.PredictorDependentParam
    if not getattr(cls, "_link_wrapped", False):
        raw = cls.estimate

        def estimate(self, params):
            eta, logp = raw(self, params)
            return self.link(eta), logp

        cls.estimate = estimate
        cls._link_wrapped = True


def time_axis(days, origin="1901-01-01"):
    """datetime64[ns] axis of integer day offsets (what the reference's time coordinate is)."""
    return (np.datetime64(origin, "ns") + np.asarray(days).astype("timedelta64[D]")).astype("datetime64[ns]")


def detrend_cell(variable, y, gmt_scaled, days, modes=4, lat=0.0, lon=0.0, seed=0, fit_window=None):
    """One cell through the reference's own classes.  ``y``: float32 observations, ``gmt_scaled``: predictor on the
    same axis, ``days``: integer day offsets.  Returns cfact, replaced (the reference's NaN / +-Inf float flags),
    y_scaled, the SciPy trace and the optimiser's evaluation count."""
    m = _modules()
    xr, model_scipy = m["xr"], m["model_scipy"]
    if variable in LINK_CORRECTED:
        _apply_link_correction(model_scipy)
    t = time_axis(days)
    predictor = xr.DataArray(np.asarray(gmt_scaled, dtype=np.float64), coords={"time": t})
    i0, i1 = (0, len(t)) if fit_window is None else fit_window
    subset_times = t[i0:i1]
    # count objective evaluations without touching the reference: scipy reports nfev on the OptimizeResult, which
    # ModelScipy.fit drops (model_scipy.py:524-527); wrap scipy.optimize.minimize as seen from that module
    nfev = {}
    real_minimize = model_scipy.minimize

    def counting_minimize(*a, **k):
        res = real_minimize(*a, **k)
        nfev["nfev"], nfev["nit"], nfev["status"] = int(res.nfev), int(res.nit), int(res.status)
        return res

    model_scipy.minimize = counting_minimize
    try:
        with np.errstate(all="ignore"):
            np.random.seed((seed + int(lat * 1e9) + int(lon * 1e6)) % 2**32)            # detrend.py:285
            data = xr.DataArray(np.array(y, dtype=np.float32), coords={"time": t}, attrs={"units": UNITS[variable]})
            data[np.isinf(data)] = np.nan                                                # :311
            var = m["create_variable"](variable, data)                                   # :313
            model = var.create_model(model_scipy.ModelScipy, predictor.sel(time=subset_times), modes=modes,
                                     window_size=None)                                   # :326-331
            trace = model.fit()                                                          # :335-350
            dist_ref = model.estimate_distribution(trace, predictor=predictor)           # :362-364
            pc = predictor.copy()
            pc[:] = 0
            dist_cf = model.estimate_distribution(trace, predictor=pc)                   # :366-370
            cfact_scaled = var.y_scaled.astype(np.float64)                               # :392-394
            cfact_scaled[:] = var.quantile_mapping(dist_ref, dist_cf)
            cfact = var.rescale(cfact_scaled)
            replaced = np.zeros_like(cfact.values)                                       # :397-411
            for test, code in ((np.isnan, np.nan), (lambda c: c == np.inf, np.inf), (lambda c: c == -np.inf, -np.inf)):
                idx = test(cfact.values)
                cfact.values[idx] = data.values[idx]
                replaced[idx] = code
            logp = model.estimate_logp(trace)                                            # :416
    finally:
        model_scipy.minimize = real_minimize
    return {"cfact": np.asarray(cfact.values), "replaced": replaced, "y_scaled": np.asarray(var.y_scaled.values),
            "cfact_scaled": np.asarray(cfact_scaled.values), "trace": {"params": np.asarray(trace["params"]),
                                                                        "logp": float(trace["logp"])},
            "logp": float(logp), "scaling": dict(var.scaling), **nfev}
