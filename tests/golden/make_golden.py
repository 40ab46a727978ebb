"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # all 12 fixtures, 8 processes

For every golden target of the reference's own tests
(``/root/reference/tests/data/20CRv3-ERA5_germany_target_<var>_lat<lat>_lon<lon>.h5``, the
fixtures of reference ``tests/test_detrend_pymc.py:63-160``; ``stop_date=2021-12-31``,
``modes=4``, ``seed=0``) this script

1. extracts the observed series ``y`` (float32) and the PyMC-produced ``cfact`` target,
2. runs the reference SciPy path on ``y`` -- ``attrici.variables.create_variable`` ->
   ``create_model(ModelScipy)`` -> ``fit`` -> 2x ``estimate_distribution`` ->
   ``quantile_mapping`` -> ``rescale`` -> NaN/Inf replacement (the glue of
   ``attrici/detrend.py:285, 311-416`` restated here because detrend.py itself needs the
   real xarray/netCDF4) -- and
3. stores inputs, the reference trace and (decimated) reference outputs.

For pr / tasrange / hurs / sfcWind the reference's dependent-parameter ``estimate`` is
wrapped AT RUN TIME to apply its link (the reference file is untouched); see SURVEY.md
section 8c defect 1 and oracle/attrici_oracle.py's header.

The reference cannot travel to the GPU box, hence the committed .npz files.
"""

import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "standins"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

FIXTURES = [
    ("tas", 50.75, 9.25), ("tas", 50.75, 9.75), ("tas", 51.25, 9.25), ("tas", 51.25, 9.75),
    ("ps", 50.75, 9.25), ("rlds", 50.75, 9.25), ("rsds", 50.75, 9.25), ("tasskew", 50.75, 9.25),
    ("pr", 50.75, 9.25), ("tasrange", 50.75, 9.25), ("hurs", 50.75, 9.25), ("sfcWind", 50.75, 9.25),
]
UNITS = {"tas": "K", "ps": "Pa", "rlds": "W m-2", "rsds": "W m-2", "tasskew": "1", "pr": "kg m-2 s-1",
         "tasrange": "K", "hurs": "%", "sfcWind": "m s-1"}
LINK_CORRECTED = {"pr", "tasrange", "hurs", "sfcWind"}
STOP = np.datetime64("2021-12-31")
DECIMATE = 8


def run_reference(args):
    var, lat, lon = args
    import xarray as xr  # the stand-in
    from attrici.estimation import model_scipy
    from attrici.variables import create_variable

    from oracle.fixtures_io import read_ssa_gmt, read_target_h5

    if var in LINK_CORRECTED:
        cls = model_scipy.AttriciGLMScipy.PredictorDependentParam
        if not getattr(cls, "_link_wrapped", False):
            raw = cls.estimate

            def estimate(self, params):
                eta, logp = raw(self, params)
                return self.link(eta), logp

            cls.estimate = estimate
            cls._link_wrapped = True

    tgt = read_target_h5(f"{REF}/tests/data/20CRv3-ERA5_germany_target_{var}_lat{lat}_lon{lon}.h5")
    _, gmt_raw = read_ssa_gmt(f"{REF}/tests/data/20CRv3-ERA5_germany_ssa_gmt.nc")
    time_ = tgt["ds"]
    # detrend.py:698-707
    t_scaled = (time_ - time_.min()) / (time_.max() - time_.min())
    g = np.interp(t_scaled, np.linspace(0, 1, len(gmt_raw)), gmt_raw)
    gmt_scaled = (g - g.min()) / (g.max() - g.min())
    predictor = xr.DataArray(gmt_scaled, coords={"time": time_})
    subset_times = time_[(time_ >= time_[0]) & (time_ <= STOP)]

    t0 = time.time()
    with np.errstate(all="ignore"):
        np.random.seed((0 + int(lat * 1e9) + int(lon * 1e6)) % 2**32)  # detrend.py:285
        data = xr.DataArray(tgt["y"].copy(), coords={"time": time_}, attrs={"units": UNITS[var]})
        data[np.isinf(data)] = np.nan  # :311
        variable = create_variable(var, data)  # :313
        model = variable.create_model(model_scipy.ModelScipy, predictor.sel(time=subset_times), modes=4,
                                      window_size=None)  # :326-331
        f0 = -sum(d.log_likelihood(model._initial_params) for d in model._distributions)
        rng = np.random.default_rng(7)
        theta_probe = rng.normal(0, 0.05, len(model._initial_params))
        f_probe = -sum(d.log_likelihood(theta_probe) for d in model._distributions)
        trace = model.fit()  # :335-350 (cache/timeout wrapper is a pass-through)
        t_fit = time.time() - t0
        dist_ref = model.estimate_distribution(trace, predictor=predictor)  # :362-364
        pc = predictor.copy()
        pc[:] = 0
        dist_cf = model.estimate_distribution(trace, predictor=pc)  # :366-370
        cfact_scaled = variable.y_scaled.astype(np.float64)  # :392-394
        cfact_scaled[:] = variable.quantile_mapping(dist_ref, dist_cf)
        cfact = variable.rescale(cfact_scaled)
        replaced = np.zeros_like(cfact.values)  # :397-411
        for test, code in ((np.isnan, np.nan), (lambda c: c == np.inf, np.inf), (lambda c: c == -np.inf, -np.inf)):
            idx = test(cfact.values)
            cfact.values[idx] = data.values[idx]
            replaced[idx] = code
    sel = slice(None) if var == "pr" else slice(3, None, DECIMATE)
    out = {
        "variable": var, "lat": lat, "lon": lon,
        "y": tgt["y"],  # float32 input (before any in-place masking)
        # days the constructor masked IN PLACE in the caller's data (hurs / tasrange / sfcWind)
        "y_out_nan_idx": np.flatnonzero(np.isnan(np.asarray(data.values)) & ~np.isnan(tgt["y"])),
        "n_fit": np.int64(len(subset_times)),
        "y_scaled": np.asarray(variable.y_scaled.values),
        "scaling_keys": np.array(sorted(variable.scaling)),
        "scaling_vals": np.array([float(np.asarray(getattr(variable.scaling[k], "values", variable.scaling[k])))
                                  for k in sorted(variable.scaling)]),
        "sample_idx": np.arange(len(time_))[sel],
        "pymc_cfact": tgt["cfact"][sel],  # the reference tests' golden target
        "pymc_logp": tgt["logp"][0],
        "ref_params": np.asarray(trace["params"]), "ref_logp": np.float64(trace["logp"]),
        "ref_f0": np.float64(f0), "theta_probe": theta_probe, "ref_f_probe": np.float64(f_probe),
        "ref_cfact": np.asarray(cfact.values)[sel], "ref_cfact_scaled": np.asarray(cfact_scaled.values)[sel],
        "ref_replaced": replaced[sel], "ref_fit_seconds": np.float64(t_fit),
        "link_corrected": np.bool_(var in LINK_CORRECTED),
    }
    for k, v in dist_ref.__dict__.items():
        out[f"ref_{k}"] = np.asarray(getattr(v, "values", v))[sel]
    for k, v in dist_cf.__dict__.items():
        out[f"ref_{k}_cfact"] = np.asarray(getattr(v, "values", v))[sel]
    name = f"{var}_lat{lat}_lon{lon}.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    return name, t_fit, float(trace["logp"])


def main():
    from oracle.fixtures_io import read_ssa_gmt

    t, g = read_ssa_gmt(f"{REF}/tests/data/20CRv3-ERA5_germany_ssa_gmt.nc")
    np.savez_compressed(os.path.join(HERE, "ssa_gmt.npz"), time=t, tas=g)
    todo = [f for f in FIXTURES if len(sys.argv) < 2 or f[0] in sys.argv[1:]]
    with mp.Pool(min(8, len(todo))) as pool:
        for name, t_fit, logp in pool.imap_unordered(run_reference, todo):
            print(f"{name}: reference fit {t_fit:.1f} s, logp {logp:.6f}", flush=True)


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    main()
