#!/usr/bin/env python
"""Golden vectors for the SSA smoothing of the predictor, made with the UNMODIFIED reference.

Run in the build container (needs /root/reference; its ``attrici.preprocessing`` and the vendored pyts class
import without any stand-in):

    python tests/golden/make_golden_ssa.py

Writes ``tests/golden/ssa_component.npz``:
  * ``raw_gmt10``      every 10th day of the reference's own fixture ``tests/data/20crv3-era5_gmt_raw.nc``
                       (float32, 4 493 values; read with oracle/fixtures_io.read_raw_gmt),
  * ``ref_f32``        ``calc_gmt_by_ssa(raw, times, 365, 10)[0]`` as shipped (float32 input => the reference's
                       matmul / eigh run in float32; this is what produced ``20CRv3-ERA5_germany_ssa_gmt.nc``),
  * ``ref_f64``        the same call on the series cast to float64,
  * ``case<i>_x / _w / _out``  small seeded synthetic series (trend + cycle + noise, pure noise, window >= number
                       of windows, window = n) and the reference's output in float64.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from attrici.preprocessing import calc_gmt_by_ssa  # noqa: E402  (the reference)

from oracle.fixtures_io import read_raw_gmt  # noqa: E402


def main():
    out = {}
    raw = read_raw_gmt("/root/reference/tests/data/20crv3-era5_gmt_raw.nc")
    t = np.arange(raw.size, dtype=np.float64)
    out["raw_gmt10"] = raw[::10].copy()
    out["ref_f32"] = calc_gmt_by_ssa(raw, t, window_size=365, subset=10)[0]
    out["ref_f64"] = calc_gmt_by_ssa(raw.astype(np.float64), t, window_size=365, subset=10)[0]
    rng = np.random.default_rng(20240611)
    cases = []
    for n, w in ((1500, 120), (800, 365), (730, 365), (365, 365), (400, 2), (900, 60)):
        tt = np.arange(n)
        x = 14.0 + 1.2 * (tt / n) ** 2 + 0.3 * np.sin(2 * np.pi * tt / 36.525) + rng.normal(0, 0.2, n)
        cases.append((x, w))
    cases.append((rng.normal(0, 1, 500), 50))   # no dominant component: small spectral gap
    for i, (x, w) in enumerate(cases):
        out[f"case{i}_x"] = x
        out[f"case{i}_w"] = np.int64(w)
        out[f"case{i}_out"] = calc_gmt_by_ssa(x, np.arange(x.size), window_size=w, subset=1)[0]
    out["n_cases"] = np.int64(len(cases))
    path = os.path.join(ROOT, "tests", "golden", "ssa_component.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
