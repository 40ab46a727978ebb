"""CPU oracle for derive-huss (TEST INFRASTRUCTURE ONLY): NumPy restatement of ``calc_huss_weedon2010``
(attrici/commands/derive_huss.py:27-99).  Pinned by ``tests/golden/huss.npz``, which
``tests/golden/make_golden_huss.py`` produces by executing the reference function's own source text (the module
itself imports xarray / netCDF4, which this image lacks).  Only ``tests/`` import this module."""

import numpy as np


def calc_huss_weedon2010(hurs_in, ps_in, tas_in):
    hurs = np.asarray(hurs_in, dtype=np.float64) / 100          # :52
    ps = np.asarray(ps_in, dtype=np.float64) / 100              # :53
    tas = np.asarray(tas_in, dtype=np.float64) - 273.15         # :54
    water = tas >= 0
    # Buck (1981) e_w4 / e_i3 with f_w4 / f_i4 (:58-71), selected by np.where(tas >= 0, ...) (:82-86)
    a = np.where(water, 6.1121, 6.1115)
    b = np.where(water, 18.729, 23.036)
    c = np.where(water, 257.87, 279.82)
    d = np.where(water, 227.3, 333.7)
    x = np.where(water, 7.2e-4, 2.2e-4)
    y = np.where(water, 3.20e-6, 3.83e-6)
    z = np.where(water, 5.9e-10, 6.4e-10)
    with np.errstate(all="ignore"):
        svp = (a * np.exp((b - tas / d) * tas / (tas + c))) * (1.0 + x + ps * (y + z * tas**2))   # :74-80
        huss = hurs * 0.62198 / (ps / svp + 0.62198 - 1.0)                                         # :89-92
    return np.minimum(np.maximum(huss, 1e-7), 1e-1)                                                # :94-97
