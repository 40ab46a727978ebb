#!/usr/bin/env python
"""Golden vectors for derive-huss, made by EXECUTING the reference's own function.

``attrici.commands.derive_huss`` imports xarray and netCDF4 at module level (absent here), so the source text of
``calc_huss_weedon2010`` (attrici/commands/derive_huss.py:27-99) is cut out of the file unmodified and executed with
``np`` in scope.   python tests/golden/make_golden_huss.py  ->  tests/golden/huss.npz"""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = "/root/reference/attrici/commands/derive_huss.py"


def reference_function():
    text = open(SRC).read()
    m = re.search(r"^def calc_huss_weedon2010\(.*?\n(?=^def )", text, re.M | re.S)
    ns = {"np": np}
    exec(m.group(0), ns)
    return ns["calc_huss_weedon2010"]


def main():
    f = reference_function()
    rng = np.random.default_rng(77)
    n = 20000
    hurs = rng.uniform(0.5, 100.0, n)
    ps = rng.uniform(50000.0, 105000.0, n)
    tas = rng.uniform(200.0, 320.0, n)
    # edges: exactly 0 degC, both clip limits, NaN in each input
    hurs[:8] = [50, 50, 1e-6, 100, np.nan, 50, 50, 100]
    ps[:8] = [101325, 101325, 101325, 30000, 101325, np.nan, 101325, 1000]
    tas[:8] = [273.15, 273.1499999, 220, 330, 280, 280, np.nan, 340]
    with np.errstate(all="ignore"):
        out = f(hurs, ps, tas)
    path = os.path.join(ROOT, "tests", "golden", "huss.npz")
    np.savez_compressed(path, hurs=hurs, ps=ps, tas=tas, huss=out)
    print("wrote", path, out[:8])


if __name__ == "__main__":
    main()
