"""Host side of the variable table: the units / bounds checks of ``attrici.variables`` (check_bounds :17-37,
check_units :134-156, the per-class calls at :319-320, 360-361, 502-503, 536-537, 571-572, 627-628, 673-674, 718-719,
766-767).  Where the reference is present (build container) the checks are compared with the reference's own
functions, run in a subprocess with the xarray stand-in; the expectations below hold everywhere."""

import json
import os
import subprocess
import sys

import numpy as np
import pytest

from attrici_b200.detrend import validate_bounds
from attrici_b200.variables import VARIABLES, check_bounds, check_units, create_variable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

EXPECTED = {  # variable: (units, lower bound, upper bound)
    "tas": ("K", 0.0, None), "pr": ("kg m-2 s-1", 0.0, None), "rlds": ("W m-2", 0.0, None), "ps": ("Pa", 0.0, None),
    "hurs": ("%", 0.0, 100.0), "tasskew": ({"1", "K"}, 0.0, 1.0), "rsds": ("W m-2", 0.0, None),
    "tasrange": ("K", 0.0, None), "sfcWind": ("m s-1", 0.0, None), "wind": ("m s-1", 0.0, None),
}


# the cases of the reference's own unit tests of check_bounds (tests/test_variables.py:7-52), as one table:
# data, lower, upper, raises
@pytest.mark.parametrize("data,lower,upper,raises", [
    ([1, 2, 3], None, None, False), ([1, 2, 3], 0, None, False), ([1, 2, 3], None, 4, False), ([1, 2, 3], 0, 4, False),
    ([1, 2, 3], 2, None, True), ([1, 2, 3], None, 2, True), ([1, 2, 3], 2, 2, True), ([], None, None, False),
    ([1], 0, 2, False),
])
def test_check_bounds_cases_of_the_reference_tests(data, lower, upper, raises):
    if raises:
        with pytest.raises(ValueError):
            check_bounds(np.array(data), lower=lower, upper=upper)
    else:
        check_bounds(np.array(data), lower=lower, upper=upper)


def test_table_of_units_and_bounds():
    assert set(VARIABLES) == set(EXPECTED)
    for name, (units, lo, hi) in EXPECTED.items():
        spec = create_variable(name)
        assert (set(spec.units) if not isinstance(spec.units, str) else spec.units) == units
        assert spec.lower_bound == lo and spec.upper_bound == hi


def test_check_units_messages():
    check_units(create_variable("tas"), "K")
    check_units(create_variable("tas"), None)          # no attribute: the reference only warns
    check_units(create_variable("tasskew"), "1")
    with pytest.raises(ValueError, match=r"Units of data are degC, but expected K\."):
        check_units(create_variable("tas"), "degC")
    with pytest.raises(ValueError, match=r"Units of data are %, but expected one of \{"):
        check_units(create_variable("tasskew"), "%")


def test_validate_bounds():
    hurs = create_variable("hurs")
    ok = np.array([[0.0, 50.0], [np.nan, 100.0]], dtype=np.float32)
    validate_bounds(hurs, ok)
    validate_bounds(hurs, np.full((3, 2), np.nan))      # all-NaN: nan < bound is false in the reference too
    validate_bounds(hurs, np.empty((4, 0)))
    with pytest.raises(ValueError) as e:
        validate_bounds(hurs, np.array([[1.0, -0.5], [np.nan, 3.0]]))
    assert e.value.args == (-0.5, "is smaller than lower bound", 0.0, ".")
    with pytest.raises(ValueError) as e:
        validate_bounds(hurs, np.array([[1.0, 100.5], [np.nan, 3.0]]))
    assert e.value.args == (100.5, "is bigger than upper bound", 100.0, ".")
    # infinite input values are NaN by the time the reference checks them (detrend.py:311) ...
    inf = np.array([[1.0, np.inf], [-np.inf, 3.0]])
    validate_bounds(hurs, inf, skip_inf=True)
    validate_bounds(hurs, np.array([[np.inf], [-np.inf]]), skip_inf=True)
    # ... but not in cfact, which is checked after the replacement step (detrend.py:392-411)
    with pytest.raises(ValueError):
        validate_bounds(hurs, inf)
    tas = create_variable("tas")
    validate_bounds(tas, np.array([[1e30]]))             # no upper bound
    import torch

    with pytest.raises(ValueError):
        validate_bounds(tas, torch.tensor([[1.0, -1.0]]))


_REFERENCE_PROBE = r"""
import json, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import xarray as xr
from attrici.variables import check_bounds, check_units
out = []
for units, expected in json.loads(sys.argv[3]):
    d = xr.DataArray(np.zeros(3, dtype=np.float32))
    d.attrs = {"units": units}
    try:
        check_units(d, set(expected) if isinstance(expected, list) else expected)
        out.append(None)
    except ValueError as e:
        out.append(str(e))
bounds = []
for values, lo, hi in json.loads(sys.argv[4]):
    d = xr.DataArray(np.array(values, dtype=np.float64))
    try:
        check_bounds(d, lo, hi)
        bounds.append(None)
    except ValueError as e:
        bounds.append([float(e.args[0])] + list(e.args[1:]))
print(json.dumps([out, bounds]))
"""


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "attrici")), reason="reference not present on this machine")
def test_checks_against_reference_functions():
    units_cases = [("K", "K"), ("degC", "K"), ("1", ["1", "K"]), ("%", ["1"])]
    bounds_cases = [([1.0, 2.0, float("nan")], 0.0, None), ([1.0, -2.0, float("nan")], 0.0, None),
                    ([1.0, 200.0], 0.0, 100.0), ([float("nan")] * 2, 0.0, 100.0)]
    probe = subprocess.run(
        [sys.executable, "-c", _REFERENCE_PROBE, REF, os.path.join(ROOT, "oracle", "standins"),
         json.dumps(units_cases), json.dumps(bounds_cases)], capture_output=True, text=True, timeout=300)
    if probe.returncode != 0:
        pytest.skip("reference functions not runnable with the stand-ins: " + probe.stderr[-300:])
    ref_units, ref_bounds = json.loads(probe.stdout.strip().splitlines()[-1])

    from attrici_b200.variables import VariableSpec

    for (units, expected), ref in zip(units_cases, ref_units):
        spec = VariableSpec("x", 0, 0, units=frozenset(expected) if isinstance(expected, list) else expected)
        try:
            check_units(spec, units)
            got = None
        except ValueError as e:
            got = str(e)
        assert got == ref
    for (values, lo, hi), ref in zip(bounds_cases, ref_bounds):
        spec = VariableSpec("x", 0, 0, lower_bound=lo, upper_bound=hi)
        try:
            validate_bounds(spec, np.array(values)[:, None])
            got = None
        except ValueError as e:
            got = [float(e.args[0])] + list(e.args[1:])
        assert got == ref
