"""Merged output / trace files (attrici_b200/io.py): the merge-output layout, read back with SciPy's NetCDF reader."""
import numpy as np
import pytest
from scipy.io import netcdf_file

from attrici_b200 import io as aio


def _cells():
    lat = np.array([50.75, 50.75, 51.25])        # (51.25, 9.75) has no cell
    lon = np.array([9.25, 9.75, 9.25])
    return lat, lon


def test_merged_output_layout(tmp_path):
    lat, lon = _cells()
    T, C = 40, 3
    rng = np.random.default_rng(0)
    y = rng.normal(280, 5, (T, C)).astype(np.float32)
    cfact = y.astype(np.float64) - 0.3
    replaced = (rng.random((T, C)) < 0.1).astype(np.uint8)
    logp = rng.normal(size=C)
    gmt = np.linspace(0, 1, T)
    path = str(tmp_path / "merged.nc")
    aio.write_merged_output(path, {"y": y, "cfact": cfact, "replaced": replaced}, lat, lon, np.arange(T, dtype=np.int32),
                            scalars={"logp": logp}, shared={"gmt_scaled": gmt}, attrs={"attrici_config": "variable = 'tas'"})
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.dimensions == {"lat": 2, "lon": 2, "time": T}
    assert nc.variables["cfact"].dimensions == ("time", "lat", "lon")        # merge_output.py:79-86
    assert nc.variables["logp"].dimensions == ("lat", "lon")
    assert nc.variables["lat"].data.dtype.kind == "f" and nc.variables["lat"].data.dtype.itemsize == 4
    np.testing.assert_array_equal(nc.variables["lat"][:], np.float32([50.75, 51.25]))
    np.testing.assert_array_equal(nc.variables["lon"][:], np.float32([9.25, 9.75]))
    assert nc.variables["time"].units.decode().startswith("days since")
    full = np.array(nc.variables["cfact"][:])
    np.testing.assert_array_equal(full[:, 0, 0], cfact[:, 0])
    np.testing.assert_array_equal(full[:, 0, 1], cfact[:, 1])
    np.testing.assert_array_equal(full[:, 1, 0], cfact[:, 2])
    assert np.isnan(full[:, 1, 1]).all()                                       # grid point without a cell
    np.testing.assert_array_equal(np.array(nc.variables["y"][:])[:, 1, 0], y[:, 2])
    rep = np.array(nc.variables["replaced"][:])
    np.testing.assert_array_equal(rep[:, 0, 1], replaced[:, 1])
    assert (rep[:, 1, 1] == 0).all() and nc.variables["replaced"]._FillValue == 0   # detrend.py:475
    np.testing.assert_array_equal(np.array(nc.variables["gmt_scaled"][:]), gmt)
    assert nc.attrici_config == b"variable = 'tas'"
    nc.close()


def test_merged_trace_round_trip(tmp_path):
    lat, lon = _cells()
    rng = np.random.default_rng(1)
    params, logp = rng.normal(size=(3, 27)), rng.normal(size=3)
    scaling = np.column_stack([rng.uniform(250, 260, 3), rng.uniform(40, 60, 3)])
    path = str(tmp_path / "trace.nc")
    aio.write_merged_trace(path, "tas", params, logp, scaling, lat, lon)
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.variables["params"].dimensions == ("lat", "lon", "params_dim_0")   # detrend.py:245-249
    assert set(nc.variables) >= {"params", "logp", "scaling_datamin", "scaling_scale", "lat", "lon"}
    nc.close()
    order = [2, 0]
    p, l, s = aio.read_merged_trace(path, lat[order], lon[order])
    np.testing.assert_array_equal(p, params[order])
    np.testing.assert_array_equal(l, logp[order])
    np.testing.assert_array_equal(s, scaling[order])
    with pytest.raises(KeyError):
        aio.read_merged_trace(path, [52.0], [9.25])


def test_shape_and_duplicate_checks(tmp_path):
    lat, lon = _cells()
    with pytest.raises(ValueError):
        aio.write_merged_output(str(tmp_path / "a.nc"), {"cfact": np.zeros((5, 2))}, lat, lon, np.arange(5))
    with pytest.raises(ValueError):
        aio.write_merged_output(str(tmp_path / "b.nc"), {}, [1.0, 1.0], [2.0, 2.0], np.arange(5))
