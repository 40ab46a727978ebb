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
    assert nc.dimensions == {"time": None, "lat": 2, "lon": 2}               # time is the record dimension
    assert nc.variables["time"].shape == (T,) and nc.variables["cfact"].shape == (T, 2, 2)
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
    assert rep.dtype.kind == "f" and rep.dtype.itemsize == 8                                             # the reference's float flags (0 / NaN / +-Inf)
    np.testing.assert_array_equal(np.isnan(rep[:, 0, 1]), replaced[:, 1] == 1)
    np.testing.assert_array_equal(rep[:, 0, 1] == 0, replaced[:, 1] == 0)
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


@pytest.mark.parametrize("block_bytes", [1, 100, 1 << 20])
def test_streamed_records(tmp_path, block_bytes):
    """Days appended block by block (and in several calls) give the same file; record slices of 1-byte variables
    are padded to 4 bytes (3 grid points), a fully covered grid needs no fill."""
    lat, lon = np.array([10.25, 10.25, 10.25]), np.array([1.25, 0.25, 0.75])   # one row, cells not in lon order
    T, C = 23, 3
    rng = np.random.default_rng(2)
    cfact = rng.normal(size=(T, C))
    replaced = rng.integers(0, 4, (T, C)).astype(np.uint8)
    time = np.arange(100, 100 + T, dtype=np.int64)
    path = str(tmp_path / "s.nc")
    with aio.MergedOutputWriter(path, lat, lon, {"cfact": np.float64, "replaced": np.uint8}, shared={"gmt_scaled": np.float64},
                                scalars={"status": np.int32([0, 2, 0])}, block_bytes=block_bytes) as w:
        for a, b in ((0, 9), (9, 10), (10, T)):
            w.append(time[a:b], {"cfact": cfact[a:b], "replaced": replaced[a:b]}, {"gmt_scaled": np.linspace(0, 1, T)[a:b]})
        with pytest.raises(ValueError):
            w.append(time[:2], {"cfact": cfact[:2]}, {"gmt_scaled": np.zeros(2)})
        with pytest.raises(ValueError):
            w.append(time[:2], {"cfact": cfact[:3], "replaced": replaced[:2]}, {"gmt_scaled": np.zeros(2)})
    for mmap in (False, True):
        nc = netcdf_file(path, "r", mmap=mmap)
        order = np.argsort(lon)
        np.testing.assert_array_equal(np.array(nc.variables["lon"][:]), np.float32(lon[order]))
        np.testing.assert_array_equal(np.array(nc.variables["cfact"][:])[:, 0, :], cfact[:, order])
        np.testing.assert_array_equal(np.array(nc.variables["replaced"][:])[:, 0, :], replaced[:, order].astype(np.int8))
        np.testing.assert_array_equal(np.array(nc.variables["time"][:]), time.astype(np.int32))
        np.testing.assert_array_equal(np.array(nc.variables["gmt_scaled"][:]), np.linspace(0, 1, T))
        np.testing.assert_array_equal(np.array(nc.variables["status"][:]), np.int32([0, 2, 0])[order][None, :])
        del nc  # mmap=True: SciPy warns if views outlive close()


def test_lone_record_variable_and_torch_input(tmp_path):
    """A file whose only record variable is ``time`` (no per-record padding rule), and CPU tensors as series."""
    import torch

    path = str(tmp_path / "t.nc")
    with aio.MergedOutputWriter(path, [1.0], [2.0], {}) as w:
        w.append(np.arange(5, dtype=np.int32), {})
    nc = netcdf_file(path, "r", mmap=False)
    np.testing.assert_array_equal(np.array(nc.variables["time"][:]), np.arange(5))
    nc.close()
    cf = torch.arange(12, dtype=torch.float64).reshape(6, 2)
    aio.write_merged_output(path, {"cfact": cf}, [1.0, 1.5], [2.0, 2.0], np.arange(6))
    nc = netcdf_file(path, "r", mmap=False)
    np.testing.assert_array_equal(np.array(nc.variables["cfact"][:])[:, :, 0], cf.numpy())
    nc.close()


def test_streamed_writer_random_layouts(tmp_path):
    """Random grids (cells in any order, any subset of the lat x lon box), dtypes, block sizes and append patterns:
    the file read back with SciPy equals a dense array filled the straightforward way."""
    rng = np.random.default_rng(42)
    path = str(tmp_path / "r.nc")
    for case in range(25):
        n_lat, n_lon, T = int(rng.integers(1, 6)), int(rng.integers(1, 7)), int(rng.integers(1, 30))
        ulat = np.sort(rng.choice(np.arange(-89.75, 90, 0.5), n_lat, replace=False))
        ulon = np.sort(rng.choice(np.arange(-179.75, 180, 0.5), n_lon, replace=False))
        n_grid = n_lat * n_lon
        C = int(rng.integers(1, n_grid + 1))
        cells = rng.permutation(n_grid)[:C]                      # any order, any subset
        lat, lon = ulat[cells // n_lon], ulon[cells % n_lon]
        dtypes = {"cfact": np.float64, "y": np.float32, "replaced": np.uint8, "count": np.int32}
        names = [k for k in dtypes if rng.random() < 0.7] or ["cfact"]
        data = {k: (rng.normal(size=(T, C)) * 50).astype(dtypes[k]) if dtypes[k] in (np.float64, np.float32)
                else rng.integers(0, 100, (T, C)).astype(dtypes[k]) for k in names}
        time = np.cumsum(rng.integers(1, 4, T)).astype(np.int32)
        with aio.MergedOutputWriter(path, lat, lon, {k: dtypes[k] for k in names},
                                    block_bytes=int(rng.choice([1, 64, 4096, 1 << 20]))) as w:
            a = 0
            while a < T:
                b = min(T, a + int(rng.integers(1, T + 1)))
                w.append(time[a:b], {k: data[k][a:b] for k in names})
                a = b
        nc = netcdf_file(path, "r", mmap=False)
        # lat / lon of the file are the unique coordinates of the cells present, which may be fewer than the box
        flat, flon = np.array(nc.variables["lat"][:]), np.array(nc.variables["lon"][:])
        assert np.array_equal(flat, np.unique(lat).astype(np.float32)) and np.array_equal(flon, np.unique(lon).astype(np.float32))
        assert np.array_equal(np.array(nc.variables["time"][:]), time)
        ilat, ilon = np.searchsorted(np.unique(lat), lat), np.searchsorted(np.unique(lon), lon)
        for k in names:
            got = np.array(nc.variables[k][:])
            cast = {np.uint8: np.int8}.get(dtypes[k], dtypes[k])
            fill = 0 if np.dtype(cast).kind == "i" else np.nan
            want = np.full((T, flat.size, flon.size), fill, dtype=cast)
            want[:, ilat, ilon] = data[k].astype(cast)
            assert got.shape == want.shape and np.array_equal(got, want, equal_nan=np.dtype(cast).kind == "f"), (case, k)
        nc.close()


def test_per_cell_files_at_the_reference_paths(tmp_path):
    """attrici/detrend.py:236-242, 287-293 (paths), 447-477 (timeseries file), 221-255 (trace file)."""
    lat, lon = np.array([50.75, 51.25, -0.25]), np.array([9.25, 9.75, 100.0])
    T, C = 12, 3
    rng = np.random.default_rng(3)
    y = rng.normal(280, 5, (T, C)).astype(np.float32)
    cfact = y.astype(np.float64) + 0.5
    codes = rng.integers(0, 4, (T, C)).astype(np.uint8)
    params, logp = rng.normal(size=(C, 27)), rng.normal(size=C)
    scaling = np.column_stack([rng.uniform(250, 260, C), rng.uniform(40, 60, C)])
    ts, tr = aio.cell_paths(tmp_path, "tas", 50.75, 9.25)
    assert ts == str(tmp_path / "timeseries" / "tas" / "lat_50.75" / "ts_lat50.75_lon9.25.nc")
    assert tr == str(tmp_path / "trace" / "tas" / "lat_50.75" / "trace_lat50.75_lon9.25.nc")
    assert aio.cell_paths(tmp_path, "tas", -0.25, 100.0)[0].endswith("lat_-0.25/ts_lat-0.25_lon100.nc")   # ":g"
    written = aio.write_cell_files(tmp_path, "tas", lat, lon, np.arange(T, dtype=np.int32),
                                   {"y": y, "cfact": cfact, "replaced": codes}, logp, params=params, scaling=scaling,
                                   shared={"gmt_scaled": np.linspace(0, 1, T)})
    assert len(written) == C and written[0] == ts
    for c in range(C):
        ts, tr = aio.cell_paths(tmp_path, "tas", lat[c], lon[c])
        nc = netcdf_file(ts, "r", mmap=False)
        assert nc.variables["cfact"].shape == (T, 1, 1) and nc.variables["cfact"].dimensions == ("time", "lat", "lon")
        np.testing.assert_array_equal(np.array(nc.variables["cfact"][:])[:, 0, 0], cfact[:, c])
        np.testing.assert_array_equal(np.array(nc.variables["y"][:])[:, 0, 0], y[:, c])
        rep = np.array(nc.variables["replaced"][:])[:, 0, 0]
        np.testing.assert_array_equal(rep, aio.replaced_flags(codes[:, c]))       # 0 / NaN / +Inf / -Inf
        assert nc.variables["replaced"].data.dtype.itemsize == 8
        assert np.array(nc.variables["logp"][:]).ravel()[0] == logp[c]
        assert np.array(nc.variables["lat"][:])[0] == np.float32(lat[c])
        nc.close()
        p, l, sc = aio.read_merged_trace(tr, lat[c:c + 1], lon[c:c + 1])
        np.testing.assert_array_equal(p[0], params[c])
        np.testing.assert_array_equal(sc[0], scaling[c])
    # existing output is skipped unless overwrite (detrend.py:296-305)
    assert aio.write_cell_files(tmp_path, "tas", lat, lon, np.arange(T), {"cfact": cfact}, logp, overwrite=False) == []
    assert len(aio.write_cell_files(tmp_path, "tas", lat, lon, np.arange(T), {"cfact": cfact}, logp, cells=[1])) == 1
    assert np.array_equal(aio.replaced_flags([0, 1, 2, 3]), [0.0, np.nan, np.inf, -np.inf], equal_nan=True)


def test_provenance_attributes(tmp_path):
    """Keys of get_data_provenance_metadata (attrici/util.py:17-41) + the configuration as TOML, stored as global
    attributes."""
    import tomllib

    from attrici_b200.detrend import Config

    cfg = Config("tas", modes=4, seed=3, report_variables=("cfact", 'mu "x"'))
    a = aio.provenance_attrs(cfg, note="synthetic")
    assert set(a) == {"attrici_version", "attrici_packages", "attrici_python_version", "created_at", "attrici_config", "note"}
    parsed = tomllib.loads(a["attrici_config"])
    assert parsed["variable"] == "tas" and parsed["modes"] == 4 and parsed["seed"] == 3 and parsed["fit_only"] is False
    assert parsed["report_variables"] == ["cfact", 'mu "x"'] and "window_size" not in parsed and "device" not in parsed
    assert "numpy==" in a["attrici_packages"]
    path = str(tmp_path / "p.nc")
    aio.write_merged_output(path, {"cfact": np.zeros((2, 1))}, [1.0], [2.0], np.arange(2), attrs=a)
    nc = netcdf_file(path, "r", mmap=False)
    assert tomllib.loads(nc.attrici_config.decode())["variable"] == "tas" and nc.note == b"synthetic"
    nc.close()


def test_writer_takes_slab_major_and_sparse_results(tmp_path):
    """``ColumnBlocks`` (per-slab [T, n_k] blocks of ``detrend_host(slab_major=True)``) and ``SparseReplaced`` (the
    (day, cell, code) list of ``replaced="sparse"``) are written like the dense [T, C] arrays they stand for."""
    import torch
    from scipy.io import netcdf_file

    from attrici_b200 import io as aio

    rng = np.random.default_rng(3)
    T, C = 37, 11
    lat, lon = np.repeat([10.25, 10.75], 6)[:C], np.tile(np.arange(6) * 0.5 + 3.25, 2)[:C]
    cfact = rng.normal(size=(T, C))
    rep = (rng.random((T, C)) < 0.05).astype(np.uint8) * rng.integers(1, 4, size=(T, C)).astype(np.uint8)
    entries = np.array([(t, c, rep[t, c]) for t in range(T) for c in range(C) if rep[t, c]], dtype=np.int32).reshape(-1, 3)
    entries = entries[rng.permutation(len(entries))]                      # the kernel appends in no particular order
    assert np.array_equal(aio.dense_replaced(entries, T, C), rep)
    blocks = aio.ColumnBlocks([torch.from_numpy(cfact[:, :4].copy()), torch.from_numpy(cfact[:, 4:9].copy()),
                               torch.from_numpy(cfact[:, 9:].copy())])
    assert blocks.shape == (T, C) and np.array_equal(blocks[5:9].numpy(), cfact[5:9])
    assert np.array_equal(blocks.columns([0, 4, 10]).numpy(), cfact[:, [0, 4, 10]])
    sparse = aio.SparseReplaced(entries, T, C)
    assert np.array_equal(sparse[3:20].numpy(), rep[3:20])
    a, b = str(tmp_path / "a.nc"), str(tmp_path / "b.nc")
    aio.write_merged_output(a, {"cfact": cfact, "replaced": rep}, lat, lon, np.arange(T, dtype=np.int32), replaced_as_flags=False)
    aio.write_merged_output(b, {"cfact": blocks, "replaced": sparse}, lat, lon, np.arange(T, dtype=np.int32),
                            replaced_as_flags=False)
    assert open(a, "rb").read() == open(b, "rb").read()
    aio.write_merged_output(b, {"cfact": blocks, "replaced": sparse}, lat, lon, np.arange(T, dtype=np.int32))
    with netcdf_file(b, mmap=False) as f:
        flags = f.variables["replaced"][:].reshape(T, -1)
    assert np.isnan(flags).sum() == (rep == 1).sum() and np.isposinf(flags).sum() == (rep == 2).sum()
