"""Merged output and trace files written directly from the batched results.

The reference writes one NetCDF file per cell (``ts_lat<lat>_lon<lon>.nc``, attrici/detrend.py:447-477; trace:
``write_trace`` :221-255) and later merges them with ``attrici merge-output`` (attrici/commands/merge_output.py:27-123)
into one file with dimensions ``(time, lat, lon)``.  With the cells solved as one batch the merged layout can be
written at once.  Files are NetCDF-3 (64-bit offset) through ``scipy.io.netcdf_file`` - this image has neither
netCDF4 nor h5py, so the reference's zlib-compressed NetCDF-4 encoding is not available; xarray / netCDF4 open
classic files transparently.  Layout (the one ``merge-output`` produces):

* dimensions ``time, lat, lon`` (lat / lon = sorted unique coordinates of the cells, float32 like :53, :61),
* ``time(time)`` with its ``units`` / ``calendar`` attributes, ``lat(lat)``, ``lon(lon)``,
* per-cell series ``<name>(time, lat, lon)`` and per-cell scalars ``<name>(lat, lon)``; grid points without a cell
  hold the fill value (NaN; 0 for ``replaced``, whose ``_FillValue`` is 0 as at detrend.py:475),
* the trace file: every entry ``k`` of the per-cell trace with dimensions ``(lat, lon, k_dim_0, ...)`` (:245-249) -
  ``params(lat, lon, params_dim_0)``, ``logp(lat, lon)``, and the variable's scaling constants ``scaling_<key>``.

Pure host code (NumPy + SciPy); nothing here touches the GPU.
"""

import numpy as np
from scipy.io import netcdf_file

from .variables import create_variable


def _grid(lat, lon):
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    if lat.ndim != 1 or lat.shape != lon.shape:
        raise ValueError("lat / lon must be 1-D arrays with one entry per cell")
    ulat, ilat = np.unique(lat, return_inverse=True)
    ulon, ilon = np.unique(lon, return_inverse=True)
    flat = ilat * ulon.size + ilon
    if np.unique(flat).size != flat.size:
        raise ValueError("two cells share one (lat, lon)")
    return ulat, ulon, ilat, ilon


def _create(path, ulat, ulon, attrs):
    nc = netcdf_file(path, "w", version=2)
    nc.createDimension("lat", ulat.size)
    nc.createDimension("lon", ulon.size)
    v = nc.createVariable("lat", "f4", ("lat",))
    v[:] = ulat.astype(np.float32)
    v.units = "degrees_north"
    v = nc.createVariable("lon", "f4", ("lon",))
    v[:] = ulon.astype(np.float32)
    v.units = "degrees_east"
    for k, val in (attrs or {}).items():
        setattr(nc, k, val)
    return nc


def _nc_type(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.uint8 or dtype == np.int8 or dtype == np.bool_:
        return "i1", np.int8
    if dtype.kind in "iu":
        return "i4", np.int32
    return ("f4", np.float32) if dtype == np.float32 else ("f8", np.float64)


def write_merged_output(path, series, lat, lon, time, time_units="days since 1901-01-01 00:00:00",
                        calendar="proleptic_gregorian", scalars=None, shared=None, attrs=None):
    """One ``(time, lat, lon)`` file for a batch of cells.

    ``series``: dict name -> [T, C] array (e.g. ``{"y": y, "cfact": res["cfact"], "replaced": res["replaced"]}``, the
    ``report_variables`` of detrend.py:462-467); ``scalars``: dict name -> [C] (e.g. ``logp``); ``shared``: dict
    name -> [T] (``gmt_scaled``)."""
    ulat, ulon, ilat, ilon = _grid(lat, lon)
    time = np.asarray(time)
    T, C = time.size, ilat.size
    nc = _create(path, ulat, ulon, attrs)
    try:
        nc.createDimension("time", T)
        code, cast = _nc_type(time.dtype)
        v = nc.createVariable("time", code, ("time",))
        v[:] = time.astype(cast)
        v.units = time_units
        v.calendar = calendar
        for name, arr in (shared or {}).items():
            arr = np.asarray(arr)
            if arr.shape != (T,):
                raise ValueError(f"shared variable {name} must have shape [T]")
            code, cast = _nc_type(arr.dtype)
            nc.createVariable(name, code, ("time",))[:] = arr.astype(cast)
        for name, arr in (series or {}).items():
            arr = np.asarray(arr)
            if arr.shape != (T, C):
                raise ValueError(f"series {name} must have shape [T, C] = [{T}, {C}], got {arr.shape}")
            code, cast = _nc_type(arr.dtype)
            v = nc.createVariable(name, code, ("time", "lat", "lon"))
            fill = 0 if code in ("i1", "i4") else np.nan
            v._FillValue = cast(fill)
            full = np.full((T, ulat.size, ulon.size), fill, dtype=cast)
            full[:, ilat, ilon] = arr
            v[:] = full
        for name, arr in (scalars or {}).items():
            arr = np.asarray(arr)
            if arr.shape != (C,):
                raise ValueError(f"scalar variable {name} must have shape [C]")
            code, cast = _nc_type(arr.dtype)
            v = nc.createVariable(name, code, ("lat", "lon"))
            fill = 0 if code in ("i1", "i4") else np.nan
            full = np.full((ulat.size, ulon.size), fill, dtype=cast)
            full[ilat, ilon] = arr
            v[:] = full
    finally:
        nc.close()


def write_merged_trace(path, variable, params, logp, scaling, lat, lon, attrs=None):
    """The traces of a batch of cells in merged form: ``params(lat, lon, params_dim_0)``, ``logp(lat, lon)`` and
    ``scaling_<key>(lat, lon)`` - the entries ``write_trace`` stores per cell (detrend.py:221-255 with the SciPy trace
    schema ``{"logp", "params"}``, model_scipy.py:524-527, plus the variable's scaling constants)."""
    spec = create_variable(variable) if isinstance(variable, str) else variable
    ulat, ulon, ilat, ilon = _grid(lat, lon)
    params = np.asarray(params, dtype=np.float64)
    C = ilat.size
    if params.ndim != 2 or params.shape[0] != C:
        raise ValueError("params must be [C, P]")
    nc = _create(path, ulat, ulon, attrs)
    try:
        nc.createDimension("params_dim_0", params.shape[1])
        full = np.full((ulat.size, ulon.size, params.shape[1]), np.nan)
        full[ilat, ilon] = params
        nc.createVariable("params", "f8", ("lat", "lon", "params_dim_0"))[:] = full
        full = np.full((ulat.size, ulon.size), np.nan)
        full[ilat, ilon] = np.asarray(logp, dtype=np.float64)
        nc.createVariable("logp", "f8", ("lat", "lon"))[:] = full
        scaling = np.asarray(scaling, dtype=np.float64)
        cols = {"datamin": 0, "scale": 1}
        for key in spec.scaling_keys:
            full = np.full((ulat.size, ulon.size), np.nan)
            full[ilat, ilon] = scaling[:, cols[key]]
            nc.createVariable(f"scaling_{key}", "f8", ("lat", "lon"))[:] = full
        nc.variable = spec.name
    finally:
        nc.close()


def read_merged_trace(path, lat, lon):
    """``(params [C, P], logp [C], scaling [C, 2])`` of the requested cells from a merged trace file (what
    ``--trace-file`` is for: skip the fit and go straight to ``CellBatchSolver.quantile_map``)."""
    nc = netcdf_file(path, "r", mmap=False)
    try:
        flat, flon = np.asarray(nc.variables["lat"][:], dtype=np.float32), np.asarray(nc.variables["lon"][:], dtype=np.float32)
        lat32, lon32 = np.asarray(lat, dtype=np.float32), np.asarray(lon, dtype=np.float32)
        ilat, ilon = np.searchsorted(flat, lat32), np.searchsorted(flon, lon32)
        if (ilat >= flat.size).any() or (ilon >= flon.size).any() or (flat[ilat] != lat32).any() or (flon[ilon] != lon32).any():
            raise KeyError("not all cells are in the trace file")
        params = np.array(nc.variables["params"][:])[ilat, ilon]
        logp = np.array(nc.variables["logp"][:])[ilat, ilon]
        scaling = np.zeros((lat32.size, 2))
        scaling[:, 1] = 1.0
        for key, col in (("datamin", 0), ("scale", 1)):
            if f"scaling_{key}" in nc.variables:
                scaling[:, col] = np.array(nc.variables[f"scaling_{key}"][:])[ilat, ilon]
        return params, logp, scaling
    finally:
        nc.close()
