"""Merged output and trace files written directly from the batched results.

The reference writes one NetCDF file per cell (``ts_lat<lat>_lon<lon>.nc``, attrici/detrend.py:447-477; trace:
``write_trace`` :221-255) and later merges them with ``attrici merge-output`` (attrici/commands/merge_output.py:27-123)
into one file with dimensions ``(time, lat, lon)``.  With the cells solved as one batch the merged layout can be
written at once.  Files are NetCDF classic (64-bit offset) - this image has neither netCDF4 nor h5py, so the
reference's zlib-compressed NetCDF-4 encoding is not available; xarray / netCDF4 open classic files transparently.
The series file is streamed by ``MergedOutputWriter`` (``time`` as the record dimension, a block of days per write);
the small trace file goes through ``scipy.io.netcdf_file``.  Layout (the one ``merge-output`` produces):

* dimensions ``time, lat, lon`` (lat / lon = sorted unique coordinates of the cells, float32 like :53, :61),
* ``time(time)`` with its ``units`` / ``calendar`` attributes, ``lat(lat)``, ``lon(lon)``,
* per-cell series ``<name>(time, lat, lon)`` and per-cell scalars ``<name>(lat, lon)``; grid points without a cell
  hold the fill value (NaN; 0 for ``replaced``, whose ``_FillValue`` is 0 as at detrend.py:475),
* the trace file: every entry ``k`` of the per-cell trace with dimensions ``(lat, lon, k_dim_0, ...)`` (:245-249) -
  ``params(lat, lon, params_dim_0)``, ``logp(lat, lon)``, and the variable's scaling constants ``scaling_<key>``.

Pure host code (NumPy + SciPy); nothing here touches the GPU.
"""

import os

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


# NetCDF classic format, 64-bit-offset variant ("CDF-2"): tags and type codes of the file header
_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 10, 11, 12
_NC_CODE = {"i1": 1, "c": 2, "i4": 4, "f4": 5, "f8": 6}
_BE = {"i1": ">i1", "i4": ">i4", "f4": ">f4", "f8": ">f8"}


def _pad4(n):
    return (n + 3) & ~3


def _be32(n):
    return int(n).to_bytes(4, "big", signed=True)


def _nc_name(name):
    b = name.encode("utf-8")
    return _be32(len(b)) + b + b"\0" * (_pad4(len(b)) - len(b))


def _nc_attributes(attrs):
    """``att_list`` of the classic format: text attributes as NC_CHAR, numbers as one value of their own type."""
    if not attrs:
        return _be32(0) + _be32(0)
    out = [_be32(_NC_ATTRIBUTE), _be32(len(attrs))]
    for key, val in attrs.items():
        out.append(_nc_name(key))
        if isinstance(val, (str, bytes)):
            b = val.encode("utf-8") if isinstance(val, str) else val
            out += [_be32(_NC_CODE["c"]), _be32(len(b)), b + b"\0" * (_pad4(len(b)) - len(b))]
        else:
            code, cast = _nc_type(np.asarray(val).dtype)
            b = np.asarray(val).astype(cast).astype(_BE[code]).tobytes()
            out += [_be32(_NC_CODE[code]), _be32(len(b) // np.dtype(cast).itemsize), b + b"\0" * (_pad4(len(b)) - len(b))]
    return b"".join(out)


class MergedOutputWriter:
    """Streams the ``merge-output`` layout to disk, a block of days at a time.

    ``time`` is the record (unlimited) dimension of a NetCDF classic file in the 64-bit-offset variant, written by
    hand: the format is a header followed by the fixed-size variables (``lat``, ``lon``, per-cell scalars) and then one
    record per day holding that day's slice of every ``(time, ...)`` variable — which is the order the results
    arrive in when cells are solved as batches, so a block of days is one contiguous write and memory stays at one
    block whatever the grid (a global 0.5 degree ``cfact(time, lat, lon)`` is 90 GB; ``scipy.io.netcdf_file``
    builds all variables in memory before writing, and the classic format caps a fixed-size variable at 4 GiB, a
    record variable only per record).  xarray / netCDF4 / SciPy read the file like any other.

    ``series``: name -> dtype of the ``(time, lat, lon)`` variables; ``shared``: name -> dtype of ``(time,)``
    variables (``gmt_scaled``); ``scalars``: name -> [C] array of ``(lat, lon)`` variables, written at once."""

    def __init__(self, path, lat, lon, series, shared=None, scalars=None, time_dtype=np.int32,
                 time_units="days since 1901-01-01 00:00:00", calendar="proleptic_gregorian", attrs=None,
                 block_bytes=64 << 20):
        ulat, ulon, ilat, ilon = _grid(lat, lon)
        self._flat = ilat * ulon.size + ilon
        self._n_cells = ilat.size
        self._identity = bool(np.array_equal(self._flat, np.arange(ulat.size * ulon.size)))
        n_grid = ulat.size * ulon.size
        self.block_bytes = int(block_bytes)
        self.n_records = 0
        self._block = None
        dims = {"time": 0, "lat": ulat.size, "lon": ulon.size}
        dim_id = {k: i for i, k in enumerate(dims)}
        # (name, dims, nc type, attributes, payload of a fixed variable or None)
        fixed = [("lat", ("lat",), "f4", {"units": "degrees_north"}, ulat.astype(">f4").tobytes()),
                 ("lon", ("lon",), "f4", {"units": "degrees_east"}, ulon.astype(">f4").tobytes())]
        for name, arr in (scalars or {}).items():
            arr = np.asarray(arr)
            if arr.shape != (self._n_cells,):
                raise ValueError(f"scalar variable {name} must have shape [C]")
            code, cast = _nc_type(arr.dtype)
            full = np.full(n_grid, 0 if code in ("i1", "i4") else np.nan, dtype=_BE[code])
            full[self._flat] = arr.astype(cast)
            fixed.append((name, ("lat", "lon"), code, {}, full.tobytes()))
        code, self._time_cast = _nc_type(time_dtype)
        records = [("time", ("time",), code, {"units": time_units, "calendar": calendar}, 1)]
        for name, dt in (shared or {}).items():
            records.append((name, ("time",), _nc_type(dt)[0], {}, 1))
        self._fill = {}
        for name, dt in (series or {}).items():
            code, cast = _nc_type(dt)
            # `replaced` has _FillValue 0 whatever its type (attrici/detrend.py:471-476)
            self._fill[name] = cast(0 if (code in ("i1", "i4") or name == "replaced") else np.nan)
            records.append((name, ("time", "lat", "lon"), code, {"_FillValue": self._fill[name]}, n_grid))
        self._casts = {name: np.dtype(_BE[code]).newbyteorder("=").type for name, _, code, _, _ in records}
        self._series = tuple(series or {})
        self._shared = tuple(shared or {})

        # record layout: every variable's slice of a record is padded to 4 bytes unless it is the only record variable
        sizes = [np.dtype(_BE[code]).itemsize * n for _, _, code, _, n in records]
        vsizes = sizes if len(records) == 1 else [_pad4(x) for x in sizes]
        offsets = np.concatenate([[0], np.cumsum(vsizes)[:-1]]).tolist()
        self._rec_size = int(sum(vsizes))
        self._rec_dtype = np.dtype({
            "names": [r[0] for r in records],
            "formats": [(_BE[code], (n,)) if len(vdims) == 3 else _BE[code] for _, vdims, code, _, n in records],
            "offsets": offsets, "itemsize": self._rec_size})

        def header(begins):
            out = [b"CDF\x02", _be32(0), _be32(_NC_DIMENSION), _be32(len(dims))]
            out += [_nc_name(k) + _be32(n) for k, n in dims.items()]
            out.append(_nc_attributes(attrs))
            out += [_be32(_NC_VARIABLE), _be32(len(fixed) + len(records))]
            for (name, vdims, code, vattrs, _), vsize, begin in zip(
                    fixed + records, [_pad4(len(f[4])) for f in fixed] + vsizes, begins):
                out += [_nc_name(name), _be32(len(vdims))] + [_be32(dim_id[d]) for d in vdims]
                out += [_nc_attributes(vattrs), _be32(_NC_CODE[code]), _be32(vsize), int(begin).to_bytes(8, "big")]
            return b"".join(out)

        n_var = len(fixed) + len(records)
        pos = len(header([0] * n_var))            # the header's length does not depend on the offsets
        begins = []
        for f in fixed:
            begins.append(pos)
            pos += _pad4(len(f[4]))
        self._rec_begin = pos
        begins += [pos + o for o in offsets]
        self._f = open(path, "wb")
        self._f.write(header(begins))
        for f in fixed:
            self._f.write(f[4] + b"\0" * (_pad4(len(f[4])) - len(f[4])))
        assert self._f.tell() == self._rec_begin

    def append(self, time, series, shared=None):
        """Write the next ``n`` days: ``time`` [n], ``series`` name -> [n, C], ``shared`` name -> [n]."""
        time = np.asarray(time)
        n = time.size
        if set(series or {}) != set(self._series) or set(shared or {}) != set(self._shared):
            raise ValueError("append needs exactly the variables the writer was created with")
        for name in self._series:
            if tuple(series[name].shape) != (n, self._n_cells):
                raise ValueError(f"series {name} must have shape [{n}, {self._n_cells}], got {tuple(series[name].shape)}")
        step = max(1, self.block_bytes // self._rec_size)
        for a in range(0, n, step):
            b = min(n, a + step)
            if self._block is None or self._block.size < b - a:
                # one record block, reused: its pages stay mapped, and the grid points without a cell keep the fill
                # value written here (every append overwrites the same cell positions)
                self._block = np.zeros(min(step, max(n, b - a)), dtype=self._rec_dtype)
                if not self._identity:
                    for name in self._series:
                        self._block[name][...] = self._fill[name]
            rec = self._block[: b - a]
            rec["time"] = time[a:b].astype(self._time_cast)
            for name in self._shared:
                v = np.asarray(shared[name])
                if v.shape != (n,):
                    raise ValueError(f"shared variable {name} must have shape [{n}]")
                rec[name] = v[a:b].astype(self._casts[name])
            for name in self._series:
                src = np.asarray(series[name][a:b])
                if self._identity:
                    rec[name] = src                      # cells already in grid order, no gaps
                    continue
                field = rec[name]                        # row-wise scatter: NumPy's 2-D fancy assignment is 6x slower
                for i in range(b - a):
                    field[i, self._flat] = src[i]
            self._f.write(rec.view(np.uint8))
        self.n_records += n

    def close(self):
        if self._f is not None:
            if self._rec_size % 4:               # a lone, unpadded record variable: the file still ends on 4 bytes
                self._f.write(b"\0" * (_pad4(self._f.tell()) - self._f.tell()))
            self._f.seek(4)
            self._f.write(_be32(self.n_records))
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class ReplacedFlagsView:
    """[T, C] view of the library's uint8 ``replaced`` codes that yields the reference's float64 flags
    (0 / NaN / +Inf / -Inf, attrici/detrend.py:397-410) a block of days at a time, so that a merged file holds what
    the reference's files hold without a second [T, C] float64 array in memory."""

    dtype = np.dtype(np.float64)

    def __init__(self, codes):
        self.codes = codes
        self.shape = tuple(codes.shape)

    def __getitem__(self, key):
        part = self.codes[key]
        return replaced_flags(part.numpy() if hasattr(part, "numpy") else np.asarray(part))

    def __array__(self, dtype=None, copy=None):
        return replaced_flags(np.asarray(self.codes))


def write_merged_output(path, series, lat, lon, time, time_units="days since 1901-01-01 00:00:00",
                        calendar="proleptic_gregorian", scalars=None, shared=None, attrs=None, replaced_as_flags=True):
    """One ``(time, lat, lon)`` file for a batch of cells.

    An integer ``replaced`` series (the library's codes 0..3) is stored as the reference's float64 flags
    0 / NaN / +Inf / -Inf (attrici/detrend.py:397-410, ``replaced_as_flags``; ``False`` keeps the one-byte codes, with
    the code table in the variable's ``flag_meanings`` attribute).

    ``series``: dict name -> [T, C] array (e.g. ``{"y": y, "cfact": res["cfact"], "replaced": res["replaced"]}``, the
    ``report_variables`` of detrend.py:462-467); ``scalars``: dict name -> [C] (e.g. ``logp``); ``shared``: dict
    name -> [T] (``gmt_scaled``).  Streams through ``MergedOutputWriter``."""
    time = np.asarray(time)
    T = time.size
    lazy = (ColumnBlocks, SparseReplaced)   # sliced by days on demand
    series = {k: (v if isinstance(v, lazy) else (v.numpy() if hasattr(v, "numpy") else np.asarray(v)))
              for k, v in (series or {}).items()}
    np_dtype = {k: (v.numpy_dtype() if isinstance(v, lazy) else v.dtype) for k, v in series.items()}
    if replaced_as_flags and "replaced" in series and np_dtype["replaced"].kind in "iub":
        series["replaced"] = ReplacedFlagsView(series["replaced"])
        np_dtype["replaced"] = np.dtype(np.float64)
    shared = {k: np.asarray(v) for k, v in (shared or {}).items()}
    n_cells = np.asarray(lat).size
    for name, arr in series.items():
        if tuple(arr.shape) != (T, n_cells):
            raise ValueError(f"series {name} must have shape [T, C] = [{T}, {n_cells}], got {tuple(arr.shape)}")
    for name, arr in shared.items():
        if arr.shape != (T,):
            raise ValueError(f"shared variable {name} must have shape [T]")
    def days_of(v, a, b):
        part = v[a:b]
        return part.numpy() if hasattr(part, "numpy") else part

    with MergedOutputWriter(path, lat, lon, np_dtype,
                            shared={k: v.dtype for k, v in shared.items()}, scalars=scalars, time_dtype=time.dtype,
                            time_units=time_units, calendar=calendar, attrs=attrs) as w:
        step = max(1, w.block_bytes // w._rec_size)
        for a in range(0, T, step):
            b = min(T, a + step)
            w.append(time[a:b], {k: days_of(v, a, b) for k, v in series.items()}, {k: v[a:b] for k, v in shared.items()})


class ColumnBlocks:
    """[T, n] series held as column blocks (the per-slab ``cfact_slabs`` of ``CellBatchSolver.detrend_host(slab_major=
    True)``): slicing a range of days concatenates the blocks' rows, so the writers can consume the slab-major result
    without a [T, n] copy."""

    def __init__(self, blocks, dtype=None):
        """``dtype``: torch dtype the rows are cast to when they are sliced (e.g. float32 for a file of the inputs' dtype)."""
        import torch

        self.blocks = [b if isinstance(b, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(b)) for b in blocks]
        self.shape = (int(self.blocks[0].shape[0]), int(sum(b.shape[1] for b in self.blocks)))
        self.dtype = dtype or self.blocks[0].dtype
        self.is_cuda = False

    def dim(self):
        return 2

    def element_size(self):
        return self.blocks[0][:0].to(self.dtype).element_size()

    def numpy_dtype(self):
        return self.blocks[0][:0].to(self.dtype).numpy().dtype

    def __getitem__(self, key):
        import torch

        return torch.cat([b[key].to(self.dtype) for b in self.blocks], dim=1)

    def columns(self, cols):
        """[T, len(cols)] copy of the given columns."""
        import torch

        edges = np.cumsum([0] + [b.shape[1] for b in self.blocks])
        out = []
        for c in cols:
            k = int(np.searchsorted(edges, c, side="right") - 1)
            out.append(self.blocks[k][:, int(c - edges[k])])
        return torch.stack(out, dim=1)


class SparseReplaced:
    """[T, n] uint8 plane of replacement codes given as the sparse ``(day, cell, code)`` rows of
    ``detrend_host(replaced="sparse")``; a range of days is materialised on demand."""

    def __init__(self, entries, n_time, n_cells):
        import torch

        e = np.asarray(entries).reshape(-1, 3)
        order = np.argsort(e[:, 0], kind="stable")
        self._e = e[order]
        self.shape = (int(n_time), int(n_cells))
        self.dtype = torch.uint8
        self.is_cuda = False

    def dim(self):
        return 2

    def element_size(self):
        return 1

    def numpy_dtype(self):
        return np.dtype(np.uint8)

    def __getitem__(self, key):
        import torch

        a, b, _ = key.indices(self.shape[0])
        out = np.zeros((b - a, self.shape[1]), dtype=np.uint8)
        lo, hi = np.searchsorted(self._e[:, 0], [a, b])
        rows = self._e[lo:hi]
        out[rows[:, 0] - a, rows[:, 1]] = rows[:, 2].astype(np.uint8)
        return torch.from_numpy(out)


def write_merged_output_sharded(path, series_local, counts, lat, lon, time, dist, group=None, dst=0,
                                block_bytes=64 << 20, replaced_as_flags=True, **kwargs):
    """The final host gather of a multi-rank run, straight into the merged file: every rank holds the series of its
    contiguous cell range (``series_local``: name -> [T, counts[rank]], the ``cfact`` / ``replaced`` of
    ``detrend_cells``); block of days by block of days the ranks' slices are gathered to rank ``dst`` over ``group``
    (host tensors: a gloo group, e.g. ``dist.new_group(backend="gloo")`` next to the NCCL one) and appended to a
    ``MergedOutputWriter`` there, so that no rank ever holds more than one block of the full ``[T, C]`` series.

    ``lat`` / ``lon`` / ``time`` and the keyword arguments (``scalars``, ``shared``, ``attrs``, ...: full-size, as for
    ``write_merged_output``) are only read on ``dst``.  Collective: all ranks of ``group`` must call it with the same
    variable names.  Integer ``replaced`` codes are stored as the reference's float64 flags unless
    ``replaced_as_flags=False``.  Returns the number of days written (on ``dst``) or None."""
    import torch

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [int(c) for c in counts]
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    names = sorted(series_local)
    local = {}
    for k in names:
        t = series_local[k]
        if not isinstance(t, (torch.Tensor, ColumnBlocks, SparseReplaced)):
            t = torch.from_numpy(np.ascontiguousarray(t))
        if t.dim() != 2 or t.shape[1] != counts[rank] or t.is_cuda:
            raise ValueError(f"series {k} must be a host array [T, {counts[rank]}] on rank {rank}")
        local[k] = t
    T = int(np.asarray(time).size) if rank == dst else 0
    n_days = torch.tensor([T if rank == dst else 0], dtype=torch.int64)
    dist.broadcast(n_days, src=dst if group is None else dist.get_global_rank(group, dst), group=group)
    T = int(n_days.item())
    for k in names:
        if local[k].shape[0] != T:
            raise ValueError(f"series {k} has {local[k].shape[0]} days on rank {rank}, the time axis {T}")
    cmax, n_cells = max(counts), sum(counts)
    row_bytes = max(1, sum(local[k].element_size() for k in names) * max(n_cells, 1))
    step = max(1, int(block_bytes) // row_bytes)
    writer = None
    if rank == dst:
        shared = {k: np.asarray(v) for k, v in (kwargs.pop("shared", None) or {}).items()}
        time = np.asarray(time)
        np_dtype = {k: (local[k].numpy_dtype() if hasattr(local[k], "numpy_dtype") else local[k][:0].numpy().dtype)
                    for k in names}
        as_flags = {k for k in names if k == "replaced" and replaced_as_flags and np_dtype[k].kind in "iub"}
        writer = MergedOutputWriter(path, lat, lon, {k: np.dtype(np.float64) if k in as_flags else np_dtype[k]
                                                     for k in names},
                                    shared={k: v.dtype for k, v in shared.items()}, time_dtype=time.dtype,
                                    block_bytes=block_bytes, **kwargs)
    try:
        for a in range(0, T, step):
            b = min(T, a + step)
            block = {}
            for k in names:
                part = local[k][a:b]
                if counts[rank] < cmax:                      # gather needs equal shapes: pad the short shards
                    pad = part.new_zeros((b - a, cmax))
                    pad[:, : counts[rank]] = part
                    part = pad
                part = part.contiguous()
                bucket = [torch.empty_like(part) for _ in range(world)] if rank == dst else None
                dist.gather(part, bucket, dst=dst if group is None else dist.get_global_rank(group, dst), group=group)
                if rank == dst:
                    block[k] = torch.cat([t[:, :c] for t, c in zip(bucket, counts)], dim=1).numpy()
                    if k in as_flags:                        # the reference's float flags (detrend.py:397-410)
                        block[k] = replaced_flags(block[k])
            if rank == dst:
                writer.append(time[a:b], block, {k: v[a:b] for k, v in shared.items()})
    finally:
        if writer is not None:
            writer.close()
    return T if rank == dst else None


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


def provenance_attrs(config=None, **other_metadata):
    """Global attributes of the output files, the keys of ``get_data_provenance_metadata`` (attrici/util.py:17-41):
    package version, installed distributions, Python version, creation time, and ``attrici_config`` = the run's
    configuration as TOML ``key = value`` lines (``Config.to_toml``, attrici/detrend.py:106-116: ``None`` omitted)."""
    import dataclasses
    import importlib.metadata
    import sys
    from datetime import datetime

    from . import __version__

    res = {
        "attrici_version": f"attrici_b200 {__version__}",
        "attrici_packages": "\n".join(sorted({f"{d.name}=={d.version}" for d in importlib.metadata.distributions()},
                                             key=str.casefold)),
        "attrici_python_version": sys.version,
        "created_at": datetime.now().isoformat(),
    }
    if config is not None:
        items = dataclasses.asdict(config) if dataclasses.is_dataclass(config) else dict(config)

        def toml(v):
            if isinstance(v, bool):
                return "true" if v else "false"
            if isinstance(v, (int, float)):
                return repr(v)
            if isinstance(v, (list, tuple)):
                return "[" + ", ".join(toml(x) for x in v) + "]"
            return '"' + str(v).replace("\\", "\\\\").replace('"', '\\"') + '"'

        res["attrici_config"] = "\n".join(f"{k} = {toml(v)}" for k, v in items.items() if v is not None) + "\n"
    res.update(other_metadata)
    return res


def dense_replaced(entries, n_time, n_cells):
    """[T, C] uint8 plane of replacement codes from the sparse ``(day, cell, code)`` rows of
    ``CellBatchSolver.detrend_host(replaced="sparse")``."""
    plane = np.zeros((n_time, n_cells), dtype=np.uint8)
    e = np.asarray(entries)
    if e.size:
        plane[e[:, 0], e[:, 1]] = e[:, 2].astype(np.uint8)
    return plane


def replaced_flags(codes):
    """The library's ``replaced`` codes (``ATTRICI_REPLACED_*``: 0 none, 1 NaN, 2 +Inf, 3 -Inf) as the float64 flags
    the reference stores (attrici/detrend.py:396-409: 0, NaN, +Inf, -Inf; ``_FillValue`` 0)."""
    table = np.array([0.0, np.nan, np.inf, -np.inf])
    return table[np.asarray(codes).astype(np.intp)]


def cell_paths(output_dir, variable, lat, lon):
    """``(timeseries file, trace file)`` of one cell, the reference's layout (attrici/detrend.py:236-242, 287-293):
    ``<out>/timeseries/<var>/lat_<lat:g>/ts_lat<lat:g>_lon<lon:g>.nc`` and
    ``<out>/trace/<var>/lat_<lat:g>/trace_lat<lat:g>_lon<lon:g>.nc``."""
    lat, lon = float(lat), float(lon)
    return (os.path.join(str(output_dir), "timeseries", variable, f"lat_{lat:g}", f"ts_lat{lat:g}_lon{lon:g}.nc"),
            os.path.join(str(output_dir), "trace", variable, f"lat_{lat:g}", f"trace_lat{lat:g}_lon{lon:g}.nc"))


def write_cell_files(output_dir, variable, lat, lon, time, series, logp, params=None, scaling=None, shared=None,
                     cells=None, overwrite=True, attrs=None, **time_kwargs):
    """Per-cell files at the reference's paths, for runs whose downstream tools expect them: for every cell one
    timeseries file with dimensions ``(time, lat, lon)`` of lengths ``(T, 1, 1)`` holding ``series`` (name -> [T, C];
    ``replaced`` codes are stored as the reference's NaN / Inf flags), ``logp(lat, lon)`` and the ``shared`` ``(time,)``
    variables (attrici/detrend.py:447-477), and - if ``params`` is given - one trace file (``write_trace``, :221-255).
    ``cells``: indices to write (default all).  An existing timeseries file is skipped unless ``overwrite``
    (:296-305).  Returns the list of timeseries paths written.  A global run should prefer ``write_merged_output``:
    67 k small files are the reference's scalability wall."""
    spec = create_variable(variable) if isinstance(variable, str) else variable
    lat, lon = np.asarray(lat, dtype=np.float64), np.asarray(lon, dtype=np.float64)
    written = []
    for c in (range(lat.size) if cells is None else cells):
        ts_path, trace_path = cell_paths(output_dir, spec.name, lat[c], lon[c])
        if params is not None:
            os.makedirs(os.path.dirname(trace_path), exist_ok=True)
            write_merged_trace(trace_path, spec, np.asarray(params)[c:c + 1], np.asarray(logp)[c:c + 1],
                               np.asarray(scaling)[c:c + 1], lat[c:c + 1], lon[c:c + 1], attrs=attrs)
        if not series:
            continue
        if os.path.exists(ts_path) and not overwrite:
            continue
        os.makedirs(os.path.dirname(ts_path), exist_ok=True)
        cell = {}
        for k, v in series.items():
            col = (v.numpy() if hasattr(v, "numpy") else np.asarray(v))[:, c:c + 1]
            cell[k] = replaced_flags(col) if k == "replaced" and col.dtype.kind in "iub" else col
        write_merged_output(ts_path, cell, lat[c:c + 1], lon[c:c + 1], time, scalars={"logp": np.asarray(logp)[c:c + 1]},
                            shared=shared, attrs=attrs, **time_kwargs)
        written.append(ts_path)
    return written
