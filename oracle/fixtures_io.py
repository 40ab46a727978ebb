"""Byte-level readers for the reference's golden fixtures (TEST INFRASTRUCTURE ONLY).

Neither h5py/netCDF4 nor xarray exist in this image, so the two fixture
containers the reference's tests use are parsed directly:

* ``tests/data/20CRv3-ERA5_germany_target_<var>_lat<lat>_lon<lon>.h5`` - pandas
  "fixed" HDFStore frames with contiguous, unfiltered datasets
  (columns ``ds`` int64 ns, ``y`` float32, ``[cfact, logp]`` float64 pairs); these are
  the targets of reference ``tests/test_detrend_pymc.py:63-160``.
* ``tests/data/20CRv3-ERA5_germany_ssa_gmt.nc`` - netCDF-4 with two chunked
  (512 x f64, unfiltered) datasets indexed by version-1 B-trees.

* ``tests/data/20crv3-era5_gmt_raw.nc`` - netCDF-4, ``tas(time, lat, lon)`` float32 with one value per
  chunk (unlimited time axis), unfiltered, indexed by a version-1 B-tree; the input of reference
  ``tests/test_ssa.py:27-44``.

Only used by ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` exists) to produce the committed ``tests/golden/*.npz``.
"""

import re
import struct

import numpy as np

N_ROWS = 44925  # 1901-01-01 ... 2023-12-31


def read_target_h5(path, n_rows=N_ROWS):
    """Return dict(ds=datetime64[ns], y=float32, cfact=float64, logp=float64)."""
    buf = open(path, "rb").read()
    out = {}
    for m in re.finditer(rb"\x08\x00\x18\x00\x00\x00\x00\x00\x03\x01", buf):
        addr, size = struct.unpack_from("<QQ", buf, m.end())
        if size < n_rows * 4 or size % n_rows:
            continue
        width = size // n_rows
        raw = buf[addr : addr + size]
        if width == 4:
            out["y"] = np.frombuffer(raw, dtype="<f4").copy()
        elif width == 16:
            pairs = np.frombuffer(raw, dtype="<f8").reshape(n_rows, 2)
            out["cfact"] = pairs[:, 0].copy()
            out["logp"] = pairs[:, 1].copy()
        elif width == 8:
            v = np.frombuffer(raw, dtype="<i8")
            if np.abs(v).max() > 10**15:
                out["ds"] = v.astype("datetime64[ns]")
    missing = {"y", "cfact", "logp", "ds"} - set(out)
    if missing:
        raise ValueError(f"{path}: columns not found: {missing}")
    return out


def read_chunked_f64_nc(path, n_points=4493):
    """Return the 1-D float64 datasets of a small netCDF-4 file, in file order.

    Walks every version-1 chunk B-tree leaf node (``TREE`` signature, node type 1,
    level 0) and copies each 512-element chunk to its offset.
    """
    buf = open(path, "rb").read()
    datasets = []
    for m in re.finditer(rb"TREE", buf):
        pos = m.start()
        node_type, level, used = struct.unpack_from("<BBH", buf, pos + 4)
        if node_type != 1 or level != 0 or used == 0 or used > 64:
            continue
        # header: sig(4) type(1) level(1) used(2) left(8) right(8) = 24 bytes
        # key: chunk_size(4) filter_mask(4) offset[ndims+1](8 each); child addr (8)
        arr = np.full(512 * used, np.nan)
        ok = True
        p = pos + 24
        for _ in range(used):
            csize, mask, off0, off1, addr = struct.unpack_from("<IIQQQ", buf, p)
            p += 32
            if csize != 4096 or mask != 0 or off1 != 0 or addr + csize > len(buf):
                ok = False
                break
            arr[off0 : off0 + 512] = np.frombuffer(buf[addr : addr + 4096], dtype="<f8")
        if ok:
            datasets.append((pos, arr[:n_points].copy()))
    datasets.sort(key=lambda t: t[0])
    return [a for _, a in datasets]


def read_ssa_gmt(path):
    """Return (time_days_since_1900, tas_K) of the SSA-smoothed GMT fixture."""
    ds = read_chunked_f64_nc(path)
    if len(ds) != 2:
        raise ValueError(f"expected 2 datasets, found {len(ds)}")
    a, b = ds
    # time is the evenly spaced one
    if np.allclose(np.diff(a), np.diff(a)[0]):
        return a, b
    return b, a


def read_raw_gmt(path):
    """Return the daily raw GMT series ``tas[time]`` (float32) of ``20crv3-era5_gmt_raw.nc``.

    Walks the leaf nodes of the 3-D chunk B-tree: key = chunk size (4 bytes), filter mask, offsets
    (time, 0, 0, 0); one float32 per chunk.
    """
    buf = open(path, "rb").read()
    vals = {}
    for m in re.finditer(rb"TREE", buf):
        pos = m.start()
        node_type, level, used = struct.unpack_from("<BBH", buf, pos + 4)
        if node_type != 1 or level != 0 or used == 0:
            continue
        p = pos + 24
        for _ in range(used):
            csize, mask, t, o1, o2, o3, addr = struct.unpack_from("<IIQQQQQ", buf, p)
            p += 48
            if csize != 4 or mask != 0 or o1 or o2 or o3 or addr + 4 > len(buf):
                break
            vals[t] = struct.unpack_from("<f", buf, addr)[0]
    n = max(vals) + 1
    if len(vals) != n:
        raise ValueError(f"{n - len(vals)} chunks missing")
    return np.array([vals[t] for t in range(n)], dtype=np.float32)
