"""Minimal stand-in for ``xarray`` (TEST SCAFFOLDING, not product code).

xarray is not installed in this image. The reference's hot-path modules
(``attrici/variables.py``, ``attrici/estimation/model_scipy.py``,
``attrici/util.py:calc_oscillations``) only ever touch 1-D arrays with a ``time``
coordinate; this module provides exactly that surface so the UNMODIFIED reference
can be imported by ``tests/golden/make_golden.py`` to pin the oracle.
"""

import numpy as np


def _unwrap(x):
    return x.values if isinstance(x, (DataArray, Scalar)) else x


class Scalar(np.lib.mixins.NDArrayOperatorsMixin):
    """0-d result of a reduction; keeps the NumPy scalar dtype (float32 stays float32)."""

    def __init__(self, value):
        self.values = value

    def item(self):
        return self.values.item() if hasattr(self.values, "item") else self.values

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        arrs = [i for i in inputs if isinstance(i, DataArray)]
        res = getattr(ufunc, method)(*[_unwrap(i) for i in inputs], **kwargs)
        if arrs and isinstance(res, np.ndarray) and res.shape == arrs[0].values.shape:
            return DataArray(res, coords={"time": arrs[0]._time})
        if isinstance(res, np.ndarray) and res.ndim > 0:
            return res
        return Scalar(res)

    def __bool__(self):
        return bool(self.values)

    def __float__(self):
        return float(self.values)

    def __repr__(self):
        return f"Scalar({self.values!r})"


class DataArray(np.lib.mixins.NDArrayOperatorsMixin):
    def __init__(self, values, coords=None, dims=None, attrs=None, **kw):
        self.values = np.asarray(values)
        if coords is not None and "time" in coords:
            t = coords["time"]
            self._time = np.asarray(getattr(t, "values", t))
        else:
            self._time = np.arange(self.values.shape[0])
        self.attrs = dict(attrs or {})
        self.dims = ("time",)

    # --- numpy protocol -------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        t = next(i._time for i in inputs if isinstance(i, DataArray))
        res = getattr(ufunc, method)(*[_unwrap(i) for i in inputs], **kwargs)
        if isinstance(res, np.ndarray) and res.shape == t.shape:
            return DataArray(res, coords={"time": t})
        if isinstance(res, np.ndarray) and res.ndim > 0:
            return res
        return Scalar(res)

    def __len__(self):
        return len(self.values)

    def __iter__(self):
        return iter(self.values)

    @property
    def shape(self):
        return self.values.shape

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def time(self):
        return DataArray(self._time, coords={"time": self._time})

    def __getattr__(self, name):
        if name == "units":
            try:
                return self.__dict__["attrs"]["units"]
            except KeyError:
                raise AttributeError(name)
        raise AttributeError(name)

    # --- selection ------------------------------------------------------
    def sel(self, time):
        tv = np.asarray(_unwrap(time))
        if tv.dtype == bool:
            return DataArray(self.values[tv], coords={"time": self._time[tv]}, attrs=self.attrs)
        idx = np.searchsorted(self._time, tv)
        if not np.array_equal(self._time[idx], tv):
            raise KeyError("time values not found")
        return DataArray(self.values[idx], coords={"time": tv}, attrs=self.attrs)

    def __getitem__(self, key):
        key = _unwrap(key)
        res = self.values[key]
        if isinstance(res, np.ndarray) and res.ndim == 1:
            return DataArray(res, coords={"time": self._time[key]}, attrs=self.attrs)
        return Scalar(res)

    def __setitem__(self, key, value):
        self.values[_unwrap(key)] = _unwrap(value)

    # --- reductions / predicates ---------------------------------------
    def min(self):
        v = self.values
        return Scalar(np.nanmin(v) if v.dtype.kind == "f" else v.min())

    def max(self):
        v = self.values
        return Scalar(np.nanmax(v) if v.dtype.kind == "f" else v.max())

    def sum(self):
        return Scalar(np.nansum(self.values))

    def any(self):
        return Scalar(np.any(self.values))

    def notnull(self):
        return DataArray(~np.isnan(self.values), coords={"time": self._time})

    def isnull(self):
        return DataArray(np.isnan(self.values), coords={"time": self._time})

    def isin(self, vals):
        return DataArray(np.isin(self.values, vals), coords={"time": self._time})

    def item(self):
        return self.values.item()

    def copy(self):
        return DataArray(self.values.copy(), coords={"time": self._time.copy()}, attrs=self.attrs)

    def astype(self, dtype):
        return DataArray(self.values.astype(dtype), coords={"time": self._time}, attrs=self.attrs)

    def __repr__(self):
        return f"DataArray({self.values!r})"


class Dataset:  # only referenced in type positions by the modules we import
    pass
