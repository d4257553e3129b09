"""DOK: the mutable dictionary-of-keys builder next to the device containers.

Mirrors sparse/numba_backend/_dok.py:94-563 (same constructor forms, item get / set rules, conversions).  Upstream's
DOK is a plain Python dict as well: it exists to ASSEMBLE an array entry by entry on the host and is converted to COO
for every computation.  The same holds here -- a DOK holds `{index tuple: scalar}` on the host; `to_coo()` uploads it
once, and every operation (arithmetic, reductions, products, slicing) runs on the COO through the device path, the
result being handed back as DOK where upstream returns one.  No arithmetic is done on the dictionary.
"""
from __future__ import annotations

import operator
from collections.abc import Iterable
from numbers import Integral

import numpy as np

from ._coo import COO, _is_scipy_sparse
from ._sparse_array import SparseArray
from ._utils import _zero_of_dtype, equivalent


class DOK(SparseArray):
    def __init__(self, shape, data=None, dtype=None, fill_value=None):
        self.data = {}
        src = None
        if isinstance(shape, COO):
            src = DOK.from_coo(shape)
        elif isinstance(shape, SparseArray) and not isinstance(shape, DOK):
            src = DOK.from_coo(shape.asformat("coo"))
        elif isinstance(shape, np.ndarray):
            src = DOK.from_numpy(shape)
        elif _is_scipy_sparse(shape):
            src = DOK.from_scipy_sparse(shape)
        if src is not None:
            self._make_shallow_copy_of(src)
            return
        if data is not None and not isinstance(data, dict) and data:
            raise ValueError("data must be a dict.")
        data = data or {}
        if dtype is None:
            dtype = np.result_type(*(np.asarray(v).dtype for v in data.values())) if len(data) else np.dtype("float64")
        self.dtype = np.dtype(dtype)
        SparseArray.__init__(self, shape, fill_value=None)
        self.fill_value = self.dtype.type(_zero_of_dtype(self.dtype) if fill_value is None else fill_value)
        for key, value in data.items():
            self[key] = value

    # ---- conversions -------------------------------------------------------------------------------------------------
    def _make_shallow_copy_of(self, other):
        self.data, self.dtype = other.data, other.dtype
        SparseArray.__init__(self, other.shape, fill_value=None)
        self.fill_value = other.fill_value

    @classmethod
    def from_coo(cls, x):
        """_dok.py:164-191: one D2H of the coordinate and value arrays, then the dictionary."""
        out = cls(x.shape, dtype=x.dtype, fill_value=x.fill_value)
        coords, values = x.coords, x.data
        out.data = {tuple(int(c) for c in coords[:, i]): values[i] for i in range(values.shape[0])}
        return out

    @classmethod
    def from_numpy(cls, x):
        x = np.asanyarray(x)
        out = cls(x.shape, dtype=x.dtype)
        idx = np.nonzero(~equivalent(x, out.fill_value)) if x.ndim else ()
        if x.ndim == 0:
            if not equivalent(x, out.fill_value):
                out.data[()] = x[()]
            return out
        vals = x[idx]
        out.data = {tuple(int(c[i]) for c in idx): vals[i] for i in range(vals.shape[0])}
        return out

    @classmethod
    def from_scipy_sparse(cls, x, /, *, fill_value=None):
        out = cls.from_coo(COO.from_scipy_sparse(x))
        if fill_value is not None:
            out.fill_value = out.dtype.type(fill_value)
        return out

    def to_coo(self):
        """Canonical COO of the dictionary (keys sorted on the host, one upload)."""
        keys = sorted(self.data)
        coords = np.array(keys, dtype=np.intp).T.reshape(self.ndim, len(keys)) if keys else \
            np.zeros((self.ndim, 0), dtype=np.intp)
        values = np.array([self.data[k] for k in keys], dtype=self.dtype)
        return COO(coords, values, shape=self.shape, has_duplicates=False, sorted=True, fill_value=self.fill_value)

    tocoo = to_coo

    def asformat(self, format, **kwargs):
        if isinstance(format, str) and format == "dok" or format is DOK:
            return self
        if isinstance(format, type) and issubclass(format, DOK):
            return self
        if (isinstance(format, str) and format == "coo") or format is COO:
            if kwargs:
                raise ValueError(f"Extra kwargs found: {kwargs}")
            return self.to_coo()
        return self.to_coo().asformat(format, **kwargs)

    def todense(self):
        return self.to_coo().todense()

    def copy(self, deep=True):
        out = DOK(self.shape, dtype=self.dtype, fill_value=self.fill_value)
        out.data = dict(self.data) if deep else self.data
        return out

    def reshape(self, shape, order="C"):
        if order not in {"C", None}:
            raise NotImplementedError("The 'order' parameter is not supported")
        return DOK.from_coo(self.to_coo().reshape(shape))

    # ---- metadata -------------------------------------------------------------------------------------------------------
    @property
    def nnz(self):
        return len(self.data)

    @property
    def format(self):
        return "dok"

    @property
    def nbytes(self):
        return self.nnz * self.dtype.itemsize

    # ---- items ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def _is_index_sequence(k):
        return isinstance(k, Iterable) and not isinstance(k, (str, bytes))

    def __getitem__(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if key and all(self._is_index_sequence(k) for k in key):
            if len(key) != self.ndim:
                raise NotImplementedError(f"Index sequences for all {self.ndim} array dimensions needed!")
            if not all(len(key[0]) == len(k) for k in key):
                raise IndexError("Unequal length of index sequences!")
            picked = {i: self.data[k] for i, k in enumerate(zip(*[[int(v) for v in ks] for ks in key], strict=True))
                      if k in self.data}
            return DOK(shape=(len(key[0]),), data=picked, dtype=self.dtype, fill_value=self.fill_value)
        out = self.to_coo()[key]
        return DOK.from_coo(out) if isinstance(out, SparseArray) else out

    def __setitem__(self, key, value):
        value = np.asarray(value, dtype=self.dtype)
        if self.ndim == 1 and self._is_index_sequence(key) and not isinstance(key, tuple) \
                and all(isinstance(i, (int, np.integer)) for i in key):
            key = (key,)
        if isinstance(key, tuple) and key and all(self._is_index_sequence(k) for k in key):
            if len(key) != self.ndim:
                raise NotImplementedError(f"Index sequences for all {self.ndim} array dimensions needed!")
            if not all(len(key[0]) == len(k) for k in key):
                raise IndexError("Unequal length of index sequences!")
            self._set_points(key, value)
            return
        if not isinstance(key, tuple):
            key = (key,)
        key = self._expand(key)
        ranges = []
        for axis, (ind, extent) in enumerate(zip(key, self.shape)):
            if isinstance(ind, slice):
                ranges.append(range(*ind.indices(extent)))
            elif isinstance(ind, Integral) and not isinstance(ind, bool):
                i = operator.index(ind)
                if not -extent <= i < extent:
                    raise IndexError(f"index {i} is out of bounds for axis {axis} with size {extent}")
                ranges.append(i + extent if i < 0 else i)
            else:
                raise IndexError("All indices must be slices or integers when setting an item.")
        sel_shape = tuple(len(r) for r in ranges if isinstance(r, range))
        if value.ndim > len(sel_shape):
            raise ValueError("setting an array element with a sequence.")
        block = np.broadcast_to(value, sel_shape)
        for pos in np.ndindex(*sel_shape):
            it = iter(pos)
            idx = tuple(r[next(it)] if isinstance(r, range) else r for r in ranges)
            self._put(idx, block[pos])

    def _expand(self, key):
        """Ellipsis and missing trailing axes filled with full slices."""
        if any(k is Ellipsis for k in key):
            at = [i for i, k in enumerate(key) if k is Ellipsis]
            if len(at) > 1:
                raise IndexError("an index can only have a single ellipsis ('...')")
            fill = (slice(None),) * (self.ndim - (len(key) - 1))
            key = key[:at[0]] + fill + key[at[0] + 1:]
        if len(key) > self.ndim:
            raise IndexError(f"too many indices for array: array is {self.ndim}-dimensional, but {len(key)} were "
                             "indexed")
        return key + (slice(None),) * (self.ndim - len(key))

    def _set_points(self, idxs, values):
        idxs = tuple(np.asanyarray(i) for i in idxs)
        if not all(np.issubdtype(i.dtype, np.integer) for i in idxs):
            raise IndexError("Indices must be sequences of integer types!")
        if idxs[0].ndim != 1:
            raise IndexError("Indices are not 1d sequences!")
        if values.ndim == 0:
            values = np.full(idxs[0].size, values, self.dtype)
        elif values.ndim > 1:
            raise ValueError(f"Dimension of values ({values.ndim}) must be 0 or 1!")
        if idxs[0].shape != values.shape:
            raise ValueError(f"Shape mismatch of indices ({idxs[0].shape}) and values ({values.shape})!")
        for idx, v in zip(zip(*[[int(c) for c in i] for i in idxs], strict=True), values, strict=True):
            self._put(idx, v)

    def _put(self, idx, value):
        if not equivalent(value, self.fill_value):
            self.data[idx] = value[()] if isinstance(value, np.ndarray) else value
        else:
            self.data.pop(idx, None)

    def __repr__(self):
        return f"<DOK: shape={self.shape!s}, dtype={self.dtype!s}, nnz={self.nnz:d}, fill_value={self.fill_value!s}>"

    __str__ = __repr__

    # ---- computations go through COO ---------------------------------------------------------------------------------------
    def reduce(self, method, axis=(0,), keepdims=False, **kwargs):
        out = self.to_coo().reduce(method, axis=axis, keepdims=keepdims, **kwargs)
        return DOK.from_coo(out) if isinstance(out, SparseArray) else out

    def astype(self, dtype, casting="unsafe", copy=True):
        return DOK.from_coo(self.to_coo().astype(dtype, casting=casting, copy=copy))

    def transpose(self, axes=None):
        return DOK.from_coo(self.to_coo().transpose(axes))

    @property
    def T(self):
        return self.transpose()

    def _make_result(self, coo):
        return DOK.from_coo(coo)
