"""SparseArray base: NumPy protocol dispatch and reductions.

Mirrors sparse/numba_backend/_sparse_array.py: `__array_ufunc__` (:322-370) routes ufunc calls to
`elemwise` and ufunc.reduce to `reduce` (:372-437); `__array_function__` (:282-308) routes
np.tensordot / np.matmul / np.dot / np.sum ... to this package's functions.
"""
from __future__ import annotations

import operator
from functools import reduce as _functools_reduce

import numpy as np
from numpy.lib.mixins import NDArrayOperatorsMixin

from ._utils import _zero_of_dtype, equivalent, normalize_axis

_reduce_super_ufunc = {np.add: np.multiply, np.multiply: np.power}


class SparseArray(NDArrayOperatorsMixin):
    __array_priority__ = 12

    def __init__(self, shape, fill_value=None):
        if not isinstance(shape, (tuple, list, np.ndarray)) and shape is not None:
            shape = (shape,)
        if not all(isinstance(s, (int, np.integer)) and int(s) >= 0 for s in shape):
            raise ValueError("shape must be an non-negative integer or a tuple of non-negative integers.")
        self.shape = tuple(int(s) for s in shape)
        if fill_value is not None:
            self.fill_value = np.asarray(fill_value)[()] if not isinstance(fill_value, np.generic) else fill_value

    # ---- shape metadata --------------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return _functools_reduce(operator.mul, self.shape, 1)

    @property
    def density(self):
        return self.nnz / self.size if self.size else 0.0

    @property
    def device(self):
        return "cuda"

    def __len__(self):
        if self.ndim == 0:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def __array__(self, dtype=None, copy=None):
        x = self.todense()
        return x.astype(dtype) if dtype is not None and x.dtype != dtype else x

    def __repr__(self):
        return (f"<{type(self).__name__}: shape={self.shape}, dtype={self.dtype}, nnz={self.nnz}, "
                f"fill_value={self.fill_value}>")

    __str__ = __repr__

    # ---- NumPy protocols ---------------------------------------------------------------------------
    def __array_function__(self, func, types, args, kwargs):
        import sparse_b200 as module

        name = func.__name__
        sparse_func = getattr(module, name, None)
        if sparse_func is None or sparse_func is func:
            sparse_func = getattr(type(self), name, None)
        if sparse_func is None or not callable(sparse_func):
            return NotImplemented
        return sparse_func(*args, **kwargs)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        from ._elemwise import elemwise

        out = kwargs.pop("out", None)
        if out is not None:
            raise NotImplementedError("sparse_b200: the `out=` argument of ufuncs is not supported on the CUDA path")
        if getattr(ufunc, "signature", None) is not None:
            return self.__array_function__(ufunc, (np.ndarray, type(self)), inputs, kwargs)
        if method == "outer":
            method = "__call__"
            cum_ndim = 0
            inputs_transformed = []
            for inp in inputs:
                inputs_transformed.append(inp[(Ellipsis,) + (None,) * cum_ndim])
                cum_ndim += inp.ndim
            inputs = tuple(inputs_transformed)
        if method == "__call__":
            result = elemwise(ufunc, *inputs, **kwargs)
        elif method == "reduce":
            result = SparseArray._reduce(ufunc, *inputs, **kwargs)
        else:
            return NotImplemented
        return result

    # ---- reductions (reference: _sparse_array.py:310-437) ------------------------------------------
    @staticmethod
    def _reduce(method, *args, **kwargs):
        assert len(args) == 1
        self = args[0]
        if isinstance(self, np.ndarray):
            return method.reduce(self, **kwargs)
        return self.reduce(method, **kwargs)

    def reduce(self, method, axis=(0,), keepdims=False, **kwargs):
        from ._reduce import reduce_impl

        return reduce_impl(self, method, axis=axis, keepdims=keepdims, **kwargs)

    def sum(self, axis=None, keepdims=False, dtype=None, out=None):
        return np.add.reduce(self, out=out, axis=axis, keepdims=keepdims, dtype=dtype)

    def max(self, axis=None, keepdims=False, out=None):
        return np.maximum.reduce(self, out=out, axis=axis, keepdims=keepdims)

    amax = max

    def min(self, axis=None, keepdims=False, out=None):
        return np.minimum.reduce(self, out=out, axis=axis, keepdims=keepdims)

    amin = min

    def prod(self, axis=None, keepdims=False, dtype=None, out=None):
        return np.multiply.reduce(self, out=out, axis=axis, keepdims=keepdims, dtype=dtype)

    def any(self, axis=None, keepdims=False, out=None):
        return np.logical_or.reduce(self, out=out, axis=axis, keepdims=keepdims)

    def all(self, axis=None, keepdims=False, out=None):
        return np.logical_and.reduce(self, out=out, axis=axis, keepdims=keepdims)

    def mean(self, axis=None, keepdims=False, dtype=None, out=None):
        """_sparse_array.py:mean -- sum / n with NumPy's dtype rules."""
        if axis is None:
            axis = tuple(range(self.ndim))
        elif not isinstance(axis, tuple):
            axis = (axis,)
        den = _functools_reduce(operator.mul, (self.shape[i] for i in axis), 1)
        if dtype is None:
            if issubclass(self.dtype.type, (np.integer, np.bool_)):
                dtype = inter_dtype = np.dtype("f8")
            else:
                dtype = self.dtype
                inter_dtype = np.dtype("f4") if issubclass(dtype.type, np.float16) else dtype
        else:
            inter_dtype = dtype
        num = self.sum(axis=axis, keepdims=keepdims, dtype=inter_dtype)
        if num.ndim:
            out = np.true_divide(num, den)
        else:
            out = (num / den) if not isinstance(num, SparseArray) else np.true_divide(num, den)
        return out.astype(dtype) if hasattr(out, "astype") else out

    # ---- scalar conversion (_sparse_array.py:970-993) ---------------------------------------------------
    def _to_scalar(self, builtin):
        if self.size != 1 or self.shape != ():
            raise ValueError(f"{builtin} can be computed for one-element arrays only.")
        return builtin(self.todense().flatten()[0])

    def __bool__(self):
        return self._to_scalar(bool)

    def __float__(self):
        return self._to_scalar(float)

    def __int__(self):
        return self._to_scalar(int)

    def __index__(self):
        return self._to_scalar(int)

    # ---- misc -----------------------------------------------------------------------------------------
    @property
    def real(self):
        return self

    def _zero_fill(self):
        return bool(equivalent(self.fill_value, _zero_of_dtype(self.dtype), loose=True))

    def _norm_axis(self, axis):
        return normalize_axis(axis, self.ndim)
