"""NaN-skipping reductions: nansum, nanprod, nanmean, nanmax, nanmin, nanreduce.

Host-side mirror of sparse/numba_backend/_coo/common.py:346-531 (the five public functions), :418-428
(`_contains_nan`), :674-693 (`_replace_nan`) and :696-732 (`nanreduce`).  The reference replaces NaN through a
three-operand `where(isnan(x), value, x)`; here that is ONE map kernel (binary op `nan_replace` of csrc/elemwise.cu
applied to the stored values, pruning what became the fill value), followed by the same fused segmented reduction
as the plain reductions (csrc/reduce_fused.cu; `fmax` / `fmin` are reduction ops there).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _kernels as Kn
from ._coo import COO
from ._sparse_array import SparseArray


def _as_coo(x, name):
    """asCOO (_coo/common.py:21-52): sparse inputs only; anything but COO is converted."""
    from ._coo import _is_scipy_sparse

    if not isinstance(x, SparseArray) and not _is_scipy_sparse(x):
        raise ValueError(f"Performing this operation would produce a dense result: {name}")
    if isinstance(x, COO):
        return x
    if isinstance(x, SparseArray):
        return x.asformat("coo")
    return COO.from_scipy_sparse(x)


def _replace_nan(array, value):
    """_coo/common.py:674-693."""
    from ._elemwise import elemwise, nan_replace

    if not np.issubdtype(array.dtype, np.floating):
        return array
    return elemwise(nan_replace, array, value)


def _contains_nan(ar):
    """_coo/common.py:418-428: dtype first, then the fill value, then the stored values (one flag kernel)."""
    if isinstance(ar, SparseArray):
        if not np.issubdtype(ar.dtype, np.floating):
            return False
        if ar.nnz != ar.size and np.isnan(ar.fill_value):
            return True
        if ar.nnz == 0:
            return False
        c = ar if isinstance(ar, COO) else ar.asformat("coo")
        return bool(Kn.any_nan(c._data_dev()))
    return bool(np.isnan(ar))


def nanreduce(x, method, identity=None, axis=None, keepdims=False, **kwargs):
    """_coo/common.py:696-732."""
    arr = _replace_nan(x, method.identity if identity is None else identity)
    return arr.reduce(method, axis, keepdims, **kwargs)


def nansum(x, axis=None, keepdims=False, dtype=None, out=None):
    """_coo/common.py:346-361."""
    assert out is None
    x = _as_coo(x, "nansum")
    return nanreduce(x, np.add, axis=axis, keepdims=keepdims, dtype=dtype)


def nanprod(x, axis=None, keepdims=False, dtype=None, out=None):
    """_coo/common.py:501-531."""
    assert out is None
    x = _as_coo(x, "nanprod")
    return nanreduce(x, np.multiply, axis=axis, keepdims=keepdims, dtype=dtype)


def nanmean(x, axis=None, keepdims=False, dtype=None, out=None):
    """Mean over the non-NaN elements of every slice (_coo/common.py:364-415 upstream): the NaN-free sum divided by
    the number of valid elements, with NumPy's dtype rules for the division; integer input has no NaNs."""
    assert out is None
    x = _as_coo(x, "nanmean")
    if x.dtype.kind != "f":
        return x.mean(axis=axis, keepdims=keepdims, dtype=dtype)
    red = tuple(range(x.ndim)) if axis is None else (axis if isinstance(axis, tuple) else (axis,))
    slice_size = 1
    for a in red:
        slice_size *= x.shape[a]
    valid = slice_size - np.isnan(x).sum(axis=axis, dtype=np.int64, keepdims=keepdims)  # per-slice count (sparse)
    if (valid == 0).any():
        warnings.warn("Mean of empty slice", RuntimeWarning, stacklevel=1)
    total = _replace_nan(x, 0).sum(axis=red, dtype=dtype, keepdims=keepdims)
    with np.errstate(invalid="ignore", divide="ignore"):
        if total.ndim:
            return np.true_divide(total, valid, casting="unsafe")
        return (total / valid).astype(x.dtype if dtype is None else dtype)


def _nan_extreme(x, method, name, axis, keepdims, dtype):
    x = _as_coo(x, name)
    ar = x.reduce(method, axis=axis, keepdims=keepdims, dtype=dtype)
    if _contains_nan(ar):
        warnings.warn("All-NaN slice encountered", RuntimeWarning, stacklevel=2)
    return ar


def nanmax(x, axis=None, keepdims=False, dtype=None, out=None):
    """_coo/common.py:431-464: reduce with np.fmax + the all-NaN warning."""
    assert out is None
    return _nan_extreme(x, np.fmax, "nanmax", axis, keepdims, dtype)


def nanmin(x, axis=None, keepdims=False, dtype=None, out=None):
    """_coo/common.py:467-500: reduce with np.fmin + the all-NaN warning."""
    assert out is None
    return _nan_extreme(x, np.fmin, "nanmin", axis, keepdims, dtype)
