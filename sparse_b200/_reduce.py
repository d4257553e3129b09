"""Reductions (`x.sum(axis)`, `np.max(x, axis)`, `x.reduce(ufunc, axis)`).

Host-side mirror of SparseArray.reduce (sparse/numba_backend/_sparse_array.py:372-437),
COO._reduce_calc/_reduce_return (_coo/core.py:693-723) and GCXS._reduce_calc/_reduce_return
(_compressed/compressed.py:354-386).  The grouped reduction itself (transpose -> 2-D -> reduceat over runs) runs on
the device: permuted linearisation + stable sort (only when the reduced axes are not trailing), then ONE single-pass
segmented-scan kernel that also applies the fill-value correction (csrc/reduce_fused.cu, `b2s_reduce_single`).
"""
from __future__ import annotations

import numpy as np

from . import _kernels as Kn
from ._coo import COO
from ._sparse_array import _reduce_super_ufunc
from ._utils import equivalent, normalize_axis, prod

_RED = {np.add: 0, np.multiply: 1, np.maximum: 2, np.minimum: 3, np.logical_and: 4, np.logical_or: 5,
        np.bitwise_and: 6, np.bitwise_or: 7, np.bitwise_xor: 8, np.fmax: 9, np.fmin: 10}
_RED_DTYPES = (np.dtype("float32"), np.dtype("float64"), np.dtype("int32"), np.dtype("int64"), np.dtype("bool"))


def reduce_impl(x, method, axis=(0,), keepdims=False, **kwargs):
    from ._gcxs import GCXS

    if x.dtype.kind == "c":
        from ._complex import reduce_complex

        ax = normalize_axis(axis, x.ndim)
        return reduce_complex(x, method, ax, keepdims, **{k: v for k, v in kwargs.items() if v is not None})
    if method not in _RED:
        raise TypeError(f"sparse_b200: reduction with {getattr(method, '__name__', method)!r} is not in the CUDA op set "
                        f"({', '.join(sorted(m.__name__ for m in _RED))}); there is no CPU fallback.")
    axis = normalize_axis(axis, x.ndim)
    red_kwargs = {k: v for k, v in kwargs.items() if v is not None}
    with np.errstate(all="ignore"):
        zero_reduce_result = method.reduce([x.fill_value, x.fill_value], **red_kwargs)
    super_ufunc = _reduce_super_ufunc.get(method)
    if not equivalent(zero_reduce_result, x.fill_value) and super_ufunc is None:
        raise ValueError(f"Performing this reduction operation would produce a dense result: {method!s}")
    if not isinstance(axis, tuple):
        axis = (axis,)
    if axis == (None,):
        axis = tuple(range(x.ndim))
    was_gcxs = isinstance(x, GCXS)
    c = x.tocoo() if was_gcxs else x

    out = _reduce_coo(c, method, axis, super_ufunc, red_kwargs)
    if keepdims:
        shape = list(x.shape)
        for ax in axis:
            shape[ax] = 1
        out = out.reshape(tuple(shape))
    if out.ndim == 0:
        # 0-D result per the Array API: a COO whose single element is its fill value (nnz = 0), for GCXS input too
        return COO.from_numpy(out.todense())
    if was_gcxs and len(set(axis)) < x.ndim:
        return GCXS.from_coo(out, None if out.ndim == 1 else (int(np.argmin(out.shape)),))
    return out  # a reduction over EVERY axis of a GCXS array goes through COO and stays COO (compressed.py:355-360)


_SIGN64 = np.int64(-(2**63))


def _reduce_coo(x, method, axis, super_ufunc, kwargs):
    op = _RED[method]
    neg_axis = tuple(ax for ax in range(x.ndim) if ax not in set(axis))
    nrows = prod(x.shape[d] for d in neg_axis)
    ncols = prod(x.shape[d] for d in axis)
    # result dtype by NumPy's own rules (e.g. add.reduce(int32) -> int64; logical_* -> bool)
    with np.errstate(all="ignore"):
        res_dt = method.reduce(np.zeros(1, dtype=x.dtype), **kwargs).dtype
    narrow_back = None
    if res_dt not in _RED_DTYPES and res_dt.kind in "iu":
        # narrow / unsigned integer results (uint8 max, the uint64 NumPy sums unsigned values into): reduced in int64,
        # cast back at the end (modular like NumPy's own arithmetic in that dtype)
        narrow_back, res_dt = res_dt, np.dtype(np.int64)
    if res_dt not in _RED_DTYPES:
        raise TypeError(f"sparse_b200: reduction dtype {res_dt} is outside the CUDA dtype matrix")
    work_dt = res_dt
    kept_shape = tuple(x.shape[d] for d in neg_axis)
    fill_in = x.fill_value
    # uint64 through int64 is modular-correct for add / multiply only; the ORDER-based reductions see values >= 2**63
    # as negative.  x ^ 2**63 maps the unsigned order onto the signed one: flip before, flip back after.
    flip = narrow_back == np.dtype(np.uint64) and op in (2, 3)
    if flip:
        fill_in = np.asarray(fill_in, dtype=np.uint64).view(np.int64)[()] ^ _SIGN64
    if op in (4, 5):  # logical reductions work on truth values
        fill_w = np.bool_(bool(fill_in))
        work_dt = np.dtype(np.bool_)
    else:
        fill_w = work_dt.type(fill_in)
    result_fill = fill_w
    if super_ufunc is not None:
        with np.errstate(all="ignore"):
            result_fill = np.asarray(super_ufunc(fill_w, ncols)).astype(work_dt)[()]
    if x.nnz == 0:
        out_dt = narrow_back or work_dt
        if flip:
            result_fill = np.int64(result_fill) ^ _SIGN64
        return COO(np.zeros((len(neg_axis), 0), dtype=np.intp), np.empty(0, dtype=out_dt), shape=kept_shape,
                   has_duplicates=False, sorted=True, fill_value=np.asarray(result_fill).astype(out_dt)[()])
    # kept axes first; keys over (kept..., reduced...) so that group id = key // ncols (sort only if not already so)
    keys, data = x._permuted_keys(neg_axis + tuple(axis))
    data = Kn.cast(data, work_dt)
    if flip:
        from ._elemwise import _BINARY

        data, _ = Kn.ew_map(_BINARY[np.bitwise_xor], 0, data, _SIGN64, 0, np.int64)
    # one fused segmented-scan pass pair: values with the fill-value contribution applied, group ids (= linear index
    # over the kept axes) and their coordinates (_grouped_reduce + _sparse_array.py:405-422 + _reduce_return)
    _, gids, vals, n_eq = Kn.reduce_fused(op, keys, data, ncols, fill_w, result_fill, kept_shape, want_coords=False)
    if n_eq:  # prune=True of _reduce_return (_coo/core.py:713-723): drop results equal to the result fill value
        flags = Kn.flag_not_fill(vals, result_fill)
        pos, total = Kn.scan_flags(flags)
        vals = Kn.compact(vals, flags, pos, total)
        gids = Kn.compact(gids, flags, pos, total)
    if flip:
        vals, _ = Kn.ew_map(_BINARY[np.bitwise_xor], 0, vals, _SIGN64, 0, np.int64)
        result_fill = np.int64(result_fill) ^ _SIGN64
    out = COO._from_device(None, vals, kept_shape, result_fill, keys=gids)  # coordinates are derived lazily
    return out.astype(narrow_back, _raw=True) if narrow_back is not None else out
