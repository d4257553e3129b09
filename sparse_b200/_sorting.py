"""Value sorting on the device: `sort`, `unique_values`, `unique_counts`.

Reference: sparse/numba_backend/_coo/common.py:1189-1346, 1401-1452 (`np.unique` on the host data array; `_sort_coo`: a
numba loop that `np.sort`s every group).  Here values are turned into order-preserving unsigned 64-bit keys with three
element-wise passes (IEEE bits: flip all bits of negatives, the sign bit of non-negatives), sorted by the device radix
sort (`b2s_sort_keys`, stable LSD) and -- for `sort` -- sorted once more by group id, which the stability of the radix sort
turns into "by (group, value)".  Positions inside a group follow from an entry's rank (position minus the group start)
and the place the block of fill values takes in the order.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO
from ._sparse_array import SparseArray
from ._utils import key_bits

_MIN64 = np.int64(-(2**63))
UniqueCountsResult = namedtuple("UniqueCountsResult", ["values", "counts"])


def _validate(x):
    from ._coo import _is_scipy_sparse

    if _is_scipy_sparse(x):
        return COO.from_scipy_sparse(x)
    if not isinstance(x, SparseArray):
        raise ValueError(f"Input must be an instance of SparseArray, but it's {type(x)}.")
    return x.asformat("coo")


def _order_keys(data, descending=False):
    """int64 keys whose UNSIGNED order is the order of `data` (NaNs last, every NaN distinct)."""
    from ._elemwise import _BINARY, _UNARY, dense_binary

    t = D.torch()
    dt = D.np_dtype(data)
    if dt.kind == "f":
        if dt.itemsize not in (4, 8):
            raise TypeError(f"sparse_b200: sorting {dt} values is outside the CUDA dtype matrix")
        bits = data.view(t.int64) if dt.itemsize == 8 else Kn.cast(data.view(t.int32), np.int64)
        sign, _ = Kn.ew_map(_BINARY[np.right_shift], 0, bits, np.int64(63), 0, np.int64)      # -1 for negatives, else 0
        mask, _ = Kn.ew_map(_BINARY[np.bitwise_or], 0, sign, _MIN64, 0, np.int64)             # all ones / sign bit only
        keys = dense_binary(np.bitwise_xor, bits, mask)
        if Kn.any_nan(data):  # NaN != NaN: spread equal NaN patterns apart so that each one stays a separate value
            is_nan, _ = Kn.ew_map(_UNARY[np.isnan], 2, data, None, False, np.bool_)
            keys = dense_binary(np.add, keys, dense_binary(np.multiply, Kn.iota(int(data.shape[0])),
                                                           Kn.cast(is_nan, np.int64)))
    elif dt.kind in "iub":
        keys, _ = Kn.ew_map(_BINARY[np.bitwise_xor], 0, Kn.cast(data, np.int64), _MIN64, 0, np.int64)
    else:
        raise TypeError(f"sparse_b200: sorting {dt} values is outside the CUDA dtype matrix")
    if descending:
        keys, _ = Kn.ew_map(_UNARY[np.invert], 2, keys, None, 0, np.int64)
    return keys


def _unique(x):
    """(sorted unique stored values, their counts) as host arrays."""
    data = x._data_dev()
    n = int(data.shape[0])
    if n == 0:
        return np.empty(0, dtype=x.dtype), np.empty(0, dtype=np.intp)
    ks, perm = Kn.sort_keys(_order_keys(data), 64)
    ds = Kn.gather(data, perm)
    heads = Kn.flag_heads(ks)
    pos, total = Kn.scan_flags(heads)
    values = D.download(Kn.compact(ds, heads, pos, total))
    starts = D.download(Kn.compact(Kn.iota(n), heads, pos, total))
    return values, np.diff(np.r_[starts, n]).astype(np.intp)


def unique_counts(x, /):
    """Unique elements and their counts, fill value included (_coo/common.py:1189-1236); host arrays like upstream."""
    x = _validate(x).flatten()
    values, counts = _unique(x)
    fill_count = x.size - x.nnz
    if fill_count > 0:
        if np.isnan(x.fill_value):  # every NaN is its own value (Array API)
            values = np.concatenate([values, np.full(fill_count, x.fill_value)])
            counts = np.concatenate([counts, np.ones(fill_count, dtype=counts.dtype)])
        else:
            at = int(np.searchsorted(values, x.fill_value))
            values = np.insert(values, at, x.fill_value)
            counts = np.insert(counts, at, fill_count)
    return UniqueCountsResult(values, counts)


def unique_values(x, /):
    """Sorted unique elements, fill value included (_coo/common.py:1239-1277)."""
    x = _validate(x).flatten()
    values, _ = _unique(x)
    fill_count = x.size - x.nnz
    if fill_count > 0:
        if np.isnan(x.fill_value):
            values = np.concatenate([values, np.full(fill_count, x.fill_value)])
        else:
            values = np.insert(values, int(np.searchsorted(values, x.fill_value)), x.fill_value)
    return values


def sort(x, /, *, axis=-1, descending=False, stable=False):
    """Sorted copy along `axis` (_coo/common.py:1280-1346).  The fill values of a group form one block placed where the
    fill value belongs in the order; `stable` changes nothing for values."""
    from ._elemwise import _BINARY, dense_binary
    from ._gcxs import GCXS

    was_gcxs = isinstance(x, GCXS)
    original = x
    x = _validate(x)
    if x.ndim == 0:
        return original
    if not isinstance(axis, (int, np.integer)) or not -x.ndim <= axis < x.ndim:
        raise IndexError(f"{axis} is out of bounds for array of dimension {x.ndim}")
    axis = int(axis) % x.ndim
    order = [d for d in range(x.ndim) if d != axis] + [axis]
    moved = x.transpose(order) if order != list(range(x.ndim)) else x
    kept = moved.shape[:-1]
    n = moved.shape[-1]
    G = int(np.prod(kept, dtype=np.int64))
    flat = moved.reshape((G, n))
    if flat.nnz == 0:
        return original
    keys = flat.sorted_keys()
    rows, _, indptr = Kn.csr_from_keys(keys, G, n, np.int64, want_rows=True)
    data = flat._data_dev()
    # order by (group, value): sort by value, then stably by group
    _, p1 = Kn.sort_keys(_order_keys(data, descending), 64)
    _, p2 = Kn.sort_keys(Kn.gather(rows, p1), key_bits(max(G, 1)))
    perm = Kn.gather(p1, p2)
    ds = Kn.gather(data, perm)
    # group ids after the sort are the original `rows` (same multiset, sorted); rank = position - group start
    nnz = int(keys.shape[0])
    start = Kn.gather(indptr, rows)
    rank = dense_binary(np.subtract, Kn.iota(nnz), start)
    stop = Kn.gather(indptr, dense_binary(np.add, rows, Kn.full(nnz, 1, np.int64)))
    free = dense_binary(np.subtract, Kn.full(nnz, n, np.int64), dense_binary(np.subtract, stop, start))  # fills per group
    # entries that come after the block of fill values: value > fill ascending, value < fill descending (NaN last)
    fill = flat.fill_value
    T = D.np_dtype(ds)
    if T == np.bool_:
        ds_cmp, fill_c = Kn.cast(ds, np.int32), np.int32(bool(fill))
    else:
        ds_cmp, fill_c = ds, T.type(fill)
    with np.errstate(all="ignore"):
        after, _ = Kn.ew_map(_BINARY[np.less if descending else np.greater], 0, ds_cmp, fill_c, False, np.bool_)
        if T.kind == "f" and not descending:  # NaNs sort last ascending (after the fills), FIRST descending
            from ._elemwise import _UNARY

            nan_e, _ = Kn.ew_map(_UNARY[np.isnan], 2, ds, None, False, np.bool_)
            after = dense_binary(np.logical_or, Kn.cast(after, np.int32), Kn.cast(nan_e, np.int32))  # bool
    shift = dense_binary(np.multiply, Kn.cast(after, np.int64), free)
    position = dense_binary(np.add, rank, shift)
    t = D.torch()
    new_keys = Kn.linearize(t.stack([rows, position]), [n, 1])
    out = COO._from_device(None, ds, (G, n), fill, keys=new_keys).reshape(kept + (n,))
    if order != list(range(x.ndim)):
        inverse = [order.index(d) for d in range(x.ndim)]
        out = out.transpose(inverse)
    return GCXS.from_coo(out) if was_gcxs else out
