"""Basic indexing (integers, slices with any step, None, Ellipsis) of COO and GCXS arrays on the device.

Host-side mirror of sparse/numba_backend/_coo/indexing.py:12-133 (`getitem`), _slicing.py (`normalize_index`) and
_compressed/indexing.py:14-174.  The reference narrows the sorted coordinate rows axis by axis with binary searches
(`_compute_mask`, numba) and then transforms the surviving coordinates; here one streaming kernel
(`b2s_coo_slice_keys`, csrc/prims.cu) unravels each entry's linear key, tests every axis against its
(start, step, count) range and emits the key over the RESULT shape, followed by the usual flag-scan-compact (and a
sort only when a negative step reverses an axis).  ONE one-dimensional advanced index (integer list / array or
boolean mask, _coo/indexing.py:136-172) is served by `_manip.take_axis` after the basic part of the index: a lookup
table over the axis extent is scattered on the device and every surviving entry gathers its slot from it.  Several
1-D advanced indices of one length are taken together: their axes are made adjacent, merged into one virtual axis by a
reshape, and the raveled index list is taken along it.
"""
from __future__ import annotations

import operator

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._utils import c_strides, key_bits, prod


def _normalize(index, shape):
    """Canonical per-entry items: ("int", i) | ("slice", start, step, count) | ("none",), Ellipsis expanded and
    missing trailing axes filled with full slices (NumPy's rules and error types)."""
    ells = [k for k, i in enumerate(index) if i is Ellipsis]
    if len(ells) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
    ndim = len(shape)
    if n_real > ndim:
        raise IndexError(f"too many indices for array: array is {ndim}-dimensional, but {n_real} were indexed")
    fill = (slice(None),) * (ndim - n_real)
    if ells:
        index = index[:ells[0]] + fill + index[ells[0] + 1:]
    else:
        index = index + fill
    items, d = [], 0
    for ind in index:
        if ind is None:
            items.append(("none",))
            continue
        extent = shape[d]
        if isinstance(ind, slice):
            start, stop, step = ind.indices(extent)
            items.append(("slice", start, step, len(range(start, stop, step))))
        elif isinstance(ind, (bool, np.bool_)) or isinstance(ind, (list, np.ndarray)) or D.is_device_tensor(ind):
            raise NotImplementedError("sparse_b200: scalar booleans and nested advanced indices are not on the CUDA "
                                      "path; integers, slices, None, Ellipsis and one 1-D integer / boolean array are")
        else:
            try:
                i = operator.index(ind)
            except TypeError:
                raise IndexError("only integers, slices (`:`), ellipsis (`...`) and None are valid indices") from None
            if i >= extent:  # upstream's wording (_slicing.py:127-132)
                raise IndexError(f"Index is not smaller than dimension {i:d} >= {extent:d}")
            if i < -extent:
                raise IndexError(f"Negative index is not greater than negative dimension {i:d} <= -{extent:d}")
            items.append(("int", i + extent if i < 0 else i))
        d += 1
    return items


def _split_advanced(index, shape):
    """(index with every advanced entry replaced by a full slice, their positions in the RESULT of that basic index,
    their integer lists) or None.  Several 1-D advanced indices must have the same length; they are taken together
    (`x[[0, 1], [0, 2]]` picks (0, 0) and (1, 2)), like `_compute_multi_axis_multi_mask` (_coo/indexing.py:292-349)."""
    adv = [k for k, i in enumerate(index)
           if isinstance(i, (list, np.ndarray)) or (D.is_device_tensor(i) and not isinstance(i, (bool, np.bool_)))]
    if not adv:
        return None
    n_real = sum(1 for i in index if i is not None and i is not Ellipsis)
    basic, positions, arrays = list(index), [], []
    for k in adv:
        arr = D.download(index[k]) if D.is_device_tensor(index[k]) else np.asarray(index[k])
        if arr.ndim != 1:
            raise IndexError("Only one-dimensional iterable indices supported.")
        # the axis of x the advanced index addresses: entries before it that consume an axis (Ellipsis expands)
        axis = 0
        for i in index[:k]:
            if i is Ellipsis:
                axis += len(shape) - n_real
            elif i is not None:
                axis += 1
        if axis >= len(shape):
            raise IndexError(f"too many indices for array: array is {len(shape)}-dimensional")
        extent = shape[axis]
        if arr.size == 0:
            arr = arr.astype(np.int64)
        if arr.dtype == np.bool_:
            if len(arr) != extent:
                raise IndexError(f"boolean index did not match indexed array; dimension is {extent:d} "
                                 f"but corresponding boolean dimension is {len(arr):d}")
            arr = np.flatnonzero(arr)
        elif arr.dtype.kind not in "iu":
            raise IndexError("only integers, slices (`:`), ellipsis (`...`), numpy.newaxis (`None`) and integer or "
                             "boolean arrays are valid indices")
        arr = arr.astype(np.int64)
        if arr.size and (arr.min() < -extent or arr.max() >= extent):
            bad = arr[(arr < -extent) | (arr >= extent)][0]
            raise IndexError(f"index {bad} is out of bounds for axis {axis} with size {extent}")
        arrays.append(np.where(arr < 0, arr + extent, arr))
        basic[k] = slice(None)
        # position of that axis in the result of the basic part: integers before it drop out, None adds one
        pos = 0
        for i in index[:k]:
            if i is Ellipsis:
                pos += len(shape) - n_real
            elif i is None or isinstance(i, (slice, list, np.ndarray)) or D.is_device_tensor(i):
                pos += 1
        positions.append(pos)
    if len({len(a) for a in arrays}) != 1:
        raise IndexError("shape mismatch: indexing arrays could not be broadcast together. Ensure all indexing arrays "
                         "are of the same length.")
    return tuple(basic), positions, arrays


def _take_advanced(y, positions, arrays):
    """Apply the advanced indices to `y` (the result of the basic part, advanced axes still in place)."""
    from ._manip import take_axis

    if len(positions) == 1:
        return take_axis(y, arrays[0], positions[0])
    # several advanced axes: make them adjacent at the first one's position, merge them into ONE virtual axis
    # (C order), and take the raveled index list along it
    first = positions[0]
    rest = [d for d in range(y.ndim) if d not in positions]
    order = [d for d in rest if d < first] + list(positions) + [d for d in rest if d > first]
    extents = tuple(y.shape[d] for d in positions)
    merged = int(np.prod(extents, dtype=np.int64)) if extents else 1
    if merged > (1 << 28):
        raise NotImplementedError("sparse_b200: the advanced axes span more than 2^28 positions; index them one at a "
                                  "time")
    z = y.transpose(order) if order != list(range(y.ndim)) else y
    lead = tuple(y.shape[d] for d in rest if d < first)
    tail = tuple(y.shape[d] for d in rest if d > first)
    z = z.reshape(lead + (merged,) + tail)
    combined = np.ravel_multi_index(tuple(arrays), extents) if len(arrays[0]) else np.empty(0, dtype=np.int64)
    return take_axis(z, combined, len(lead))


def coo_getitem(x, index):
    """COO.__getitem__ (_coo/indexing.py:12-133)."""
    from ._coo import COO

    if isinstance(index, str):
        if x.dtype.names is None:
            raise IndexError("only integers, slices (`:`), ellipsis (`...`), numpy.newaxis (`None`) and integer or "
                             "boolean arrays are valid indices")
        raise NotImplementedError("sparse_b200: structured dtypes are outside the CUDA dtype matrix")
    if not isinstance(index, tuple):
        index = (index,)
    split = _split_advanced(index, x.shape)
    if split is not None:
        basic, positions, arrays = split
        return _take_advanced(coo_getitem(x, basic), positions, arrays)
    last_ellipsis = len(index) > 0 and index[-1] is Ellipsis
    items = _normalize(index, x.shape)
    if len(index) != 0 and all(it[0] == "slice" and it[1:] == (0, 1, ext) for it, ext in zip(items, x.shape)) \
            and len(items) == x.ndim:
        return x
    new_shape = tuple(1 if it[0] == "none" else it[3] for it in items if it[0] != "int")
    real = [it for it in items if it[0] != "none"]

    # leading integer on a canonical array: contiguous run of the sorted entries (the batched-matmul slice a[i])
    if (x.ndim > 1 and real[0][0] == "int" and x._coords is not None
            and all(it[0] == "slice" and it[1:] == (0, 1, ext) for it, ext in zip(items[1:], x.shape[1:]))
            and len(items) == x.ndim):
        return x._take_leading(real[0][1])

    if x.nnz == 0 or x.ndim == 0:
        data = x._data_dev() if x.nnz else None
        total = x.nnz
        keys = None
    else:
        st_new = c_strides(new_shape)
        start, step, count, ostride = [], [], [], []
        p = 0
        for it in items:
            if it[0] == "none":
                p += 1
            elif it[0] == "int":
                start.append(it[1]); step.append(1); count.append(1); ostride.append(0)
            else:
                start.append(it[1]); step.append(it[2]); count.append(it[3]); ostride.append(st_new[p])
                p += 1
        flags, keys = Kn.slice_keys(x.sorted_keys(), x.shape, start, step, count, ostride)
        pos, total = Kn.scan_flags(flags)
        data = x._data_dev()
        if total != x.nnz:
            keys = Kn.compact(keys, flags, pos, total)
            data = Kn.compact(data, flags, pos, total)
        if total > 1 and any(s < 0 for s in step):
            unsorted, _ = Kn.keys_flags(keys)
            if unsorted:
                keys, perm = Kn.sort_keys(keys, key_bits(prod(new_shape)))
                data = Kn.gather(data, perm)

    if not new_shape:
        if not last_ellipsis:  # a single element: the stored value, else the fill value
            return D.download(data[:1])[0] if total else x.fill_value
        host = D.download(data[:total]) if total else np.empty(0, dtype=x.dtype)
        return COO(np.empty((0, total), dtype=np.intp), host, shape=(), has_duplicates=False, sorted=True,
                   fill_value=x.fill_value)
    if total == 0:
        return COO(np.zeros((len(new_shape), 0), dtype=np.intp), np.empty(0, dtype=x.dtype), shape=new_shape,
                   has_duplicates=False, sorted=True, fill_value=x.fill_value)
    if keys is None:  # 0-D input with None axes
        keys = Kn.full(total, 0, np.int64)
    return COO._from_device(None, data, new_shape, x.fill_value, keys=keys)


def gcxs_getitem(x, index):
    """GCXS.__getitem__ (_compressed/indexing.py:14-174): same selection, result compressed along the surviving
    compressed axes of the operand."""
    from ._gcxs import GCXS

    if not isinstance(index, tuple):
        index = (index,)
    if x.ndim == 1:
        r = coo_getitem(x.tocoo(), index)
        return GCXS.from_coo(r) if hasattr(r, "nnz") else r
    split = _split_advanced(index, x.shape)
    if split is not None and len(split[1]) > 1:  # several advanced axes merge into one: default compression
        r = coo_getitem(x.tocoo(), index)
        return GCXS.from_coo(r) if hasattr(r, "nnz") else r
    layout_index = split[0] if split is not None else index  # an advanced index keeps its axis, like a slice
    items = _normalize(layout_index, x.shape)
    if split is None and len(index) != 0 and len(items) == x.ndim \
            and all(it[0] == "slice" and it[1:] == (0, 1, ext) for it, ext in zip(items, x.shape)):
        return x
    real = [it for it in items if it[0] != "none"]
    if all(it[0] == "int" for it in real) and len(real) == len(items):
        return coo_getitem(x.tocoo(), tuple(it[1] for it in real))  # get_single_element: always a scalar
    comp, n_comp_kept, n_uncomp_kept, pos = [], 0, 0, 0
    for axis, it in enumerate(real):
        if it[0] == "int":
            continue
        if axis in x.compressed_axes:
            comp.append(pos)
            n_comp_kept += 1
        else:
            n_uncomp_kept += 1
        pos += 1
    if n_comp_kept == 0 or n_uncomp_kept == 0:
        comp = [0]
    for k, it in enumerate(items):
        if it[0] == "none":
            comp = [c + 1 if c >= k else c for c in comp]
    r = coo_getitem(x.tocoo(), index)
    return GCXS.from_coo(r, None if r.ndim == 1 else tuple(comp))
