"""Structure-changing operations next to the hot path: joining, axis juggling, diagonals, padding, tiling, `take`.

Host-side mirror of the array-manipulation half of the reference's namespace -- `concatenate` / `stack`
(sparse/numba_backend/_coo/common.py:132-250, _compressed/common.py:6-96), `moveaxis` / `pad` / `outer` / `repeat` /
`tile` / `unstack` / `squeeze` (_common.py:1895-2062, 2788-2805, 3121-3232), `kron`, `triu` / `tril`, `roll`,
`diagonal` / `diagonalize`, `expand_dims`, `flip`, `take`, `matrix_transpose` (_coo/common.py:67-130, 252-332, 735-935,
1074-1187, 1349-1384, 1571-1597).

Every function is a re-keying of stored entries: the reference edits coordinate rows with NumPy and lets
`COO.__init__` re-sort; here the entries' linear keys over the RESULT shape are produced directly on the device by
`b2s_coo_linearize` (key = sum_d coord_d * stride_d, where a constant offset is one extra coordinate row of ones),
filtered with the flag / scan / compact primitives and re-sorted (radix sort on the significant key bits) only when the
new key order differs from the stored one.  Nothing here touches values on the host and there is no CPU fallback.
"""
from __future__ import annotations

import builtins
from collections.abc import Iterable

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO, as_coo
from ._sparse_array import SparseArray
from ._utils import c_strides, check_zero_fill_value, equivalent, key_bits, normalize_axis, prod


# ---------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------
def check_consistent_fill_value(arrays):
    """_utils.py:599-641."""
    arrays = list(arrays)
    if not builtins.all(isinstance(s, SparseArray) for s in arrays):
        raise ValueError("All arrays must be instances of SparseArray.")
    if len(arrays) == 0:
        raise ValueError("At least one array required.")
    fv = arrays[0].fill_value
    for i, arg in enumerate(arrays):
        if not equivalent(fv, arg.fill_value):
            raise ValueError("This operation requires consistent fill-values, "
                             f"but argument {i:d} had a fill value of {arg.fill_value!s}, which "
                             f"is different from a fill_value of {fv!s} in the first argument.")


def _as_coo(x):
    if isinstance(x, COO):
        return x
    if isinstance(x, SparseArray):
        return x.asformat("coo")
    return as_coo(x)


def _rekey(x, strides, offset=0):
    """Linear keys of x's entries under `strides` (one per axis of x) plus a constant `offset`."""
    coords = x._dev()[0]
    if offset:
        t = D.torch()
        ones = Kn.full(x.nnz, 1, D.np_dtype(coords))
        return Kn.linearize(t.cat([ones[None, :], coords], dim=0), [int(offset), *[int(s) for s in strides]])
    return Kn.linearize(coords, [int(s) for s in strides])


def _sorted_result(keys, data, shape, fill_value):
    """Canonical COO from (keys over `shape`, data): sort only when the new key order is not the stored one."""
    if int(keys.shape[0]) > 1:
        unsorted, _ = Kn.keys_flags(keys)
        if unsorted:
            keys, perm = Kn.sort_keys(keys, key_bits(prod(shape)))
            data = Kn.gather(data, perm)
    return COO._from_device(None, data, tuple(shape), fill_value, keys=keys)


def _empty(shape, dtype, fill_value):
    return COO(np.zeros((len(shape), 0), dtype=np.intp), np.empty(0, dtype=dtype), shape=tuple(shape),
               has_duplicates=False, sorted=True, fill_value=fill_value)


def _join(parts_keys, parts_data, shape, dtype, fill_value):
    t = D.torch()
    parts = [(k, d) for k, d in zip(parts_keys, parts_data) if int(k.shape[0])]
    if not parts:
        return _empty(shape, dtype, fill_value)
    keys = t.cat([k for k, _ in parts]) if len(parts) > 1 else parts[0][0]
    data = t.cat([Kn.cast(d, dtype) for _, d in parts]) if len(parts) > 1 else Kn.cast(parts[0][1], dtype)
    return _sorted_result(keys, data, shape, fill_value)


def _wrap_like(arrays, out, axis, compressed_axes=None):
    """All-GCXS input gives GCXS output compressed along `axis` unless told otherwise (_compressed/common.py).  A
    narrow index dtype shared by the inputs is kept (and widened by the usual rule when the result outgrows it)."""
    from ._gcxs import GCXS

    vis = {getattr(a, "_idx_vis", None) for a in arrays}
    if len(vis) == 1 and None not in vis and out.ndim:
        (v,) = vis
        out._idx_vis = v if np.can_cast(np.min_scalar_type(max(out.shape)), v) else np.dtype(
            np.min_scalar_type(max(out.shape)))
    if builtins.all(isinstance(a, GCXS) for a in arrays):
        if arrays[0].ndim == 1:
            return out  # 1-D GCXS inputs are joined as COO upstream (_compressed/common.py:18-22)
        if out.ndim < 2:
            return GCXS.from_coo(out)
        return GCXS.from_coo(out, tuple(compressed_axes) if compressed_axes is not None else (axis,))
    return out


# ---------------------------------------------------------------------------------------------------------------
# joining
# ---------------------------------------------------------------------------------------------------------------
def concatenate(arrays, axis=0, compressed_axes=None):
    """numpy.concatenate for sparse arrays (_common.py:1518-1558)."""
    arrays = list(arrays)
    check_consistent_fill_value(arrays)  # also: every operand must be a SparseArray, at least one of them
    orig = arrays
    if axis is None:
        axis = 0
        arrays = [a.reshape(-1) for a in arrays]
    coos = [_as_coo(a) for a in arrays]
    axis = normalize_axis(axis, coos[0].ndim)
    first = coos[0]
    for x in coos:
        if x.ndim != first.ndim or builtins.any(x.shape[ax] != first.shape[ax] for ax in range(first.ndim) if ax != axis):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    shape = list(first.shape)
    shape[axis] = builtins.sum(x.shape[axis] for x in coos)
    dtype = np.result_type(*[x.dtype for x in coos])
    st = c_strides(shape)
    keys, data, off = [], [], 0
    for x in coos:
        if x.nnz:
            keys.append(_rekey(x, st, off * st[axis]))
            data.append(x._data_dev())
        off += x.shape[axis]
    out = _join(keys, data, shape, dtype, dtype.type(first.fill_value))
    return _wrap_like(orig, out, axis, compressed_axes)


def stack(arrays, axis=0, compressed_axes=None):
    """numpy.stack for sparse arrays (_common.py:1479-1515)."""
    arrays = list(arrays)
    check_consistent_fill_value(arrays)
    coos = [_as_coo(a) for a in arrays]
    first = coos[0]
    if len({x.shape for x in coos}) != 1:
        raise ValueError("all input arrays must have the same shape")
    axis = normalize_axis(axis, first.ndim + 1)
    shape = list(first.shape)
    shape.insert(axis, len(coos))
    dtype = np.result_type(*[x.dtype for x in coos])
    st = c_strides(shape)
    st_in = [s for d, s in enumerate(st) if d != axis]
    keys, data = [], []
    for i, x in enumerate(coos):
        if x.nnz:
            if x.ndim == 0:
                keys.append(Kn.full(1, i * st[axis], np.int64))
            else:
                keys.append(_rekey(x, st_in, i * st[axis]))
            data.append(x._data_dev())
    out = _join(keys, data, shape, dtype, dtype.type(first.fill_value))
    return _wrap_like(arrays, out, axis, compressed_axes)


def unstack(x, axis=0):
    """Tuple of the slices along `axis` (_common.py:3203-3231)."""
    if not isinstance(x, SparseArray):
        raise TypeError("`a` must be a SparseArray.")
    ndim = x.ndim
    if not isinstance(axis, (int, np.integer)) or not (-ndim <= axis < ndim):
        raise ValueError(f"axis must be in range [-{ndim}, {ndim}), got {axis}")
    axis = int(axis) % ndim
    x = x.transpose((axis,) + tuple(i for i in range(ndim) if i != axis))
    return tuple(x[i] for i in range(x.shape[0]))


# ---------------------------------------------------------------------------------------------------------------
# axes
# ---------------------------------------------------------------------------------------------------------------
def moveaxis(a, source, destination):
    """numpy.moveaxis (_common.py:1954-1999)."""
    if not isinstance(source, Iterable):
        source = (source,)
    if not isinstance(destination, Iterable):
        destination = (destination,)
    source = normalize_axis(tuple(source), a.ndim)
    destination = normalize_axis(tuple(destination), a.ndim)
    if len(source) != len(destination):
        raise ValueError("`source` and `destination` arguments must have the same number of elements")
    order = [n for n in range(a.ndim) if n not in source]
    for dest, src in sorted(zip(destination, source, strict=True)):
        order.insert(dest, src)
    return a.transpose(order)


def swapaxes(x, axis1, axis2):
    """numpy.swapaxes (`COO.swapaxes`, _coo/core.py)."""
    if not -x.ndim <= axis1 < x.ndim or not -x.ndim <= axis2 < x.ndim:
        raise ValueError(f"Invalid axis values. Axis values must be in the range [-{x.ndim}, {x.ndim}) but got "
                         f"axis1={axis1}, axis2={axis2}")
    axes = list(range(x.ndim))
    axes[axis1], axes[axis2] = axes[axis2], axes[axis1]
    return x.transpose(axes)


def permute_dims(x, /, axes=None):
    return x.transpose(axes)


def matrix_transpose(x, /):
    """Swap the last two axes (_coo/common.py:1571-1597)."""
    if not hasattr(x, "ndim"):
        x = np.asarray(x)
    if x.ndim < 2:
        raise ValueError("`x.ndim >= 2` must hold.")
    if isinstance(x, np.ndarray):
        return np.swapaxes(x, -1, -2)
    return x.transpose(list(range(x.ndim - 2)) + [x.ndim - 1, x.ndim - 2])


def squeeze(x, /, axis=None):
    """numpy.squeeze (_common.py:2788-2805, `COO.squeeze`)."""
    if axis is None:
        axis = tuple(d for d, s in enumerate(x.shape) if s == 1)
    if not isinstance(axis, Iterable):
        axis = (axis,)
    for d in axis:
        if not isinstance(d, (int, np.integer)):
            raise ValueError(f"Invalid axis parameter: `{d}`.")
    axis = tuple(int(d) + x.ndim if d < 0 else int(d) for d in axis)
    for d in axis:
        if not 0 <= d < x.ndim:
            raise ValueError(f"Invalid axis index {d} for ndim={x.ndim}")
        if x.shape[d] != 1:
            raise ValueError(f"Specified axis `{d}` has a size greater than one: {x.shape[d]}")
    return x.reshape(tuple(s for d, s in enumerate(x.shape) if d not in axis))


def expand_dims(x, /, axis=0):
    """numpy.expand_dims (_coo/common.py:1074-1133)."""
    if not isinstance(axis, (int, np.integer, tuple, list)):
        raise IndexError(f"Invalid axis position: type={type(axis)}, axis={axis}")
    axes = (axis,) if isinstance(axis, (int, np.integer)) else tuple(axis)
    out_ndim = x.ndim + len(axes)
    norm = []
    for a in axes:
        if not -out_ndim <= a < out_ndim:
            raise IndexError(f"Invalid axis position: axis={a} for an array of {out_ndim} dimensions")
        norm.append(a % out_ndim)
    if len(set(norm)) != len(norm):
        raise ValueError("repeated axis")
    it = iter(x.shape)
    return x.reshape(tuple(1 if d in norm else next(it) for d in range(out_ndim)))


def flip(x, /, axis=None):
    """numpy.flip (_coo/common.py:1136-1186): a slice with step -1 on every flipped axis."""
    if axis is None:
        axis = tuple(range(x.ndim))
    if not isinstance(axis, Iterable):
        axis = (axis,)
    axis = normalize_axis(tuple(axis), x.ndim)
    return x[tuple(slice(None, None, -1) if d in axis else slice(None) for d in range(x.ndim))]


def roll(a, shift, axis=None):
    """numpy.roll (_coo/common.py:735-812): per axis, the two halves of the array swap places."""
    a = a if isinstance(a, SparseArray) else as_coo(a)
    if axis is None:
        return roll(a.reshape((-1,)), shift, 0).reshape(a.shape)
    if np.ndim(shift) > 1 or np.ndim(axis) > 1:
        raise ValueError("'shift' and 'axis' must be integers or 1D sequences.")
    if isinstance(shift, Iterable) and not isinstance(axis, Iterable):
        raise ValueError("If 'shift' is a 1D sequence, 'axis' must have equal length.")
    if not isinstance(shift, Iterable):
        shift = (shift,)
    if not isinstance(axis, Iterable):
        axis = (axis,)
    shift, axis = tuple(int(v) for v in shift), tuple(axis)
    if len(shift) == 1 and len(axis) > 1:
        shift = shift * len(axis)
    if len(shift) != len(axis):
        raise ValueError("If 'shift' is a 1D sequence, 'axis' must have equal length.")
    axis = normalize_axis(axis, a.ndim)
    # upstream adds the shift to the coordinate rows in their own dtype (_coo/common.py:788-806): a shift that a narrow
    # or unsigned index dtype cannot hold is refused here as well
    idt = np.dtype(_as_coo(a)._idx_dtype()) if isinstance(a, SparseArray) else np.dtype(np.intp)
    from ._utils import can_store

    if not can_store(idt, max(max(a.shape), max(abs(v) for v in shift))) or (idt.kind == "u" and min(shift) < 0):
        raise ValueError(f"cannot roll with coords.dtype {idt} and shift {shift}. Try casting coords to a larger "
                         "signed dtype.")
    out = a
    for sh, ax in zip(shift, axis):
        n = out.shape[ax]
        k = int(sh) % n if n else 0
        if k == 0:
            continue
        full = (slice(None),) * ax
        out = concatenate([out[full + (slice(n - k, None),)], out[full + (slice(None, n - k),)]], axis=ax)
    return out


# ---------------------------------------------------------------------------------------------------------------
# selections on coordinates
# ---------------------------------------------------------------------------------------------------------------
def _select(x, keep_flags):
    pos, total = Kn.scan_flags(keep_flags)
    if total == x.nnz:
        return x
    if total == 0:
        return _empty(x.shape, x.dtype, x.fill_value)
    return COO._from_device(None, Kn.compact(x._data_dev(), keep_flags, pos, total), x.shape, x.fill_value,
                            keys=Kn.compact(x.sorted_keys(), keep_flags, pos, total))


def _band_flags(x, k, upper):
    """uint8 flags of the entries with (col - row >= k) [upper] or (col - row <= k) [lower] on the last two axes."""
    from ._elemwise import _BINARY

    strides = [0] * x.ndim
    strides[-2], strides[-1] = -1, 1
    diff = _rekey(x, strides)  # col - row, int64
    op = _BINARY[np.greater_equal] if upper else _BINARY[np.less_equal]
    _, flags = Kn.ew_map(op, 0, diff, np.int64(k), False, np.bool_)
    return flags


def triu(x, k=0):
    """Upper triangle of the last two axes (_coo/common.py:252-290)."""
    check_zero_fill_value(x)
    if not x.ndim >= 2:
        raise NotImplementedError("sparse.triu is not implemented for scalars or 1-D arrays.")
    c = _as_coo(x)
    if c.nnz == 0:
        return c
    return _select(c, _band_flags(c, int(k), True))


def tril(x, k=0):
    """Lower triangle of the last two axes (_coo/common.py:293-331)."""
    check_zero_fill_value(x)
    if not x.ndim >= 2:
        raise NotImplementedError("sparse.tril is not implemented for scalars or 1-D arrays.")
    c = _as_coo(x)
    if c.nnz == 0:
        return c
    return _select(c, _band_flags(c, int(k), False))


def diagonal(a, offset=0, axis1=0, axis2=1):
    """numpy.diagonal (_coo/common.py:815-878): entries with coord[axis2] - coord[axis1] == offset; the diagonal
    becomes the LAST axis of the result."""
    from ._elemwise import _BINARY

    a = _as_coo(a)
    if a.ndim < 2:
        raise ValueError("diagonal requires at least two dimensions.")
    axis1, axis2 = normalize_axis((axis1, axis2), a.ndim)
    if axis1 == axis2:
        raise ValueError("axis1 and axis2 cannot be the same.")
    offset = int(offset)
    n1, n2 = a.shape[axis1], a.shape[axis2]
    dlen = max(min(n1, n2 - offset) if offset >= 0 else min(n1 + offset, n2), 0)
    rest = [d for d in range(a.ndim) if d not in (axis1, axis2)]
    shape = tuple(a.shape[d] for d in rest) + (dlen,)
    if a.nnz == 0 or dlen == 0:
        return _empty(shape, a.dtype, a.fill_value)
    strides = [0] * a.ndim
    strides[axis1], strides[axis2] = -1, 1
    _, flags = Kn.ew_map(_BINARY[np.equal], 0, _rekey(a, strides), np.int64(offset), False, np.bool_)
    pos, total = Kn.scan_flags(flags)
    if total == 0:
        return _empty(shape, a.dtype, a.fill_value)
    st = c_strides(shape)
    strides = [0] * a.ndim
    for p, d in enumerate(rest):
        strides[d] = st[p]
    strides[axis1 if offset >= 0 else axis2] = st[-1]  # position along the diagonal = the smaller coordinate
    keys = _rekey(a, strides)
    data = a._data_dev()
    if total != a.nnz:
        keys, data = Kn.compact(keys, flags, pos, total), Kn.compact(data, flags, pos, total)
    return _sorted_result(keys, data, shape, a.fill_value)


def diagonalize(a, axis=0):
    """Inverse of `diagonal` (_coo/common.py:881-934): `axis` is duplicated as a new last axis."""
    a = _as_coo(a)
    axis = normalize_axis(axis, a.ndim)
    shape = tuple(a.shape) + (a.shape[axis],)
    if a.nnz == 0:
        return _empty(shape, a.dtype, a.fill_value)
    st = c_strides(shape)
    strides = list(st[:-1])
    strides[axis] += st[-1]
    return _sorted_result(_rekey(a, strides), a._data_dev(), shape, a.fill_value)


def pad(array, pad_width, mode="constant", **kwargs):
    """numpy.pad, constant mode with the array's own fill value (_common.py:2002-2061): a pure shift of the keys."""
    from ._gcxs import GCXS

    if not isinstance(array, SparseArray):
        raise NotImplementedError("Input array is not compatible.")
    if mode.lower() != "constant":
        raise NotImplementedError(f"Mode '{mode}' is not yet supported.")
    if not equivalent(kwargs.pop("constant_values", _zero_like(array)), array.fill_value, loose=True):
        raise ValueError("constant_values can only be equal to fill value.")
    if kwargs:
        raise NotImplementedError("Additional Unsupported arguments present.")
    pw = np.broadcast_to(np.asarray(pad_width, dtype=np.int64), (array.ndim, 2))
    if (pw < 0).any():
        raise ValueError("index can't contain negative values")
    c = _as_coo(array)
    shape = tuple(int(s + b + e) for s, (b, e) in zip(c.shape, pw))
    if c.nnz == 0:
        out = _empty(shape, c.dtype, c.fill_value)
    else:
        st = c_strides(shape)
        off = builtins.sum(int(b) * s for (b, _), s in zip(pw, st))
        out = COO._from_device(None, c._data_dev(), shape, c.fill_value, keys=_rekey(c, st, off))
    if isinstance(array, GCXS):
        return GCXS.from_coo(out, array.compressed_axes if out.ndim > 1 else None)
    return out


def _zero_like(x):
    return x.dtype.type(0)


# ---------------------------------------------------------------------------------------------------------------
# products of shapes
# ---------------------------------------------------------------------------------------------------------------
def repeat(a, repeats, axis=None):
    """numpy.repeat with an integer count (_common.py:3121-3161): new unit axis, broadcast, merge."""
    from ._elemwise import broadcast_to

    if not isinstance(a, SparseArray):
        raise TypeError("`a` must be a SparseArray.")
    if not isinstance(repeats, int):
        raise ValueError("`repeats` must be an integer, uneven repeats are not yet Implemented.")
    c = _as_coo(a)
    if axis is None:
        c = c.reshape((-1,))
        axis = 0
    axis = normalize_axis(axis, c.ndim)
    new_shape = list(c.shape)
    new_shape[axis] *= repeats
    wide = c.reshape(c.shape[:axis + 1] + (1,) + c.shape[axis + 1:])
    wide = broadcast_to(wide, c.shape[:axis + 1] + (repeats,) + c.shape[axis + 1:])
    return wide.reshape(tuple(new_shape))


def tile(a, reps):
    """numpy.tile (_common.py:3164-3200): interleave unit axes, broadcast them to `reps`, merge pairwise."""
    from ._elemwise import broadcast_to

    a = _as_coo(a) if isinstance(a, SparseArray) else as_coo(a)
    reps = (reps,) if isinstance(reps, (int, np.integer)) else tuple(reps)
    if a.ndim == 0:
        a = a.reshape((1,))
    if len(reps) < a.ndim:
        reps = (1,) * (a.ndim - len(reps)) + reps
    elif len(reps) > a.ndim:
        a = a.reshape((1,) * (len(reps) - a.ndim) + a.shape)
    shape = a.shape
    inter = tuple(v for s in shape for v in (1, s))
    wide = tuple(v for r, s in zip(reps, shape) for v in (int(r), s))
    return broadcast_to(a.reshape(inter), wide).reshape(tuple(int(r) * s for r, s in zip(reps, shape)))


def outer(a, b, out=None):
    """numpy.outer of the flattened operands (_common.py:1895-1925)."""
    from ._gcxs import GCXS

    if isinstance(a, SparseArray):
        a = _as_coo(a)
    if isinstance(b, SparseArray):
        b = _as_coo(b)
    gcxs = isinstance(out, GCXS)
    if gcxs:
        out = out.tocoo()
    kw = {"out": out} if out is not None else {}
    res = np.multiply.outer(a.flatten() if hasattr(a, "flatten") else np.ravel(a),
                            b.flatten() if hasattr(b, "flatten") else np.ravel(b), **kw)
    return GCXS.from_coo(res) if gcxs else res


def kron(a, b):
    """Kronecker product (_coo/common.py:67-129): out[i*Bm + k, j*Bn + l] = a[i, j] * b[k, l], i.e. the broadcast
    product of a[i, 1, j, 1] and b[1, k, 1, l] with each axis pair merged."""
    from ._elemwise import _is_scipy_sparse

    def conv(x):
        if _is_scipy_sparse(x):
            return COO.from_scipy_sparse(x)
        return x

    a, b = conv(a), conv(b)
    if not (isinstance(a, SparseArray) or isinstance(b, SparseArray)):
        raise ValueError("Performing this operation would produce a dense result: kron")
    check_zero_fill_value(*(x for x in (a, b) if isinstance(x, SparseArray)))
    if np.isscalar(a) or np.isscalar(b) or getattr(a, "ndim", 1) == 0 or getattr(b, "ndim", 1) == 0:
        return a * b
    a, b = _as_coo(a), _as_coo(b)  # a dense operand contributes its non-zero entries (as_coo prunes)
    nd = max(a.ndim, b.ndim)
    a = a.reshape((1,) * (nd - a.ndim) + tuple(a.shape))
    b = b.reshape((1,) * (nd - b.ndim) + tuple(b.shape))
    out_shape = tuple(sa * sb for sa, sb in zip(a.shape, b.shape))
    from ._elemwise import _BINARY, dense_binary

    dtr = np.result_type(a.dtype, b.dtype)
    na, nb = a.nnz, b.nnz
    if na == 0 or nb == 0:
        return _empty(out_shape, dtr, dtr.type(0))
    # stored x stored only, like upstream (a broadcast product would also store negative * fill = -0.0):
    # pair (p, q) -> coordinate a_coords[:, p] * b.shape + b_coords[:, q], value a.data[p] * b.data[q]
    t = D.torch()
    pair = Kn.iota(na * nb)
    ia, _ = Kn.ew_map(_BINARY[np.floor_divide], 0, pair, np.int64(nb), 0, np.int64)
    ib, _ = Kn.ew_map(_BINARY[np.remainder], 0, pair, np.int64(nb), 0, np.int64)
    ka = Kn.gather(_rekey(a, [sb * st for sb, st in zip(b.shape, c_strides(out_shape))]), ia)
    kb = Kn.gather(_rekey(b, c_strides(out_shape)), ib)
    keys = dense_binary(np.add, ka, kb)
    # narrow / unsigned integers and bool are storage-only on the device: the product runs in the wider signed type and
    # the cast back wraps exactly like NumPy's arithmetic in the narrow type (bool: any non-zero product is True)
    from ._elemwise import _WIDE_FOR

    wide = np.dtype("int32") if dtr == np.bool_ else _WIDE_FOR.get(dtr, dtr)
    data = dense_binary(np.multiply, Kn.cast(Kn.gather(a._data_dev(), ia), wide),
                        Kn.cast(Kn.gather(b._data_dev(), ib), wide))
    if wide != dtr:
        data = Kn.cast(data, dtr)
    del t
    return _sorted_result(keys, data, out_shape, dtr.type(0))


# ---------------------------------------------------------------------------------------------------------------
# take / advanced index along one axis
# ---------------------------------------------------------------------------------------------------------------
def take_axis(x, idx, axis):
    """x[..., idx, ...] along `axis` for a host list of non-negative integers (duplicates allowed).

    A lookup table over the axis extent (position of each requested index, -1 elsewhere) is scattered on the device;
    every stored entry gathers its slot from the table, entries whose slot is -1 are compacted away and the survivors
    are re-keyed with the slot as their new coordinate.  Repeated indices are handled in rounds (the r-th round serves
    the r-th occurrence of every value), so the cost is O(rounds * nnz + extent)."""
    from ._elemwise import _BINARY

    x = _as_coo(x)
    idx = np.asarray(idx, dtype=np.int64).reshape(-1)
    n = x.shape[axis]
    shape = tuple(x.shape[:axis]) + (int(idx.size),) + tuple(x.shape[axis + 1:])
    if idx.size and (idx.min() < 0 or idx.max() >= n):
        raise IndexError(f"index out of bounds for axis {axis} with size {n}")
    if x.nnz == 0 or idx.size == 0:
        return _empty(shape, x.dtype, x.fill_value)
    # occurrence number of every requested index (host work on the INDEX LIST only)
    order = np.argsort(idx, kind="stable")
    sorted_idx = idx[order]
    starts = np.r_[0, np.flatnonzero(np.diff(sorted_idx)) + 1]
    occ = np.empty(idx.size, dtype=np.int64)
    occ[order] = np.arange(idx.size) - np.repeat(starts, np.diff(np.r_[starts, idx.size]))
    coords, data = x._dev()
    t = D.torch()
    axis_row = Kn.cast(coords[axis].contiguous(), np.int64)
    st = c_strides(shape)
    keys_parts, data_parts = [], []
    for r in range(int(occ.max()) + 1):
        sel = np.flatnonzero(occ == r)
        table = Kn.full(n, -1, np.int64)
        Kn.scatter(D.upload(sel.astype(np.int64)), D.upload(idx[sel]), table)
        slot = Kn.gather(table, axis_row)
        _, flags = Kn.ew_map(_BINARY[np.greater_equal], 0, slot, np.int64(0), False, np.bool_)
        pos, total = Kn.scan_flags(flags)
        if total == 0:
            continue
        new_coords = t.cat([coords[:axis], Kn.cast(slot, D.np_dtype(coords))[None, :], coords[axis + 1:]], dim=0)
        keys = Kn.linearize(new_coords, st)
        d = data
        if total != x.nnz:
            keys, d = Kn.compact(keys, flags, pos, total), Kn.compact(data, flags, pos, total)
        keys_parts.append(keys)
        data_parts.append(d)
    return _join(keys_parts, data_parts, shape, x.dtype, x.fill_value)


def take(x, indices, /, axis=None):
    """numpy.take along one axis (_coo/common.py:1349-1383)."""
    from ._gcxs import GCXS

    if not isinstance(x, SparseArray):
        raise ValueError(f"Input must be an instance of SparseArray, but it's {type(x)}.")
    was_gcxs = isinstance(x, GCXS)
    c = _as_coo(x)
    if axis is None:
        c = c.reshape((-1,))
        axis = 0
    axis = normalize_axis(axis, c.ndim)
    idx = np.asarray(indices)
    if idx.dtype == np.bool_:
        idx = np.flatnonzero(idx)
    idx = idx.astype(np.int64)
    idx = np.where(idx < 0, idx + c.shape[axis], idx)
    out = take_axis(c, idx, axis)
    return GCXS.from_coo(out) if was_gcxs else out
