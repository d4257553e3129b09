"""`argmax` / `argmin` along one axis (or over the flattened array) on the device.

Reference: sparse/numba_backend/_coo/common.py:614-671, 1455-1568 (`_arg_minmax_common`, `_compute_minmax_args`: a numba
loop over every output position that masks the whole coordinate array each time).  Semantics = NumPy's: the FIRST index
along the axis that holds the extreme value, fill values included.  With the reduced axis moved last (groups g of
length n), the answer for a group is the smaller of

  P_g  the smallest stored coordinate whose value equals the group's extreme value m_g (m_g from the segmented-scan
       reduction, which already folds the fill value in for incomplete groups), and
  F_g  the first coordinate NOT stored in the group, counted only when the fill value itself is extreme (m_g == fill).

P is one three-operand `where` + one min-reduction; F uses the rank of an entry inside its group (position minus the
group's start from `b2s_csr_from_keys`): stored coordinates are increasing, so coordinate - rank is 0 on the gap-free
prefix and the first entry where it is positive has rank = first missing coordinate (else the group's count).  Every step
is an existing kernel (element-wise maps, gathers, the reduction); nothing runs on the host but the final dense-to-COO
re-fill of the (small) result.
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO
from ._sparse_array import SparseArray

_BIG = np.int64(2**62)


def _arg_minmax(x, axis, keepdims, mode):
    from ._coo import _is_scipy_sparse
    from ._elemwise import dense_binary, where

    if _is_scipy_sparse(x):
        x = COO.from_scipy_sparse(x)
    elif not isinstance(x, SparseArray):
        raise ValueError(f"Input must be an instance of SparseArray, but it's {type(x)}.")
    x = x.asformat("coo")
    if not isinstance(axis, (int, np.integer, type(None))) or isinstance(axis, bool):
        raise ValueError(f"`axis` must be `int` or `None`, but it's: {type(axis)}.")
    if axis is not None and (axis >= x.ndim or axis < -x.ndim):
        raise ValueError(f"`axis={axis}` is out of bounds for array of dimension {x.ndim}.")
    if x.ndim == 0:
        raise ValueError("Input array must be at least 1-D, but it's 0-D.")
    if x.dtype.kind == "c":
        raise TypeError("sparse_b200: argmax / argmin of a complex array is outside the CUDA op set")
    orig_ndim = x.ndim
    if axis is None:
        flat, kept_shape = x.reshape((1, x.size)), ()
        axis_n = None
    else:
        axis_n = int(axis) % x.ndim
        moved = x.transpose([d for d in range(x.ndim) if d != axis_n] + [axis_n])
        kept_shape = moved.shape[:-1]
        flat = moved.reshape((int(np.prod(kept_shape, dtype=np.int64)), moved.shape[-1]))
    G, n = flat.shape
    if n == 0:
        raise ValueError("attempt to get argmax of an empty sequence" if mode == "max"
                         else "attempt to get argmin of an empty sequence")
    fill = flat.fill_value
    m = (flat.max if mode == "max" else flat.min)(axis=1, keepdims=True)  # (G, 1): extreme value incl. fill values
    if flat.nnz == 0:
        res = np.zeros(G, dtype=np.int64)
    else:
        keys = flat.sorted_keys()
        rows, cols, indptr = Kn.csr_from_keys(keys, G, n, np.int64, want_rows=True)
        # P: smallest stored coordinate holding the extreme value
        coord_vals = COO._from_device(None, cols, (G, n), _BIG, keys=keys)        # value = own coordinate, BIG elsewhere
        with np.errstate(all="ignore"):
            hit = flat == m
            if flat.dtype.kind == "f":  # NaN == NaN is False, but a NaN extreme (NumPy: the first NaN wins) must match
                hit = np.logical_or(hit, np.logical_and(np.isnan(flat), np.isnan(m)))
        P = where(hit, coord_vals, _BIG).min(axis=1)
        # F: first coordinate not stored in the group
        iota = Kn.iota(int(keys.shape[0]))
        rank = dense_binary(np.subtract, iota, Kn.gather(indptr, rows))
        gap = dense_binary(np.subtract, cols, rank)                                 # 0 on the gap-free prefix of a group
        from ._elemwise import _BINARY

        is_prefix, _ = Kn.ew_map(_BINARY[np.equal], 0, gap, np.int64(0), False, np.bool_)
        penalty, _ = Kn.ew_map(_BINARY[np.multiply], 0, Kn.cast(is_prefix, np.int64), _BIG, 0, np.int64)
        first_gap = COO._from_device(None, dense_binary(np.add, rank, penalty), (G, n), _BIG, keys=keys).min(axis=1)
        count = COO._from_device(None, Kn.full(int(keys.shape[0]), 1, np.int64), (G, n), np.int64(0),
                                 keys=keys).sum(axis=1)
        F = np.minimum(first_gap, count)                                            # no gap inside: right after the last
        # the fill value competes only where it is the extreme value and the group has a free slot
        m1 = m.reshape((G,))
        with np.errstate(all="ignore"):
            fill_wins = (m1 == fill) if fill == fill else np.isnan(m1)
            fill_wins = np.logical_and(fill_wins, count < n)
        res = where(fill_wins, np.minimum(P, F), P).todense()
    # fill value 0 with the non-zero indices stored, like upstream (also for the 0-D result of axis=None)
    out = COO.from_numpy(np.asarray(res, dtype=np.intp).reshape(-1)).reshape(kept_shape if axis_n is not None else ())
    if keepdims:
        shape = [1] * orig_ndim if axis_n is None else list(x.shape)
        if axis_n is not None:
            shape[axis_n] = 1
        out = out.reshape(tuple(shape))
    return out


def argmax(x, /, *, axis=None, keepdims=False):
    """Index of the first maximum along `axis` (flattened array if None); COO of intp, fill value 0."""
    return _arg_minmax(x, axis, keepdims, "max")


def argmin(x, /, *, axis=None, keepdims=False):
    """Index of the first minimum along `axis` (flattened array if None); COO of intp, fill value 0."""
    return _arg_minmax(x, axis, keepdims, "min")
