"""Host-side helpers mirroring sparse/numba_backend/_utils.py semantics (metadata only, no data path)."""
from __future__ import annotations

import operator
from collections.abc import Iterable
from functools import reduce

import numpy as np


def prod(xs) -> int:
    return int(reduce(operator.mul, (int(x) for x in xs), 1))


def _zero_of_dtype(dtype):
    """_utils.py:_zero_of_dtype."""
    return np.zeros((), dtype=dtype)[()]


def equivalent(x, y, /, loose=False):
    """`equivalent`, _utils.py:406-452 -- for host scalars (fill values) only."""
    x = np.asarray(x)
    y = np.asarray(y)
    dt = np.result_type(x.dtype, y.dtype)
    if not any(np.issubdtype(dt, t) for t in [np.floating, np.complexfloating]):
        return x == y
    if loose:
        return (x == y) | ((x != x) & (y != y))
    if x.size == 0 or y.size == 0:
        return np.empty(np.broadcast_shapes(x.shape, y.shape), dtype=np.bool_)
    x, y = np.broadcast_arrays(x[..., None], y[..., None])
    return (x.astype(dt).view(np.uint8) == y.astype(dt).view(np.uint8)).all(axis=-1)


def check_zero_fill_value(*args, loose=True):
    """_utils.py:562-596: tensordot / matmul / dot require zero fill values."""
    for i, arg in enumerate(args):
        if getattr(arg, "size", 1) == 0:
            continue
        if hasattr(arg, "fill_value") and not equivalent(arg.fill_value, _zero_of_dtype(arg.dtype), loose=loose):
            raise ValueError(
                f"This operation requires zero fill values, but argument {i:d} had a fill value of {arg.fill_value!s}."
            )


def check_fill_value(x, /, *, accept_fv=None):
    """_utils.py:531-560: raise unless the fill value is one of `accept_fv` (default: zero only)."""
    if accept_fv is None:
        accept_fv = [0]
    if not isinstance(accept_fv, Iterable):
        accept_fv = [accept_fv]
    if not any(equivalent(fv, x.fill_value, loose=True) for fv in accept_fv):
        raise ValueError(f"{x.fill_value=} but should be in {accept_fv}.")


def normalize_axis(axis, ndim):
    """_utils.py:normalize_axis."""
    if axis is None:
        return None
    if isinstance(axis, (int, np.integer)):
        axis = int(axis)
        if axis < -ndim or axis >= ndim:
            raise ValueError(f"Invalid axis index {axis} for ndim={ndim}")
        return axis % ndim if ndim else axis
    if isinstance(axis, Iterable):
        return tuple(normalize_axis(a, ndim) for a in axis)
    raise ValueError(f"axis {axis} not understood")


def can_store(dtype, scalar) -> bool:
    """_utils.py:651-658."""
    try:
        with np.errstate(all="raise"):
            return bool(np.array(scalar, dtype=dtype) == np.array(scalar))
    except (ValueError, OverflowError, FloatingPointError):
        return False


def check_compressed_axes(ndim, compressed_axes):
    """_utils.py:check_compressed_axes."""
    if compressed_axes is None:
        return
    if isinstance(ndim, Iterable):
        ndim = len(ndim)
    if not isinstance(compressed_axes, Iterable):
        raise ValueError("compressed_axes must be an iterable")
    if len(compressed_axes) == ndim:
        raise ValueError("cannot compress all axes")
    if len(set(compressed_axes)) != len(compressed_axes):
        raise ValueError("axes must be unique")
    if not all(isinstance(a, (int, np.integer)) for a in compressed_axes):
        raise ValueError("axes must be integers")
    if min(compressed_axes) < 0 or max(compressed_axes) >= ndim:
        raise ValueError("axis out of range")


def c_strides(shape):
    """Element strides of a C-ordered array of `shape`."""
    st = [1] * len(shape)
    for d in range(len(shape) - 2, -1, -1):
        st[d] = st[d + 1] * int(shape[d + 1])
    return st


def key_bits(size) -> int:
    """Number of low bits that can be set in a linear index < size."""
    return max(1, int(max(int(size) - 1, 1)).bit_length())


def check_linear_range(shape):
    if prod(shape) >= 2**63:
        raise ValueError(f"shape {tuple(shape)} is too large for a 63-bit linear index")


def isscalar(x):
    return np.isscalar(x) or (isinstance(x, np.ndarray) and x.ndim == 0) or isinstance(x, np.generic)
