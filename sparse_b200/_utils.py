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
    """Equality in the sense the reference prunes by (_utils.py:406-452 upstream): integers and booleans compare by
    value; floating-point and complex values compare by BIT PATTERN, so +0.0 and -0.0 differ and a NaN equals the
    identical NaN.  `loose=True` is ordinary equality that also lets any NaN equal any NaN.  Fill values are host
    scalars, but arrays broadcast as usual."""
    x, y = np.asarray(x), np.asarray(y)
    common = np.result_type(x.dtype, y.dtype)
    if common.kind not in "fc":
        return x == y
    if loose:
        return (x == y) | (np.isnan(x) & np.isnan(y))
    xb, yb = np.broadcast_arrays(x.astype(common), y.astype(common))
    raw = [np.ascontiguousarray(v).reshape(-1).view(np.uint8).reshape(-1, common.itemsize) for v in (xb, yb)]
    return (raw[0] == raw[1]).all(axis=1).reshape(xb.shape)[()]


def check_zero_fill_value(*args, loose=True):
    """tensordot / matmul / dot / kron are only defined for zero fill values (_utils.py:562-596 upstream; same
    message).  Operands without a fill value (ndarrays) and empty operands pass."""
    for position, operand in enumerate(args):
        fv = getattr(operand, "fill_value", None)
        if fv is None or getattr(operand, "size", 1) == 0:
            continue
        if not equivalent(fv, _zero_of_dtype(operand.dtype), loose=loose):
            raise ValueError(
                f"This operation requires zero fill values, but argument {position:d} had a fill value of {fv!s}."
            )


def check_fill_value(x, /, *, accept_fv=None):
    """Raise unless x's fill value is (loosely) one of `accept_fv` -- a scalar or a collection, default zero
    (_utils.py:531-560 upstream; same message)."""
    allowed = [0] if accept_fv is None else (list(accept_fv) if isinstance(accept_fv, Iterable) else [accept_fv])
    for fv in allowed:
        if equivalent(fv, x.fill_value, loose=True):
            return
    raise ValueError(f"{x.fill_value=} but should be in {allowed}.")


def normalize_axis(axis, ndim):
    """None stays None; an integer becomes its non-negative form; a collection is normalised item by item
    (`normalize_axis` upstream; ValueError with upstream's wording for anything else)."""
    if axis is None:
        return None
    if isinstance(axis, Iterable) and not isinstance(axis, (str, bytes)):
        return tuple(normalize_axis(a, ndim) for a in axis)
    try:
        a = operator.index(axis)
    except TypeError:
        raise ValueError(f"axis {axis} not understood") from None
    if not -ndim <= a < ndim:
        # upstream's message shows the index after its `+= ndim` (_utils.py:389-393): -5 with ndim 4 reads -1
        raise ValueError(f"Invalid axis index {a + ndim if a < 0 else a} for ndim={ndim}")
    return a + ndim if a < 0 else a


def can_store(dtype, scalar) -> bool:
    """_utils.py:651-658."""
    try:
        with np.errstate(all="raise"):
            return bool(np.array(scalar, dtype=dtype) == np.array(scalar))
    except (ValueError, OverflowError, FloatingPointError):
        return False


def check_compressed_axes(ndim, compressed_axes):
    """Validate GCXS `compressed_axes` against an ndim (or a shape): a collection of distinct in-range integers that
    leaves at least one axis uncompressed (`check_compressed_axes` upstream; same messages)."""
    if compressed_axes is None:
        return
    nd = len(ndim) if isinstance(ndim, Iterable) else ndim
    if not isinstance(compressed_axes, Iterable):
        raise ValueError("compressed_axes must be an iterable")
    axes = list(compressed_axes)
    rules = (
        (lambda: len(axes) == nd, "cannot compress all axes"),
        (lambda: len(set(axes)) != len(axes), "axes must be unique"),
        (lambda: any(not isinstance(a, (int, np.integer)) for a in axes), "axes must be integers"),
        (lambda: min(axes) < 0 or max(axes) >= nd, "axis out of range"),
    )
    for broken, message in rules:
        if broken():
            raise ValueError(message)


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
