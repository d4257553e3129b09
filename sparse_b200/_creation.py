"""Array creation and the Array-API style function forms of the reference's namespace.

Mirrors sparse/numba_backend/_common.py:1561-1928 (`eye`, `full[_like]`, `zeros[_like]`, `ones[_like]`,
`empty[_like]`, `can_cast`, `asnumpy`), :2077-2136 (`asarray`), :2267-3118 (`permute_dims`, `std`, `var`, `abs`,
`reshape`, `astype`, `equal`, `round`, `isinf`, `isnan`, `nonzero`, `imag`, `real`, `vecdot`) and
_coo/common.py:584-611, 937-999 (`argwhere`, `isposinf`, `isneginf`, `result_type`).  A constant array is a sparse
array with no stored entries and the constant as its fill value, so creation never touches the device; everything
else forwards to the device-backed methods of the containers.
"""
from __future__ import annotations

import builtins

import numpy as np

from ._coo import COO, as_coo
from ._sparse_array import SparseArray
from ._utils import check_zero_fill_value, normalize_axis


def _check_device(device, method=False):
    """Arrays live on the CUDA device; `"cpu"` is accepted so that Array-API code written against upstream (which only
    knows `"cpu"`, _common.py:33-41, _sparse_array.py:55-57) keeps running."""
    if device in (None, "cpu", "cuda") or str(device).startswith("cuda"):
        return
    if method:
        raise ValueError(f"Only `device='cuda'` (or 'cpu', for compatibility) is supported. Got {device!r}.")
    raise ValueError(f"Device must be `'cuda'`, `'cpu'` (accepted for compatibility) or `None`. Got {device!r}.")


def full(shape, fill_value, dtype=None, format="coo", order="C", *, device=None, **kwargs):
    """_common.py:1629-1681."""
    _check_device(device)
    if dtype is None:
        dtype = np.array(fill_value).dtype
    if not isinstance(shape, tuple):
        shape = tuple(shape) if hasattr(shape, "__iter__") else (shape,)
    if order not in {"C", None}:
        raise NotImplementedError("Currently, only 'C' and None are supported.")
    out = COO(np.empty((len(shape), 0), dtype=np.intp), np.empty(0, dtype=dtype), shape=shape,
              fill_value=np.dtype(dtype).type(fill_value), has_duplicates=False, sorted=True)
    return out.asformat(format, **kwargs)


def full_like(a, fill_value, dtype=None, shape=None, format=None, *, device=None, **kwargs):
    """_common.py:1684-1724."""
    if format is None and not isinstance(a, np.ndarray):
        format = type(a).__name__.lower()
    elif format is None:
        format = "coo"
    kwargs.pop("compressed_axes", None)
    return full(a.shape if shape is None else shape, fill_value, dtype=(a.dtype if dtype is None else dtype),
                format=format, device=device, **kwargs)


def zeros(shape, dtype=float, format="coo", *, device=None, **kwargs):
    return full(shape, 0, np.dtype(dtype), format=format, device=device, **kwargs)


def zeros_like(a, dtype=None, shape=None, format=None, *, device=None, **kwargs):
    return full_like(a, 0, dtype=dtype, shape=shape, format=format, device=device, **kwargs)


def ones(shape, dtype=float, format="coo", *, device=None, **kwargs):
    return full(shape, 1, np.dtype(dtype), format=format, device=device, **kwargs)


def ones_like(a, dtype=None, shape=None, format=None, *, device=None, **kwargs):
    return full_like(a, 1, dtype=dtype, shape=shape, format=format, device=device, **kwargs)


empty = zeros
empty_like = zeros_like


def eye(N, M=None, k=0, dtype=float, format="coo", *, device=None, **kwargs):
    """Ones on the k-th diagonal (_common.py:1561-1626)."""
    _check_device(device)
    M = N if M is None else M
    N, M, k = int(N), int(M), int(k)
    n = builtins.min(N, M)
    if k > 0:
        n = builtins.max(builtins.min(n, M - k), 0)
    elif k < 0:
        n = builtins.max(builtins.min(n, N + k), 0)
    if n == 0:
        return zeros((N, M), dtype=dtype, format=format, device=device)
    base = np.arange(n, dtype=np.intp)
    coords = np.stack([base - builtins.min(k, 0), base + builtins.max(k, 0)])
    return COO(coords, np.ones(n, dtype=dtype), shape=(N, M), has_duplicates=False, sorted=True).asformat(format,
                                                                                                         **kwargs)


def asarray(obj, /, *, dtype=None, format=None, copy=False, device=None, **kwargs):
    """_common.py:2077-2136."""
    from ._coo import _is_scipy_sparse
    from ._gcxs import CSC, CSR, GCXS

    _check_device(device)
    from ._dok import DOK

    if format not in {None, "coo", "dok", "gcxs", "csc", "csr"}:
        raise ValueError(f"{format} format not supported.")
    classes = {"coo": COO, "dok": DOK, "gcxs": GCXS, "csc": CSC, "csr": CSR}
    if isinstance(obj, SparseArray):
        if copy:
            obj = obj.copy()
        out = obj.asformat(format, **kwargs) if format is not None else obj
        return out.astype(dtype, copy=False) if dtype is not None else out
    format = "coo" if format is None else format
    if _is_scipy_sparse(obj):
        out = classes[format].from_scipy_sparse(obj)
    elif np.isscalar(obj) or isinstance(obj, (np.ndarray, list, tuple)) or hasattr(obj, "__iter__"):
        out = classes[format].from_numpy(np.asarray(obj))
    else:
        raise ValueError(f"{type(obj)} not supported.")
    return out.astype(dtype, copy=False) if dtype is not None and out.dtype != np.dtype(dtype) else out


def asnumpy(a, dtype=None, order=None):
    """Dense NumPy array of `a` (_common.py:1928-1951)."""
    if isinstance(a, SparseArray):
        a = a.todense()
    return np.array(a, dtype=dtype, copy=False, order=order) if np.lib.NumpyVersion(np.__version__) < "2.0.0" \
        else np.asarray(a, dtype=dtype, order=order)


def can_cast(from_, to, /, *, casting="safe"):
    """_common.py:1863-1892."""
    from_ = from_.dtype if isinstance(from_, (SparseArray, np.ndarray)) else np.dtype(from_)
    to = to.dtype if isinstance(to, (SparseArray, np.ndarray)) else np.dtype(to)
    return np.can_cast(from_, to, casting=casting)


def result_type(*arrays_and_dtypes):
    """numpy.result_type; a sparse array stands in by its dtype, a 0-D one by its (dense) scalar value because NumPy
    promotes 0-D arrays differently (_coo/common.py:991-1009)."""
    def arg(x):
        if not isinstance(x, SparseArray):
            return x
        return x.dtype if x.ndim > 0 else x.todense()

    return np.result_type(*(arg(x) for x in arrays_and_dtypes))


# ---- function forms -------------------------------------------------------------------------------------------
def std(x, /, *, axis=None, correction=0.0, keepdims=False):
    return x.std(axis=axis, ddof=correction, keepdims=keepdims)


def var(x, /, *, axis=None, correction=0.0, keepdims=False):
    return x.var(axis=axis, ddof=correction, keepdims=keepdims)


def abs(x, /):
    return x.__abs__()


def reshape(x, /, shape, *, copy=None):
    return x.reshape(shape=shape)


def astype(x, dtype, /, *, copy=True, device=None):
    _check_device(device)
    return x.astype(dtype, copy=copy)


def equal(x1, x2, /):
    return x1 == x2


def round(x, /, decimals=0, out=None):
    return x.round(decimals=decimals, out=out)


def isinf(x, /):
    return x.isinf()


def isnan(x, /):
    return x.isnan()


def isposinf(x, out=None):
    """_coo/common.py:937-961."""
    res = np.logical_and(np.isinf(x), np.greater(x, 0))
    if out is not None:
        out._make_shallow_copy_of(res)
        return out
    return res


def isneginf(x, out=None):
    """_coo/common.py:964-988."""
    res = np.logical_and(np.isinf(x), np.less(x, 0))
    if out is not None:
        out._make_shallow_copy_of(res)
        return out
    return res


def nonzero(x, /):
    """Tuple of coordinate arrays of the non-zero entries (_common.py:3009-3038)."""
    check_zero_fill_value(x)
    c = x.asformat("coo") if isinstance(x, SparseArray) else as_coo(x)
    return tuple(c.coords)


def argwhere(a):
    """(nnz, ndim) coordinates of the non-zero entries (_coo/common.py:584-611)."""
    return np.transpose(nonzero(a))


def imag(x, /):
    return x.imag


def real(x, /):
    return x.real


def vecdot(x1, x2, /, *, axis=-1):
    """Sum of conj(x1) * x2 along `axis` (_common.py:3095-3118)."""
    ndmin = builtins.min(x1.ndim, x2.ndim)
    if not (-ndmin <= axis < ndmin) or x1.shape[axis] != x2.shape[axis]:
        raise ValueError("Shapes must match along `axis`.")
    if np.issubdtype(x1.dtype, np.complexfloating):
        x1 = np.conj(x1)
    normalize_axis(axis, ndmin)
    return np.sum(x1 * x2, axis=axis, dtype=np.result_type(x1, x2))


def diff(x, axis=-1, n=1, prepend=None, append=None):
    """n-th discrete difference along `axis` (_common.py:3234-3264): concatenations and two shifted slices."""
    from ._manip import concatenate

    if not isinstance(x, SparseArray):
        raise TypeError("`x` must be a SparseArray.")
    if axis < 0:
        axis = x.ndim + axis
    if prepend is not None:
        x = concatenate([prepend, x], axis=axis)
    if append is not None:
        x = concatenate([x, append], axis=axis)
    lead = (slice(None),) * axis
    for _ in range(n):
        x = x[lead + (slice(1, None),)] - x[lead + (slice(None, -1),)]
    return x


def interp(x, xp, fp, left=None, right=None, period=None):
    """One-dimensional piecewise-linear interpolation of the stored values and of the fill value (_common.py:3267-3349);
    the result is pruned against the new fill value.

    `np.interp` is not a ufunc, so upstream runs it with NumPy on the host data array.  Here every segment [xp[j], xp[j+1])
    is applied with element-wise device passes in NumPy's own operation order -- `slope * (v - xp[j]) + fp[j]`, multiply
    then add, slope computed once on the host in double like `np.interp` does -- selected by the two comparisons and
    merged with raw-bit ORs, so the values are bit-identical to `np.interp` for real `fp`."""
    from . import _device as D
    from . import _kernels as Kn
    from ._elemwise import _BINARY, _UNARY, _bitor_raw, dense_binary
    from ._elemwise import _sel_x as _sel_x_op
    from ._gcxs import GCXS

    if isinstance(xp, SparseArray):
        xp = xp.todense()
    if isinstance(fp, SparseArray):
        fp = fp.todense()
    if not isinstance(x, SparseArray):
        return np.interp(x, xp, fp, left=left, right=right, period=period)
    if period is not None:
        raise NotImplementedError("sparse_b200.interp: `period` is not on the CUDA path")
    xp, fp = np.asarray(xp, dtype=np.float64), np.asarray(fp)
    if np.iscomplexobj(fp):
        raise TypeError("sparse_b200.interp: complex `fp` is outside the CUDA op set")
    fp = fp.astype(np.float64)
    if xp.ndim != 1 or fp.ndim != 1 or len(xp) != len(fp) or len(xp) == 0:
        raise ValueError("fp and xp are not of the same length." if xp.ndim == 1 and fp.ndim == 1
                         else "Data points must be 1-D sequences")
    from ._dok import DOK

    if isinstance(x, DOK):
        return DOK.from_coo(interp(x.to_coo(), xp, fp, left=left, right=right))
    was_gcxs = isinstance(x, GCXS)
    c = x.asformat("coo")
    new_fill = np.float64(np.interp(np.float64(c.fill_value), xp, fp, left=left, right=right))
    if c.nnz == 0:
        out = COO(np.zeros((c.ndim, 0), dtype=np.intp), np.empty(0, dtype=np.float64), shape=c.shape,
                  has_duplicates=False, sorted=True, fill_value=new_fill)
        return GCXS.from_coo(out, x.compressed_axes) if was_gcxs and out.ndim > 1 else out
    v = Kn.cast(c._data_dev(), np.float64)
    lo = np.float64(fp[0] if left is None else left)
    hi = np.float64(fp[-1] if right is None else right)

    def pick(flag_bool, values):
        """values where the flag is set, +0 elsewhere (device op `_sel_x`)."""
        return dense_binary(_sel_x_op, Kn.cast(flag_bool, np.float64), values)

    def cmp(op, scalar):
        return Kn.ew_map(_BINARY[op], 0, v, np.float64(scalar), False, np.bool_)[0]

    n = int(v.shape[0])
    acc = pick(cmp(np.less, xp[0]), Kn.full(n, lo, np.float64))
    acc = dense_binary(_bitor_raw, acc, pick(cmp(np.greater, xp[-1]), Kn.full(n, hi, np.float64)))
    acc = dense_binary(_bitor_raw, acc, pick(cmp(np.equal, xp[-1]), Kn.full(n, fp[-1], np.float64)))
    for j in range(len(xp) - 1):
        inside = dense_binary(np.logical_and, Kn.cast(cmp(np.greater_equal, xp[j]), np.int32),
                              Kn.cast(cmp(np.less, xp[j + 1]), np.int32))
        with np.errstate(all="ignore"):
            slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])
        seg, _ = Kn.ew_map(_BINARY[np.subtract], 0, v, np.float64(xp[j]), 0, np.float64)
        seg, _ = Kn.ew_map(_BINARY[np.multiply], 0, seg, np.float64(slope), 0, np.float64)
        seg, _ = Kn.ew_map(_BINARY[np.add], 0, seg, np.float64(fp[j]), 0, np.float64)
        acc = dense_binary(_bitor_raw, acc, pick(inside, seg))
    # NaN inputs stay NaN (no comparison selects them): v + 0 where v is NaN
    is_nan, _ = Kn.ew_map(_UNARY[np.isnan], 2, v, None, False, np.bool_)
    acc = dense_binary(_bitor_raw, acc, pick(is_nan, v))
    keys = c.sorted_keys()
    flags = Kn.flag_not_fill(acc, new_fill)
    pos, total = Kn.scan_flags(flags)
    if total != n:
        acc, keys = Kn.compact(acc, flags, pos, total), Kn.compact(keys, flags, pos, total)
    out = COO._from_device(None, acc, c.shape, new_fill, keys=keys)
    return GCXS.from_coo(out, x.compressed_axes) if was_gcxs and out.ndim > 1 else out


def asCOO(x, name="asCOO", check=True):
    """COO of `x`; a dense input is refused when `check` is set (_coo/common.py:21-53)."""
    from ._coo import _is_scipy_sparse

    if check and not (isinstance(x, SparseArray) or _is_scipy_sparse(x)):
        raise ValueError(f"Performing this operation would produce a dense result: {name}")
    return x if isinstance(x, COO) else COO(x)


def broadcast_shapes(*shapes):
    """numpy.broadcast_shapes (_coo/common.py:1599-1617)."""
    return np.broadcast_shapes(*shapes)


def broadcast_arrays(*arrays):
    """Every array broadcast to the common shape (_common.py:2835-2869); dense inputs become COO."""
    shape = np.broadcast_shapes(*[a.shape for a in arrays])
    out = []
    for a in arrays:
        if isinstance(a, (np.generic, np.ndarray)):
            a = COO.from_numpy(np.asarray(a))
        out.append(a.broadcast_to(shape))
    return tuple(out)
