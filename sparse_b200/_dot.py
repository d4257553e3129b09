"""tensordot / matmul / dot and the `_dot` format dispatch.

Host-side mirror of sparse/numba_backend/_common.py:95-503 (same names, argument meaning, result types and
errors); every kernel call goes to libsparse_b200 (K1 spmm, K3 sparse-out, K4 spgemm).  The dispatch table
is the reference's; how each row maps to a device kernel is noted inline.
"""
from __future__ import annotations

import builtins
import warnings
from itertools import chain

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO, _is_scipy_sparse, as_coo
from ._gcxs import GCXS
from ._sparse_array import SparseArray
from ._utils import check_zero_fill_value


def _dot_dtype(dt1, dt2):
    """_common.py:635-636."""
    return (np.zeros((), dtype=dt1) * np.zeros((), dtype=dt2)).dtype


def _is_dense(x):
    return isinstance(x, np.ndarray) or D.is_device_tensor(x)


def _dense_dtype(x):
    return x.dtype if isinstance(x, np.ndarray) else D.np_dtype(x)


def _dense_dev(x, dtype):
    """Dense operand on the device, cast to `dtype`, 2-D."""
    if isinstance(x, np.ndarray):
        xd = D.upload(np.ascontiguousarray(x))
    else:
        xd = x
    return Kn.cast(xd, dtype)


def _narrow_idx(*tensors, limit):
    """int32 index arrays when every value fits (saves 4 B/nnz of HBM traffic), else int64 for all."""
    t = D.torch()
    if limit < 2**31 - 1:
        return [x if x.dtype == t.int32 else Kn.cast(x, np.int32) for x in tensors]
    return [x if x.dtype == t.int64 else Kn.cast(x, np.int64) for x in tensors]


def _csr_arrays(g, dtr):
    """(data, indices, indptr) of a 2-D GCXS on the device: data in dtr, indices narrowed when legal."""
    data, indices, indptr = g._dev()
    nrows, ncols = g._compressed_shape
    indices, indptr = _narrow_idx(indices, indptr, limit=max(nrows, ncols, g.nnz))
    return Kn.cast(data, dtr), indices, indptr


def _coo_as_csr(c, dtr, by_col=False):
    """2-D canonical COO -> CSR arrays (by_col: CSR of the transpose, i.e. CSC)."""
    data = c._data_dev()
    M, N = c.shape
    if not by_col:
        keys = c.sorted_keys()
        _, indices, indptr = Kn.csr_from_keys(keys, M, N, np.int64)
        d = data
    else:
        keys = Kn.linearize(c._dev()[0], [1, M])  # key over (col, row)
        unsorted, _ = Kn.keys_flags(keys)
        d = data
        if unsorted:
            from ._utils import key_bits

            keys, perm = Kn.sort_keys(keys, key_bits(M * N))
            d = Kn.gather(data, perm)
        _, indices, indptr = Kn.csr_from_keys(keys, N, M, np.int64)
    indices, indptr = _narrow_idx(indices, indptr, limit=max(M, N, c.nnz))
    return Kn.cast(d, dtr), indices, indptr


def _return_dense(out_dev, like_inputs):
    """np.ndarray unless a dense operand was a device tensor (then stay on the device)."""
    if builtins.any(D.is_device_tensor(x) for x in like_inputs):
        return out_dev
    return D.download(out_dev)


# --------------------------------------------------------------------------------------------------------------
def tensordot(a, b, axes=2, *, return_type=None):
    """Equivalent of numpy.tensordot (reference: _common.py:95-215)."""
    check_zero_fill_value(a, b)
    if _is_scipy_sparse(a):
        a = GCXS.from_scipy_sparse(a)
    if _is_scipy_sparse(b):
        b = GCXS.from_scipy_sparse(b)
    # any other SparseArray (the DOK builder) computes as COO (_common.py:132-135: `asformat` to a supported type)
    if isinstance(a, SparseArray) and not isinstance(a, (COO, GCXS)):
        a = a.asformat("coo")
    if isinstance(b, SparseArray) and not isinstance(b, (COO, GCXS)):
        b = b.asformat("coo")
    try:
        iter(axes)
    except TypeError:
        axes_a = list(range(-axes, 0))
        axes_b = list(range(axes))
    else:
        axes_a, axes_b = axes
    try:
        na = len(axes_a)
        axes_a = list(axes_a)
    except TypeError:
        axes_a = [axes_a]
        na = 1
    try:
        nb = len(axes_b)
        axes_b = list(axes_b)
    except TypeError:
        axes_b = [axes_b]
        nb = 1

    as_ = tuple(a.shape)
    nda = len(as_)
    bs = tuple(b.shape)
    ndb = len(bs)
    equal = True
    if nda == 0 or ndb == 0:
        if axes_a == [] and axes_b == []:
            if nda == 0 and isinstance(a, SparseArray):
                a = a.todense()
            if ndb == 0 and isinstance(b, SparseArray):
                b = b.todense()
            return a * b
        pos = int(nda != 0)
        raise ValueError(f"Input {pos} operand does not have enough dimensions")
    if na != nb:
        equal = False
    else:
        for k in range(na):
            if as_[axes_a[k]] != bs[axes_b[k]]:
                equal = False
                break
            if axes_a[k] < 0:
                axes_a[k] += nda
            if axes_b[k] < 0:
                axes_b[k] += ndb
    if not equal:
        raise ValueError("shape-mismatch for sum")

    notin = [k for k in range(nda) if k not in axes_a]
    newaxes_a = notin + axes_a
    N2 = 1
    for axis in axes_a:
        N2 *= as_[axis]
    newshape_a = (-1, N2)
    olda = [as_[axis] for axis in notin]

    notin = [k for k in range(ndb) if k not in axes_b]
    newaxes_b = axes_b + notin
    N2 = 1
    for axis in axes_b:
        N2 *= bs[axis]
    newshape_b = (N2, -1)
    oldb = [bs[axis] for axis in notin]

    if builtins.any(dim == 0 for dim in chain(newshape_a, newshape_b)):
        dt = np.result_type(_any_dtype(a), _any_dtype(b))
        res = COO(np.empty((len(olda) + len(oldb), 0), dtype=np.uintp), data=np.empty(0, dtype=dt),
                  shape=tuple(olda + oldb), has_duplicates=False, sorted=True)
        if _is_dense(a) or _is_dense(b):
            res = np.zeros(tuple(olda + oldb), dtype=dt)
        return res

    at = _transpose_reshape(a, newaxes_a, newshape_a)
    bt = _transpose_reshape(b, newaxes_b, newshape_b)
    res = _dot(at, bt, return_type)
    return res.reshape(tuple(olda + oldb))


def _any_dtype(x):
    return _dense_dtype(x) if _is_dense(x) else x.dtype


def _transpose_reshape(x, axes, newshape):
    """x.transpose(axes).reshape(newshape) (_common.py:212-213); one fused device pass for COO."""
    if isinstance(x, COO):
        n = x.size
        shape = tuple(n // newshape[1] if s == -1 else s for s in newshape) if newshape[0] == -1 else tuple(
            n // newshape[0] if s == -1 else s for s in newshape)
        if tuple(axes) == tuple(range(x.ndim)):
            return x.reshape(shape)
        return x._permute_reshape(tuple(axes), shape)
    if isinstance(x, GCXS):
        n = x.size
        shape = tuple(n // newshape[1] if s == -1 else s for s in newshape) if newshape[0] == -1 else tuple(
            n // newshape[0] if s == -1 else s for s in newshape)
        if tuple(axes) == tuple(range(x.ndim)) and shape == tuple(x.shape):
            return x
        if x.ndim == 2 and tuple(axes) == (1, 0) and shape == tuple(x.shape[::-1]):
            return x._2d_transpose()
        identity = tuple(axes) == tuple(range(x.ndim))
        c = x.tocoo()
        c = c.reshape(shape) if identity else c._permute_reshape(tuple(axes), shape)
        # the reference keeps GCXS here.  Compressed axis of the 2-D view: a 2-D operand keeps the axis its (O(1))
        # transpose has (compressed.py:728-729,761; reshape to the same ndim keeps it: :665-667); any other operand
        # gets reshape's default, the shorter axis (:671)
        if x.ndim == 2:
            ca = (int(x.compressed_axes[0]) if identity else (int(x.compressed_axes[0]) + 1) % 2,)
        else:
            ca = (int(np.argmin(shape)),)
        return GCXS.from_coo(c, ca)
    if isinstance(x, np.ndarray):
        return x.transpose(axes).reshape(newshape)
    # device tensor
    return x.permute(*axes).reshape(newshape)


def matmul(a, b):
    """Equivalent of numpy.matmul (reference: _common.py:218-293)."""
    check_zero_fill_value(a, b)
    if not hasattr(a, "ndim") or not hasattr(b, "ndim"):
        raise TypeError(f"Cannot perform dot product on types {type(a)}, {type(b)}")
    if isinstance(a, SparseArray) and not isinstance(a, (COO, GCXS)):
        a = a.asformat("coo")
    if isinstance(b, SparseArray) and not isinstance(b, (COO, GCXS)):
        b = b.asformat("coo")
    if _check_nan(a) or _check_nan(b):
        warnings.warn("Nan will not be propagated in matrix multiplication", RuntimeWarning, stacklevel=1)

    if b.ndim <= 2:
        return dot(a, b)
    if a.ndim <= 2:
        res = dot(a, b)
        axes = list(range(res.ndim))
        axes.insert(-1, axes.pop(0))
        return res.transpose(axes)
    if a.ndim <= b.ndim and np.prod(a.shape[:-1]) == 1:
        res = dot(a.reshape(-1), b)
        shape = list(res.shape)
        shape.insert(-1, 1)
        return res.reshape(shape)
    if b.ndim <= a.ndim and np.prod(b.shape[:-2]) == 1:
        return dot(a, b.reshape(b.shape[-2:]))

    if a.ndim < b.ndim:
        a = a[(None,) * (b.ndim - a.ndim)]
    if a.ndim > b.ndim:
        b = b[(None,) * (a.ndim - b.ndim)]
    for i, j in zip(a.shape[:-2], b.shape[:-2], strict=True):
        if i != 1 and j != 1 and i != j:
            raise ValueError("shapes of a and b are not broadcastable")

    def _matmul_recurser(a, b):
        if a.ndim == 2:
            return dot(a, b)
        res = []
        for i in range(builtins.max(a.shape[0], b.shape[0])):
            a_i = a[0] if a.shape[0] == 1 else a[i]
            b_i = b[0] if b.shape[0] == 1 else b[i]
            res.append(_matmul_recurser(a_i, b_i))
        mask = [isinstance(x, SparseArray) for x in res]
        if builtins.all(mask):
            return stack(res)
        res = [x.todense() if isinstance(x, SparseArray) else x for x in res]
        return np.stack(res)

    return _matmul_recurser(a, b)


def _check_nan(x):
    """check_class_nan / nan_check (_common.py:51-92): any-NaN scan of the stored data."""
    if isinstance(x, COO):
        if x.dtype.kind != "f" or x.nnz == 0:
            return False
        return Kn.any_nan(x._data_dev())
    if isinstance(x, GCXS):
        if x.dtype.kind != "f" or x.nnz == 0:
            return False
        return Kn.any_nan(x._dev()[0])
    if isinstance(x, np.ndarray):
        return x.dtype.kind == "f" and bool(np.isnan(np.min(x))) if x.size else False
    if D.is_device_tensor(x):
        return Kn.any_nan(x.contiguous().reshape(-1))
    return False


def stack(arrays, axis=0):
    """Stack the sparse results of the batched matmul (_common.py:288): device re-keying in `_manip.stack`."""
    from ._manip import stack as _stack

    return _stack(arrays, axis)


def dot(a, b):
    """Equivalent of numpy.dot (reference: _common.py:296-336)."""
    check_zero_fill_value(a, b)
    if not hasattr(a, "ndim") or not hasattr(b, "ndim"):
        raise TypeError(f"Cannot perform dot product on types {type(a)}, {type(b)}")
    if a.ndim == 1 and b.ndim == 1:
        if isinstance(a, SparseArray):
            a = as_coo(a)
        if isinstance(b, SparseArray):
            b = as_coo(b)
        return (a * b).sum()
    a_axis = -1
    b_axis = -2
    if b.ndim == 1:
        b_axis = -1
    return tensordot(a, b, axes=(a_axis, b_axis))


# --------------------------------------------------------------------------------------------------------------
def _wrap_gcxs(data, indices, indptr, shape, compressed_axes, return_type):
    out = GCXS._from_device(data, indices, indptr, shape, compressed_axes)
    if return_type == np.ndarray:
        return out.todense()
    if return_type == COO:
        return out.tocoo()
    return out


def _spgemm_csr(a, b, out_shape, dtr, *, wide=False):
    """CSR(a) @ CSR(b) with pruning (the GCXS(..., prune=True) of _common.py:374-379 fused into the finish pass)."""
    ad, ai, ap = _csr_arrays(a, dtr)
    bd, bi, bp = _csr_arrays(b, dtr)
    t = D.torch()
    if ai.dtype != bi.dtype:
        ai, ap, bi, bp = [Kn.cast(x, np.int64) for x in (ai, ap, bi, bp)]
    M, K = a._compressed_shape
    n_col = b._compressed_shape[1]
    indptr, indices, _, data, _ = Kn.spgemm(ap, ai, ad, bp, bi, bd, M, K, n_col, sorted_order=False, wide=wide,
                                            prune=True)
    return data, indices, indptr


def _dot(a, b, return_type=None):
    """Format dispatch of _common.py:339-503 (2-D operands)."""
    from . import _complex as C

    if (C.is_complex(a) or C.is_complex(b)) and not (isinstance(a, np.ndarray) and isinstance(b, np.ndarray)):
        return C.dot_complex(_dot, a, b, return_type)  # four real products on the planes
    from ._elemwise import _WIDE_FOR

    dts = [_dense_dtype(x) if _is_dense(x) else x.dtype for x in (a, b)]
    dtr = _dot_dtype(*dts)
    if dtr in _WIDE_FOR and not (isinstance(a, np.ndarray) and isinstance(b, np.ndarray)):
        # narrow / unsigned integer products: exact in the wider signed type, cast back = NumPy's modular arithmetic
        W = _WIDE_FOR[dtr]
        host_in = not builtins.any(D.is_device_tensor(x) for x in (a, b))

        def widen(x):
            if isinstance(x, SparseArray):
                if x.dtype == W:
                    return x
                return x.astype(W, _keep_format=True) if isinstance(x, GCXS) else x.astype(W, _raw=True)
            return Kn.cast(D.upload(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x, W)

        out = _dot(widen(a), widen(b), return_type)
        if isinstance(out, SparseArray):
            return out.astype(dtr, _keep_format=True) if isinstance(out, GCXS) else out.astype(dtr, _raw=True)
        out = Kn.cast(out if D.is_device_tensor(out) else D.upload(np.ascontiguousarray(out)), dtr)
        return D.download(out) if host_in else out
    out_shape = (a.shape[0], b.shape[1])
    if builtins.all(isinstance(arr, SparseArray) for arr in [a, b]) and builtins.any(
            isinstance(arr, GCXS) for arr in [a, b]):
        a = a.asformat("gcxs")
        b = b.asformat("gcxs", compressed_axes=a.compressed_axes)

    # ---- GCXS @ GCXS -> K4 --------------------------------------------------------------------------------
    if isinstance(a, GCXS) and isinstance(b, GCXS):
        if a.nbytes > b.nbytes:
            b = b.change_compressed_axes(a.compressed_axes)
        else:
            a = a.change_compressed_axes(b.compressed_axes)
        dtr = _dot_dtype(a.dtype, b.dtype)
        if a.compressed_axes == (0,):  # csr @ csr
            compressed_axes = (0,)
            data, indices, indptr = _spgemm_csr(a, b, out_shape, dtr)
        else:  # csc @ csc: a @ b = (b.T @ a.T).T with the same arrays (_common.py:362-373)
            compressed_axes = (1,)
            data, indices, indptr = _spgemm_csr(b._2d_transpose(), a._2d_transpose(), out_shape[::-1], dtr)
        return _wrap_gcxs(data, indices, indptr, out_shape, compressed_axes, return_type)

    # ---- GCXS @ dense ----------------------------------------------------------------------------------------
    if isinstance(a, GCXS) and _is_dense(b):
        dtr = _dot_dtype(a.dtype, _dense_dtype(b))
        M, N = out_shape
        K = a.shape[1]
        dense_out = return_type is None or return_type == np.ndarray
        if (a.compressed_axes == (0,) and dense_out and a._data is None and isinstance(b, np.ndarray)
                and a.dtype == dtr and b.dtype == dtr and a.nnz > 0):
            # both operands live on the host: streamed H2D / K1 / D2H pipeline of the host-buffer ABI call
            # (use a.to_device() first to keep A resident in HBM across calls instead)
            out = D.pinned_empty((M, N), dtr)
            return Kn.spmm_csr_dense_host(a._data_np, a._indices_np, a._indptr_np, b, out=out)
        bd = _dense_dev(b, dtr)
        if a.compressed_axes == (0,):  # csr @ ndarray
            ad, ai, ap = _csr_arrays(a, dtr)
            if dense_out:
                return _return_dense(Kn.spmm_csr_dense(ad, ai, ap, bd, M, K, N, long_rows=a._has_long_rows()), [b])
            out, flags = Kn.spmm_csr_dense_flagged(ad, ai, ap, bd, M, K, N)
            _, cols, data, indptr = Kn.dense_to_csr(out, flags=flags)
            g = GCXS._from_device(data, cols, indptr, out_shape, (0,))
            g._prune()
            return g.tocoo() if return_type == COO else g
        # csc @ ndarray
        if dense_out:
            # out[r,:] = sum over columns c (ascending) of A[r,c] * b[c,:]  == CSR(A) @ b in stored order
            csr = a.change_compressed_axes((0,))
            ad, ai, ap = _csr_arrays(csr, dtr)
            return _return_dense(Kn.spmm_csr_dense(ad, ai, ap, bd, M, K, N), [b])
        # sparse out, compressed by output column: out^T = sparsify(b^T) @ A^T, float64 accumulator (:835)
        data, indices, indptr = _csc_dense_sparse(a, bd, dtr)
        g = GCXS._from_device(data, indices, indptr, out_shape, (1,))
        return g.tocoo() if return_type == COO else g

    # ---- dense @ GCXS ----------------------------------------------------------------------------------------
    if _is_dense(a) and isinstance(b, GCXS):
        dtr = _dot_dtype(_dense_dtype(a), b.dtype)
        M, N = out_shape
        K = b.shape[0]
        ad = _dense_dev(a, dtr)
        dense_out = return_type is None or return_type == np.ndarray
        bt = b._2d_transpose()  # (N, K); constant-time
        if dense_out:
            # out^T (N x M) = CSR(b^T) @ a^T in ascending-k order for both layouts of b
            csr_bt = bt if bt.compressed_axes == (0,) else bt.change_compressed_axes((0,))
            td, ti, tp = _csr_arrays(csr_bt, dtr)
            at = Kn.transpose_dense(ad)  # (K, M)
            out_t = Kn.spmm_csr_dense(td, ti, tp, at, N, K, M)
            return _return_dense(Kn.transpose_dense(out_t), [a])
        if b.compressed_axes == (0,):
            # _dot_csc_ndarray_sparse(bt, at): rows of out = sparsify(a) @ CSR(b), wide accumulator
            data, indices, indptr = _csc_dense_sparse(bt, Kn.transpose_dense(ad), dtr)
            g = GCXS._from_device(data, indices, indptr, out_shape, (0,))
            return g.tocoo() if return_type == COO else g
        # b is CSC: _dot_csr_ndarray_sparse(bt (CSR of b^T), at) -> (N x M) CSR == GCXS ca=(1,) of out
        td, ti, tp = _csr_arrays(bt, dtr)
        at = Kn.transpose_dense(ad)
        out, flags = Kn.spmm_csr_dense_flagged(td, ti, tp, at, N, K, M)
        _, cols, data, indptr = Kn.dense_to_csr(out, flags=flags)
        g = GCXS._from_device(data, cols, indptr, out_shape, (1,))
        g._prune()
        return g.tocoo() if return_type == COO else g

    # ---- COO @ COO -> K4 (sorted order = canonical COO, no global sort needed) -------------------------------
    if isinstance(a, COO) and isinstance(b, COO):
        dtr = _dot_dtype(a.dtype, b.dtype)
        ad, ai, ap = _coo_as_csr(a, dtr)
        bd, bi, bp = _coo_as_csr(b, dtr)
        if ai.dtype != bi.dtype:
            ai, ap, bi, bp = [Kn.cast(x, np.int64) for x in (ai, ap, bi, bp)]
        M, K = a.shape
        _, cols, rows, data, _ = Kn.spgemm(ap, ai, ad, bp, bi, bd, M, K, b.shape[1], sorted_order=True, prune=True,
                                           want_indptr=False, want_rows=True)
        t = D.torch()
        coords = t.stack([rows, cols])
        out = COO._from_device(coords, data, out_shape)
        if return_type == np.ndarray:
            return out.todense()
        if return_type == GCXS:
            return out.asformat("gcxs")
        return out

    # ---- COO @ dense ---------------------------------------------------------------------------------------------
    if isinstance(a, COO) and _is_dense(b):
        dtr = _dot_dtype(a.dtype, _dense_dtype(b))
        M, N = out_shape
        K = a.shape[1]
        bd = _dense_dev(b, dtr)
        ad, ai, ap = _coo_as_csr(a, dtr)
        out = Kn.spmm_csr_dense(ad, ai, ap, bd, M, K, N)  # same per-element order as _dot_coo_ndarray (:979-1014)
        if return_type is None or return_type == np.ndarray:
            return _return_dense(out, [b])
        return _dense_to_coo(out, out_shape, return_type)  # `if data_curr != 0` (:1062)

    # ---- dense @ COO ---------------------------------------------------------------------------------------------
    if _is_dense(a) and isinstance(b, COO):
        dtr = _dot_dtype(_dense_dtype(a), b.dtype)
        M, N = out_shape
        K = b.shape[0]
        ad = _dense_dev(a, dtr)
        td, ti, tp = _coo_as_csr(b, dtr, by_col=True)  # CSR of b^T (N x K), ascending k per column of b
        at = Kn.transpose_dense(ad)
        out = Kn.transpose_dense(Kn.spmm_csr_dense(td, ti, tp, at, N, K, M))
        if return_type is None or return_type == np.ndarray:
            return _return_dense(out, [a])
        return _dense_to_coo(out, out_shape, return_type)  # `if data_curr != 0` (:1149), then prune=True (:497)

    if _is_dense(a) and _is_dense(b):
        if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
            return np.dot(a, b)
        raise TypeError("sparse_b200: dense @ dense is not part of the sparse hot path")

    raise TypeError("Unsupported types.")


def _dense_to_coo(out_dev, out_shape, return_type):
    rows, cols, data, _ = Kn.dense_to_csr(out_dev.contiguous(), mode=0, want_rows=True, want_indptr=False)
    t = D.torch()
    out = COO._from_device(t.stack([rows, cols]), data, out_shape)
    if return_type == GCXS:
        return out.asformat("gcxs")
    return out


def _csc_dense_sparse(a_csc, bd, dtr):
    """_dot_csc_ndarray_sparse (_common.py:807-866): A given by columns (GCXS ca=(1,), shape (R, C)), b dense (C x N).
    Result compressed by output column j: row j of out^T = sum_c b[c,j] * A[:,c] over the NON-ZERO b[c,j] in
    ascending c -- a Gustavson product sparsify(b^T) @ CSR(A^T) with a float64 accumulator; entries whose
    sum is 0 are skipped.  Returns (data, indices, indptr) with indptr over the N output columns."""
    data, indices, indptr = a_csc._dev()
    R, C = a_csc.shape
    N = int(bd.shape[1])
    bt = Kn.transpose_dense(bd)  # (N x C)
    _, bcols, bdata, bptr = Kn.dense_to_csr(bt, mode=0)  # structural test `u != 0` (:842)
    ai, ap = _narrow_idx(indices, indptr, limit=2**40)  # keep int64: mixes with dense_to_csr's int64 output
    out_ptr, out_idx, _, out_data, _ = Kn.spgemm(bptr, bcols, bdata, ap, ai, Kn.cast(data, dtr), N, C, R,
                                                 sorted_order=False, wide=True, prune=True)
    return out_data, out_idx, out_ptr
