"""Complex dtypes on the real kernels: every complex operation is split into its real and imaginary planes.

The CUDA dtype matrix is real (float32 / float64 / int32 / int64); the reference's dot kernels are compiled per dtype
pair by numba and therefore also run for complex operands (tests/test_dot.py:303-335 upstream).  A complex array is
stored as complex64 / complex128 (the structural kernels are element-size generic) and arithmetic is done on planes:

    (ar + i ai) @ (br + i bi) = (ar @ br - ai @ bi) + i (ar @ bi + ai @ br)        four real products (K1 / K3 / K4)
    add / subtract / multiply / divide / conj / abs / equal, sum                     the same identities per element

each real product / element-wise step being the existing device path.  Planes of a sparse operand share its
coordinates (zeros stay stored in a plane); results of different planes can differ in structure, so `combine` aligns
the two on their UNION with two passes of the merge kernel (raw-bit OR with +0: exact) whose prune test is disabled by
an impossible fill value,
interleaves the aligned values (a strided copy) and finally prunes the entries equal to the complex fill value.

Values differ from the reference's complex multiply-accumulate only by rounding (sum of products taken per plane):
tolerance parity, like the reductions.  The interleave / de-interleave copies are torch views; no arithmetic on the
host, no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO
from ._sparse_array import SparseArray


def is_complex(x) -> bool:
    dt = getattr(x, "dtype", None)
    if dt is None:
        return isinstance(x, complex)
    if D.is_device_tensor(x):
        return x.is_complex()
    return np.dtype(dt).kind == "c"


def plane_dtype(cdtype) -> np.dtype:
    return np.dtype(np.float32) if np.dtype(cdtype) == np.complex64 else np.dtype(np.float64)


def complex_dtype(rdtype) -> np.dtype:
    return np.dtype(np.complex64) if np.dtype(rdtype) == np.float32 else np.dtype(np.complex128)


def _interleave(re, im):
    t = D.torch()
    return t.view_as_complex(t.stack([re, im], dim=-1).contiguous())


def _deinterleave(c):
    r = D.torch().view_as_real(c)
    return r[..., 0].contiguous(), r[..., 1].contiguous()


# ---------------------------------------------------------------------------------------------------------------
# planes of operands
# ---------------------------------------------------------------------------------------------------------------
def planes(x, rdt=None):
    """(real, imag) of an operand, each of dtype `rdt` (default: the operand's own plane precision).

    COO / GCXS: two arrays sharing the operand's coordinates; dense: two device tensors; scalars: two scalars.
    A real operand has imag = None."""
    from ._gcxs import GCXS

    if isinstance(x, SparseArray):
        cplx = x.dtype.kind == "c"
        rdt = np.dtype(rdt or (plane_dtype(x.dtype) if cplx else x.dtype))
        fill = np.asarray(x.fill_value)
        if isinstance(x, GCXS):
            data, indices, indptr = x._dev()
            parts = _deinterleave(data) if cplx else (data, None)
            mk = lambda d, f: GCXS._from_device(Kn.cast(d, rdt), indices, indptr, x.shape, x.compressed_axes, rdt.type(f))
        else:
            data = x._data_dev()
            parts = _deinterleave(data) if cplx else (data, None)
            coords = x._coords
            keys = x.sorted_keys() if coords is None else x._keys
            mk = lambda d, f: COO._from_device(coords, Kn.cast(d, rdt), x.shape, rdt.type(f), keys=keys)
        re = mk(parts[0], fill.real)
        im = mk(parts[1], fill.imag) if cplx else None
        return re, im
    if D.is_device_tensor(x):
        if x.is_complex():
            re, im = _deinterleave(x)
            return (Kn.cast(re, rdt), Kn.cast(im, rdt)) if rdt is not None else (re, im)
        return (Kn.cast(x, rdt) if rdt is not None else x), None
    if isinstance(x, np.ndarray) and x.ndim > 0:
        xd = D.upload(np.ascontiguousarray(x))
        return planes(xd, rdt)
    s = np.asarray(x)[()]
    if np.iscomplexobj(s):
        return (s.real if rdt is None else rdt.type(s.real)), (s.imag if rdt is None else rdt.type(s.imag))
    return (s if rdt is None else rdt.type(s)), None


def _weird_nan(rdt):
    """A NaN bit pattern no arithmetic produces: used as the merge kernel's fill so that nothing is pruned."""
    if np.dtype(rdt) == np.float32:
        return np.array([0x7FC0BEEF], dtype=np.uint32).view(np.float32)[0]
    return np.array([0x7FF8DEAD0000BEEF], dtype=np.uint64).view(np.float64)[0]


def combine(re, im, cdtype=None):
    """Complex array re + i*im from two REAL arrays of the same shape (COO / GCXS, dense device tensors or scalars).

    Sparse planes may differ in structure: both are aligned on the union of their stored positions, interleaved and
    pruned of entries equal to the complex fill value."""
    from ._gcxs import GCXS

    if not isinstance(re, SparseArray) and not isinstance(im, SparseArray):
        if D.is_device_tensor(re) or D.is_device_tensor(im):
            shape = re.shape if D.is_device_tensor(re) else im.shape
            rdt = D.np_dtype(re if D.is_device_tensor(re) else im)
            if not D.is_device_tensor(re):
                re = Kn.full(int(np.prod(shape)), re, rdt).reshape(shape)
            if not D.is_device_tensor(im):
                im = Kn.full(int(np.prod(shape)), im, rdt).reshape(shape)
            return _interleave(re.contiguous(), im.contiguous())
        rdt = np.result_type(re, im)
        return complex_dtype(rdt).type(complex(re, im))
    was_gcxs = isinstance(re, GCXS) or isinstance(im, GCXS)
    ca = next((p.compressed_axes for p in (re, im) if isinstance(p, GCXS)), None)
    template = re if isinstance(re, SparseArray) else im
    shape = template.shape
    rdt = template.dtype

    def as_coo(p):
        if isinstance(p, SparseArray):
            return p.asformat("coo").astype(rdt, copy=False) if p.dtype != rdt else p.asformat("coo")
        # scalar / dense plane next to a sparse one: a constant plane is a fill value
        if np.ndim(p) == 0:
            return COO(np.zeros((len(shape), 0), dtype=np.intp), np.empty(0, dtype=rdt), shape=shape,
                       has_duplicates=False, sorted=True, fill_value=rdt.type(p))
        raise TypeError("sparse_b200: cannot combine a sparse plane with a dense plane")

    re, im = as_coo(re), as_coo(im)
    cdt = np.dtype(cdtype or complex_dtype(rdt))
    cfill = cdt.type(complex(re.fill_value, im.fill_value))
    if re.nnz == 0 and im.nnz == 0:
        out = COO(np.zeros((len(shape), 0), dtype=np.intp), np.empty(0, dtype=cdt), shape=shape, has_duplicates=False,
                  sorted=True, fill_value=cfill)
        return GCXS.from_coo(out, ca) if was_gcxs else out
    kr, ki = re.sorted_keys(), im.sorted_keys()
    dr, di = re._data_dev(), im._data_dev()
    if kr is ki:
        keys, vr, vi = kr, dr, di
    else:
        # align both planes on the union: OR of the raw bit patterns with +0 keeps every value bit for bit (-0.0, NaN
        # payloads; op 17 of the merge kernel), and the impossible fill value switches the prune test off
        weird = _weird_nan(rdt)
        bitor = 17
        zr, zi = Kn.full(int(kr.shape[0]), 0, rdt), Kn.full(int(ki.shape[0]), 0, rdt)
        _, vr, keys = Kn.ew_merge_fused(bitor, kr, dr, 1, ki, zi, 1, re.fill_value, rdt.type(0), weird, rdt, shape,
                                        want_coords=False)
        _, vi, keys2 = Kn.ew_merge_fused(bitor, kr, zr, 1, ki, di, 1, rdt.type(0), im.fill_value, weird, rdt, shape,
                                         want_coords=False)
        assert int(keys.shape[0]) == int(keys2.shape[0])
    cdata = _interleave(vr, vi)
    if cdt != complex_dtype(rdt):
        raise TypeError(f"sparse_b200: plane dtype {rdt} does not match {cdt}")
    flags = Kn.flag_not_fill(cdata, cfill)
    pos, total = Kn.scan_flags(flags)
    if total != int(keys.shape[0]):
        keys, cdata = Kn.compact(keys, flags, pos, total), Kn.compact(cdata, flags, pos, total)
    out = COO._from_device(None, cdata, shape, cfill, keys=keys)
    return GCXS.from_coo(out, ca) if was_gcxs and out.ndim else out


def segment_sum(data, heads, pos, total):
    """Sum of every run of equal keys (COO._sum_duplicates) for any value dtype: complex runs are summed per plane."""
    if not data.is_complex():
        return Kn.segment_sum(data, heads, pos, total)
    re, im = _deinterleave(data)
    return _interleave(Kn.segment_sum(re, heads, pos, total), Kn.segment_sum(im, heads, pos, total))


# ---------------------------------------------------------------------------------------------------------------
# real <-> complex casts (astype)
# ---------------------------------------------------------------------------------------------------------------
def cast_values(data, src, dst):
    """Device cast of a value array where `src` or `dst` is complex (Kn.cast covers the real matrix)."""
    src, dst = np.dtype(src), np.dtype(dst)
    if src.kind == "c" and dst.kind == "c":
        re, im = _deinterleave(data)
        p = plane_dtype(dst)
        return _interleave(Kn.cast(re, p), Kn.cast(im, p))
    if dst.kind == "c":
        p = plane_dtype(dst)
        re = Kn.cast(data, p)
        return _interleave(re, Kn.full(int(re.shape[0]), 0, p))
    re, _ = _deinterleave(data)  # complex -> real: the imaginary part is discarded (NumPy's ComplexWarning case)
    return Kn.cast(re, dst)


# ---------------------------------------------------------------------------------------------------------------
# element-wise
# ---------------------------------------------------------------------------------------------------------------
def _add(x, y):
    """x + y on planes; None stands for the exactly-zero imaginary plane of a real operand."""
    if x is None:
        return y
    if y is None:
        return x
    return np.add(x, y)


def _sub(x, y):
    if y is None:
        return x
    if x is None:
        return np.negative(y)
    return np.subtract(x, y)


def _mul(x, y):
    if x is None or y is None:
        return None
    return np.multiply(x, y)


def _zero_like_plane(p):
    return np.multiply(p, p.dtype.type(0)) if isinstance(p, SparseArray) else (p * 0)


def elemwise_complex(func, args):
    """`func` over operands of which at least one is complex; returns NotImplemented for ufuncs outside the set."""
    from ._elemwise import _get_nary_broadcast_shape, broadcast_to

    def stand_in(a):
        if isinstance(a, SparseArray):
            return np.empty(0, dtype=a.dtype)
        if D.is_device_tensor(a):
            return np.empty(0, dtype=D.np_dtype(a))
        if isinstance(a, np.ndarray) and a.ndim > 0:
            return np.empty(0, dtype=a.dtype)
        return a  # scalar: NEP-50 weak typing

    with np.errstate(all="ignore"):
        probe = func if func not in (np.equal, np.not_equal, np.absolute, np.isnan, np.isinf, np.isfinite) else np.add
        res_dt = probe(*[stand_in(a) for a in args] * (2 if len(args) == 1 and probe is np.add else 1)).dtype
    rdt = plane_dtype(res_dt) if res_dt.kind == "c" else res_dt
    pl = [planes(a, rdt) for a in args]
    with np.errstate(all="ignore"):
        if len(args) == 1:
            (re, im), = pl
            if func in (np.positive,):
                return combine(np.positive(re), np.positive(im))
            if func is np.negative:
                return combine(np.negative(re), np.negative(im))
            if func is np.conjugate:
                return combine(np.positive(re), np.negative(im))
            if func is np.absolute:
                return np.sqrt(np.add(np.multiply(re, re), np.multiply(im, im)))
            if func is np.square:
                return combine(_sub(np.multiply(re, re), np.multiply(im, im)), np.multiply(np.multiply(re, im), rdt.type(2)))
            if func is np.isnan:
                return np.logical_or(np.isnan(re), np.isnan(im))
            if func is np.isinf:
                return np.logical_or(np.isinf(re), np.isinf(im))
            if func is np.isfinite:
                return np.logical_and(np.isfinite(re), np.isfinite(im))
            return NotImplemented
        if len(args) != 2:
            return NotImplemented
        (ar, ai), (br, bi) = pl
        if func is np.add:
            re, im = _add(ar, br), _add(ai, bi)
        elif func is np.subtract:
            re, im = _sub(ar, br), _sub(ai, bi)
        elif func is np.multiply:
            re = _sub(_mul(ar, br), _mul(ai, bi))
            im = _add(_mul(ar, bi), _mul(ai, br))
        elif func is np.true_divide:
            if bi is None:
                re, im = np.true_divide(ar, br), (np.true_divide(ai, br) if ai is not None else None)
            else:
                den = np.add(np.multiply(br, br), np.multiply(bi, bi))
                re = np.true_divide(_add(_mul(ar, br), _mul(ai, bi)), den)
                im = np.true_divide(_sub(_mul(ai, br), _mul(ar, bi)), den)
        elif func in (np.equal, np.not_equal):
            eq_re = np.equal(ar, br)
            a_i = ai if ai is not None else rdt.type(0)
            b_i = bi if bi is not None else rdt.type(0)
            if ai is None and bi is None:
                eq = eq_re
            elif ai is None:
                eq = np.logical_and(eq_re, np.equal(b_i, a_i))
            else:
                eq = np.logical_and(eq_re, np.equal(a_i, b_i))
            return eq if func is np.equal else np.logical_not(eq)
        else:
            return NotImplemented
    if im is None:  # cannot happen for a complex result, kept for symmetry
        im = _zero_like_plane(re)
    if re is None:
        re = _zero_like_plane(im)
    # planes of broadcast operands may come back narrower than the result (a scalar plane): broadcast explicitly
    shape = _get_nary_broadcast_shape(*[tuple(a.shape) if hasattr(a, "shape") else () for a in args])
    fix = lambda p: broadcast_to(p.asformat("coo"), shape) if isinstance(p, SparseArray) and tuple(p.shape) != shape else p
    return combine(fix(re), fix(im), res_dt)


# ---------------------------------------------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------------------------------------------
def reduce_complex(x, method, axis, keepdims, **kwargs):
    """ufunc.reduce over a complex array: only `add` distributes over the planes."""
    if method is not np.add:
        raise TypeError(f"sparse_b200: reduction {getattr(method, '__name__', method)!r} of a complex array is outside "
                        "the CUDA op set (add only)")
    dtype = kwargs.pop("dtype", None)
    rdt = plane_dtype(dtype) if dtype is not None and np.dtype(dtype).kind == "c" else None
    re, im = planes(x, rdt)
    sr = re.reduce(np.add, axis=axis, keepdims=keepdims, **kwargs)
    si = im.reduce(np.add, axis=axis, keepdims=keepdims, **kwargs)
    return combine(sr, si)


# ---------------------------------------------------------------------------------------------------------------
# 2-D products
# ---------------------------------------------------------------------------------------------------------------
def _spgemm_complex_gcxs(a, b, dtr, rdt, return_type):
    """GCXS @ GCXS with complex operands, in the reference's layout (`_dot_csr_csr`, _common.py:639-717: columns of a
    row in reverse first-touch order, then `prune=True`).  The structure of a Gustavson product depends on the index
    arrays only, and the planes of an operand share them: the (up to) four real products are run WITHOUT pruning, so
    their outputs are aligned entry for entry; values are combined on the aligned arrays and only then are the entries
    equal to the complex zero dropped -- exactly where upstream drops them."""
    from ._dot import _csr_arrays, _wrap_gcxs
    from ._elemwise import dense_binary
    from ._gcxs import GCXS

    out_shape = (a.shape[0], b.shape[1])
    a = a.asformat("gcxs")
    b = b.asformat("gcxs", compressed_axes=a.compressed_axes)
    if a.nbytes > b.nbytes:
        b = b.change_compressed_axes(a.compressed_axes)
    else:
        a = a.change_compressed_axes(b.compressed_axes)
    if a.compressed_axes == (0,):
        A, B, ca = a, b, (0,)
    else:  # csc @ csc: a @ b = (b.T @ a.T).T on the same arrays (_common.py:362-373)
        A, B, ca = b._2d_transpose(), a._2d_transpose(), (1,)
    (Ar, Ai), (Br, Bi) = planes(A, rdt), planes(B, rdt)
    M, K = A._compressed_shape
    n_col = B._compressed_shape[1]

    def product(x, y):
        if x is None or y is None:
            return None
        xd, xi, xp = _csr_arrays(x, rdt)
        yd, yi, yp = _csr_arrays(y, rdt)
        if xi.dtype != yi.dtype:
            xi, xp, yi, yp = [Kn.cast(t, np.int64) for t in (xi, xp, yi, yp)]
        indptr, indices, _, data, _ = Kn.spgemm(xp, xi, xd, yp, yi, yd, M, K, n_col, sorted_order=False, prune=False)
        return data, indices, indptr

    rr, ii, ri, ir = product(Ar, Br), product(Ai, Bi), product(Ar, Bi), product(Ai, Br)
    _, indices, indptr = rr
    re = rr[0] if ii is None else dense_binary(np.subtract, rr[0], ii[0])
    im = ri[0] if ir is None else (ir[0] if ri is None else dense_binary(np.add, ri[0], ir[0]))
    if im is None:
        im = Kn.full(int(re.shape[0]), 0, rdt)
    g = GCXS._from_device(_interleave(re, im), indices, indptr, out_shape, ca)
    g._prune()
    if return_type == np.ndarray:
        return g.todense()
    return g.tocoo() if return_type == COO else g


def dot_complex(dot2d, a, b, return_type):
    """2-D product with at least one complex operand, on top of the real `_dot` dispatch (`dot2d`)."""
    from ._dot import _dense_dtype, _dot_dtype, _is_dense
    from ._gcxs import GCXS

    adt = _dense_dtype(a) if _is_dense(a) else a.dtype
    bdt = _dense_dtype(b) if _is_dense(b) else b.dtype
    dtr = _dot_dtype(adt, bdt)
    rdt = plane_dtype(dtr)
    ar, ai = planes(a, rdt)
    br, bi = planes(b, rdt)
    if isinstance(a, SparseArray) and isinstance(b, SparseArray) and (isinstance(a, GCXS) or isinstance(b, GCXS)):
        return _spgemm_complex_gcxs(a, b, dtr, rdt, return_type)
    dense_in = _is_dense(a) or _is_dense(b)
    dense_out = (dense_in and return_type is None) or return_type == np.ndarray
    # real products; dense operands are device tensors here, so dense results stay on the device
    rt = return_type
    prod = lambda x, y: None if x is None or y is None else dot2d(x, y, rt)
    rr, ii, ri, ir = prod(ar, br), prod(ai, bi), prod(ar, bi), prod(ai, br)
    if dense_out:
        def dev(p):
            if p is None:
                return None
            if isinstance(p, SparseArray):
                return p.todense_device() if hasattr(p, "todense_device") else D.upload(p.todense())
            return p if D.is_device_tensor(p) else D.upload(np.ascontiguousarray(p))

        rr, ii, ri, ir = dev(rr), dev(ii), dev(ri), dev(ir)
        from ._elemwise import dense_binary

        re = rr if ii is None else dense_binary(np.subtract, rr, ii)
        im = ri if ir is None else (ir if ri is None else dense_binary(np.add, ri, ir))
        if im is None:
            im = Kn.full(int(re.numel()), 0, rdt).reshape(re.shape)
        out = _interleave(re.contiguous(), im.contiguous())
        keep_dev = any(D.is_device_tensor(x) for x in (a, b))
        return out if keep_dev else D.download(out)
    re = rr if ii is None else np.subtract(rr, ii)
    im = ri if ir is None else (ir if ri is None else np.add(ri, ir))
    if im is None:
        im = np.multiply(re, rdt.type(0))
    return combine(re, im, dtr)
