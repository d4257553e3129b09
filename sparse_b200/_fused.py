"""Fused SDDMM and MTTKRP: the reference's two example paths as single device kernels.

* sddmm(s, a, b)   == s * (a @ b)                       (examples/sddmm_example.py:51-52)
* mttkrp(B, D, C)  == sparse.sum(B[:, :, :, None] * D[None, None, :, :] * C[None, :, None, :], axis=(1, 2))
                                                        (examples/mttkrp_example.py:51-52)
The unfused expressions also run (through elemwise + reduce); these entry points avoid the dense (a @ b)
intermediate and the nnz x J broadcast products.  Floating-point association differs from the reference's
BLAS / reduceat order, so parity is tolerance-based (rtol 1e-5 f32, 1e-12 f64 in the tests).
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _kernels as Kn
from ._coo import COO
from ._dot import _coo_as_csr, _dense_dev, _narrow_idx
from ._sparse_array import SparseArray
from ._utils import check_zero_fill_value


def _dense_dt(x):
    return x.dtype if isinstance(x, np.ndarray) else D.np_dtype(x)


def sddmm(s, a, b, *, b_transposed=False):
    """Sampled dense-dense matrix product: ``s * (a @ b)`` for sparse 2-D ``s`` and dense ``a`` (M,K), ``b`` (K,N).
    With ``b_transposed=True`` `b` is given as b^T (N,K), the layout the kernel gathers from."""
    from ._gcxs import GCXS

    if not isinstance(s, SparseArray) or s.ndim != 2:
        raise TypeError("sddmm: `s` must be a 2-D sparse_b200 array")
    check_zero_fill_value(s)
    M, N = s.shape
    if b_transposed:
        bN, bK = int(b.shape[0]), int(b.shape[1])
    else:
        bK, bN = int(b.shape[0]), int(b.shape[1])
    if a.ndim != 2 or b.ndim != 2 or a.shape[0] != M or bN != N or a.shape[1] != bK:
        raise ValueError(f"sddmm: shape mismatch s{s.shape}, a{tuple(a.shape)}, b{tuple(b.shape)}")
    K = int(a.shape[1])
    T = np.result_type(s.dtype, _dense_dt(a), _dense_dt(b))
    if T not in (np.dtype("float32"), np.dtype("float64")):
        raise TypeError(f"sddmm: dtype {T} is outside the CUDA dtype matrix (float32, float64)")
    was_gcxs = isinstance(s, GCXS)
    c = s.tocoo() if was_gcxs else s
    vals, cols, indptr = _coo_as_csr(c, T)
    ad = _dense_dev(a, T)
    bt = _dense_dev(b, T) if b_transposed else Kn.transpose_dense(_dense_dev(b, T))  # (N, K): contiguous gathers
    out = Kn.sddmm(indptr, cols, vals, ad, bt, M, N, K)
    fill = T.type(0)
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and M and N and K:
        # the unfused expression's fill value is `0 * (a @ b)[0, 0]` (_umath.py:520-527: the first element of
        # func(fill, ndarray)), i.e. -0.0 when that corner is negative -- and the prune test below is by bit pattern,
        # so the sign decides WHICH zeros stay stored.  K multiply-adds on the host; device operands keep +0.0
        # (no read-back on the asynchronous path)
        with np.errstate(all="ignore"):
            corner = np.dot(a[0].astype(T, copy=False), (b[0] if b_transposed else b[:, 0]).astype(T, copy=False))
            if np.isfinite(corner):
                fill = T.type(T.type(0) * (corner + T.type(0)))  # a sum that starts at +0.0 is never -0.0
    res = COO._from_device(c._coords, out, s.shape, fill, keys=c.sorted_keys())
    res._canonicalise(check_sort=False, sum_dups=False, prune=True)  # s * dense drops exact zeros (_umath.py:627-633)
    return res.asformat("gcxs", compressed_axes=s.compressed_axes) if was_gcxs else res


def mttkrp(B, Dm, Cm):
    """Matricised tensor times Khatri-Rao product: out[i, j] = sum_{k,l} B[i,k,l] * D[l,j] * C[k,j]."""
    from ._gcxs import GCXS

    if not isinstance(B, SparseArray) or B.ndim != 3:
        raise TypeError("mttkrp: `B` must be a 3-D sparse_b200 array")
    check_zero_fill_value(B)
    I_, K_, L_ = B.shape
    if Dm.ndim != 2 or Cm.ndim != 2 or Dm.shape[0] != L_ or Cm.shape[0] != K_ or Dm.shape[1] != Cm.shape[1]:
        raise ValueError(f"mttkrp: shape mismatch B{B.shape}, D{tuple(Dm.shape)}, C{tuple(Cm.shape)}")
    J = int(Dm.shape[1])
    T = np.result_type(B.dtype, _dense_dt(Dm), _dense_dt(Cm))
    if T not in (np.dtype("float32"), np.dtype("float64")):
        raise TypeError(f"mttkrp: dtype {T} is outside the CUDA dtype matrix (float32, float64)")
    was_gcxs = isinstance(B, GCXS)
    c = B.tocoo() if was_gcxs else B
    coords, data = c._dev()
    c.sorted_keys()
    lead = coords[0].contiguous()
    indptr = Kn.indptr_from_sorted(lead, I_, D.np_dtype(coords))
    kk, ll, indptr = _narrow_idx(coords[1].contiguous(), coords[2].contiguous(), indptr, limit=max(K_, L_, c.nnz))
    out = Kn.mttkrp(indptr, kk, ll, Kn.cast(data, T), _dense_dev(Dm, T), _dense_dev(Cm, T), I_, J)
    rows, cols, vals, _ = Kn.dense_to_csr(out, mode=1, want_rows=True, want_indptr=False)
    t = D.torch()
    res = COO._from_device(t.stack([rows, cols]), vals, (I_, J), T.type(0))
    return GCXS.from_coo(res) if was_gcxs else res
