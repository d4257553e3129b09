"""Python-side call wrappers around the C ABI (device tensors in, device tensors out).

One function per ABI entry point; each cites the reference closure it replaces.
"""
from __future__ import annotations

import numpy as np

from . import _device as D
from . import _lib
from ._lib import i32, i64, vp


def _idx_bytes(t) -> int:
    return 8 if t.dtype == D.torch().int64 else 4


def spmm_csr_dense(a_data, a_indices, a_indptr, b, M: int, K: int, N: int, out=None):
    """out[M,N] = CSR(A) @ B -- replaces _dot_csr_ndarray (_common.py:720-755).

    All arguments are CUDA tensors; `b` is (K, N) row-major (row stride may exceed N).
    Bit-identical to the reference loop (stored order, unfused multiply/add).
    """
    lib = _lib.load()
    t = D.torch()
    dt = D.np_dtype(a_data)
    assert b.dtype == a_data.dtype, "promote operands to _dot_dtype first"
    assert a_indices.dtype == a_indptr.dtype
    if b.dim() != 2 or b.stride(1) != 1:
        b = b.contiguous()
    if out is None:
        out = t.empty((M, N), dtype=a_data.dtype, device=a_data.device)
    ldb = b.stride(0) if K > 1 else max(N, 1)
    ldc = out.stride(0) if M > 1 else max(N, 1)
    rc = lib.b2s_spmm_csr_dense(
        i32(D.dtype_code(dt)), i32(_idx_bytes(a_indices)), i64(M), i64(K), i64(N), vp(D.ptr(a_data)),
        vp(D.ptr(a_indices)), vp(D.ptr(a_indptr)), vp(D.ptr(b)), i64(max(ldb, N)), vp(D.ptr(out)), i64(max(ldc, N)),
        vp(D.stream_ptr()),
    )
    _lib.check(rc, "b2s_spmm_csr_dense")
    return out


def spmm_csr_dense_host(a_data: np.ndarray, a_indices: np.ndarray, a_indptr: np.ndarray, b: np.ndarray,
                        out: np.ndarray | None = None) -> np.ndarray:
    """Host-buffer form of K1 (H2D + kernel + D2H inside the call)."""
    _lib.require_device()
    lib = _lib.load()
    dt = a_data.dtype
    assert b.dtype == dt
    M = len(a_indptr) - 1
    K, N = b.shape
    a_indices = np.ascontiguousarray(a_indices, dtype=np.int64)
    a_indptr = np.ascontiguousarray(a_indptr, dtype=np.int64)
    b = np.ascontiguousarray(b)
    if out is None:
        out = np.empty((M, N), dtype=dt)
    rc = lib.b2s_spmm_csr_dense_host(
        i32(D.dtype_code(dt)), i64(M), i64(K), i64(N), i64(len(a_data)), vp(a_data.ctypes.data),
        vp(a_indices.ctypes.data), vp(a_indptr.ctypes.data), vp(b.ctypes.data), vp(out.ctypes.data),
    )
    _lib.check(rc, "b2s_spmm_csr_dense_host")
    return out


def spmm_set_variant(variant: int = 1, unroll: int = 8):
    _lib.check(_lib.load().b2s_spmm_set_variant(i32(variant), i32(unroll)))
