"""Shared helpers for the parity tests."""
import numpy as np


def rand_csr(rng, M, K, density, dtype=np.float32, idx_dtype=np.int64, empty_rows=True):
    """Random CSR with sorted unique columns per row (uniform positions)."""
    nnz_target = int(round(M * K * density))
    if M * K == 0 or nnz_target == 0:
        return (np.zeros(0, dtype), np.zeros(0, idx_dtype), np.zeros(M + 1, idx_dtype))
    lin = np.unique(rng.integers(0, M * K, size=nnz_target, dtype=np.int64))
    rows, cols = lin // K, lin % K
    indptr = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=indptr[1:])
    if np.issubdtype(np.dtype(dtype), np.integer):
        data = rng.integers(-5, 6, size=len(lin)).astype(dtype)
    else:
        data = (rng.random(len(lin)) * 2 - 1).astype(dtype)
    return data, cols.astype(idx_dtype), indptr.astype(idx_dtype)


def rand_dense(rng, shape, dtype=np.float32, zero_frac=0.0):
    if np.issubdtype(np.dtype(dtype), np.integer):
        b = rng.integers(-5, 6, size=shape).astype(dtype)
    else:
        b = (rng.random(shape) * 2 - 1).astype(dtype)
    if zero_frac:
        b[rng.random(shape) < zero_frac] = 0
    return b


def bits_equal(x, y):
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    return x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x.view(np.uint8), y.view(np.uint8))
