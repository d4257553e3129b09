"""CPU oracle for the pydata/sparse hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the reference algorithms that
``sparse_b200`` implements in CUDA.  It exists to *check* the CUDA path
(tests/, ``__graft_entry__.smoke()``) and to give ``bench.py`` its
``cpu_baseline`` / ``--impl reference`` leg.  The product package never
imports it; the product has no CPU fallback.

* ``dot_oracle.c`` (compiled to ``liboracle.so``): the numba kernels of
  ``sparse/numba_backend/_common.py:543-1158`` plus ``_match_arrays``
  (``_umath.py:53-92``) and ``_calc_counts_invidx`` (``_coo/core.py:1601-1628``).
* the NumPy-level glue around those kernels (``_umath.py``, ``_sparse_array.py:372-437``,
  ``_coo/core.py``, ``_compressed/compressed.py``) is not restated here: for it the
  checker is the reference itself, through the golden input / output fixtures of
  ``tests/golden/`` (2500+ cases written by ``tests/golden/make_golden.py``).

Parity status: PINNED against golden vectors generated from the reference
itself (``tests/golden/make_golden.py``), checked in ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

I64P = ctypes.POINTER(ctypes.c_int64)
_CT = {
    "f32": ctypes.c_float,
    "f64": ctypes.c_double,
    "i32": ctypes.c_int32,
    "i64": ctypes.c_int64,
}
_SUF = {np.dtype("float32"): "f32", np.dtype("float64"): "f64", np.dtype("int32"): "i32", np.dtype("int64"): "i64"}


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc, -ffp-contract=off).  Falls back to no OpenMP."""
    src = os.path.join(_HERE, "dot_oracle.c")
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src):
        return _LIB_PATH
    base = ["-O3", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wno-unknown-pragmas", "-shared"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            cmd = [cc, *base, *omp, "-o", _LIB_PATH, src]
            try:
                subprocess.run(cmd, check=True, capture_output=True, text=True)
                return _LIB_PATH
            except (OSError, subprocess.CalledProcessError) as e:  # try next
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_max_threads.restype = ctypes.c_int
        for name in ("orc_csr_csr_count", "orc_match_arrays", "orc_counts_invidx"):
            getattr(_lib, name).restype = ctypes.c_int64
        for suf in _CT:
            for name in (
                "orc_csr_dense_sparse_count",
                "orc_csc_dense_sparse_count",
                "orc_csc_dense_sparse_fill",
                "orc_coo_dense_sparse",
                "orc_dense_coo_sparse",
            ):
                getattr(_lib, f"{name}_{suf}").restype = ctypes.c_int64
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> int:
    """Fix the OpenMP thread count of the row-parallel kernels; returns the count now in effect."""
    lib().orc_set_threads(ctypes.c_int(int(n)))
    return max_threads()


def _i64(x):
    return np.ascontiguousarray(x, dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _L(v):
    return ctypes.c_int64(int(v))


def dot_dtype(dt1, dt2):
    """_dot_dtype, _common.py:635-636."""
    return (np.zeros((), dtype=dt1) * np.zeros((), dtype=dt2)).dtype


def _suffix(dt):
    dt = np.dtype(dt)
    if dt not in _SUF:
        raise TypeError(f"oracle: unsupported dtype {dt}")
    return _SUF[dt]


# --------------------------------------------------------------------------
# Kernel-seam functions: same argument meaning as the reference factories
# called from _dot (_common.py:357-498).
# --------------------------------------------------------------------------
def dot_csr_ndarray(out_shape, a_data, a_indices, a_indptr, b):
    """_dot_csr_ndarray (_common.py:720-755)."""
    dtr = dot_dtype(a_data.dtype, b.dtype)
    suf = _suffix(dtr)
    M, N = map(int, out_shape)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b = np.ascontiguousarray(b, dtype=dtr)
    ai, ap = _i64(a_indices), _i64(a_indptr)
    out = np.empty((M, N), dtype=dtr)
    getattr(lib(), f"orc_csr_dense_{suf}")(_L(M), _L(N), _p(a_data), _p(ai), _p(ap), _p(b), _p(out))
    return out


def dot_csc_ndarray(a_shape, b_shape, a_data, a_indices, a_indptr, b):
    """_dot_csc_ndarray (_common.py:869-904)."""
    dtr = dot_dtype(a_data.dtype, b.dtype)
    suf = _suffix(dtr)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b = np.ascontiguousarray(b, dtype=dtr)
    ai, ap = _i64(a_indices), _i64(a_indptr)
    out = np.empty((int(a_shape[0]), int(b_shape[1])), dtype=dtr)
    getattr(lib(), f"orc_csc_dense_{suf}")(
        _L(a_shape[0]), _L(a_shape[1]), _L(b_shape[1]), _p(a_data), _p(ai), _p(ap), _p(b), _p(out)
    )
    return out


def csr_csr_count_nnz(out_shape, a_indices, b_indices, a_indptr, b_indptr):
    """_csr_csr_count_nnz (_common.py:543-570)."""
    ai, bi, ap, bp = _i64(a_indices), _i64(b_indices), _i64(a_indptr), _i64(b_indptr)
    return int(lib().orc_csr_csr_count(_L(out_shape[0]), _L(out_shape[1]), _p(ai), _p(bi), _p(ap), _p(bp)))


def dot_csr_csr(out_shape, a_data, b_data, a_indices, b_indices, a_indptr, b_indptr):
    """_dot_csr_csr (_common.py:639-717) -> (data, indices, indptr)."""
    dtr = dot_dtype(a_data.dtype, b_data.dtype)
    suf = _suffix(dtr)
    n_row, n_col = map(int, out_shape)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b_data = np.ascontiguousarray(b_data, dtype=dtr)
    ai, bi, ap, bp = _i64(a_indices), _i64(b_indices), _i64(a_indptr), _i64(b_indptr)
    nnz = csr_csr_count_nnz(out_shape, ai, bi, ap, bp)
    data = np.empty(nnz, dtype=dtr)
    indices = np.empty(nnz, dtype=np.int64)
    indptr = np.empty(n_row + 1, dtype=np.int64)
    getattr(lib(), f"orc_csr_csr_{suf}")(
        _L(n_row), _L(n_col), _p(a_data), _p(b_data), _p(ai), _p(bi), _p(ap), _p(bp),
        _p(data), _p(indices), _p(indptr), ctypes.c_void_p(0), ctypes.c_int(1),
    )
    return data, indices, indptr


def dot_coo_coo(out_shape, a_coords, b_coords, a_data, b_data, a_indptr, b_indptr):
    """_dot_coo_coo (_common.py:907-976) -> (coords[2,nnz], data)."""
    dtr = dot_dtype(a_data.dtype, b_data.dtype)
    suf = _suffix(dtr)
    n_row, n_col = map(int, out_shape)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b_data = np.ascontiguousarray(b_data, dtype=dtr)
    ai, bi, ap, bp = _i64(a_coords[1]), _i64(b_coords[1]), _i64(a_indptr), _i64(b_indptr)
    nnz = csr_csr_count_nnz(out_shape, ai, bi, ap, bp)
    data = np.empty(nnz, dtype=dtr)
    coords = np.empty((2, nnz), dtype=np.int64)
    rows, cols = coords[0], coords[1]
    getattr(lib(), f"orc_csr_csr_{suf}")(
        _L(n_row), _L(n_col), _p(a_data), _p(b_data), _p(ai), _p(bi), _p(ap), _p(bp),
        _p(data), _p(cols), ctypes.c_void_p(0), _p(rows), ctypes.c_int(0),
    )
    return coords, data


def dot_csr_ndarray_sparse(out_shape, a_data, a_indices, a_indptr, b):
    """_dot_csr_ndarray_sparse (_common.py:758-804) -> (data, indices, indptr)."""
    dtr = dot_dtype(a_data.dtype, b.dtype)
    suf = _suffix(dtr)
    M, N = map(int, out_shape)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b = np.ascontiguousarray(b, dtype=dtr)
    ai, ap = _i64(a_indices), _i64(a_indptr)
    indptr = np.empty(M + 1, dtype=np.int64)
    nnz = int(getattr(lib(), f"orc_csr_dense_sparse_count_{suf}")(_L(M), _L(N), _p(ai), _p(ap), _p(b), _p(indptr)))
    data = np.empty(nnz, dtype=dtr)
    indices = np.empty(nnz, dtype=np.int64)
    getattr(lib(), f"orc_csr_dense_sparse_fill_{suf}")(
        _L(M), _L(N), _p(a_data), _p(ai), _p(ap), _p(b), _p(data), _p(indices)
    )
    return data, indices, indptr


def dot_csc_ndarray_sparse(a_shape, b_shape, a_data, a_indices, a_indptr, b):
    """_dot_csc_ndarray_sparse (_common.py:807-866) -> (data, indices, indptr).

    Like the reference, ``data``/``indices`` are sized by the structural count
    while entries whose float64 sum is exactly 0 are skipped, so the tail of
    the arrays beyond the last written entry is uninitialised in the reference
    (np.empty).  We return the written prefix length as a fourth value.
    """
    dtr = dot_dtype(a_data.dtype, b.dtype)
    suf = _suffix(dtr)
    a_data = np.ascontiguousarray(a_data, dtype=dtr)
    b = np.ascontiguousarray(b, dtype=dtr)
    ai, ap = _i64(a_indices), _i64(a_indptr)
    N = int(b_shape[1])
    indptr = np.empty(N + 1, dtype=np.int64)
    indptr[0] = 0
    nnz = int(
        getattr(lib(), f"orc_csc_dense_sparse_count_{suf}")(
            _L(a_shape[0]), _L(a_shape[1]), _L(N), _p(ai), _p(ap), _p(b), _p(indptr)
        )
    )
    data = np.zeros(nnz, dtype=dtr)
    indices = np.zeros(nnz, dtype=np.int64)
    written = int(
        getattr(lib(), f"orc_csc_dense_sparse_fill_{suf}")(
            _L(a_shape[0]), _L(a_shape[1]), _L(N), _p(a_data), _p(ai), _p(ap), _p(b), _p(data), _p(indices)
        )
    )
    return data, indices, indptr, written


def _strides_elems(arr):
    return arr.strides[0] // arr.itemsize, arr.strides[1] // arr.itemsize


def dot_coo_ndarray(coords1, data1, array2, out_shape):
    """_dot_coo_ndarray dense (_common.py:979-1014); array2 is b.T (any strides)."""
    dtr = dot_dtype(data1.dtype, array2.dtype)
    suf = _suffix(dtr)
    data1 = np.ascontiguousarray(data1, dtype=dtr)
    if array2.dtype != dtr:
        array2 = array2.astype(dtr)
    rows, cols = _i64(coords1[0]), _i64(coords1[1])
    M, N = map(int, out_shape)
    s0, s1 = _strides_elems(array2)
    out = np.empty((M, N), dtype=dtr)
    getattr(lib(), f"orc_coo_dense_{suf}")(
        _L(len(data1)), _p(rows), _p(cols), _p(data1), _p(array2), _L(s0), _L(s1), _L(M), _L(N), _p(out)
    )
    return out


def dot_coo_ndarray_sparse(coords1, data1, array2, out_shape):
    """_dot_coo_ndarray sparse (_common.py:1017-1072) -> (coords[2,n], data)."""
    dtr = dot_dtype(data1.dtype, array2.dtype)
    suf = _suffix(dtr)
    data1 = np.ascontiguousarray(data1, dtype=dtr)
    if array2.dtype != dtr:
        array2 = array2.astype(dtr)
    rows, cols = _i64(coords1[0]), _i64(coords1[1])
    N = int(out_shape[1])
    s0, s1 = _strides_elems(array2)
    f = getattr(lib(), f"orc_coo_dense_sparse_{suf}")
    null = ctypes.c_void_p(0)
    n = int(f(_L(len(data1)), _p(rows), _p(cols), _p(data1), _p(array2), _L(s0), _L(s1), _L(N), null, null, null))
    coords = np.empty((2, n), dtype=np.int64)
    data = np.empty(n, dtype=dtr)
    f(_L(len(data1)), _p(rows), _p(cols), _p(data1), _p(array2), _L(s0), _L(s1), _L(N),
      _p(coords[0]), _p(coords[1]), _p(data))
    return coords, data


def dot_ndarray_coo(array1, coords2, data2, out_shape):
    """_dot_ndarray_coo dense (_common.py:1075-1103)."""
    dtr = dot_dtype(array1.dtype, data2.dtype)
    suf = _suffix(dtr)
    array1 = np.ascontiguousarray(array1, dtype=dtr)
    data2 = np.ascontiguousarray(data2, dtype=dtr)
    rows, cols = _i64(coords2[0]), _i64(coords2[1])
    M, N = map(int, out_shape)
    K = int(array1.shape[1])
    out = np.empty((M, N), dtype=dtr)
    getattr(lib(), f"orc_dense_coo_{suf}")(
        _L(M), _L(K), _L(N), _p(array1), _L(len(data2)), _p(rows), _p(cols), _p(data2), _p(out)
    )
    return out


def dot_ndarray_coo_sparse(array1, coords2, data2, out_shape):
    """_dot_ndarray_coo sparse (_common.py:1106-1158); coords2/data2 belong to b.T."""
    dtr = dot_dtype(array1.dtype, data2.dtype)
    suf = _suffix(dtr)
    array1 = np.ascontiguousarray(array1, dtype=dtr)
    data2 = np.ascontiguousarray(data2, dtype=dtr)
    first, second = _i64(coords2[0]), _i64(coords2[1])
    M = int(out_shape[0])
    K = int(array1.shape[1])
    f = getattr(lib(), f"orc_dense_coo_sparse_{suf}")
    null = ctypes.c_void_p(0)
    n = int(f(_L(M), _L(K), _p(array1), _L(len(data2)), _p(first), _p(second), _p(data2), null, null, null))
    coords = np.empty((2, n), dtype=np.int64)
    data = np.empty(n, dtype=dtr)
    f(_L(M), _L(K), _p(array1), _L(len(data2)), _p(first), _p(second), _p(data2),
      _p(coords[0]), _p(coords[1]), _p(data))
    return coords, data


def match_arrays(a, b):
    """_match_arrays (_umath.py:53-92) -> (a_idx, b_idx) as uintp."""
    a, b = _i64(a), _i64(b)
    null = ctypes.c_void_p(0)
    n = int(lib().orc_match_arrays(_p(a), _L(len(a)), _p(b), _L(len(b)), null, null))
    ia = np.empty(n, dtype=np.int64)
    ib = np.empty(n, dtype=np.int64)
    lib().orc_match_arrays(_p(a), _L(len(a)), _p(b), _L(len(b)), _p(ia), _p(ib))
    return ia.astype(np.uintp), ib.astype(np.uintp)


def calc_counts_invidx(groups):
    """_calc_counts_invidx (_coo/core.py:1601-1628) -> (inv_idx, counts)."""
    g = _i64(groups)
    inv = np.empty(len(g), dtype=np.int64)
    cnt = np.empty(len(g), dtype=np.int64)
    n = int(lib().orc_counts_invidx(_p(g), _L(len(g)), _p(inv), _p(cnt)))
    return inv[:n].copy(), cnt[:n].copy()


def uncompress_dimension(indptr):
    """uncompress_dimension (_compressed/convert.py:81-87)."""
    ip = _i64(indptr)
    rows = np.empty(int(ip[-1]) if len(ip) else 0, dtype=np.int64)
    lib().orc_uncompress(_p(ip), _L(len(ip) - 1), _p(rows))
    return rows
