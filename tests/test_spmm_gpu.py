"""K1 parity: CUDA CSR x dense (through the C ABI) vs the oracle and the golden vectors.
Bit-exact for every dtype (same operation order as _dot_csr_ndarray, _common.py:720-755)."""
import numpy as np
import pytest

import oracle
from _golden import cases
from _util import bits_equal, rand_csr, rand_dense

pytestmark = pytest.mark.gpu


def _run_dev(a_data, a_indices, a_indptr, b, idx32=True):
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    M = len(a_indptr) - 1
    K, N = b.shape
    idt = np.int32 if idx32 else np.int64
    out = Kn.spmm_csr_dense(D.upload(a_data), D.upload(a_indices.astype(idt)), D.upload(a_indptr.astype(idt)),
                            D.upload(b), M, K, N)
    return D.download(out)


@pytest.mark.parametrize("c", cases("dot_kernels", kernel="csr_ndarray"),
                         ids=lambda c: f"{c['dtype']}-{c['a_shape']}x{c['b_shape']}")
def test_golden_reference_vectors(c):
    a = c.arr
    got = _run_dev(a["a_data"], a["a_indices"], a["a_indptr"], np.ascontiguousarray(a["b"]))
    assert bits_equal(got, a["out"])


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("M,K,N,density", [
    (300, 500, 128, 0.05), (257, 300, 64, 0.1), (100, 100, 32, 0.2), (64, 80, 16, 0.3), (50, 70, 8, 0.3),
    (33, 40, 4, 0.5), (40, 50, 1, 0.5), (40, 50, 3, 0.5), (20, 30, 130, 0.4), (1000, 2000, 256, 0.01),
    (17, 19, 129, 0.5), (1, 1, 1, 1.0), (5, 7, 128, 0.0),
])
@pytest.mark.parametrize("idx32", [True, False])
def test_vs_oracle(dtype, M, K, N, density, idx32):
    rng = np.random.default_rng(hash((M, K, N)) % 2**32)
    a_data, a_indices, a_indptr = rand_csr(rng, M, K, density, dtype)
    b = rand_dense(rng, (K, N), dtype)
    want = oracle.dot_csr_ndarray((M, N), a_data, a_indices, a_indptr, b)
    got = _run_dev(a_data, a_indices, a_indptr, b, idx32)
    assert bits_equal(got, want)


@pytest.mark.parametrize("variant,unroll", [(1, 4), (1, 16), (1, 32), (2, 8)])
def test_kernel_variants_are_bit_identical(variant, unroll):
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(5)
    M, K, N = 3000, 4000, 128
    a_data, a_indices, a_indptr = rand_csr(rng, M, K, 0.02, np.float32)
    # a few long and empty rows
    b = rand_dense(rng, (K, N), np.float32)
    want = oracle.dot_csr_ndarray((M, N), a_data, a_indices, a_indptr, b)
    try:
        Kn.spmm_set_variant(variant, unroll)
        got = _run_dev(a_data, a_indices, a_indptr, b)
    finally:
        Kn.spmm_set_variant(1, 8)
    assert bits_equal(got, want)


def test_host_buffer_entry_point():
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(9)
    M, K, N = 2000, 3000, 128
    a_data, a_indices, a_indptr = rand_csr(rng, M, K, 0.01, np.float32)
    b = rand_dense(rng, (K, N), np.float32)
    got = Kn.spmm_csr_dense_host(a_data, a_indices, a_indptr, b)
    assert bits_equal(got, oracle.dot_csr_ndarray((M, N), a_data, a_indices, a_indptr, b))


def test_skewed_rows_and_strided_b():
    """Power-law row lengths; B given with a row stride larger than N."""
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(11)
    M, K, N = 500, 4000, 128
    lens = np.minimum((rng.pareto(1.2, M) * 20).astype(np.int64), K)
    lens[::7] = 0
    indptr = np.zeros(M + 1, np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in lens]).astype(np.int64)
    data = (rng.random(len(indices)) - 0.5).astype(np.float32)
    bfull = rand_dense(rng, (K, N + 32), np.float32)
    want = oracle.dot_csr_ndarray((M, N), data, indices, indptr, np.ascontiguousarray(bfull[:, :N]))
    bt = D.upload(bfull)[:, :N]
    out = Kn.spmm_csr_dense(D.upload(data), D.upload(indices.astype(np.int32)), D.upload(indptr.astype(np.int32)), bt,
                            M, K, N)
    assert bits_equal(D.download(out), want)


def test_large_linearity_and_row_checksum():
    """Full-width property check at a size the oracle cannot cover quickly:
    (A @ [B1 | B2]) splits exactly, and scaling A's values by 2 scales C exactly (power of two)."""
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    t = D.torch()
    g = t.Generator(device="cuda").manual_seed(1)
    M = K = 200_000
    nnz = 4_000_000
    lin = t.unique(t.randint(0, M * K, (nnz,), generator=g, device="cuda", dtype=t.int64))
    rows, cols = lin // K, (lin % K).to(t.int32)
    indptr = t.zeros(M + 1, dtype=t.int64, device="cuda")
    indptr[1:] = t.cumsum(t.bincount(rows, minlength=M), 0)
    indptr = indptr.to(t.int32)
    vals = t.rand(len(lin), generator=g, device="cuda", dtype=t.float32)
    B = t.rand((K, 128), generator=g, device="cuda", dtype=t.float32)
    C = Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128)
    C2 = Kn.spmm_csr_dense(vals * 2, cols, indptr, B, M, K, 128)
    assert t.equal(C2, C * 2)
    Cl = Kn.spmm_csr_dense(vals, cols, indptr, B[:, :64], M, K, 64)
    assert t.equal(Cl, C[:, :64])
    # sampled rows against the oracle
    sel = np.arange(0, M, 9973)
    ip = D.download(indptr).astype(np.int64)
    ci, vd, Bh = D.download(cols).astype(np.int64), D.download(vals), D.download(B)
    sub_ptr = np.zeros(len(sel) + 1, np.int64)
    np.cumsum(ip[sel + 1] - ip[sel], out=sub_ptr[1:])
    take = np.concatenate([np.arange(ip[r], ip[r + 1]) for r in sel])
    want = oracle.dot_csr_ndarray((len(sel), 128), vd[take], ci[take], sub_ptr, Bh)
    assert bits_equal(D.download(C)[sel], want)


@pytest.mark.parametrize("n_panels", [2, 3, 8, 0])
@pytest.mark.parametrize("dtype,N", [(np.float32, 128), (np.float32, 256), (np.float64, 64), (np.float32, 36)])
def test_column_panel_passes_are_bit_identical(n_panels, dtype, N):
    """K1p: panel passes carry partial sums through C; results must equal the one-pass kernel and the oracle."""
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(77)
    M, K = 3000, 5000
    a_data, a_indices, a_indptr = rand_csr(rng, M, K, 0.02, dtype)
    # long rows spanning every panel and a few empty rows
    b = rand_dense(rng, (K, N), dtype)
    want = oracle.dot_csr_ndarray((M, N), a_data, a_indices, a_indptr, b)
    ad, ai, ap, bd = (D.upload(a_data), D.upload(a_indices.astype(np.int32)), D.upload(a_indptr.astype(np.int32)),
                      D.upload(b))
    assert Kn.csr_rows_sorted(ai, ap, M)
    got = Kn.spmm_csr_dense(ad, ai, ap, bd, M, K, N, n_panels=n_panels, rows_sorted=True)
    assert bits_equal(D.download(got), want)


def test_panel_mode_refuses_unsorted_rows():
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(78)
    M, K, N = 200, 300, 128
    a_data, a_indices, a_indptr = rand_csr(rng, M, K, 0.1, np.float32)
    for r in range(M):  # reverse every row: still a valid CSR, no longer column-sorted
        s, e = a_indptr[r], a_indptr[r + 1]
        a_indices[s:e] = a_indices[s:e][::-1].copy()
        a_data[s:e] = a_data[s:e][::-1].copy()
    b = rand_dense(rng, (K, N), np.float32)
    ad, ai, ap, bd = (D.upload(a_data), D.upload(a_indices.astype(np.int32)), D.upload(a_indptr.astype(np.int32)),
                      D.upload(b))
    assert not Kn.csr_rows_sorted(ai, ap, M)
    got = Kn.spmm_csr_dense(ad, ai, ap, bd, M, K, N, n_panels=4)  # falls back to the one-pass kernel
    assert bits_equal(D.download(got), oracle.dot_csr_ndarray((M, N), a_data, a_indices, a_indptr, b))


@pytest.mark.parametrize("dtype,N", [(np.float32, 128), (np.float64, 64), (np.float32, 256), (np.int64, 128),
                                     (np.float32, 132), (np.float64, 72), (np.float32, 129)])
def test_long_rows_take_the_column_split_kernel(dtype, N):
    """nnz-balanced mode: rows longer than max(512, 4 x mean) are computed by the column-split kernel on a side stream
    (shared-memory ring of cp.async copies when B's rows are 16-byte aligned -- incl. a zero-filled tail panel --, the
    register-staged kernel otherwise: N = 129); the row-split kernel skips them.  Bit-identical to the oracle."""
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(91)
    M, K = 5000, 30000
    lens = rng.integers(0, 40, M)
    long_rows = [3, 77, 1234, 4999]
    for r, n in zip(long_rows, (4097, 9000, 25000, 30000)):
        lens[r] = n
    indptr = np.zeros(M + 1, np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in lens]).astype(np.int64)
    if np.issubdtype(np.dtype(dtype), np.integer):
        data = rng.integers(-3, 4, len(indices)).astype(dtype)
    else:
        data = (rng.random(len(indices)) - 0.5).astype(dtype)
    b = rand_dense(rng, (K, N), dtype)
    want = oracle.dot_csr_ndarray((M, N), data, indices, indptr, b)
    ipd = D.upload(indptr.astype(np.int32))
    assert Kn.csr_max_row_nnz(ipd, M) == 30000
    got = Kn.spmm_csr_dense(D.upload(data), D.upload(indices.astype(np.int32)), ipd, D.upload(b), M, K, N,
                            long_rows=True)
    assert bits_equal(D.download(got), want)
