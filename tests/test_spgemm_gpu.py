"""K4 stress: every row bin of the SpGEMM (warp kernels with 128/256/512-slot tables, CTA-per-row kernel with the
global-memory hash) against the oracle of _dot_csr_csr / _dot_coo_coo -- bit-exact values AND column order."""
import numpy as np
import pytest

import oracle
from _util import rand_csr

pytestmark = pytest.mark.gpu


def _run(a, b, shape_a, shape_b, sorted_order=False, prune=False):
    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    (ad, ai, ap), (bd, bi, bp) = a, b
    up = lambda x, dt=None: D.upload(x.astype(dt) if dt is not None else x)
    indptr, indices, rows, data, n_struct = Kn.spgemm(
        up(ap, np.int32), up(ai, np.int32), up(ad), up(bp, np.int32), up(bi, np.int32), up(bd), shape_a[0], shape_a[1],
        shape_b[1], sorted_order=sorted_order, prune=prune, want_indptr=True, want_rows=sorted_order)
    return (D.download(indptr), D.download(indices), D.download(rows) if rows is not None else None,
            D.download(data), n_struct)


def _bits(x):
    return np.ascontiguousarray(x).view(np.uint8)


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64])
@pytest.mark.parametrize("M,K,N,da,db", [
    (300, 400, 500, 0.02, 0.02),      # tiny rows: 128-slot warp kernel
    (200, 600, 3000, 0.03, 0.02),     # P ~ 100-300: 256/512-slot warp kernels
    (120, 2000, 3000, 0.05, 0.05),    # P ~ 15 000: CTA kernel, several 2048-product tiles, global hash
    (40, 700, 64, 0.5, 0.6),          # P >> n_col: heavy collisions, output completely dense -> row flip (:709-714)
    (64, 64, 64, 1.0, 1.0),           # fully dense operands
    (50, 300, 100000, 0.2, 0.001),    # wide output, long sparse B rows
])
def test_ref_order_bit_exact(dtype, M, K, N, da, db):
    rng = np.random.default_rng(hash((M, K, N)) % 2**31)
    a = rand_csr(rng, M, K, da, dtype)
    b = rand_csr(rng, K, N, db, dtype)
    wd, wi, wp = oracle.dot_csr_csr((M, N), a[0], b[0], a[1], b[1], a[2], b[2])
    ip, ix, _, d, n_struct = _run(a, b, (M, K), (K, N))
    assert n_struct == len(wi)
    assert np.array_equal(ip, wp)
    assert np.array_equal(ix, wi), "column order differs from the reference's reverse-first-touch order"
    assert np.array_equal(_bits(d), _bits(wd))


@pytest.mark.parametrize("t0,t1", [(1, 1), (8, 16), (64, 256)])
def test_every_bin_agrees(t0, t1):
    """Lower the binning thresholds so the same rows go through different kernels; results must not change."""
    from sparse_b200 import _kernels as Kn

    rng = np.random.default_rng(5)
    M, K, N = 150, 500, 800
    a = rand_csr(rng, M, K, 0.04, np.float64)
    b = rand_csr(rng, K, N, 0.03, np.float64)
    wd, wi, wp = oracle.dot_csr_csr((M, N), a[0], b[0], a[1], b[1], a[2], b[2])
    try:
        Kn.spgemm_set_thresholds(t0, t1)
        ip, ix, _, d, _ = _run(a, b, (M, K), (K, N))
    finally:
        Kn.spgemm_set_thresholds(64, 256)
    assert np.array_equal(ip, wp) and np.array_equal(ix, wi) and np.array_equal(_bits(d), _bits(wd))


@pytest.mark.parametrize("M,K,N,da,db", [(200, 300, 400, 0.05, 0.05), (60, 1500, 2500, 0.06, 0.05)])
def test_sorted_order_and_prune(M, K, N, da, db):
    """COO x COO mode: ascending columns per row, exact values, +0 sums dropped."""
    rng = np.random.default_rng(8)
    a = rand_csr(rng, M, K, da, np.float64)
    b = rand_csr(rng, K, N, db, np.float64)
    # make some products cancel exactly: duplicate a column of A with negated values into B rows
    wd, wi, wp = oracle.dot_csr_csr((M, N), a[0], b[0], a[1], b[1], a[2], b[2])
    rows = np.repeat(np.arange(M), np.diff(wp))
    order = np.lexsort((wi, rows))
    keep = _bits(wd[order]).reshape(-1, 8).any(axis=1)
    ip, ix, rr, d, _ = _run(a, b, (M, K), (K, N), sorted_order=True, prune=True)
    assert np.array_equal(rr, rows[order][keep])
    assert np.array_equal(ix, wi[order][keep])
    assert np.array_equal(_bits(d), _bits(wd[order][keep]))


def test_duplicate_columns_inside_a_b_row_keep_stored_order():
    """A B row holding the same column twice (non-canonical input): adds must stay in stored order."""
    a = (np.array([1.0, 2.0]), np.array([0, 1]), np.array([0, 2]))
    bd = np.array([1e16, 1.0, -1e16, 3.0, 5.0])
    b = (bd, np.array([2, 2, 2, 0, 2]), np.array([0, 3, 5]))
    wd, wi, wp = oracle.dot_csr_csr((1, 4), a[0], b[0], a[1], b[1], a[2], b[2])
    ip, ix, _, d, _ = _run(a, b, (1, 2), (2, 4))
    assert np.array_equal(ix, wi) and np.array_equal(_bits(d), _bits(wd))
