"""Parity at BASELINE.json's full configuration sizes (C3, C4, C5) -- GPU only.

Where the oracle (C port / NumPy restatement, pinned to the reference by the golden tests) finishes in seconds the
comparison is exact and complete; otherwise size-independent properties are checked (linearity in a power of two,
idempotence, sampled entries against a float64 restatement)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def _sp():
    import _mock_kernels
    import sparse_b200

    _mock_kernels.uninstall()
    return sparse_b200


def _union_oracle(ka, da, kb, db, f):
    """NumPy restatement of elemwise on canonical key streams (zero fill values): union, f, drop zeros."""
    keys = np.union1d(ka, kb)
    va = np.zeros(len(keys), da.dtype)
    vb = np.zeros(len(keys), db.dtype)
    va[np.searchsorted(keys, ka)] = da
    vb[np.searchsorted(keys, kb)] = db
    r = f(va, vb)
    keep = r.view(np.uint64) != 0
    return keys[keep], r[keep]


def test_c3_full_size_broadcast_add_and_multiply():
    """C3: COO (512,512,512,64) + COO (512,512,512,1) at density 1e-4 (broadcast on axis 3), fp64."""
    sp = _sp()
    rng = np.random.default_rng(0)
    shape_a, shape_b = (512, 512, 512, 64), (512, 512, 512, 1)
    a = sp.random(shape_a, nnz=858_993, random_state=rng)
    b = sp.random(shape_b, nnz=13_421, random_state=rng)
    ka = np.ravel_multi_index(a.coords, shape_a)
    kb3 = np.ravel_multi_index(b.coords[:3], shape_b[:3])
    kb = (kb3[:, None] * 64 + np.arange(64)[None, :]).reshape(-1)
    db = np.repeat(b.data, 64)
    for f in (np.add, np.multiply, np.maximum):
        got = f(a, b)
        wk, wd = _union_oracle(ka, a.data, kb, db, f)
        assert got.shape == shape_a and got.nnz == len(wk)
        assert np.array_equal(np.ravel_multi_index(got.coords, shape_a), wk)
        assert np.array_equal(got.data.view(np.uint64), wd.view(np.uint64))
    # reductions on the same tensor against float64 NumPy
    s3 = a.sum(axis=3)
    dense_keys = np.ravel_multi_index(a.coords[:3], shape_a[:3])
    uk, inv = np.unique(dense_keys, return_inverse=True)
    want = np.bincount(inv, weights=a.data)
    assert np.array_equal(np.ravel_multi_index(s3.coords, shape_a[:3]), uk)
    assert np.allclose(s3.data, want, rtol=1e-12, atol=0)
    m0 = a.max(axis=0)
    k123 = np.ravel_multi_index(a.coords[1:], shape_a[1:])
    uk2 = np.unique(k123)
    want_max = np.zeros(len(uk2))
    np.maximum.at(want_max, np.searchsorted(uk2, k123), a.data)
    assert np.array_equal(np.ravel_multi_index(m0.coords, shape_a[1:]), uk2)
    assert np.array_equal(m0.data, want_max)


def test_c3_full_size_matches_the_reference_digests():
    """C3 against the REFERENCE ITSELF at full size: tests/golden/fullsize_digests.json holds, for the seeded inputs of
    tests/golden/fullsize_inputs.py, nnz and SHA-256 (coordinates + value bit patterns) of the reference's add /
    multiply / maximum results in float64 and float32, and coordinates digest + sampled values of its reductions."""
    import json
    import os
    import sys

    sp = _sp()
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, gdir)
    import fullsize_inputs as FI

    with open(os.path.join(gdir, "fullsize_digests.json")) as f:
        want = json.load(f)["cases"]
    for dt in (np.float64, np.float32):
        (ca, da), (cb, db) = FI.c3_inputs(dt)
        name = np.dtype(dt).name
        a = sp.COO(ca, da, shape=FI.C3_SHAPE_A, has_duplicates=False, sorted=True)
        b = sp.COO(cb, db, shape=FI.C3_SHAPE_B, has_duplicates=False, sorted=True)
        w = want[f"c3_inputs_{name}"]
        assert (a.nnz, b.nnz) == (w["a_nnz"], w["b_nnz"]) and FI.coo_digest(a.coords, a.data) == w["a"]
        for f in (np.add, np.multiply, np.maximum):
            r = f(a, b)
            w = want[f"c3_{f.__name__}_{name}"]
            assert isinstance(r, sp.COO) and r.dtype == dt and r.nnz == w["nnz"], (f.__name__, name, r.nnz)
            assert FI.coo_digest(r.coords, r.data) == w["sha256"], (f.__name__, name)
        if dt == np.float64:
            for axis in ((3,), (0, 1), (0,)):
                for red in ("sum", "max"):
                    r = getattr(a, red)(axis=axis)
                    w = want[f"c3_{red}_axis{''.join(map(str, axis))}_float64"]
                    assert r.nnz == w["nnz"] and FI.digest(np.asarray(r.coords, dtype=np.int64)) == w["coords_sha256"]
                    got = r.data[:: w["sample_step"]]
                    assert np.allclose(got, w["values_sample"], rtol=1e-12, atol=0), (red, axis)
                    assert np.isclose(float(np.sum(r.data)), w["total"], rtol=1e-11)


def test_c2_full_size_sampled_rows_bit_exact():
    """C2 at the configured size (1e6 x 1e6, nnz 1e8, N = 128, fp32): 12 000 rows of the product -- a contiguous block
    and a random sample -- are bit-identical to the oracle's _dot_csr_ndarray on the same arrays."""
    import torch

    import bench
    from sparse_b200 import _kernels as Kn

    _sp()
    dev = torch.device("cuda", 0)
    M = K = 1_000_000
    N = 128
    vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, N, 77, dev)
    C = Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, N)
    ip = indptr.cpu().numpy().astype(np.int64)
    Bh = B.cpu().numpy()
    rng = np.random.default_rng(5)
    rows = np.unique(np.concatenate([np.arange(499_000, 503_000), rng.choice(M, 8000, replace=False)]))
    seg = [np.arange(ip[r], ip[r + 1]) for r in rows]
    sel = np.concatenate(seg)
    sub_ptr = np.zeros(len(rows) + 1, np.int64)
    np.cumsum([len(s) for s in seg], out=sub_ptr[1:])
    sel_t = torch.from_numpy(sel).to(dev)
    want = oracle.dot_csr_ndarray((len(rows), N), vals[sel_t].cpu().numpy(), cols[sel_t].cpu().numpy().astype(np.int64),
                                  sub_ptr, Bh)
    got = C[torch.from_numpy(rows).to(dev)].cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # linearity in a power of two holds exactly over the whole product (size-independent property)
    C2 = Kn.spmm_csr_dense(vals * 2.0, cols, indptr, B, M, K, N)
    assert torch.equal(C2, C * 2.0)


def test_c5_full_size_spgemm_bit_exact():
    """C5: (1e6 x 1e6, density 1e-5)^2 -- indptr/indices/data identical to the oracle of _dot_csr_csr
    (reverse-first-touch column order included)."""
    sp = _sp()
    rng = np.random.default_rng(3)
    n = 1_000_000
    A = sp.random((n, n), nnz=10_000_000, random_state=rng, format="gcxs", compressed_axes=(0,))
    A = A.astype(np.float32)
    got = sp.tensordot(A, A, axes=1)
    d, i, p = oracle.dot_csr_csr((n, n), A.data, A.data, A.indices, A.indices, A.indptr, A.indptr)
    keep = d.view(np.uint32) != 0  # prune=True
    assert isinstance(got, sp.GCXS) and got.compressed_axes == (0,)
    assert got.nnz == int(keep.sum())
    if keep.all():
        assert np.array_equal(got.indptr, p)
    assert np.array_equal(got.indices, i[keep])
    assert np.array_equal(got.data.view(np.uint32), d[keep].view(np.uint32))


def test_c4_sddmm_full_size_properties():
    """C4 at the configured size (1e6 x 1e6 mask, nnz 1e8, K = 256, fp32): exact linearity in a power of two,
    coordinates preserved, sampled entries vs a float64 restatement."""
    sp = _sp()
    import torch

    import bench

    rng = np.random.default_rng(4)
    n, K = 1_000_000, 256
    vals, cols, indptr, _ = bench.make_workload(torch, n, n, 100_000_000, 1, 21, torch.device("cuda", 0))
    s = sp.GCXS((vals, cols, indptr), shape=(n, n), compressed_axes=(0,)).tocoo()
    del vals, cols, indptr
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand((n, K), generator=g, device="cuda", dtype=torch.float32)
    b = torch.rand((K, n), generator=g, device="cuda", dtype=torch.float32)
    r1 = sp.sddmm(s, a, b)
    r2 = sp.sddmm(s * np.float32(2.0), a, b)
    # uniform[0,1) float32 values hit exactly 0.0 a few times in 1e8 draws: those products are +0 and are pruned,
    # like upstream's `s * (a @ b)`
    nz = s.data != 0
    assert r1.nnz == int(nz.sum()) and np.array_equal(r1.coords, s.coords[:, nz])
    assert np.array_equal(r2.data, r1.data * np.float32(2.0))
    sel = rng.choice(r1.nnz, 2000, replace=False)
    ii, jj = r1.coords[0, sel], r1.coords[1, sel]
    ah = a[torch.from_numpy(ii).cuda()].double().cpu().numpy()
    bh = b[:, torch.from_numpy(jj).cuda()].double().cpu().numpy().T
    want = s.data[nz][sel].astype(np.float64) * np.einsum("nk,nk->n", ah, bh)
    assert np.allclose(r1.data[sel], want, rtol=2e-5, atol=0)


def test_c1_coo_tensordot_matches_oracle():
    """C1: COO(1000^2 @ 0.01) . COO -> COO, benchmark seed 42 (benchmarks/conftest.py:6-8 upstream)."""
    sp = _sp()
    rng = np.random.default_rng(42)
    a = sp.random((1000, 1000), density=0.01, random_state=rng)
    b = sp.random((1000, 1000), density=0.01, random_state=rng)
    got = sp.tensordot(a, b, axes=1)
    ip_a = np.searchsorted(a.coords[0], np.arange(1001))
    ip_b = np.searchsorted(b.coords[0], np.arange(1001))
    co, d = oracle.dot_coo_coo((1000, 1000), a.coords, b.coords, a.data, b.data, ip_a, ip_b)
    order = np.lexsort((co[1], co[0]))
    assert np.array_equal(got.coords, co[:, order])
    assert np.array_equal(got.data.view(np.uint64), d[order].view(np.uint64))
    dense = sp.tensordot(a, b.todense(), axes=1)
    assert np.array_equal(dense.view(np.uint64),
                          oracle.dot_coo_ndarray(a.coords, a.data, b.todense().T, (1000, 1000)).view(np.uint64))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
def test_reductions_with_runs_spanning_many_tiles(dtype):
    """Single-pass reduction kernel: groups that span hundreds of 2048-entry tiles (carry-in runs folded through
    several look-back windows), groups that end exactly on tile boundaries, and a full reduction to one value."""
    sp = _sp()
    rng = np.random.default_rng(12)
    shape = (6, 1_500_000)
    nnz = 3_000_000
    x = sp.random(shape, nnz=nnz, random_state=rng)
    if dtype == np.int64:
        x = sp.COO(x.coords, rng.integers(-5, 6, x.nnz), shape=shape, has_duplicates=False, sorted=True)
    else:
        x = x.astype(dtype)
    rows = x.coords[0]
    d = x.data
    tol = dict(rtol=1e-5 if dtype == np.float32 else 1e-11, atol=0)
    s1 = x.sum(axis=1).todense()
    want = np.array([d[rows == r].astype(np.float64 if dtype != np.int64 else np.int64).sum() for r in range(6)])
    assert np.allclose(s1, want, **tol) if dtype != np.int64 else np.array_equal(s1, want)
    m1 = x.max(axis=1).todense()
    want_max = np.array([max(d[rows == r].max(), 0) for r in range(6)], dtype=dtype)
    assert np.array_equal(m1, want_max)
    tot = x.sum()
    ref = d.astype(np.float64).sum() if dtype != np.int64 else d.sum()
    assert np.allclose(tot.todense(), ref, **tol) if dtype != np.int64 else tot.todense() == ref
    # every group exactly one tile long (2048 entries), fully stored: runs end on every tile boundary
    y = sp.COO.from_numpy(np.arange(1, 2048 * 300 + 1, dtype=dtype).reshape(300, 2048))
    want_rows = np.arange(1, 2048 * 300 + 1, dtype=np.int64).reshape(300, 2048).sum(axis=1)
    got_rows = y.sum(axis=1).todense()
    assert np.array_equal(got_rows, want_rows) if dtype == np.int64 else np.allclose(got_rows, want_rows, **tol)
    assert np.array_equal(y.min(axis=1).todense(), np.arange(300, dtype=dtype) * 2048 + 1)


def test_elemwise_union_over_thousands_of_tiles_int_exact():
    """Single-pass merge kernel at ~3300 tiles with integer data (exact), equal keys in both operands on every tile
    boundary pattern, results pruned where a + b == 0."""
    sp = _sp()
    rng = np.random.default_rng(13)
    shape = (3000, 4000)
    a = sp.random(shape, nnz=3_000_000, random_state=rng)
    b = sp.random(shape, nnz=3_000_000, random_state=rng)
    a = sp.COO(a.coords, rng.integers(-3, 4, a.nnz), shape=shape, has_duplicates=False, sorted=True, prune=True)
    b = sp.COO(b.coords, rng.integers(-3, 4, b.nnz), shape=shape, has_duplicates=False, sorted=True, prune=True)
    for f in (np.add, np.multiply, np.maximum, np.not_equal):
        got = f(a, b)
        want = f(a.todense(), b.todense())
        assert np.array_equal(got.todense(), want), f.__name__
        assert got.nnz == np.count_nonzero(want), f.__name__


def test_flag_scan_beyond_2_31_elements():
    """The scan / compaction primitives switch to 64-bit offsets when an array has 2^31 or more elements (round 1
    refused them): positions of the set flags of a 2^31 + 4099 element array, set every 2^20 elements."""
    import torch

    from sparse_b200 import _kernels as Kn

    _sp()
    n = 2**31 + 4099
    flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
    flags[:: 2**20] = 1
    flags[n - 1] = 1
    pos, total = Kn.scan_flags(flags)
    want = (n + 2**20 - 1) // 2**20 + 1
    assert total == want
    assert int(pos[n - 1].item()) == want - 1 and int(pos[2**31].item()) == 2**11 and int(pos[2**31 + 1].item()) == 2**11 + 1
    del flags, pos
    torch.cuda.empty_cache()


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64])
@pytest.mark.parametrize("ncols", [1, 7, 64, 2048])
def test_both_reduction_kernels_agree(dtype, ncols):
    """b2s_reduce_single has two forms: count / scan / emit without inter-tile communication (runs no longer than a
    tile) and the single-pass look-back kernel.  Same group ids, values equal (integers) or to rounding, on streams
    whose runs start and end on every position relative to the 2048-element tiles (incl. completely filled groups of
    exactly one tile)."""
    import torch

    from sparse_b200 import _device as D
    from sparse_b200 import _kernels as Kn

    _sp()
    rng = np.random.default_rng(ncols)
    groups = 40_000 if ncols <= 64 else 600
    keep = rng.random(groups * ncols) < (0.6 if ncols > 1 else 0.8)
    if ncols == 2048:
        keep[: 5 * ncols] = True  # five groups that fill whole tiles
    keys = np.nonzero(keep)[0].astype(np.int64)
    vals = rng.integers(-4, 5, len(keys)).astype(dtype) if dtype == np.int64 else (rng.random(len(keys)) - 0.3).astype(dtype)
    kd, vd = D.upload(keys), D.upload(vals)
    fill = dtype(0.5) if dtype != np.int64 else dtype(2)
    res = {}
    for op in (0, 2, 3):  # add, max, min
        for form in (1, 2):
            Kn.reduce_set_form(form)
            try:
                _, gids, out, neq = Kn.reduce_fused(op, kd, vd, ncols, fill, fill, (groups,), want_coords=False)
            finally:
                Kn.reduce_set_form(0)
            res[form] = (D.download(gids), D.download(out), neq)
        assert np.array_equal(res[1][0], res[2][0]) and np.array_equal(res[1][0], np.unique(keys // ncols))
        if dtype == np.int64 or op != 0:
            assert np.array_equal(res[1][1], res[2][1]), (op, ncols)
        else:
            assert np.allclose(res[1][1], res[2][1], rtol=1e-5 if dtype == np.float32 else 1e-12, atol=1e-6 if dtype == np.float32 else 1e-13)
        assert res[1][2] == res[2][2]
        # against NumPy (fill-value contribution of the reference: op(v, fill) where the group is incomplete)
        gk = keys // ncols
        uk, start = np.unique(gk, return_index=True)
        cnt = np.diff(np.r_[start, len(gk)])
        f = {0: np.add, 2: np.maximum, 3: np.minimum}[op]
        want = f.reduceat(vals.astype(np.float64 if dtype != np.int64 else np.int64), start)
        if op == 0:
            want = want + (ncols - cnt) * (np.float64(fill) if dtype != np.int64 else np.int64(fill))
        else:
            want = np.where(cnt < ncols, f(want, fill), want)
        got = res[2][1].astype(np.float64 if dtype != np.int64 else np.int64)
        assert np.allclose(got, want, rtol=2e-5 if dtype == np.float32 else 1e-11, atol=1e-5 if dtype == np.float32 else 1e-12)
