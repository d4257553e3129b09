"""Complex operands on the dot / element-wise / reduction path (sparse_b200/_complex.py: real and imaginary planes on
the real kernels) and the storage-only value dtypes.  Upstream: tests/test_dot.py:303-335 (`test_complex`,
`test_dot_dense`), tests/test_coo.py:1318-1332 (`test_complex_methods`).  Tolerance parity (1e-6 relative for
complex64, 1e-12 for complex128): the per-plane sums round differently from a complex multiply-accumulate."""
import numpy as np
import pytest

from _api import sp  # noqa: F401


def _tol(dt):
    return dict(rtol=2e-6, atol=1e-6) if np.dtype(dt) == np.complex64 else dict(rtol=1e-12, atol=1e-13)


def _crand(sp, shape, dtype, rng, fmt="coo", density=0.5):
    re = sp.random(shape, density=density, random_state=rng)
    im = sp.random(shape, density=density, random_state=rng)
    x = (re + im * 1j).astype(dtype)
    return x.asformat(fmt) if fmt != "dense" else x.todense()


def _dense(x):
    return x.todense() if hasattr(x, "todense") else x


@pytest.mark.parametrize("dt1", [np.complex64, np.complex128])
@pytest.mark.parametrize("dt2", [np.complex64, np.complex128, np.float64])
@pytest.mark.parametrize("f1", ["coo", "gcxs", "dense"])
@pytest.mark.parametrize("f2", ["coo", "gcxs", "dense"])
@pytest.mark.parametrize("nd", [(2, 2), (2, 1), (1, 2)])
def test_complex_matmul(sp, dt1, dt2, f1, f2, nd):
    rng = np.random.default_rng(11)
    a = _crand(sp, (20,) * nd[0], dt1, rng, f1)
    if np.dtype(dt2).kind == "c":
        b = _crand(sp, (20,) * nd[1], dt2, rng, f2)
    else:
        b = sp.random((20,) * nd[1], density=0.5, random_state=rng).astype(dt2)
        b = b.asformat(f2) if f2 != "dense" else b.todense()
    want = _dense(a) @ _dense(b)
    got = a @ b
    res = _dense(got)
    assert res.dtype == want.dtype and res.shape == want.shape
    assert np.allclose(res, want, **_tol(want.dtype))
    if f1 != "dense" and f2 != "dense" and want.ndim:
        assert isinstance(got, sp.SparseArray)
        assert got.nnz == int(np.count_nonzero(res))  # canonical: no stored zeros


def test_complex_tensordot_nd_and_einsum(sp):
    rng = np.random.default_rng(3)
    t3 = _crand(sp, (3, 4, 5), np.complex128, rng)
    m = _crand(sp, (6, 5), np.complex128, rng)
    d3, dm = t3.todense(), m.todense()
    assert np.allclose(sp.tensordot(t3, m, axes=([2], [1])).todense(), np.tensordot(d3, dm, axes=([2], [1])), rtol=1e-12)
    assert np.allclose(sp.tensordot(t3, dm, axes=([2], [1])), np.tensordot(d3, dm, axes=([2], [1])), rtol=1e-12)
    assert np.allclose(sp.einsum("ijk,lk->ijl", t3, m).todense(), np.einsum("ijk,lk->ijl", d3, dm), rtol=1e-12)
    v = _crand(sp, (5,), np.complex64, rng)
    assert np.allclose(_dense(sp.dot(v, v)), np.dot(v.todense(), v.todense()), rtol=1e-5)
    assert np.allclose(sp.vecdot(m, m, axis=1).todense(), np.sum(dm.conj() * dm, axis=1), rtol=1e-12)


@pytest.mark.parametrize("fmt", ["coo", "gcxs"])
def test_complex_elementwise_and_sum(sp, fmt):
    rng = np.random.default_rng(4)
    a = _crand(sp, (4, 5), np.complex128, rng, fmt)
    b = _crand(sp, (4, 5), np.complex128, rng, fmt)
    y = sp.random((4, 5), density=0.5, random_state=rng).asformat(fmt)
    d, e, f = a.todense(), b.todense(), y.todense()
    for got, want in (
        (a + b, d + e), (a - b, d - e), (a * b, d * e), (a * y, d * f), (y * a, f * d), (a + y, d + f), (a - y, d - f),
        (a * 2.5, d * 2.5), (a * (1 - 2j), d * (1 - 2j)), (a / (2 + 1j), d / (2 + 1j)), (a / 4.0, d / 4.0),
        (-a, -d), (a.conj(), d.conj()), (np.conjugate(a), d.conj()), (abs(a), abs(d)), (np.square(a), d * d),
        (a * np.full((4, 5), 2.0), d * 2.0), (a * np.full((4, 5), 1j), d * 1j),
    ):
        assert isinstance(got, type(a)), type(got)
        res = got.todense()
        assert res.dtype == want.dtype and np.allclose(res, want, rtol=1e-12, atol=1e-14)
        assert got.nnz == int(np.count_nonzero(res))
    assert np.array_equal((a == a).todense(), d == d) and np.array_equal((a != b).todense(), d != e)
    assert np.array_equal(a.real.todense(), d.real) and np.array_equal(a.imag.todense(), d.imag)
    assert a.real.nnz == np.count_nonzero(d.real) and a.imag.dtype == np.float64
    for axis in (None, 0, 1, (0, 1)):
        assert np.allclose(a.sum(axis=axis).todense(), d.sum(axis=axis), rtol=1e-12)
    assert np.allclose(a.mean(axis=0).todense(), d.mean(axis=0), rtol=1e-12)
    with pytest.raises(TypeError):
        np.maximum(a, b)
    with pytest.raises(TypeError):
        a.prod()


def test_complex_methods_upstream_cases(sp):
    for x in (np.array([1, 2, 0, 0, 0]), np.array([1 + 2j, 2 - 1j, 0, 1, 0])):
        s = sp.COO.from_numpy(x)
        for got, want in ((s.imag, x.imag), (s.real, x.real), (s.conj(), x.conj())):
            assert got.dtype == want.dtype and np.array_equal(got.todense(), want)
            assert got.nnz == np.count_nonzero(want)


def test_complex_casts_and_structure(sp):
    rng = np.random.default_rng(6)
    r = sp.random((5, 6), density=0.5, random_state=rng)
    c = r.astype(np.complex64)
    assert c.dtype == np.complex64 and c.nnz == r.nnz and np.array_equal(c.todense(), r.todense().astype(np.complex64))
    c2 = (c * (1 + 1j)).astype(np.complex128)
    d = c2.todense()
    assert c2.dtype == np.complex128
    assert np.array_equal(c2.T.todense(), d.T) and np.array_equal(c2[1:, ::2].todense(), d[1:, ::2])
    assert np.array_equal(c2.reshape((6, 5)).todense(), d.reshape(6, 5))
    g = sp.GCXS(c2, compressed_axes=(1,))
    assert g.dtype == np.complex128 and np.array_equal(g.todense(), d) and np.array_equal(g.tocoo().todense(), d)
    assert np.array_equal(sp.concatenate([c2, c2], axis=1).todense(), np.concatenate([d, d], axis=1))
    assert np.array_equal(sp.COO.from_numpy(d).todense(), d)


@pytest.mark.parametrize("dt", [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64, np.float16])
def test_storage_only_dtypes(sp, dt):
    """Narrow / unsigned values keep their dtype (no silent upcast) and every structural operation moves them.  uint64
    and float16 are storage-only (arithmetic raises; `astype` is the way into the compute matrix)."""
    a = np.array([[1, 0, 3], [0, 0, 2]], dtype=dt)
    x = sp.COO(a)
    assert x.dtype == dt and np.array_equal(x.todense(), a)
    assert np.array_equal(x.T.todense(), a.T) and np.array_equal(x[:, 1:].todense(), a[:, 1:])
    g = sp.GCXS(x)
    assert g.dtype == dt and np.array_equal(g.todense(), a)
    assert np.array_equal(sp.concatenate([x, x]).todense(), np.concatenate([a, a]))
    if np.dtype(dt) in (np.uint64, np.float16):
        with pytest.raises(TypeError):
            x + x
    if np.dtype(dt) != np.float16:
        y = x.astype(np.int64)
        assert np.array_equal((y + y).todense(), a.astype(np.int64) * 2)
        assert np.array_equal(y.astype(dt).todense(), a)


def test_index_dtype_is_a_host_view(sp):
    """tests/test_compressed.py:397-405 upstream (`test_upcast`): uint8 coordinates, uint16 once they overflow."""
    a = sp.random((50, 50, 50), density=0.1, format="coo", idx_dtype=np.uint8, random_state=1)
    assert a.coords.dtype == np.uint8
    b = a.asformat("gcxs")
    assert b.indices.dtype == np.uint16 and b.indptr.dtype == np.uint16
    assert np.array_equal(b.todense(), a.todense())
    c = sp.COO(a.coords.astype(np.int16), a.data, shape=a.shape, has_duplicates=False, sorted=True)
    assert c.coords.dtype == np.int16 and np.array_equal((c + c).todense(), a.todense() * 2)
    with pytest.raises(ValueError):
        sp.COO.from_numpy(np.arange(300), idx_dtype=np.int8)
    with pytest.raises(ValueError):
        sp.GCXS.from_coo(sp.random((25, 25, 25), density=0.01, random_state=2), idx_dtype=np.int8)
