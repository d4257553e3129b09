"""NumPy-protocol details of the element-wise / reduction path found by running the reference's own tests against this
package (tools/run_reference_tests.py): `out=` and the in-place operators (tests/test_elemwise.py:113-205 upstream),
`var` / `std` (tests/test_coo.py:44-93), `round` / `clip` / `astype` (tests/test_coo.py:1160-1330), the 2-D classes
(tests/test_compressed_2d.py) and SciPy input that is not canonical (tests/test_conversion.py:46-60).  Compared with
dense NumPy; mock kernels on CPU, CUDA kernels under `-m gpu`."""
import operator

import numpy as np
import pytest
import scipy.sparse as sps

from _api import sp  # noqa: F401


@pytest.fixture
def xy(sp):
    rng = np.random.default_rng(1)
    x = sp.random((4, 5, 3), density=0.5, random_state=rng)
    y = sp.random((4, 5, 3), density=0.5, random_state=rng)
    return x, y, x.todense(), y.todense()


@pytest.mark.parametrize("op", [operator.iadd, operator.isub, operator.imul])
@pytest.mark.parametrize("fmt", ["coo", "gcxs"])
def test_inplace_operators(sp, xy, op, fmt):
    x, y, d, e = xy
    x, y = x.asformat(fmt), y.asformat(fmt)
    x0, d = x, d.copy()
    x = op(x, y)
    d = op(d, e)
    assert x is x0 and type(x) is type(y)
    assert np.array_equal(x.todense(), d)
    x = op(x, 2.5)
    d = op(d, 2.5)
    assert x is x0 and np.array_equal(x.todense(), d)


def test_out_argument(sp, xy):
    x, y, d, e = xy
    z = sp.COO.from_numpy(np.zeros(d.shape))
    r = np.multiply(x, y, out=z)
    assert r is z and np.array_equal(z.todense(), d * e)
    r = np.sqrt(x, out=z)
    assert r is z and np.array_equal(z.todense(), np.sqrt(d))
    s = np.add.reduce(x, axis=0, out=sp.COO.from_numpy(np.zeros(d.shape[1:])))
    assert np.allclose(s.todense(), d.sum(0), rtol=1e-13)
    z32 = sp.COO.from_numpy(np.zeros(d.shape, dtype=np.float32))
    np.add(x, y, out=z32)  # same-kind cast into out's dtype, computed in float32 like NumPy's dtype= loop
    assert z32.dtype == np.float32 and np.allclose(z32.todense(), (d + e).astype(np.float32), rtol=1e-6)
    xi = sp.COO.from_numpy(np.arange(6))
    with pytest.raises(TypeError):
        xi += 1.5  # float result into an integer array: NumPy's casting error
    with pytest.raises(ValueError):
        np.add(x, y, out=sp.COO.from_numpy(np.zeros((2, 2))))
    with pytest.raises(TypeError):
        np.add(x, y, out=d.copy())  # a dense `out` for sparse operands: NotImplemented, as upstream


def test_outer_order(sp):
    a = sp.COO.from_numpy(np.array([1.0, 2.0, 0.0]))
    b = sp.COO.from_numpy(np.array([0.0, 3.0, 4.0, 5.0]))
    assert np.array_equal(np.multiply.outer(a, b).todense(), np.multiply.outer(a.todense(), b.todense()))


@pytest.mark.parametrize("axis", [None, 0, (1, 2), -1])
@pytest.mark.parametrize("ddof", [0, 1])
@pytest.mark.parametrize("fmt", ["coo", "gcxs"])
def test_var_std(sp, xy, axis, ddof, fmt):
    x, _, d, _ = xy
    x = x.asformat(fmt)
    for keepdims in (False, True):
        got = x.var(axis=axis, ddof=ddof, keepdims=keepdims)
        assert np.allclose(got.todense(), d.var(axis=axis, ddof=ddof, keepdims=keepdims), rtol=1e-12, atol=1e-15)
        got = x.std(axis=axis, ddof=ddof, keepdims=keepdims)
        assert np.allclose(got.todense(), d.std(axis=axis, ddof=ddof, keepdims=keepdims), rtol=1e-12, atol=1e-15)
    xi = (x * 10).astype(np.int64)
    assert np.allclose(sp.var(xi, axis=axis, correction=ddof).todense(), xi.todense().var(axis=axis, ddof=ddof),
                       rtol=1e-12)
    assert np.allclose(np.std(x, axis=axis).todense(), d.std(axis=axis), rtol=1e-12)


def test_round_clip_conj(sp, xy):
    x, _, d, _ = xy
    y = x * 100 - 20
    e = y.todense()
    for dec in (0, 1, -1, 2):
        assert np.array_equal(y.round(dec).todense(), e.round(dec)), dec
        assert np.array_equal(np.round(y, dec).todense(), e.round(dec)), dec
    yi = y.astype(np.int64)
    for dec in (0, 1, -1):
        assert np.array_equal(yi.round(dec).todense(), yi.todense().round(dec)), dec
    assert np.array_equal(y.clip(0, 30).todense(), e.clip(0, 30))
    assert np.array_equal(x.clip(max=0.5).todense(), d.clip(max=0.5))
    assert np.array_equal(sp.clip(x, min=0.25).todense(), d.clip(min=0.25))
    assert np.array_equal(np.clip(y, -5, 5).todense(), e.clip(-5, 5))
    with pytest.raises(ValueError):
        x.clip()
    assert np.array_equal(x.conj().todense(), d) and np.array_equal(np.conj(x).todense(), d)
    assert x.real is x and x.imag.nnz == 0 and x.imag.dtype == x.dtype
    w = sp.COO.from_numpy(np.array([0.0, np.inf, -np.inf, 2.0, np.nan]))
    assert np.array_equal(sp.isposinf(w).todense(), np.isposinf(w.todense()))
    assert np.array_equal(sp.isneginf(w).todense(), np.isneginf(w.todense()))


@pytest.mark.parametrize("fmt", ["coo", "gcxs"])
def test_astype_prunes_like_elemwise(sp, fmt):
    d = np.array([[0.4, 1.5, 0.0], [-0.2, 0.0, 2.0]])
    x = sp.COO.from_numpy(d).asformat(fmt)
    y = x.astype(np.int64)
    assert y.dtype == np.int64 and y.nnz == 2 and np.array_equal(y.todense(), d.astype(np.int64))
    t = sp.COO.from_numpy(np.array([1e-60, 1.0])).asformat(fmt).astype(np.float32)
    assert t.nnz == 1 and t.dtype == np.float32  # underflow to 0 is pruned too
    assert x.astype(np.float64, copy=False) is x
    with pytest.raises(TypeError):
        x.astype(np.int32, casting="safe")


def test_scipy_input_is_canonicalised(sp):
    """gh-602 upstream: CSR with duplicate and unsorted column indices."""
    data = np.array((2.0, 1.0, 3.0, 3.0, 1.0))
    indices = np.array((1, 0, 0, 1, 1), dtype=int)
    indptr = np.array((0, 2, 5), dtype=int)
    ref = np.array(((1.0, 2.0), (3.0, 4.0)))
    for make in (sps.csr_array, sps.csr_matrix):
        x = make((data, indices, indptr), shape=(2, 2))
        for cls in (sp.GCXS, sp.COO, sp.CSR, sp.CSC):
            g = cls(x) if cls is not sp.COO else sp.COO.from_scipy_sparse(x)
            assert np.array_equal(g.todense(), ref), cls
            assert np.array_equal(g[:1].todense(), ref[:1]) and np.array_equal(g[1:].todense(), ref[1:])
        assert x.nnz == 5  # the caller's matrix is left alone
    z = sp.COO.from_scipy_sparse(sps.random_array((5, 6), density=0.3, random_state=np.random.default_rng(0)),
                                 fill_value=0.0)
    assert z.fill_value == 0.0


def test_csr_csc_classes(sp):
    rng = np.random.default_rng(2)
    s = sp.random((20, 30), density=0.25, random_state=rng)
    d = s.todense()
    for cls, other, ca in ((sp.CSR, sp.CSC, (0,)), (sp.CSC, sp.CSR, (1,))):
        a = cls(s)
        assert type(a) is cls and a.compressed_axes == ca and np.array_equal(a.todense(), d)
        t = a.transpose()
        assert type(t) is other and t.shape == (30, 20) and np.array_equal(t.todense(), d.T)
        assert t.data is a.data and t.indices is a.indices and t.indptr is a.indptr  # O(1), arrays shared
        tc = a.transpose(copy=True)
        assert tc.data is not a.data and np.array_equal(tc.todense(), d.T)
        assert type(t.transpose()) is cls and a.transpose(axes=(0, 1)) is a
        assert np.array_equal((a @ a.T).todense(), d @ d.T) or np.allclose((a @ a.T).todense(), d @ d.T)
        for fmt, typ in (("csr", sp.CSR), ("csc", sp.CSC), ("gcxs", sp.GCXS), ("coo", sp.COO)):
            r = a.asformat(fmt)
            assert isinstance(r, typ) and np.array_equal(r.todense(), d)
        with pytest.raises(ValueError, match="Invalid transpose axes"):
            a.transpose(axes=0)
        with pytest.raises(ValueError, match="compressed axis"):
            cls(s, compressed_axes=(1 - ca[0],))
        for n in (0, 1, 3):
            with pytest.raises(ValueError, match=f"{n}-d"):
                cls(np.ones((5,) * n))
        assert np.array_equal(cls.from_scipy_sparse(sps.coo_array(d)).todense(), d)
        assert np.array_equal(cls.from_numpy(d).todense(), d)
    full = sp.full((10, 20), fill_value=1.0, format="csr")
    assert full.mT.shape == (20, 10) and full.mT.fill_value == 1.0


def test_integer_shifts(sp):
    """tests/test_elemwise.py:534-640 upstream (bitshift binary / scalar / in-place)."""
    rng = np.random.default_rng(0)
    x = (sp.random((4, 5), density=0.5, random_state=rng) * 100).astype(np.int64)
    y = (sp.random((4, 5), density=0.5, random_state=rng) * 10).astype(np.int64)
    d, e = x.todense(), y.todense()
    assert np.array_equal((x << y).todense(), d << e)
    assert np.array_equal((x >> y).todense(), d >> e)
    assert np.array_equal((x << 3).todense(), d << 3) and np.array_equal((x >> 1).todense(), d >> 1)
    big = sp.COO.from_numpy(np.array([1, -8, 5, 0], dtype=np.int64))
    cnt = sp.COO.from_numpy(np.array([70, 65, 63, 2], dtype=np.int64))  # counts >= 64: NumPy gives 0 / -1
    with np.errstate(all="ignore"):
        assert np.array_equal((big << cnt).todense(), big.todense() << cnt.todense())
        assert np.array_equal((big >> cnt).todense(), big.todense() >> cnt.todense())
    x <<= y
    assert np.array_equal(x.todense(), d << e)
    with pytest.raises(TypeError):
        x << 1.5
