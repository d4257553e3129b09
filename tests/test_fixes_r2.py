"""Round-2 regression tests for the advisor's findings on the in-scope path (elemwise broadcasting against a dense
operand, reductions of uint64, the result cache after in-place updates) and kron's stored-entries-only product.
Dense NumPy is the judge, as in the reference's own tests."""
import numpy as np
import pytest

from _api import sp  # noqa: F401  (fixture: mock kernels on CPU, CUDA kernels under -m gpu)


@pytest.mark.parametrize("sshape,dshape", [((4,), (3, 4)), ((3, 1), (3, 4)), ((1,), (5,)), ((1, 4), (3, 4)),
                                           ((3, 4), (3, 4))])
@pytest.mark.parametrize("func", [np.add, np.subtract, np.true_divide, np.maximum, np.less, np.not_equal])
def test_sparse_operand_broadcast_up_to_a_dense_one_gives_the_dense_result(sp, sshape, dshape, func):
    """_umath.py:463-465: when func(fill, ndarray) is not constant the result is func(a.todense(), b), a dense array --
    also when the sparse operand has to be broadcast to the dense operand's shape."""
    rng = np.random.default_rng(7)
    a = sp.random(sshape, density=0.6, random_state=rng)
    d = rng.random(dshape) + 0.5
    with np.errstate(all="ignore"):
        want = func(a.todense(), d)
        got = func(a, d)
        want_r = func(d, a.todense())
        got_r = func(d, a)
    for g, w in ((got, want), (got_r, want_r)):
        # dense when func(fill, ndarray) varies, sparse when it is constant (0 / d, 0 < d ...): upstream's rule
        constant = len(np.unique(func(np.zeros_like(d), d) if g is got else func(d, np.zeros_like(d)))) == 1
        assert isinstance(g, sp.COO if constant else np.ndarray), type(g)
        g = g.todense() if constant else g
        assert g.shape == w.shape and np.array_equal(g, w, equal_nan=True)


def test_kron_stores_products_of_stored_entries_only(sp):
    """_coo/common.py:67-129: nnz(kron(a, b)) = nnz(a) * nnz(b); negative * fill must not leave -0.0 entries."""
    rng = np.random.default_rng(11)
    a = sp.random((6, 5), nnz=16, random_state=rng) - 0.5
    a = sp.COO.from_numpy(np.where(a.todense() == -0.5, 0.0, a.todense()))
    b = sp.random((4, 3), nnz=8, random_state=rng)
    k = sp.kron(a, b)
    want = np.kron(a.todense(), b.todense())
    assert k.nnz == a.nnz * b.nnz
    assert np.array_equal(k.todense(), want) and not np.signbit(k.todense()[want == 0]).any()
    assert np.array_equal(sp.kron(a, b.todense()).todense(), want)
    v = sp.kron(sp.COO.from_numpy(np.array([0.0, -2.0, 3.0])), sp.COO.from_numpy(np.array([[1.0, 0.0], [0.0, 4.0]])))
    assert v.shape == (2, 6) and np.array_equal(v.todense(), np.kron(np.array([0.0, -2.0, 3.0]), [[1.0, 0.0], [0.0, 4.0]]))
    e = sp.kron(sp.COO.from_numpy(np.zeros((2, 2))), b)
    assert e.nnz == 0 and e.shape == (8, 6)


def test_result_cache_is_dropped_by_in_place_updates(sp):
    d = np.arange(12.0).reshape(3, 4)
    x = sp.COO.from_numpy(d)
    x.enable_caching()
    assert np.array_equal(x.T.todense(), d.T) and np.array_equal(x.reshape((4, 3)).todense(), d.reshape(4, 3))
    x += x
    assert np.array_equal(x.T.todense(), (2 * d).T)
    assert np.array_equal(x.reshape((4, 3)).todense(), (2 * d).reshape(4, 3))
    assert np.array_equal(x.tocsr().todense(), 2 * d)
    y = sp.COO.from_numpy(d)
    y.enable_caching()
    _ = y.T
    np.multiply(y, 3.0, out=y)
    assert np.array_equal(y.T.todense(), (3 * d).T)


def test_uint64_max_min_order_values_above_2_63(sp):
    d = np.array([[0, 2**63 + 5, 3, 0, 7], [2**64 - 1, 1, 0, 0, 2**63]], dtype=np.uint64)
    x = sp.COO.from_numpy(d)
    assert x.max().todense() == d.max() and x.min().todense() == d.min()
    assert np.array_equal(x.max(axis=1).todense(), d.max(axis=1)) and x.max(axis=1).dtype == np.uint64
    assert np.array_equal(x.min(axis=0).todense(), d.min(axis=0))
    assert np.array_equal(x.max(axis=0).todense(), d.max(axis=0))
    assert np.array_equal(x.sum(axis=1).todense(), d.sum(axis=1))  # modular, like NumPy
    f = sp.COO.from_numpy(d, fill_value=np.uint64(2**63))
    assert np.array_equal(f.max(axis=1).todense(), d.max(axis=1)) and np.array_equal(f.min(axis=1).todense(), d.min(axis=1))
