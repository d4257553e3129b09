"""DOK, the host-side dictionary-of-keys builder (sparse_b200/_dok.py), vs dense NumPy.  Upstream: tests/test_dok.py.
Item assignment happens on the host dictionary; every computation converts to COO and runs on the device path."""
import numpy as np
import pytest
import scipy.sparse as sps

from _api import sp  # noqa: F401


def test_build_by_assignment_matches_numpy(sp):
    rng = np.random.default_rng(0)
    d = np.zeros((4, 5, 3))
    s = sp.DOK((4, 5, 3))
    for key, value in [((1, 2, 0), 5.0), ((0, slice(None), 1), [1, 2, 0, 4, 5]), ((slice(1, 3), 4, slice(None)), 7.0),
                       ((3,), rng.random((5, 3))), ((slice(None, None, 2), 0, slice(None, None, -1)), rng.random((2, 3))),
                       ((Ellipsis, 2), 0.0), ((2, 1), [1.0, 0.0, 3.0]), ((-1, -1, -1), 2.5), ((1, 2, 0), 0.0)]:
        s[key] = value
        d[key] = value
        assert np.array_equal(s.todense(), d), key
    assert s.nnz == np.count_nonzero(d) and s.format == "dok"
    s[[0, 1, 3], [0, 0, 4], [2, 2, 2]] = [9.0, 8.0, 0.0]  # point-wise (all axes indexed by sequences)
    d[[0, 1, 3], [0, 0, 4], [2, 2, 2]] = [9.0, 8.0, 0.0]
    assert np.array_equal(s.todense(), d)
    assert s[1, 2, 0] == d[1, 2, 0] and s[0, 0, 2] == 9.0
    assert np.array_equal(s[[0, 1, 2], [0, 0, 0], [2, 2, 2]].todense(), d[[0, 1, 2], [0, 0, 0], [2, 2, 2]])
    sub = s[1:, ::2]
    assert isinstance(sub, sp.DOK) and np.array_equal(sub.todense(), d[1:, ::2])
    with pytest.raises(IndexError):
        s[4, 0, 0] = 1.0
    with pytest.raises(ValueError):
        s[0, 0, 0] = [1.0, 2.0]
    with pytest.raises(IndexError):
        s[0, 0, [1, 2]] = 1.0


def test_conversions(sp):
    rng = np.random.default_rng(1)
    x = sp.random((6, 7), density=0.4, random_state=rng)
    d = x.todense()
    k = sp.DOK(x)
    assert isinstance(k, sp.DOK) and k.nnz == x.nnz and np.array_equal(k.todense(), d)
    for src in (d, sps.csr_array(d), x.asformat("gcxs")):
        assert np.array_equal(sp.DOK(src).todense(), d)
    assert np.array_equal(k.to_coo().coords, x.coords) and np.array_equal(k.to_coo().data, x.data)
    for fmt, typ in (("coo", sp.COO), ("gcxs", sp.GCXS), ("dok", sp.DOK)):
        assert isinstance(k.asformat(fmt), typ) and np.array_equal(k.asformat(fmt).todense(), d)
        assert isinstance(x.asformat(fmt).asformat("dok"), sp.DOK)
    assert isinstance(sp.random((3, 4), density=0.5, format="dok", random_state=2), sp.DOK)
    assert isinstance(sp.zeros((2, 2), format="dok"), sp.DOK) and sp.asarray(d, format="dok").nnz == x.nnz
    e = sp.DOK((2, 3), {(0, 1): 3, (1, 2): 4}, dtype=np.int64, fill_value=0)
    assert e.dtype == np.int64 and np.array_equal(e.todense(), [[0, 3, 0], [0, 0, 4]])
    assert np.array_equal(k.reshape((7, 6)).todense(), d.reshape(7, 6)) and np.array_equal(k.T.todense(), d.T)
    c = k.copy()
    c[0, 0] = 42.0
    assert k[0, 0] == d[0, 0] and c[0, 0] == 42.0
    with pytest.raises(ValueError):
        sp.DOK((2, 2), data=[1, 2])


def test_computations_run_on_coo_and_come_back_as_dok(sp):
    rng = np.random.default_rng(3)
    a = sp.random((5, 6), density=0.5, format="dok", random_state=rng)
    b = sp.random((5, 6), density=0.5, format="dok", random_state=rng)
    da, db = a.todense(), b.todense()
    for got, want in ((a + b, da + db), (a * b, da * db), (a * 2.0, da * 2), (-a, -da)):
        assert isinstance(got, sp.DOK) and np.array_equal(got.todense(), want)
    # transcendental ufuncs: tolerance parity (DESIGN s4), the device sin is not libm's
    got = np.sin(a)
    assert isinstance(got, sp.DOK) and np.allclose(got.todense(), np.sin(da), rtol=1e-12, atol=0)
    assert isinstance(a + b.to_coo(), sp.COO)  # mixed formats fall back to COO like upstream
    assert isinstance(a.sum(axis=0), sp.DOK) and np.allclose(a.sum(axis=0).todense(), da.sum(0))
    assert np.allclose(a.mean(axis=1).todense(), da.mean(1)) and np.allclose(a.max().todense(), da.max())
    assert np.allclose((a @ b.T).todense(), da @ db.T) and np.allclose(a @ db.T, da @ db.T)
    assert np.allclose(sp.tensordot(a, b, axes=([0, 1], [0, 1])).todense(), np.tensordot(da, db, axes=([0, 1], [0, 1])))
    assert np.array_equal(sp.concatenate([a, b]).todense(), np.concatenate([da, db]))
    assert np.array_equal(a.astype(np.float32).todense(), da.astype(np.float32))
