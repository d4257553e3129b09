"""Empty and degenerate inputs through the public API vs dense NumPy: no stored entries, zero-length axes, 0-D arrays,
all-fill results.  Upstream: tests/test_elemwise.py:143-156, 697-726 (`*_empty`), tests/test_coo.py:853-857, 1575-1580,
tests/test_dot.py:83-97.  On the GPU these are the launches with n == 0 (every ABI entry returns early) and the
grid-size-one cases."""
import numpy as np
import pytest

from _api import sp  # noqa: F401


def _eq(got, want):
    got = got.todense() if hasattr(got, "todense") else np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert np.array_equal(got, want, equal_nan=want.dtype.kind == "f")


@pytest.mark.parametrize("fa", ["coo", "gcxs", "dense"])
@pytest.mark.parametrize("fb", ["coo", "gcxs", "dense"])
def test_products_with_no_stored_entries(sp, fa, fb):
    if fa == fb == "dense":
        pytest.skip("dense @ dense is NumPy's")
    rng = np.random.default_rng(0)
    full = sp.random((6, 5), density=0.6, random_state=rng)
    zero_a, zero_b = sp.zeros((4, 6)), sp.zeros((5, 3))

    def conv(x, f):
        return x.todense() if f == "dense" else x.asformat(f)

    for a, b in ((zero_a, full), (full, zero_b), (zero_a, sp.zeros((6, 2)))):
        _eq(conv(a, fa) @ conv(b, fb), a.todense() @ b.todense())
    v = sp.zeros((6,))
    _eq(conv(zero_a, fa) @ conv(v, fb), np.zeros(4))
    _eq(sp.tensordot(conv(full, fa), conv(zero_b, fb), axes=1), full.todense() @ np.zeros((5, 3)))


@pytest.mark.parametrize("shape_a,shape_b", [((0, 4), (4, 3)), ((3, 0), (0, 5)), ((2, 4), (4, 0)), ((0, 0), (0, 0))])
def test_products_with_zero_length_axes(sp, shape_a, shape_b):
    a = sp.COO.from_numpy(np.zeros(shape_a))
    b = sp.COO.from_numpy(np.zeros(shape_b))
    want = np.zeros(shape_a) @ np.zeros(shape_b)
    _eq(a @ b, want)
    _eq(a.asformat("gcxs") @ b.asformat("gcxs"), want)
    _eq(a @ np.zeros(shape_b), want)
    _eq(np.zeros(shape_a) @ b, want)


def test_elementwise_on_empty_and_zero_size(sp):
    rng = np.random.default_rng(1)
    x = sp.random((3, 4), density=0.5, random_state=rng)
    z = sp.zeros((3, 4))
    d = x.todense()
    for got, want in ((x + z, d), (z + x, d), (x * z, d * 0), (z * z, np.zeros((3, 4))), (z - x, -d), (z > x, 0 > d),
                      (np.maximum(z, x), np.maximum(0, d)), (z + 1.5, np.full((3, 4), 1.5)), (np.exp(z), np.ones((3, 4))),
                      (z * np.ones((3, 4)), np.zeros((3, 4))), (-z, np.zeros((3, 4)))):
        _eq(got, want)
    e = sp.COO.from_numpy(np.zeros((0, 4)))
    _eq(e + e, np.zeros((0, 4)))
    _eq(e * 2.0, np.zeros((0, 4)))
    _eq(np.sqrt(e), np.zeros((0, 4)))
    _eq(e + np.zeros((0, 4)), np.zeros((0, 4)))
    _eq(e + sp.COO.from_numpy(np.zeros((1, 4))), np.zeros((0, 4)))  # broadcasting against a zero-length axis
    s0 = sp.COO.from_numpy(np.array(3.0))  # 0-D
    _eq(s0 + s0, np.array(6.0))
    _eq(x * s0, d * 3.0)
    _eq(np.add(s0, 1.0), np.array(4.0))


def test_reductions_of_empty_arrays(sp):
    z = sp.zeros((3, 4))
    f = sp.full((3, 4), 2.0)
    for axis in (None, 0, 1, (0, 1)):
        _eq(z.sum(axis=axis), np.zeros((3, 4)).sum(axis=axis))
        _eq(z.max(axis=axis), np.zeros((3, 4)).max(axis=axis))
        _eq(f.sum(axis=axis), np.full((3, 4), 2.0).sum(axis=axis))  # fill-value contribution only
        _eq(f.prod(axis=axis), np.full((3, 4), 2.0).prod(axis=axis))
        _eq(z.any(axis=axis), np.zeros((3, 4)).any(axis=axis))
        with pytest.raises(ValueError, match="dense result"):
            f.all(axis=axis)  # logical_and(2.0, 2.0) = True != fill: refused like upstream
        _eq(z.mean(axis=axis), np.zeros((3, 4)).mean(axis=axis))
    e = sp.COO.from_numpy(np.zeros((0, 4)))
    _eq(e.sum(axis=0), np.zeros(4))
    _eq(e.sum(axis=1), np.zeros(0))
    _eq(e.sum(), np.array(0.0))
    g = z.asformat("gcxs")
    _eq(g.sum(axis=0), np.zeros(4))
    assert isinstance(g.sum(axis=0), sp.GCXS)


def test_structure_ops_on_empty(sp):
    z = sp.zeros((3, 4, 2))
    d = np.zeros((3, 4, 2))
    _eq(z.T, d.T)
    _eq(z.reshape((4, 6)), d.reshape(4, 6))
    _eq(z[1:, ::2], d[1:, ::2])
    _eq(z[[2, 0]], d[[2, 0]])
    assert z[1, 2, 1] == 0.0
    _eq(sp.concatenate([z, z], axis=1), np.concatenate([d, d], axis=1))
    _eq(sp.stack([z, z]), np.stack([d, d]))
    _eq(sp.GCXS(z), d)
    _eq(sp.GCXS(z).tocoo(), d)
    _eq(sp.GCXS(z, compressed_axes=(1, 2)).change_compressed_axes((0,)), d)
    _eq(sp.triu(z[0]), d[0])
    _eq(sp.diagonal(z), np.diagonal(d))
    _eq(sp.roll(z, 1, 0), d)
    _eq(sp.pad(z, 1), np.pad(d, 1))
    _eq(z.astype(np.float32), d.astype(np.float32))
    assert z.nonzero()[0].size == 0 and sp.argwhere(z).shape == (0, 3)
    e = sp.COO.from_numpy(np.zeros((0, 3)))
    _eq(e.T, np.zeros((3, 0)))
    _eq(e[:, 1:], np.zeros((0, 2)))
    _eq(sp.concatenate([e, sp.zeros((2, 3))]), np.zeros((2, 3)))


def test_single_entry_and_single_row(sp):
    """The smallest non-empty launches (one entry, one row, one column)."""
    a = sp.COO(np.array([[2], [1]]), np.array([3.0]), shape=(4, 3))
    b = sp.COO(np.array([[1], [0]]), np.array([-2.0]), shape=(3, 1))
    da, db = a.todense(), b.todense()
    _eq(a @ b, da @ db)
    _eq(a.asformat("gcxs") @ db, da @ db)
    _eq(a + a, da + da)
    _eq(a.sum(axis=0), da.sum(0))
    _eq(a.max(axis=1), da.max(1))
    _eq(a.T @ a, da.T @ da)
    r = sp.COO.from_numpy(np.array([[0.0, 5.0, 0.0]]))
    _eq(r @ r.T, np.array([[25.0]]))
    _eq(r.reshape((3, 1)) @ r, r.todense().T @ r.todense())
