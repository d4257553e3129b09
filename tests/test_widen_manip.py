"""Array manipulation next to the hot path (sparse_b200/_manip.py, _creation.py) vs dense NumPy.

Same bar as the reference's own tests for these functions (tests/test_coo.py upstream: `assert_eq(numpy_result,
sparse_result)` -- equal values, canonical COO, nnz == number of non-fill entries).  Every body runs on the NumPy mock
of the kernel layer (host logic, no GPU) and on the CUDA kernels (`-m gpu`)."""
import numpy as np
import pytest

from _api import sp  # noqa: F401


def assert_eq(s, d, sp=None):
    """Result `s` (sparse or dense) equals dense `d`; a COO must be canonical and hold no fill values."""
    d = np.asarray(d)
    if hasattr(s, "todense"):
        got = s.todense()
        assert s.shape == d.shape, (s.shape, d.shape)
        assert s.dtype == d.dtype, (s.dtype, d.dtype)
        assert np.array_equal(got, d, equal_nan=d.dtype.kind == "f"), (got, d)
        c = s.asformat("coo")
        if c.ndim:
            lin = np.ravel_multi_index(tuple(c.coords), c.shape) if c.nnz else np.empty(0, np.int64)
            assert np.all(np.diff(lin) > 0), "coordinates not sorted / not unique"
        assert c.nnz == int(np.sum(d != c.fill_value)), "stored fill values"
    else:
        assert np.array_equal(np.asarray(s), d)


@pytest.fixture
def xy(sp):
    rng = np.random.default_rng(0)
    x = sp.random((3, 4, 5), density=0.4, random_state=rng)
    y = sp.random((3, 4, 5), density=0.4, random_state=rng)
    return x, y, x.todense(), y.todense()


@pytest.mark.parametrize("axis", [0, 1, 2, -1])
def test_concatenate(sp, xy, axis):
    x, y, d, e = xy
    assert_eq(sp.concatenate([x, y, x], axis=axis), np.concatenate([d, e, d], axis=axis))
    assert_eq(np.concatenate([x, y], axis=axis), np.concatenate([d, e], axis=axis))
    g = sp.concatenate([sp.GCXS(x), sp.GCXS(y)], axis=axis)
    assert isinstance(g, sp.GCXS) and g.compressed_axes == (axis % 3,)
    assert_eq(g, np.concatenate([d, e], axis=axis))


def test_concatenate_flat_mixed_dtype_and_errors(sp, xy):
    x, y, d, e = xy
    assert_eq(sp.concatenate([x, y], axis=None), np.concatenate([d, e], axis=None))
    xi = (x * 10).astype(np.int64)
    assert_eq(sp.concatenate([xi, y], axis=1), np.concatenate([xi.todense(), e], axis=1))
    with pytest.raises(ValueError):
        sp.concatenate([d, x])  # dense operand
    with pytest.raises(ValueError):
        sp.concatenate([])
    with pytest.raises(ValueError):
        sp.concatenate([x, x + 1.0])  # fill values differ
    with pytest.raises(ValueError):
        sp.concatenate([x, y[:, :2, :3]], axis=0)


@pytest.mark.parametrize("axis", [0, 1, 2, 3, -1])
def test_stack(sp, xy, axis):
    x, y, d, e = xy
    assert_eq(sp.stack([x, y, x], axis=axis), np.stack([d, e, d], axis=axis))
    assert_eq(np.stack([x, y], axis=axis), np.stack([d, e], axis=axis))
    g = sp.stack([sp.GCXS(x), sp.GCXS(y)], axis=axis)
    assert isinstance(g, sp.GCXS)
    assert_eq(g, np.stack([d, e], axis=axis))


def test_unstack(sp, xy):
    x, _, d, _ = xy
    for axis in (0, 1, -1):
        parts = sp.unstack(x, axis=axis)
        want = [np.take(d, i, axis=axis) for i in range(d.shape[axis])]
        assert len(parts) == len(want)
        for p, w in zip(parts, want):
            assert_eq(p, w)
    with pytest.raises(ValueError):
        sp.unstack(x, axis=3)
    with pytest.raises(TypeError):
        sp.unstack(d, axis=0)


def test_axes(sp, xy):
    x, _, d, _ = xy
    assert_eq(sp.moveaxis(x, 0, 2), np.moveaxis(d, 0, 2))
    assert_eq(np.moveaxis(x, [0, 1], [2, 0]), np.moveaxis(d, [0, 1], [2, 0]))
    assert_eq(x.swapaxes(0, 2), d.swapaxes(0, 2))
    assert_eq(np.swapaxes(x, -1, 0), np.swapaxes(d, -1, 0))
    assert_eq(sp.matrix_transpose(x), np.swapaxes(d, -1, -2))
    assert_eq(sp.permute_dims(x, (1, 2, 0)), d.transpose(1, 2, 0))
    assert_eq(sp.squeeze(x[:1]), np.squeeze(d[:1]))
    assert_eq(x[:, :1].squeeze(1), d[:, :1].squeeze(1))
    assert_eq(sp.expand_dims(x, axis=1), np.expand_dims(d, 1))
    assert_eq(sp.expand_dims(x, axis=(0, -1)), np.expand_dims(d, (0, -1)))
    assert_eq(x.flatten(), d.flatten())
    with pytest.raises(ValueError):
        x.swapaxes(0, 3)
    with pytest.raises(ValueError):
        sp.squeeze(x, 0)
    with pytest.raises(ValueError):
        x.transpose((0, 0, 1))
    with pytest.raises(ValueError):
        x.transpose(1)


def test_flip_roll(sp, xy):
    x, _, d, _ = xy
    assert_eq(sp.flip(x, axis=1), np.flip(d, 1))
    assert_eq(sp.flip(x), np.flip(d))
    assert_eq(np.flip(x, (0, 2)), np.flip(d, (0, 2)))
    for shift, axis in ((1, 0), (-2, 1), (7, 2), (3, None), ((1, 2), (0, 2)), (0, 1)):
        assert_eq(sp.roll(x, shift, axis), np.roll(d, shift, axis))


@pytest.mark.parametrize("k", [-2, -1, 0, 1, 3])
def test_triangles_and_diagonals(sp, xy, k):
    x, _, d, _ = xy
    assert_eq(sp.triu(x, k), np.triu(d, k))
    assert_eq(sp.tril(x, k), np.tril(d, k))
    for a1, a2 in ((0, 1), (1, 2), (2, 0), (0, 2)):
        assert_eq(sp.diagonal(x, k, a1, a2), np.diagonal(d, k, a1, a2))


def test_diagonalize_pad(sp, xy):
    x, _, d, _ = xy
    z = sp.diagonalize(x, axis=1)
    assert z.shape == (3, 4, 5, 4) and z.nnz == x.nnz
    for i in range(4):
        assert_eq(z[:, i, :, i], d[:, i, :])
    assert_eq(sp.diagonal(z, 0, 1, 3), np.moveaxis(d, 1, -1))
    assert_eq(sp.pad(x, ((1, 2), (0, 1), (3, 0))), np.pad(d, ((1, 2), (0, 1), (3, 0))))
    assert_eq(np.pad(x, 2), np.pad(d, 2))
    g = sp.pad(sp.GCXS(x, compressed_axes=(1,)), 1)
    assert isinstance(g, sp.GCXS) and g.compressed_axes == (1,)
    assert_eq(g, np.pad(d, 1))
    with pytest.raises(NotImplementedError):
        sp.pad(x, 1, mode="reflect")
    with pytest.raises(ValueError):
        sp.pad(x, 1, constant_values=3)


def test_repeat_tile_kron_outer(sp, xy):
    x, y, d, e = xy
    assert_eq(sp.repeat(x, 3, axis=1), np.repeat(d, 3, axis=1))
    assert_eq(sp.repeat(x, 2), np.repeat(d, 2))
    assert_eq(sp.tile(x, (2, 1, 3)), np.tile(d, (2, 1, 3)))
    assert_eq(sp.tile(x, 2), np.tile(d, 2))
    assert_eq(sp.tile(x[0], (2, 1, 2)), np.tile(d[0], (2, 1, 2)))
    assert_eq(sp.kron(x[0], y[1]), np.kron(d[0], e[1]))
    assert_eq(sp.kron(x, y[1]), np.kron(d, e[1]))
    assert_eq(sp.kron(x[0], e[1]), np.kron(d[0], e[1]))
    assert_eq(sp.outer(x[0], y[1, 0]), np.outer(d[0], e[1, 0]))
    assert_eq(np.multiply.outer(x[0, 0], y[1, 1]), np.multiply.outer(d[0, 0], e[1, 1]))
    with pytest.raises(ValueError):
        sp.repeat(x, [1, 2], axis=0)


ADV = [([1, 0], 0), (1, [0, 2]), (0, [1, 0], 0), (1, [2, 0], 0), (1, [], 0),
       ([True, False, True], slice(1, None), slice(-2, None)),
       (slice(1, None), slice(-2, None), [True, False, True, False, True]), ([1, 0],), (Ellipsis, [2, 1, 3]),
       (slice(None), [2, 1, 2]), (1, [2, 0, 1]), (None, [2, 1], None), (slice(None, None, -1), [3, 3, 0, 3]),
       ([-1, 0, -3],)]


@pytest.mark.parametrize("index", ADV, ids=[str(i).replace(" ", "") for i in ADV])
def test_one_advanced_index(sp, xy, index):
    """The reference's test_advanced_indexing list (tests/test_coo.py:477-497 upstream) plus duplicates / negatives."""
    x, _, d, _ = xy
    assert_eq(x[index], d[index])
    g = sp.GCXS(x)[index]
    assert isinstance(g, sp.GCXS)
    assert_eq(g, d[index])


def test_take(sp, xy):
    x, _, d, _ = xy
    assert_eq(sp.take(x, [2, 0, 2], axis=1), np.take(d, [2, 0, 2], axis=1))
    assert_eq(sp.take(x, [5, 1, 59]), np.take(d, [5, 1, 59]))
    assert_eq(np.take(x, [-1, 0], axis=2), np.take(d, [-1, 0], axis=2))
    with pytest.raises(IndexError):
        sp.take(x, [5], axis=0)


def test_take_large(sp):
    rng = np.random.default_rng(5)
    x = sp.random((300, 40, 50), density=0.01, random_state=rng)
    d = x.todense()
    idx = rng.integers(0, 300, size=500)  # many repeated indices: several rounds
    assert_eq(x[idx], d[idx])
    idx = rng.permutation(40)[:17]
    assert_eq(x[:, idx, ::2], d[:, idx, ::2])


def test_creation(sp, xy):
    x, _, d, _ = xy
    assert_eq(sp.eye(4, 5, k=1), np.eye(4, 5, k=1))
    assert_eq(sp.eye(4, 5, k=-2), np.eye(4, 5, k=-2))
    assert_eq(sp.eye(3, dtype=np.int64, format="gcxs"), np.eye(3, dtype=np.int64))
    assert_eq(sp.full((2, 3), 7.5), np.full((2, 3), 7.5))
    assert_eq(sp.zeros((2, 3), dtype=np.float32), np.zeros((2, 3), np.float32))
    assert_eq(sp.ones_like(x), np.ones_like(d))
    assert_eq(sp.zeros_like(x, shape=(2, 2)), np.zeros((2, 2)))
    assert isinstance(sp.full_like(sp.GCXS(x), 2.0), sp.GCXS)
    assert sp.asarray(d, format="gcxs").nnz == x.nnz
    assert sp.asarray(x) is x
    assert sp.result_type(x, np.float32) == np.float64
    assert sp.can_cast(x, np.float32, casting="same_kind") and not sp.can_cast(x, np.int32)
    assert np.array_equal(sp.asnumpy(x), d)
    assert np.array_equal(sp.argwhere(x), np.argwhere(d))
    for got, want in zip(x.nonzero(), d.nonzero()):
        assert np.array_equal(got, want)
    with pytest.raises(RuntimeError):
        np.asarray(x)  # no silent densification


def test_diff_vecdot(sp, xy):
    x, y, d, e = xy
    assert_eq(sp.diff(x, axis=1), np.diff(d, axis=1))
    assert_eq(sp.diff(x, n=2, axis=0), np.diff(d, n=2, axis=0))
    assert_eq(sp.diff(x, axis=2, prepend=y[:, :, :1], append=y[:, :, :2]),
              np.diff(d, axis=2, prepend=e[:, :, :1], append=e[:, :, :2]))
    got = sp.vecdot(x, y, axis=1)
    assert np.allclose(got.todense(), np.sum(d * e, axis=1), rtol=1e-12)
