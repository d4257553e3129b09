"""argmax / argmin (sparse_b200/_argreduce.py) vs NumPy: first index of the extreme value, fill values included.
Upstream: tests/test_coo.py:1647-1687 (`test_argmax_argmin`, `_3D`, `_constraint`)."""
import numpy as np
import pytest

from _api import sp  # noqa: F401


@pytest.mark.parametrize("arr", [np.array([[0, 3, 0], [1, 2, 0]]), np.array([[[0, 0], [1, 0]], [[5, 0], [0, -3]]])],
                         ids=["2d", "3d"])
@pytest.mark.parametrize("axis", [None, 0, 1, -1])
@pytest.mark.parametrize("keepdims", [True, False])
def test_upstream_cases(sp, arr, axis, keepdims):
    s = sp.COO.from_numpy(arr)
    for f, g in ((sp.argmax, np.argmax), (sp.argmin, np.argmin)):
        got = f(s, axis=axis, keepdims=keepdims)
        assert isinstance(got, sp.COO) and got.fill_value == 0
        np.testing.assert_equal(got.todense(), g(arr, axis=axis, keepdims=keepdims))
    assert np.array_equal(np.argmax(s, axis=axis).todense(), np.argmax(arr, axis=axis))  # __array_function__


def test_mostly_empty_3d(sp):
    d = np.zeros((100, 55, 3))
    d[10, 10, 0] = 3
    d[10, 10, 1] = 3
    d[10, 9, 0] = -2
    s = sp.COO.from_numpy(d)
    for axis in (None, 0, 1, 2):
        np.testing.assert_equal(sp.argmax(s, axis=axis).todense(), np.argmax(d, axis=axis))
        np.testing.assert_equal(sp.argmin(s, axis=axis).todense(), np.argmin(d, axis=axis))


@pytest.mark.parametrize("seed", range(12))
def test_random_shapes_and_fill_values(sp, seed):
    rng = np.random.default_rng(seed)
    shape = tuple(rng.integers(1, 6, size=rng.integers(1, 4)))
    fill = float(rng.choice([0.0, 2.0, -1.0]))
    d = np.full(shape, fill)
    mask = rng.random(shape) < rng.choice([0.0, 0.3, 0.7, 1.0])
    d[mask] = rng.integers(-3, 4, size=shape).astype(float)[mask]
    for x in (sp.COO.from_numpy(d, fill_value=fill), sp.GCXS.from_numpy(d, fill_value=fill)):
        for axis in [None] + list(range(len(shape))):
            for keepdims in (False, True):
                np.testing.assert_equal(sp.argmax(x, axis=axis, keepdims=keepdims).todense(),
                                        np.argmax(d, axis=axis, keepdims=keepdims))
                np.testing.assert_equal(sp.argmin(x, axis=axis, keepdims=keepdims).todense(),
                                        np.argmin(d, axis=axis, keepdims=keepdims))


def test_nan_and_errors(sp):
    d = np.array([[1.0, np.nan, 3.0], [0, 0, 0], [np.nan, 0, 0]])
    s = sp.COO.from_numpy(d)
    np.testing.assert_equal(sp.argmax(s, axis=1).todense(), np.argmax(d, axis=1))
    np.testing.assert_equal(sp.argmin(s, axis=0).todense(), np.argmin(d, axis=0))
    full = sp.COO.from_numpy(np.full((2, 2), 2), fill_value=2)
    for f in (sp.argmax, sp.argmin):
        with pytest.raises(ValueError, match="`axis=2` is out of bounds for array of dimension 2."):
            f(full, axis=2)
        with pytest.raises(ValueError):
            f(np.ones((2, 2)), axis=0)
        with pytest.raises(ValueError):
            f(s, axis=(0, 1))
