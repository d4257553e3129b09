"""Basic indexing of COO / GCXS vs golden outputs of the reference (tests/test_coo.py:408-474 upstream: the same
index list).  Everything is exact: result type, shape, fill value, coordinates / indices / indptr and values."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("indexing_api")


def _dec_index(enc):
    out = []
    for i in enc:
        if i == "...":
            out.append(Ellipsis)
        elif isinstance(i, list):
            out.append(slice(*i))
        else:
            out.append(i)
    return tuple(out) if len(out) != 1 else out[0]


def _id(i, c):
    return f"{i}-{c['fmt']}-{c['index']}".replace(" ", "")


@pytest.mark.parametrize("c", CASES, ids=[_id(i, c) for i, c in enumerate(CASES)])
def test_getitem(sp, c):
    x = dec(sp, c, "a_", c["fmt"], ca=c.arr.get("a_ca") if c["fmt"] == "gcxs" else None)
    got = x[_dec_index(c["index"])]
    if c.get("check") == "dense":
        w = c.sub("out_")
        assert isinstance(got, sp.GCXS) and got.compressed_axes == tuple(int(a) for a in w["ca"])
        want = sp.GCXS((w["data"], w["indices"], w["indptr"]), shape=tuple(int(s) for s in w["shape"]),
                       compressed_axes=got.compressed_axes).todense()
        assert np.array_equal(got.todense(), want)
        return
    check_result(sp, got, c, exact=True)


def test_indexing_errors(sp):
    x = sp.random((2, 3, 4), density=0.5, random_state=3)
    for bad in ((Ellipsis, Ellipsis), (1, 1, 1, 1), (slice(None),) * 4, 5, (0, 3), -3, (0, 0, -5), 1.5):
        with pytest.raises(IndexError):
            x[bad]
    with pytest.raises(IndexError):
        x[[0, 1], [1, 2, 0]]  # advanced indices of different lengths
    with pytest.raises(IndexError):
        x[[[0, 1]]]  # only one-dimensional advanced indices
    with pytest.raises(IndexError):
        x[np.array([True, False, True])]  # boolean mask of the wrong length
    with pytest.raises(IndexError):
        x[[0, 2]]


def test_large_slices_match_numpy(sp):
    rng = np.random.default_rng(9)
    x = sp.random((40, 50, 60), density=0.02, random_state=rng)
    d = x.todense()
    for index in ((slice(3, 37, 3), slice(None, None, -2), slice(10, 50)), (17,), (slice(None), 20), (Ellipsis, 59),
                  (slice(None, None, -1), 3, slice(None, None, -7)), (None, 5, slice(2, 4), None)):
        got = x[index]
        assert np.array_equal(got.todense(), d[index])
        assert np.array_equal(got.coords, np.stack(np.nonzero(d[index])))
