"""Reductions vs golden outputs of the reference (tests/test_coo.py:44-193, tests/test_compressed.py:36-131 upstream).
Coordinates / result types exact; values to 1e-6 (f32) / 1e-12 (f64) because NumPy's reduceat order is unspecified
(max/min/any/all and integer sums are compared exactly)."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("reduce_api")


def _id(c):
    return f"{c['op']}-{c['dtype']}-{c['fmt']}-ax{c['axis']}-{'kd' if c['keepdims'] else ''}{c.get('note','')}".replace(" ", "")


@pytest.mark.parametrize("c", CASES, ids=[f"{i}-{_id(c)}" for i, c in enumerate(CASES)])
def test_reduce(sp, c):
    x = dec(sp, c, "a_", c["fmt"], ca=c.arr.get("a_ca") if c["fmt"] == "gcxs" else None)
    axis = c["axis"]
    if isinstance(axis, list):
        axis = tuple(axis)
    with np.errstate(all="ignore"):
        got = getattr(x, c["op"])(axis=axis, keepdims=c["keepdims"])
    exact = c["op"] in ("max", "min", "any", "all") or c["dtype"] == "int64" and c["op"] != "mean"
    tol = 2e-6 if c["dtype"] == "float32" else 1e-12
    check_result(sp, got, c, exact=exact, rtol=tol, atol=tol)


def test_dense_result_reduction_raises(sp):
    x = sp.COO(np.array([[0, 1]]), np.array([1.0, 2.0]), shape=(3,), has_duplicates=False, sorted=True, fill_value=1.0)
    with pytest.raises(TypeError):
        x.reduce(np.arctan2, axis=0)
