"""Broadcasting elemwise through NumPy ufuncs / operators vs golden outputs of the reference
(cases modelled on tests/test_elemwise.py:13-44, 79-111, 143-156, 206-238, 252-305, 387-413 upstream).
Coordinates are compared exactly; data bit-exactly for arithmetic/comparison ops, to 1e-6/1e-12 for transcendentals."""
import numpy as np
import pytest

from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("elemwise_api")
TRANSCENDENTAL = {"expm1", "sin", "sqrt_abs"}


def _id(c):
    return f"{c['op']}-{c['dtype']}-{c['rhs']}-{c.get('note','')}{'-swap' if c.get('swap') else ''}"


def _apply(sp, c):
    name = c["op"]
    a = dec(sp, c, "a_", "gcxs" if c["rhs"] == "gcxs" else "coo",
            ca=c.arr.get("a_ca") if c["rhs"] == "gcxs" else None)
    if c["rhs"] == "unary":
        if name == "sqrt_abs":
            return np.sqrt(np.abs(a))
        return getattr(np, name)(a)
    f = getattr(np, name)
    if c["rhs"] == "scalar":
        return f(a, c.arr["scalar"][()])
    if c["rhs"] == "dense":
        d = np.array(c.arr["dense"])
        return f(d, a) if c.get("swap") else f(a, d)
    b = dec(sp, c, "b_", "gcxs" if c["rhs"] == "gcxs" else "coo", ca=c.arr.get("b_ca") if c["rhs"] == "gcxs" else None)
    return f(a, b)


@pytest.mark.parametrize("c", CASES, ids=[f"{i}-{_id(c)}" for i, c in enumerate(CASES)])
def test_elemwise(sp, c):
    if "error" in c:
        with pytest.raises(ValueError):
            _apply(sp, c)
        return
    with np.errstate(all="ignore"):
        got = _apply(sp, c)
    exact = c["op"] not in TRANSCENDENTAL
    tol = 1e-6 if c["dtype"] == "float32" else 1e-12
    check_result(sp, got, c, exact=exact, rtol=tol, atol=tol * 1e-3)


def test_unsupported_function_raises(sp):
    x = sp.COO(np.array([[0, 1]]), np.array([1.0, 2.0]), shape=(3,), has_duplicates=False, sorted=True)
    assert np.array_equal(sp.elemwise(lambda a, b: a + b, x, x).todense(), [2.0, 4.0, 0.0])  # composite: evaluated
    with pytest.raises(TypeError, match="CUDA op set"):
        np.arctan2(x, x)
    with pytest.raises(TypeError, match="CUDA op set"):
        sp.elemwise(lambda a, b: np.arctan2(a, b), x, x)
    with pytest.raises(ValueError, match="could not be broadcast"):
        y = sp.COO(np.array([[0, 1]]), np.array([1.0, 2.0]), shape=(4,), has_duplicates=False, sorted=True)
        x + y
    with pytest.raises(TypeError, match="CUDA op set"):
        sp.elemwise(np.add, x, x, x)
