"""n-ary / user-defined functions through `elemwise` vs golden outputs of the reference (tests/test_elemwise.py:252-305
upstream: trinary broadcasting, incl. the NaN / inf "pathological" operands).  Arithmetic composites and `where` are
exact: the same IEEE operations run in the same order, element for element (the one `**` case is compared to 1e-12)."""
import numpy as np
import pytest

import _nary_funcs as NF
from _api import check_result, dec, sp  # noqa: F401
from _golden import load

CASES = load("elemwise_nary_api")


@pytest.mark.parametrize("c", CASES, ids=[f"{i}-{c['op']}-f{c['func']}-{c['note']}" for i, c in enumerate(CASES)])
def test_elemwise_composite(sp, c):
    ops = [dec(sp, c, f"op{i}_") for i in range(c["n"])]
    if c["op"] == "where":
        ops = [o[()] if isinstance(o, np.ndarray) and o.ndim == 0 else o for o in ops]
        with np.errstate(all="ignore"):
            got = sp.where(*ops)
        check_result(sp, got, c, exact=True)
        return
    f = (NF.TRINARY if c["op"] == "trinary" else NF.UNARY_BINARY)[c["func"]]
    with np.errstate(all="ignore"):
        got = sp.elemwise(f, *ops)
    uses_pow = c["op"] == "unary_binary" and c["func"] == 1  # ** 0.5: transcendental, 1e-12 like the other pow cases
    check_result(sp, got, c, exact=not uses_pow, rtol=1e-12, atol=1e-15)


def test_where_argument_checks(sp):
    s = sp.random((2, 3, 4), density=0.5, random_state=3)
    got = sp.where(s)
    want = np.where(s.todense())
    assert len(got) == len(want) and all(np.array_equal(g, w) for g, w in zip(got, want))
    with pytest.raises(ValueError):
        sp.where(np.ones((2, 3)))
    with pytest.raises(ValueError):
        sp.where(s.astype(np.bool_), s)
    assert isinstance(np.where(s.astype(np.bool_), s, s * 2.0), sp.COO)  # through __array_function__


def test_nary_ufunc_and_opaque_functions_raise(sp):
    a = sp.random((3, 4), density=0.5, random_state=1)
    got = sp.elemwise(np.clip, a, 0.1, 0.5)  # not a ufunc: evaluated through __array_function__ -> COO.clip
    assert np.array_equal(got.todense(), np.clip(a.todense(), 0.1, 0.5))
    with pytest.raises(TypeError):
        sp.elemwise(lambda x: "not an array", a)
