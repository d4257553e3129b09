"""Regression tests for the differences `tools/fuzz_vs_reference.py` found between this package's host layer and the
reference itself (differential fuzzing over formats, dtypes, fill values, shapes with zero-length axes, 0-D operands).
Expected values are what the reference returns (cited file:line); the host layer is what is under test, so the bodies
run on the NumPy mock of the kernel layer (no GPU needed, `-m "not gpu"`)."""
import numpy as np
import pytest
import torch

import _mock_kernels

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="host-layer tests on the mock backend")


@pytest.fixture
def sp():
    import sparse_b200

    _mock_kernels.install()
    yield sparse_b200
    _mock_kernels.uninstall()


def _gcxs(sp, d, ca=None, fill=0):
    c = sp.COO.from_numpy(d, fill_value=np.asarray(fill, dtype=d.dtype)[()])
    return c.asformat("gcxs", **({"compressed_axes": ca} if ca is not None else {}))


def test_tensordot_2d_gcxs_operand_keeps_its_transposed_compressed_axis(sp):
    """_common.py:212-214 + compressed.py:665-667,728-729: a 2-D GCXS operand is transposed in O(1) (compressed axis
    flips) and a reshape to another 2-D shape keeps that axis -- it decides which CSR/CSC kernel runs and therefore
    the compressed axis of the product."""
    a = _gcxs(sp, np.arange(1.0, 5.0, dtype=np.float32).reshape(1, 4), ca=(0,))
    b = sp.COO.from_numpy(np.arange(16, dtype=np.int32).reshape(2, 4, 2, 1) % 3)
    r = sp.tensordot(a, b, ([1, 0], [1, 3]))
    assert isinstance(r, sp.GCXS) and tuple(r.compressed_axes) == (1,)
    assert np.array_equal(r.todense(), np.tensordot(a.todense(), b.todense(), ([1, 0], [1, 3])))


def test_elemwise_of_0d_gcxs_stays_gcxs_and_stores_nothing(sp):
    """_umath.py:438-439,480-503: 0-D sparse operands become scalars, the result has no stored entry, its value is the
    fill value, and `asformat(out_type)` keeps GCXS."""
    x = _gcxs(sp, np.array(3, dtype=np.int64))
    r = np.add(np.float64(-1.0), x)
    assert isinstance(r, sp.GCXS) and r.shape == () and r.nnz == 0 and r.fill_value == 2.0 and r.dtype == np.float64
    y = sp.COO.from_numpy(np.array(3.0, dtype=np.float32))
    r = y.clip(-1, 2)  # Python bounds are weak scalars: float32 stays float32
    assert isinstance(r, sp.COO) and r.dtype == np.float32 and r.fill_value == np.float32(2)


def test_elemwise_with_a_zero_length_axis_returns_coo_whatever_the_operands(sp):
    """_umath.py:467-477: the empty result is returned before `asformat`."""
    x = _gcxs(sp, np.zeros((0, 3), dtype=np.float32), fill=2)
    r = np.negative(x)
    assert isinstance(r, sp.COO) and r.shape == (0, 3) and r.fill_value == np.float32(-2)
    # next to an empty ndarray the fill value is func(fill, zero of the ndarray's dtype) (:529-534)
    s = sp.COO.from_numpy(np.ones((2, 0, 3), dtype=bool), fill_value=True)
    r = np.logical_or(s, np.zeros((2, 0, 3), dtype=np.int32))
    assert isinstance(r, sp.COO) and r.fill_value == np.True_ and r.dtype == np.bool_


def test_0d_sparse_next_to_a_dense_array(sp):
    """_umath.py:536-546: func(fill, ndarray) constant -> an array without stored entries whose fill value is that
    constant; otherwise the dense result."""
    x = sp.COO.from_numpy(np.array(2.0))
    r = np.logical_and(np.zeros(3, dtype=np.float32), x)
    assert isinstance(r, sp.COO) and r.shape == (3,) and r.nnz == 0 and r.fill_value == np.False_
    r = np.add(np.arange(3.0), x)
    assert isinstance(r, np.ndarray) and np.array_equal(r, np.arange(3.0) + 2)


@pytest.mark.parametrize("name,want", [("add", np.logical_or), ("multiply", np.logical_and),
                                       ("maximum", np.logical_or), ("minimum", np.logical_and),
                                       ("fmax", np.logical_or), ("fmin", np.logical_and)])
def test_arithmetic_on_two_boolean_operands_is_the_logical_op(sp, name, want):
    rng = np.random.default_rng(5)
    da, db = rng.random((3, 4)) < 0.5, rng.random((3, 4)) < 0.5
    a, b = sp.COO.from_numpy(da), sp.COO.from_numpy(db)
    r = getattr(np, name)(a, b)
    assert r.dtype == np.bool_ and np.array_equal(r.todense(), getattr(np, name)(da, db))
    assert np.array_equal(r.todense(), want(da, db))
    r = getattr(np, name)(a, True)
    assert r.dtype == np.bool_ and np.array_equal(r.todense(), getattr(np, name)(da, True))


def test_invert_and_abs_of_bool(sp):
    d = np.array([[True, False, True], [False, False, True]])
    x = sp.COO.from_numpy(d, fill_value=True)
    r = np.invert(x)
    assert r.dtype == np.bool_ and r.fill_value == np.False_ and np.array_equal(r.todense(), ~d)
    r = np.abs(x)
    assert r.dtype == np.bool_ and np.array_equal(r.todense(), d)


def test_clip(sp):
    """`sparse.clip` converts to COO first (_coo/common.py:1070-1071), the method keeps the format; numpy.clip takes
    Python integers outside the dtype's range."""
    d = np.array([[5, 0, 1], [0, 3, 0]], dtype=np.uint32)
    g = _gcxs(sp, d)
    assert isinstance(sp.clip(g, 1, 2), sp.COO) and isinstance(g.clip(1, 2), sp.GCXS)
    r = sp.clip(g, -1, 2)
    assert r.dtype == np.uint32 and np.array_equal(r.todense(), np.clip(d, -1, 2))
    with pytest.raises(ValueError, match="dense result"):
        sp.clip(d, 0, 1)


@pytest.mark.parametrize("axis", [None, (0, 1, 2)])
def test_reduction_over_every_axis_of_gcxs_goes_through_coo(sp, axis):
    """compressed.py:355-360: `flatten().tocoo().reduce(...)`; with keepdims the reshaped COO is returned."""
    d = (np.arange(24).reshape(2, 3, 4) % 5 == 0).astype(np.float64)
    g = _gcxs(sp, d)
    r = g.sum(axis=axis, keepdims=True)
    assert isinstance(r, sp.COO) and r.shape == (1, 1, 1) and r.todense().item() == d.sum()
    r = g.sum(axis=(0, 2), keepdims=True)
    assert isinstance(r, sp.GCXS) and np.array_equal(r.todense(), d.sum(axis=(0, 2), keepdims=True))


def test_broadcast_to_a_zero_length_axis_is_empty(sp):
    """A length-1 axis stretched to length 0 (numpy.broadcast_to allows it): no entry survives."""
    x = sp.COO.from_numpy(np.array([[1.0, 0.0, 2.0]]))
    r = sp.broadcast_to(x, (2, 0, 3))
    assert isinstance(r, sp.COO) and r.shape == (2, 0, 3) and r.nnz == 0 and r.coords.shape == (3, 0)
    assert r.todense().shape == (2, 0, 3)


def test_astype_of_an_empty_gcxs_is_the_empty_coo(sp):
    """astype is an elemwise call upstream (_sparse_array.py:626-658) and elemwise returns the empty COO as it is."""
    g = _gcxs(sp, np.zeros((0, 2)), ca=(0,))
    assert isinstance(g.astype("float32"), sp.COO)
    g = _gcxs(sp, np.ones((3, 2)), ca=(0,))
    r = g.astype("float32")
    assert isinstance(r, sp.GCXS) and tuple(r.compressed_axes) == (0,)


def test_conj_of_bool_is_int8_as_in_numpy(sp):
    d = np.array([[True, False], [False, True]])
    r = np.conj(sp.COO.from_numpy(d, fill_value=True))
    assert r.dtype == np.conj(d).dtype == np.int8 and np.array_equal(r.todense(), np.conj(d)) and r.fill_value == 1


def test_einsum_with_a_fully_summed_sparse_operand_returns_a_host_array(sp):
    """The sparse operand collapses to a scalar factor, the product is dense: an ndarray for host operands, as upstream."""
    a = np.array([2, -1], dtype=np.int64)
    b = (np.arange(48).reshape(4, 4, 3) % 5 - 2).astype(np.int64)
    r = sp.einsum("k,jil->ij", sp.COO.from_numpy(a), b)
    assert isinstance(r, np.ndarray) and np.array_equal(r, np.einsum("k,jil->ij", a, b))


def test_einsum_reduces_a_device_side_product_of_two_dense_operands(sp):
    """'j,ik,lj->...': the two dense operands are multiplied first (on the device); the single-term step that follows
    must accept that intermediate."""
    a = np.array([1.0, -2.0])
    c = np.array([[3.0, 1.0], [0.0, 2.0]])
    g = np.array([[1.0, 0, 0, 2], [0, 0, 3, 0], [0, 0, 0, 0], [4, 0, 0, 5]])
    for expr in ("j,ik,lj->", "j,ik,lj->il", "j,lj,ik->k"):
        ops = (a, _gcxs(sp, g), c) if expr.startswith("j,ik") else (a, c, _gcxs(sp, g))
        dense = tuple(np.asarray(o.todense()) if hasattr(o, "todense") else o for o in ops)
        r = sp.einsum(expr, *ops)
        assert np.array_equal(r.todense() if hasattr(r, "todense") else r, np.einsum(expr, *dense))


def test_sddmm_keeps_the_zero_sign_of_the_unfused_expression(sp):
    """`s * (a @ b)` upstream: the result's fill value is `0 * (a @ b)[0, 0]` (_umath.py:520-527) -- -0.0 under a
    negative corner -- and pruning is by bit pattern, so that sign decides which zero products stay stored.  The
    fused entry point gives the same arrays as the expression."""
    s = sp.COO.from_numpy(np.array([[0.0, -3.0, 0.0], [-1.0, 2.0, 0.0]], dtype=np.float32))
    a = np.array([[-3.0], [2.0]], dtype=np.float32)
    b = np.array([[1.0, 0.0, -1.0]], dtype=np.float32)  # a @ b = [[-3, +0, 3], [2, +0, -2]]: negative corner
    fused, plain = sp.sddmm(s, a, b), s * (a @ b)
    for r in (fused, plain):
        # -3 * +0.0 = -0.0 equals the -0.0 fill: pruned; 2 * +0.0 = +0.0 does not: stored
        assert np.signbit(r.fill_value) and r.nnz == 2
        assert np.array_equal(r.coords, [[1, 1], [0, 1]]) and np.array_equal(r.data, np.float32([-2.0, 0.0]))
        assert not np.signbit(r.data[1])
    b2 = np.array([[-1.0, 0.0, 1.0]], dtype=np.float32)  # corner +3: +0.0 fill, now the -0.0 product stays stored
    fused, plain = sp.sddmm(s, a, b2), s * (a @ b2)
    for r in (fused, plain):
        assert not np.signbit(r.fill_value) and r.nnz == 2
        assert np.array_equal(r.coords, [[0, 1], [1, 0]]) and r.data[0] == 0 and np.signbit(r.data[0]) and r.data[1] == 2


def test_where_decides_sparse_or_dense_on_its_inputs(sp):
    """Upstream's three-argument where is ONE elemwise call: the result is sparse iff where(fills | ndarrays) is one
    constant (_umath.py:536-546), whatever a pass-by-pass evaluation would see on the way."""
    a = sp.COO.from_numpy(np.array([-4.0, 0.0, 2.0, 0.0], dtype=np.float32), fill_value=np.float32(2))
    b = np.array([0.0, 2.0, 0.0, 0.0], dtype=np.float32)
    c = sp.COO.from_numpy(np.array([0, 0, 0, -4], dtype=np.int64))
    r = sp.where(a != 0, b, c)  # fill of the condition is True -> where(True, b, 0) = b: not constant -> dense
    assert isinstance(r, np.ndarray) and np.array_equal(r, np.where(a.todense() != 0, b, c.todense()))
    # constant probe, but a single pass would need a dense intermediate: still sparse, fill value from the probe
    cond = np.array([-1, -2, 0, -4])
    x0 = sp.COO.from_numpy(np.array(3))  # 0-D, value 3
    y = _gcxs(sp, np.array([[-3]]), fill=3)
    r = sp.where(cond, x0, y)
    assert isinstance(r, sp.COO) and r.shape == (1, 4) and r.fill_value == 3 and r.nnz == 1
    assert np.array_equal(r.todense(), np.where(cond, 3, np.array([[-3]])))
    # the sparse operands would have to be broadcast up to a non-constant dense result: the reference's error
    with pytest.raises(ValueError, match="mixed sparse-dense"):
        sp.where(np.array([1, 0, 1]), sp.COO.from_numpy(np.ones((2, 3))), np.array([5.0, 6.0, 7.0]))


def test_where_result_format_and_degenerate_operands(sp):
    g1 = _gcxs(sp, np.array([[1.0, 0.0], [0.0, 2.0]]), ca=(1,))
    g2 = _gcxs(sp, np.array([[0.0, 5.0], [6.0, 0.0]]), ca=(1,))
    r = sp.where(g1 != 0, g1, g2)  # every sparse operand GCXS with the same compressed axis: kept (_umath.py:419-422)
    assert isinstance(r, sp.GCXS) and tuple(r.compressed_axes) == (1,)
    assert isinstance(sp.where(g1 != 0, g1, g2.tocoo()), sp.COO)
    # all operands 0-D: host arithmetic, a 0-D array without stored entries in the operands' format
    z = sp.where(_gcxs(sp, np.array(True)), _gcxs(sp, np.array(2.0)), np.float64(5.0))
    assert isinstance(z, sp.GCXS) and z.shape == () and z.nnz == 0 and z.fill_value == 2.0
    # a zero-length axis: the empty COO; its fill value comes from the probe on fills and ndarrays (:516-534)
    e = sp.where(sp.COO.from_numpy(np.zeros((0,), dtype=bool)), _gcxs(sp, np.zeros(1, dtype=np.float32)), np.int64(3))
    assert isinstance(e, sp.COO) and e.shape == (0,) and e.fill_value == 3.0 and e.dtype == np.float64


def test_nary_elemwise_next_to_ndarrays_follows_upstreams_probe(sp):
    """A user-defined function with ndarray operands: upstream probes func(fill values | ndarrays) once
    (_umath.py:505-546) -- constant: sparse with that fill value, the stored set being the visited positions whose
    value differs from it BY BIT PATTERN; not constant: dense, or the error when the sparse operands would have to be
    broadcast up.  Signed zeros survive ((x + y) * 0 is -0.0 where x + y < 0)."""
    x = np.array([[0.0, 0.0, 0.0, 4.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 2.0, 0.0]])
    yd = np.array([[-3.0, 3.0, 3.0, 0.0], [-3.0, 4.0, -4.0, -2.0], [3.0, 0.0, -3.0, 1.0]])
    y = sp.COO.from_numpy(yd, fill_value=2.0)
    z = _gcxs(sp, np.zeros((3, 4), dtype=np.float32))
    r = sp.elemwise(lambda a, b, c: (a + b) * c, x, y, z)
    want = (x + yd) * np.zeros((3, 4), dtype=np.float32)
    assert isinstance(r, sp.COO) and not np.signbit(r.fill_value) and r.nnz == int(np.signbit(want).sum()) == 5
    assert np.array_equal(np.signbit(r.todense()), np.signbit(want)) and np.array_equal(r.todense(), want)
    # not constant and nothing to broadcast up: the dense result
    d = sp.elemwise(lambda a, b, c: a + b * c, sp.COO.from_numpy(yd), x + 1.0, np.full((3, 4), 0.5))
    assert isinstance(d, np.ndarray) and np.array_equal(d, yd + (x + 1.0) * 0.5)
    # constant probe (0 * b + 0.5): sparse, fill value 0.5
    s5 = sp.elemwise(lambda a, b, c: a * b + c, sp.COO.from_numpy(yd), x + 1.0, np.full((3, 4), 0.5))
    assert isinstance(s5, sp.COO) and s5.fill_value == 0.5 and np.array_equal(s5.todense(), yd * (x + 1.0) + 0.5)
    with pytest.raises(ValueError, match="mixed sparse-dense"):
        sp.elemwise(lambda a, b, c: a + b * c, sp.COO.from_numpy(yd), np.array([1.0, 2.0, 3.0, 4.0]), np.float64(1.0))


def test_astype_of_a_0d_array_returns_the_value_as_fill_value(sp):
    """astype is an elemwise call upstream and a 0-D operand is a scalar to it (_umath.py:438-439): the cast value
    comes back as the fill value of an array without stored entries."""
    x = sp.COO(np.empty((0, 1), dtype=np.intp), np.array([0], dtype=np.uint32), shape=(), fill_value=np.uint32(1))
    assert x.nnz == 1 and x.todense() == 0
    r = x.astype(np.int32)
    assert isinstance(r, sp.COO) and r.nnz == 0 and r.fill_value == 0 and r.dtype == np.int32
    g = x.asformat("gcxs").astype(np.float64)
    assert isinstance(g, sp.GCXS) and g.nnz == 0 and g.fill_value == 0.0
    # full reductions of narrow integers go through an internal cast that must NOT do this (fill value preserved)
    y = sp.COO.from_numpy(np.array([[1, 4, 1]], dtype=np.uint8), fill_value=np.uint8(1))
    m = y.max(axis=None, keepdims=True)
    assert m.shape == (1, 1) and m.fill_value == 1 and m.todense().item() == 4 and m.dtype == np.uint8


def test_imag_of_a_real_array_is_all_positive_zero(sp):
    """numpy.imag of a real array is +0 everywhere: no stored entry (x * 0 would leave -0.0 behind negative values)."""
    d = np.array([[0.0, 3.0, 0.0], [0.0, -2.0, 0.0]])
    for x in (sp.COO.from_numpy(d), _gcxs(sp, d, ca=(1,))):
        r = x.imag
        assert type(r) is type(x) and r.nnz == 0 and r.dtype == d.dtype and not np.signbit(r.todense()).any()
    g = _gcxs(sp, d, ca=(1,)).imag
    assert tuple(g.compressed_axes) == (1,)
    z = sp.COO.from_numpy(np.array(-3, dtype=np.int32), fill_value=np.int32(1))  # 0-D: value -3 stored, fill 1
    r = z.real
    assert r.nnz == 0 and r.fill_value == -3  # elemwise form of a 0-D array (_umath.py:438-439)


def test_kron_of_narrow_integers_and_join_of_1d_gcxs(sp):
    a = np.array([[16, 0], [-3, 2]], dtype=np.int8)
    b = np.array([[16, 1]], dtype=np.int8)
    k = sp.kron(sp.COO.from_numpy(a), sp.COO.from_numpy(b))
    want = np.kron(a, b)  # 16 * 16 wraps to 0 in int8; the product of two stored entries stays stored, as upstream
    assert k.dtype == np.int8 and np.array_equal(k.todense(), want) and k.nnz == 3 * 2
    kb = sp.kron(sp.COO.from_numpy(a != 0), sp.COO.from_numpy(b != 0))
    assert kb.dtype == np.bool_ and np.array_equal(kb.todense(), np.kron(a != 0, b != 0))
    g = _gcxs(sp, np.array([0.0, 1.0, 0.0]))
    assert isinstance(sp.concatenate([g, g]), sp.COO) and isinstance(sp.stack([g, g]), sp.COO)  # 1-D GCXS joins as COO
    g2 = _gcxs(sp, np.array([[0.0, 1.0], [2.0, 0.0]]))
    assert isinstance(sp.concatenate([g2, g2]), sp.GCXS) and isinstance(sp.stack([g2, g2]), sp.GCXS)
