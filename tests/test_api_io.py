"""save_npz / load_npz against files written by the reference itself (tests/golden/ref_saved_*.npz, made by
make_golden.py:gen_io) and round trips; the node names and dtypes written match the reference's byte for byte."""
import io
import os

import numpy as np
import pytest

from _api import sp  # noqa: F401

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_load_reference_coo_file(sp):
    x = sp.load_npz(os.path.join(HERE, "ref_saved_coo.npz"))
    with np.load(os.path.join(HERE, "ref_saved_coo.npz")) as fp:
        assert isinstance(x, sp.COO) and x.shape == tuple(fp["shape"]) and x.fill_value == fp["fill_value"][()]
        assert np.array_equal(x.coords, fp["coords"]) and np.array_equal(x.data, fp["data"])
    # the loaded array is usable on the data path: a reduction matches NumPy on the dense form
    assert np.allclose(x.sum(axis=0).todense(), x.todense().sum(axis=0), rtol=1e-12)


def test_load_reference_gcxs_file(sp):
    g = sp.load_npz(os.path.join(HERE, "ref_saved_gcxs.npz"))
    with np.load(os.path.join(HERE, "ref_saved_gcxs.npz")) as fp:
        assert isinstance(g, sp.GCXS) and g.compressed_axes == tuple(fp["compressed_axes"])
        assert np.array_equal(g.indices, fp["indices"]) and np.array_equal(g.indptr, fp["indptr"])
        assert np.array_equal(g.data, fp["data"]) and g.dtype == np.float32
    assert np.array_equal(g.tocoo().todense(), g.todense())


@pytest.mark.parametrize("compressed", [True, False])
@pytest.mark.parametrize("fmt", ["coo", "gcxs"])
def test_round_trip_and_node_layout(sp, fmt, compressed):
    x = sp.random((4, 5, 6), density=0.3, random_state=11, format=fmt)
    buf = io.BytesIO()
    sp.save_npz(buf, x, compressed=compressed)
    buf.seek(0)
    ref = np.load(os.path.join(HERE, f"ref_saved_{fmt}.npz"))
    mine = np.load(io.BytesIO(buf.getvalue()))
    assert set(mine.files) == set(ref.files)
    for k in ref.files:
        assert mine[k].dtype.kind == ref[k].dtype.kind and mine[k].ndim == ref[k].ndim, k
    y = sp.load_npz(buf)
    assert type(y) is type(x) and y.shape == x.shape
    assert np.array_equal(y.todense(), x.todense())


def test_load_invalid_file_raises(sp, tmp_path):
    p = tmp_path / "bad.npz"
    np.savez(p, a=np.arange(3))
    with pytest.raises(RuntimeError):
        sp.load_npz(p)


# ---- SciPy interop (tests/test_coo.py:1113-1172, tests/test_compressed.py upstream) ---------------------------------------
def test_scipy_round_trips(sp):
    import scipy.sparse as ss

    rng = np.random.default_rng(4)
    m = ss.random(30, 40, density=0.1, format="csr", random_state=rng, dtype=np.float64)
    x = sp.COO.from_scipy_sparse(m)
    assert np.array_equal(x.todense(), m.toarray())
    back = x.to_scipy_sparse()
    assert back.format == "coo" and (abs(back - m)).nnz == 0
    csr, csc = x.tocsr(), x.tocsc()
    assert csr.format == "csr" and csc.format == "csc"
    assert np.array_equal(csr.indptr, m.indptr) and np.array_equal(csr.indices, m.indices)
    assert np.array_equal(csr.data, m.data) and np.array_equal(csc.toarray(), m.toarray())
    for ca, fmt in (((0,), "csr"), ((1,), "csc")):
        g = sp.GCXS.from_scipy_sparse(m.asformat(fmt))
        assert g.compressed_axes == ca
        s = g.to_scipy_sparse()
        assert s.format == fmt and np.array_equal(s.toarray(), m.toarray())
    # scipy operands are accepted by the products directly (_common.py:128-131)
    d = rng.random((40, 5))
    assert np.allclose(sp.tensordot(x, d, axes=1), m @ d, rtol=1e-12)
    assert np.allclose(sp.matmul(m, sp.COO.from_numpy(d)).todense(), m @ d, rtol=1e-12)


def test_scipy_export_checks(sp):
    x = sp.random((3, 4, 5), density=0.5, random_state=1)
    nd = x.to_scipy_sparse()  # n-D coo_array (SciPy >= 1.13), like upstream (_coo/core.py:1166-1198)
    assert nd.shape == (3, 4, 5) and nd.nnz == x.nnz
    with pytest.raises(ValueError):
        x.tocsr()
    y = sp.random((3, 4), density=0.5, random_state=1, fill_value=1.0)
    with pytest.raises(ValueError):
        y.to_scipy_sparse()
    assert y.to_scipy_sparse(accept_fv=1.0).shape == (3, 4)
