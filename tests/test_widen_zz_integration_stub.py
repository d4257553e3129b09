"""The ctypes stub shown in INTEGRATION.md (what a maintainer would add to the reference) is real code: it is cut out of
the document, checked against the exported ABI on any box, and on a GPU box executed against the library and compared
bit for bit with the oracle's restatement of `_dot_csr_ndarray` (_common.py:720-755)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# sparse/numba_backend/_b200\.py.*?)```", text, flags=re.S)
    assert m, "INTEGRATION.md lost its ctypes stub"
    return m.group(1)


def test_stub_compiles_and_names_exported_symbols():
    from sparse_b200 import _lib

    src = _stub_source()
    compile(src, "INTEGRATION.md:_b200.py", "exec")
    lib = _lib.load()
    for sym in set(re.findall(r"_lib\.(b2s_[a-z0-9_]+)", src)):
        assert hasattr(lib, sym), sym
    header = open(os.path.join(ROOT, "include", "sparse_b200.h")).read()
    proto = re.search(r"int b2s_spmm_csr_dense_host\((.*?)\);", header, flags=re.S).group(1)
    call = re.search(r"b2s_spmm_csr_dense_host\(\n(.*?)\)\)\n", src, flags=re.S).group(1)
    n_args_call = len([a for a in re.split(r",\s*(?![^()]*\))", re.sub(r"#.*", "", call)) if a.strip()])
    assert n_args_call == len(proto.split(",")), (n_args_call, proto)


@pytest.mark.gpu
def test_stub_runs_and_matches_the_oracle():
    import oracle
    from sparse_b200 import _lib

    src = _stub_source().replace('ctypes.CDLL("libsparse_b200.so")', f"ctypes.CDLL({_lib.LIB_PATH!r})")
    ns = {}
    exec(compile(src, "INTEGRATION.md:_b200.py", "exec"), ns)
    rng = np.random.default_rng(5)
    M, K, N = 3000, 4000, 64
    lin = np.unique(rng.integers(0, M * K, size=60_000, dtype=np.int64))
    rows, cols = lin // K, lin % K
    indptr = np.zeros(M + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=indptr[1:])
    for dt in (np.float32, np.float64):
        data = rng.random(len(lin)).astype(dt)
        b = rng.random((K, N)).astype(dt)
        got = ns["_dot_csr_ndarray_type"](dt, dt)((M, N), data, cols, indptr, b)
        want = oracle.dot_csr_ndarray((M, N), data, cols, indptr, b)
        assert got.dtype == want.dtype and np.array_equal(got.view(np.uint8), want.view(np.uint8))
