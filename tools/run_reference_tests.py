#!/usr/bin/env python
"""Run the REFERENCE's own test files against sparse_b200 (drop-in check of the host layer).

Authoring-container tool (needs /root/reference; the GPU box never runs it):

    python tools/run_reference_tests.py test_dot.py [--patch OLD=NEW] [-k expr] [pytest args...]

A throw-away shim package named ``sparse`` is written to a temp dir; it re-exports sparse_b200 under the names the
reference's tests import (``sparse``, ``sparse.numba_backend._utils.assert_eq`` ...).  The reference's test helpers
(``_utils.py``: assert_eq & co) are executed from where they lie under /root/reference -- nothing is copied into this
repository.  On a box without a GPU the kernel layer is the NumPy mock of tests/_mock_kernels.py (host-logic check); with
a GPU the real CUDA kernels run.  The output is a gap list: every failure is an API or semantics difference on (or next
to) the hot path.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SPARSE_REFERENCE", "/root/reference")

SHIM_INIT = r'''
import importlib.util, os, sys, types
ROOT = {root!r}
REF = {ref!r}
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
if not torch.cuda.is_available():
    import _mock_kernels
    _mock_kernels.install()
import numpy as np
from numpy import *  # noqa: F401,F403  (the ufunc / dtype names the namespace re-exports)
import sparse_b200 as _m
from sparse_b200 import *  # noqa: F401,F403
from sparse_b200 import sum, max, min, prod, mean, any, all, random, COO, GCXS  # noqa: F401
from sparse_b200 import _coo as _coo_mod, _gcxs as _gcxs_mod


DOK = _m.DOK


def __getattr__(name):
    return getattr(_m, name)


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


nb = _module("sparse.numba_backend")
nb.__path__ = []
for k in dir(_m):
    if not k.startswith("__"):
        setattr(nb, k, getattr(_m, k))
_coo_pkg = _module("sparse.numba_backend._coo", COO=COO, as_coo=_m.as_coo)
_coo_pkg.__path__ = []
_module("sparse.numba_backend._coo.core", COO=COO, as_coo=_m.as_coo)
_cmp = _module("sparse.numba_backend._compressed", GCXS=GCXS, CSR=_m.CSR, CSC=_m.CSC)
_cmp.__path__ = []
_module("sparse.numba_backend._compressed.compressed", GCXS=GCXS, CSR=_m.CSR, CSC=_m.CSC)
nb._compressed = _cmp
_compressed = _cmp
numba_backend = nb
from sparse_b200 import _settings as _settings_mod
sys.modules["sparse.numba_backend._settings"] = _settings_mod
nb._settings = _settings_mod
nb.DOK = DOK
_module("sparse.numba_backend._dok", DOK=DOK)
_module("sparse.numba_backend._sparse_array", SparseArray=_m.SparseArray)
from sparse_b200 import _elemwise as _elemwise_mod
sys.modules["sparse.numba_backend._umath"] = _elemwise_mod
nb._umath = _elemwise_mod
# the reference's own test helpers, executed in place (relative imports resolve to the shim modules above)
spec = importlib.util.spec_from_file_location("sparse.numba_backend._utils",
                                              os.path.join(REF, "sparse", "numba_backend", "_utils.py"))
_utils = importlib.util.module_from_spec(spec)
sys.modules["sparse.numba_backend._utils"] = _utils
spec.loader.exec_module(_utils)
_utils.random = _m.random
nb._utils = _utils
'''


def main(argv):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: this tool only runs in the authoring container")
    tmp = tempfile.mkdtemp(prefix="b2s_refshim_")
    os.makedirs(os.path.join(tmp, "sparse"))
    with open(os.path.join(tmp, "sparse", "__init__.py"), "w") as f:
        f.write(SHIM_INIT.replace("{root!r}", repr(ROOT)).replace("{ref!r}", repr(REF)))
    # the test files are staged in the scratch dir (never in this repository) so that pytest does not pick up the
    # reference's package-level conftest / __init__ chain, which would import the reference itself
    tdir = os.path.join(tmp, "reftests")
    os.makedirs(tdir)
    with open(os.path.join(tdir, "conftest.py"), "w") as f:
        f.write("import pytest\n\n\n@pytest.fixture(scope='session')\ndef rng():\n"
                "    from sparse.numba_backend._utils import default_rng\n    return default_rng\n")
    files, rest, patches = [], [], []
    it = iter(argv)
    for a in it:
        cand = os.path.join(REF, "sparse", "numba_backend", "tests", a)
        if a == "--patch":  # --patch OLD=NEW : textual substitution in the staged copy (to get past a collection error)
            patches.append(next(it).split("=", 1))
        elif os.path.exists(cand):
            shutil.copy(cand, os.path.join(tdir, a))
            files.append(os.path.join(tdir, a))
        else:
            rest.append(a)
    for path in files:
        text = open(path).read()
        for old, new in patches:
            text = text.replace(old, new)
        open(path, "w").write(text)
    if not files:
        raise SystemExit(__doc__)
    sys.path.insert(0, tmp)
    import pytest

    return pytest.main(["-c", os.devnull, "-p", "no:cacheprovider", "--rootdir", tmp, "-q", *files, *rest])


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
