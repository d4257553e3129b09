"""Process-wide switches read from the environment once at import (upstream keeps the same two switches in
sparse/numba_backend/_settings.py).

SPARSE_AUTO_DENSIFY=1       np.asarray(x) on a sparse array densifies instead of raising RuntimeError
SPARSE_WARN_ON_TOO_DENSE=1  warn when a COO takes no less memory than the equivalent dense array
"""
import os

import numpy as np


def _flag(name: str) -> bool:
    value = os.environ.get(name, "").strip().lower()
    return value not in ("", "0", "false", "no", "off")


AUTO_DENSIFY = _flag("SPARSE_AUTO_DENSIFY")
WARN_ON_TOO_DENSE = _flag("SPARSE_WARN_ON_TOO_DENSE")
NEP18_ENABLED = True  # __array_function__ dispatch is always on in the NumPy versions this package supports
IS_NUMPY2 = np.lib.NumpyVersion(np.__version__) >= "2.0.0a1"
