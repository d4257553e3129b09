"""Process-wide switches read from the environment (sparse/numba_backend/_settings.py:1-8 upstream).

SPARSE_AUTO_DENSIFY=1       np.asarray(x) on a sparse array densifies instead of raising RuntimeError
SPARSE_WARN_ON_TOO_DENSE=1  warn when a COO takes no less memory than the equivalent dense array
"""
import os

AUTO_DENSIFY = bool(int(os.environ.get("SPARSE_AUTO_DENSIFY", "0")))
WARN_ON_TOO_DENSE = bool(int(os.environ.get("SPARSE_WARN_ON_TOO_DENSE", "0")))
NEP18_ENABLED = True
