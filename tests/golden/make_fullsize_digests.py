#!/usr/bin/env python
"""Full-size golden DIGESTS from the reference itself (authoring container only: needs /root/reference + numba).

    python tests/golden/make_fullsize_digests.py        -> tests/golden/fullsize_digests.json

The outputs of BASELINE.json's config 3 (1.7e6 entries x 40 B) are too large to commit, so the fixture stores, per
case, nnz and a SHA-256 over the reference's result (int64 coordinates + value bit patterns) for the ops whose values
are order-independent (one IEEE operation per output: add, multiply, maximum, in float64 and float32), and for the
reductions of the same tensor (whose summation order NumPy's reduceat leaves unspecified) the exact coordinates digest,
the total, and every k-th value (1024 samples) for a tolerance comparison.  Inputs come from the plain-NumPy recipe in
fullsize_inputs.py, fed to the reference's COO constructor; the GPU tests feed the same arrays to sparse_b200.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fullsize_inputs as FI  # noqa: E402
from make_golden import sparse  # noqa: E402  (imports the reference from a scratch copy)


def main():
    out = {"reference": "pydata/sparse @ /root/reference (numba backend)", "cases": {}}
    for dt in (np.float64, np.float32):
        (ca, da), (cb, db) = FI.c3_inputs(dt)
        a = sparse.COO(ca, da, shape=FI.C3_SHAPE_A, has_duplicates=False, sorted=True)
        b = sparse.COO(cb, db, shape=FI.C3_SHAPE_B, has_duplicates=False, sorted=True)
        name = np.dtype(dt).name
        out["cases"][f"c3_inputs_{name}"] = {"a_nnz": int(a.nnz), "b_nnz": int(b.nnz),
                                             "a": FI.coo_digest(a.coords, a.data), "b": FI.coo_digest(b.coords, b.data)}
        for f in (np.add, np.multiply, np.maximum):
            t0 = time.perf_counter()
            r = f(a, b)
            dt_s = time.perf_counter() - t0
            assert isinstance(r, sparse.COO) and r.dtype == dt
            out["cases"][f"c3_{f.__name__}_{name}"] = {"nnz": int(r.nnz), "sha256": FI.coo_digest(r.coords, r.data),
                                                       "reference_seconds": round(dt_s, 3)}
            print(f.__name__, name, r.nnz, round(dt_s, 3), flush=True)
        if dt == np.float64:
            for axis in ((3,), (0, 1), (0,)):
                for red in ("sum", "max"):
                    r = getattr(a, red)(axis=axis)
                    step = max(1, r.nnz // 1024)
                    out["cases"][f"c3_{red}_axis{''.join(map(str, axis))}_float64"] = {
                        "nnz": int(r.nnz), "coords_sha256": FI.digest(np.asarray(r.coords, dtype=np.int64)),
                        "sample_step": step, "values_sample": [float(v) for v in r.data[::step]],
                        "total": float(np.sum(r.data))}
                    print(red, axis, r.nnz, flush=True)
    with open(os.path.join(HERE, "fullsize_digests.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
