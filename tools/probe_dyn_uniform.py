#!/usr/bin/env python
"""Does the dynamic-row variant of K1 (nnz-balanced mode) help on the UNIFORM C2 matrix too?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sparse_b200 import _kernels as Kn, _lib
from tools.tune_k1 import timeit

dev = torch.device("cuda", 0)
_lib.load()
M = K = 1_000_000
vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, 128, 1234, dev, b_seed=4321)
C = torch.empty((M, 128), dtype=torch.float32, device=dev)
C2 = torch.empty_like(C)
for flag in (False, True, False, True):
    ms = timeit(lambda: Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128, out=C if not flag else C2, long_rows=flag), reps=10)
    print(f"uniform C2, long_rows={flag}: {ms:.3f} ms", flush=True)
print("bit-identical:", bool(torch.equal(C, C2)))
