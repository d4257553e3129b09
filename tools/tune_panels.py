#!/usr/bin/env python
"""Sweep the number of column panels of K1p on the C2 workload; writes gpurun_out/tune_panels.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sparse_b200 import _kernels as Kn
from tools.tune_k1 import timeit

dev = torch.device("cuda", 0)
M = K = 1_000_000
vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, 128, 1234, dev)
C = torch.empty((M, 128), dtype=torch.float32, device=dev)
peak, _ = bench.peaks()
alg = bench.algorithmic_bytes(int(vals.numel()), M, 128)
ref = Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128).clone()
res = []
for p in (1, 2, 4, 6, 8, 10, 12, 16, 24, 32, 0):
    ms = timeit(lambda: Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128, out=C, n_panels=p, rows_sorted=True))
    same = bool(torch.equal(C, ref))
    r = {"n_panels": p, "ms": round(ms, 4), "gnnz_s": round(vals.numel() / ms / 1e6, 3), "frac": round(alg / ms / 1e6 / peak, 4),
         "bit_identical_to_one_pass": same}
    print(r, flush=True)
    res.append(r)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_panels.json"), "w"), indent=1)
