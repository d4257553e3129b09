#!/usr/bin/env python
"""K1 at C2: the forms of the dynamic-row kernel (b2s_spmm_set_variant 10..13) against the static grid (3)."""
import json, os, sys
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
ref = None
res = {}
names = {10: "dynamic rows, U=8, 4 CTAs/SM (default)", 11: "dynamic rows, U=16, 4 CTAs/SM", 12: "dynamic rows, U=8, 5 CTAs/SM",
         13: "dynamic rows, U=16, 3 CTAs/SM", 3: "static one-row-per-warp grid, U=8 (round 1)"}
for rep in range(2):
    for v in (10, 11, 12, 13, 3):
        Kn.spmm_set_variant(v, 8)
        ms = timeit(lambda: Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128, out=C), reps=10)
        same = True if ref is None else bool(torch.equal(C, ref))
        if ref is None:
            ref = C.clone()
        res.setdefault(names[v], []).append(round(ms, 4))
        print(f"{names[v]:50s} {ms:.3f} ms  bit-identical={same}", flush=True)
Kn.spmm_set_variant(0, 8)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_k1_dyn.json"), "w"), indent=1)
