#!/usr/bin/env python
"""Two C5 products (CSR(1e6 x 1e6, density 1e-5) squared, fp32) for an ncu capture of the SpGEMM kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc
import sparse_b200 as sp
A = bc.rand_csr(1_000_000, 1_000_000, 10_000_000, 3, torch.float32)
for _ in range(2):
    out = sp.tensordot(A, A, axes=1)
torch.cuda.synchronize()
print(out.nnz)
