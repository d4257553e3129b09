#!/usr/bin/env python
"""Break down the end-to-end (host buffer) K1 path: raw PCIe rates vs the pipelined ABI call vs the public API."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import sparse_b200 as sp
from sparse_b200 import _kernels as Kn, _device as D

dev = torch.device("cuda", 0)
M = K = 1_000_000
vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, 128, 1234, dev)
h_vals = vals.cpu().pin_memory(); h_cols = cols.cpu().to(torch.int64).pin_memory(); h_ptr = indptr.cpu().to(torch.int64).pin_memory()
h_B = B.cpu().pin_memory(); h_C = torch.empty((M, 128), dtype=torch.float32).pin_memory()
def t(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.median(ts))
d_buf = torch.empty(h_cols.numel(), dtype=torch.int64, device=dev)
print("H2D 0.8GB pinned ms", t(lambda: d_buf.copy_(h_cols, non_blocking=True)))
d_c = torch.empty((M, 128), dtype=torch.float32, device=dev)
print("D2H 0.512GB pinned ms", t(lambda: h_C.copy_(d_c, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_buf.copy_(h_cols, non_blocking=True)
    with torch.cuda.stream(s2): h_C.copy_(d_c, non_blocking=True)
print("H2D 0.8GB + D2H 0.512GB concurrent ms", t(both))
npv = (h_vals.numpy(), h_cols.numpy(), h_ptr.numpy(), h_B.numpy(), h_C.numpy())
print("ABI host call (int64 idx) ms", t(lambda: Kn.spmm_csr_dense_host(*npv[:4], out=npv[4])))
c32 = h_cols.to(torch.int32).pin_memory(); p32 = h_ptr.to(torch.int32).pin_memory()
print("ABI host call (int32 idx) ms", t(lambda: Kn.spmm_csr_dense_host(npv[0], c32.numpy(), p32.numpy(), npv[3], out=npv[4])))
def api():
    A = sp.GCXS((npv[0], npv[1], npv[2]), shape=(M, K), compressed_axes=(0,))
    return sp.tensordot(A, npv[3], axes=1)
print("public API ms", t(api))
print("pinned_empty ms", t(lambda: D.pinned_empty((M, 128), np.float32)))
