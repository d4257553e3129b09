#!/usr/bin/env python
"""Where does the time go on power-law matrices?  Sweep the row-length clamp with the long-row path on / off."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparse_b200 import _kernels as Kn  # noqa: E402
from sparse_b200 import _lib  # noqa: E402
from tools.tune_k1 import timeit  # noqa: E402

dev = torch.device("cuda", 0)
M = K = 1_000_000
nnz = 100_000_000
g = torch.Generator(device=dev).manual_seed(7)
B = torch.rand((K, 128), generator=g, device=dev, dtype=torch.float32)
C = torch.empty((M, 128), dtype=torch.float32, device=dev)
u = torch.rand(M, generator=g, device=dev, dtype=torch.float64)
raw = (u.pow(-1 / 1.5) * 20)
lib = _lib.load()
for clamp in (500, 2000, 4000, 8000, 50000):
    lens = raw.clamp(max=clamp).to(torch.int64)
    lens = (lens.double() * (nnz / lens.sum().item())).to(torch.int64).clamp(min=0, max=K)
    indptr = torch.zeros(M + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(lens, 0)
    n2 = int(indptr[-1].item())
    cols = torch.randint(0, K, (n2,), generator=g, device=dev, dtype=torch.int32)
    vals = torch.rand(n2, generator=g, device=dev, dtype=torch.float32)
    ip32 = indptr.to(torch.int32)
    out = {}
    for skew in (0, 1):
        out[skew] = timeit(lambda: Kn.spmm_csr_dense(vals, cols, ip32, B, M, K, 128, out=C, long_rows=bool(skew)),
                           reps=5)
    print(f"clamp {clamp:6d} max_len {int(lens.max().item()):6d} rows>4096 {int((lens > 4096).sum().item()):6d} "
          f"nnz_in_long {int(lens[lens > 4096].sum().item()):9d}  skew_off {out[0]:.3f} ms  skew_on {out[1]:.3f} ms", flush=True)
