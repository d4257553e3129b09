#!/usr/bin/env python
"""Timeline of the host-buffer pipeline (B2S_HOST_TRACE=1): when B lands, when each chunk lands / finishes / is back."""
import os, sys
os.environ["B2S_HOST_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sparse_b200 import _kernels as Kn
dev = torch.device("cuda", 0)
M = K = 1_000_000
vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, 128, 1234, dev)
h = [x.cpu().pin_memory().numpy() for x in (vals, cols.to(torch.int64), indptr.to(torch.int64), B)]
out = torch.empty((M, 128), dtype=torch.float32).pin_memory().numpy()
import time
def run(tag, args, env=None, reps=5):
    os.environ.pop("B2S_HOST_SKIP", None)
    if env:
        os.environ["B2S_HOST_SKIP"] = env
    os.environ.pop("B2S_HOST_TRACE", None)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); Kn.spmm_csr_dense_host(*args, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{tag}: min {min(ts):.2f} ms median {sorted(ts)[len(ts)//2]:.2f} ms", flush=True)
h32 = [h[0], torch.from_numpy(h[1]).to(torch.int32).pin_memory().numpy(),
       torch.from_numpy(h[2]).to(torch.int32).pin_memory().numpy(), h[3]]
Kn.spmm_host_set_threads(8)
run("int64 in, host narrow 8 threads", h)
run("int32 in (no host work, same bytes)", h32)
run("int32 in, no D2H", h32, "d2h")
run("int32 in, no K1", h32, "k1")
run("int32 in, no K1, no D2H (pure H2D)", h32, "k1,d2h")
run("int64 host narrow, no K1, no D2H", h, "k1,d2h")
Kn.spmm_host_set_threads(0)
run("int64 raw upload, device narrow", h)
run("int64 raw upload, no K1 no D2H", h, "k1,d2h")
for chunks in (4, 8, 32):
    Kn.spmm_host_set_pipeline(chunks, 4)
    run(f"int32 in, chunks={chunks}", h32)
