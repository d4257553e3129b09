#!/usr/bin/env python
"""Host-side index narrowing in b2s_spmm_csr_dense_host: sweep the thread count (0 = device-side narrowing of the raw
int64 upload) at C2 and check the result stays bit-identical to the device-resident kernel."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from sparse_b200 import _kernels as Kn

dev = torch.device("cuda", 0)
M = K = 1_000_000
vals, cols, indptr, B = bench.make_workload(torch, M, K, 100_000_000, 128, 1234, dev)
ref = Kn.spmm_csr_dense(vals, cols, indptr, B, M, K, 128).cpu().numpy()
h_vals = vals.cpu().pin_memory(); h_cols = cols.cpu().to(torch.int64).pin_memory(); h_ptr = indptr.cpu().to(torch.int64).pin_memory()
h_B = B.cpu().pin_memory(); h_C = torch.empty((M, 128), dtype=torch.float32).pin_memory()
npv = (h_vals.numpy(), h_cols.numpy(), h_ptr.numpy(), h_B.numpy(), h_C.numpy())
def t(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3), round(float(np.median(ts)), 3)
out = {"hardware_concurrency": os.cpu_count()}
dst = torch.empty(h_cols.numel(), dtype=torch.int32).pin_memory().numpy()
for n in (1, 4, 8, 16, 32):
    Kn.spmm_host_set_threads(n)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); Kn.host_narrow(npv[1], dst); ts.append((time.perf_counter() - t0) * 1e3)
    out[f"narrow_only_threads={n}"] = round(min(ts), 3)
    print("host narrow 1e8 elements, threads", n, "ms", round(min(ts), 3), flush=True)
for chunks, slots in ((16, 4), (16, 16), (32, 8), (32, 32), (64, 16), (8, 8)):
    Kn.spmm_host_set_threads(8)
    Kn.spmm_host_set_pipeline(chunks, slots)
    ms = t(lambda: Kn.spmm_csr_dense_host(*npv[:4], out=npv[4]))
    out[f"chunks={chunks},slots={slots},threads=8"] = ms
    print("chunks", chunks, "slots", slots, ms, flush=True)
Kn.spmm_host_set_pipeline(16, 4)
for n in (0, 8, 16):
    Kn.spmm_host_set_threads(n)
    h_C.zero_()
    ms = t(lambda: Kn.spmm_csr_dense_host(*npv[:4], out=npv[4]))
    same = bool(np.array_equal(npv[4].view(np.uint32), ref.view(np.uint32)))
    out[f"threads={n}"] = {"min_ms": ms[0], "median_ms": ms[1], "bit_identical": same}
    print(n, ms, same, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "host_narrow.json"), "w"), indent=1)
