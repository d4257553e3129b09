#!/usr/bin/env python
"""One large trailing-axis reduction for an ncu capture of reduce_tile_kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc
a = bc.rand_coo((512, 512, 512, 64), 85_899_345, 10)
for _ in range(2):
    out = a.sum(axis=3)
torch.cuda.synchronize()
print(out.nnz)
