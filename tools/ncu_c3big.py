#!/usr/bin/env python
"""One C3-large broadcast add and one large reduction for an ncu capture of the merge / reduce kernels."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc
a = bc.rand_coo((512, 512, 512, 64), 85_899_345, 10)
b = bc.rand_coo((512, 512, 512, 1), 1_342_177, 11)
for _ in range(2):
    out = a + b
torch.cuda.synchronize()
print(out.nnz)
