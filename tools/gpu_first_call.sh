#!/bin/bash
# What to run with the first GPU call of a round (everything the last round-1 session could not measure):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# 1. the full GPU test-suite (the tests/test_widen_zz_*.py files have only run on the NumPy mock),
# 2. the headline bench line, 3. the per-config timings.  Multi-GPU (C4 / C5 over row blocks) is a second call:
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29511 tools/bench_multi.py > gpurun_out/multi_2.log 2>&1'
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.json
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -20 gpurun_out/configs.log
