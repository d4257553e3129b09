#!/bin/bash
# One single-GPU call: the GPU test-suite (as the driver runs it), the bench line (with the configs block), the
# reference arm.  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'
set -u
mkdir -p gpurun_out
# new kernels first, under a short timeout (a hang must not eat the call)
timeout 300 python -m pytest tests/test_spgemm_gpu.py tests/test_large_scale_gpu.py -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_new.log 2>&1
echo "pytest(new) rc=$?" >> gpurun_out/gpu_tests_new.log
tail -15 gpurun_out/gpu_tests_new.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 6000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
