#!/bin/bash
# compute-sanitizer over the GPU tests of the kernels written in round 2 (SpGEMM row kernel, long-row ring kernel and
# dynamic-row K1, both reduction forms): memcheck on all of them, racecheck (shared-memory hazards: the atomics-free hash
# table relies on __syncwarp) on the SpGEMM and reduction tests.
set -u
mkdir -p gpurun_out
T="tests/test_spgemm_gpu.py tests/test_spmm_gpu.py tests/test_api_reduce.py"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest $T -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log; tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_spgemm_gpu.py "tests/test_large_scale_gpu.py::test_both_reduction_kernels_agree" "tests/test_spmm_gpu.py::test_long_rows_take_the_column_split_kernel" -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log; tail -4 gpurun_out/r02_sanitizer_racecheck.log
