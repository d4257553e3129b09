#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_spgemm_gpu.py "tests/test_large_scale_gpu.py::test_both_reduction_kernels_agree" "tests/test_spmm_gpu.py::test_long_rows_take_the_column_split_kernel" -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log; tail -4 gpurun_out/r02_sanitizer_racecheck.log
timeout 300 python tools/bench_configs.py c5 2>&1 | tail -3 | cut -c1-250
