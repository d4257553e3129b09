#!/bin/bash
# kernel iteration call: targeted GPU tests + per-config timings + ncu captures of the three kernels being tuned
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spgemm_gpu.py tests/test_spmm_gpu.py tests/test_large_scale_gpu.py tests/test_api_reduce.py tests/test_api_nanreduce.py tests/test_api_tensordot.py -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_new.log 2>&1
echo "pytest(new) rc=$?" >> gpurun_out/gpu_tests_new.log
tail -8 gpurun_out/gpu_tests_new.log
timeout 600 python tools/bench_configs.py ${B2S_CONFIGS:-c5 c3big} > gpurun_out/configs.log 2>&1
cat gpurun_out/configs.log | cut -c1-400
timeout 600 python tools/probe_skew.py > gpurun_out/skew.log 2>&1
cat gpurun_out/skew.log | tail -8
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:spgemm_rows_kernel -c 1 -f -o gpurun_out/r02_spgemm python tools/ncu_c5.py > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 300 $NCU -k regex:ew_merge_fused -c 1 -f -o gpurun_out/r02_ewmerge python tools/ncu_c3big.py > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 300 $NCU -k regex:reduce_tile -c 1 -f -o gpurun_out/r02_reduce python tools/ncu_reduce_big.py > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
