#!/bin/bash
# kernel iteration call: targeted GPU tests + per-config timings (+ optional ncu captures: B2S_NCU=1)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spgemm_gpu.py tests/test_spmm_gpu.py tests/test_large_scale_gpu.py tests/test_api_reduce.py tests/test_api_elemwise.py tests/test_api_tensordot.py -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests_new.log 2>&1
echo "pytest(new) rc=$?" >> gpurun_out/gpu_tests_new.log
tail -8 gpurun_out/gpu_tests_new.log
timeout 600 python tools/bench_configs.py ${B2S_CONFIGS:-c5 c3big} > gpurun_out/configs.log 2>&1
cat gpurun_out/configs.log | cut -c1-400
if [ "${B2S_SKEW:-0}" = "1" ]; then
timeout 600 python tools/probe_skew.py > gpurun_out/skew.log 2>&1
cat gpurun_out/skew.log | tail -8
fi
if [ "${B2S_NCU:-0}" = "1" ]; then
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:spgemm_rows_kernel -c 1 -f -o gpurun_out/r02_spgemm python tools/ncu_c5.py > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
fi
