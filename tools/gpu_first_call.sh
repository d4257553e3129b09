#!/bin/bash
# What to run with the first GPU call of a round (one B200, ~5 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_first_call.sh'
# = tools/gpu_call.sh: the GPU test-suite as the driver runs it (-x), the reference arm, the bench line with the configs
# block, smoke(), the static-vs-dynamic-rows A/B of K1 and the ncu evidence for profiles/.
# Multi-GPU (bench weak / strong / e2e with both gather transports + bench_multi over row blocks) is one call per N:
#   /usr/local/graft/bin/gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_multi.sh N'      (N = 2, 4, 8; charged N x)
# then `python tools/make_scaling_table.py r0X` collects gpurun_out/ into profiles/.
# Kernel iteration: tools/gpu_kernels.sh (targeted tests + tools/bench_configs.py, optional ncu / skew probe);
# compute-sanitizer: tools/gpu_sanitizer.sh.
exec bash "$(dirname "$0")/gpu_call.sh"
