#!/bin/bash
# One single-GPU call: the GPU test-suite (as the driver runs it), the reference arm, the bench line (with the configs
# block), and the ncu evidence for profiles/.  Usage: gpurun --timeout 1800 -- 'bash tools/gpu_call.sh'
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 600 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
# A/B of the two row-assignment forms of K1 (kernel-only): default = dynamic rows, --variant 3 = static grid
timeout 300 python bench.py --variant 3 --no-e2e --no-cpu --no-configs > gpurun_out/bench_n1_static_rows.json 2>/dev/null; tail -c 700 gpurun_out/bench_n1_static_rows.json | head -c 500; echo
if [ "${B2S_NCU:-1}" = "1" ]; then
# every launch of a short bench run with its device time (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:spmm_csr_dense_dyn -s 3 -c 1 -f -o gpurun_out/r02_k1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-configs > gpurun_out/ncu0.log 2>&1; tail -1 gpurun_out/ncu0.log
timeout 300 $NCU -k regex:spgemm_rows_kernel -c 1 -f -o gpurun_out/r02_spgemm python tools/ncu_c5.py > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_c5_launches.csv python tools/ncu_c5.py > gpurun_out/ncu1b.log 2>&1
fi
