#!/bin/bash
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:ew_merge_fused -c 1 -f -o gpurun_out/r02_ewmerge python tools/ncu_c3big.py > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
timeout 300 $NCU -k regex:rd_emit -c 1 -f -o gpurun_out/r02_rd_emit python tools/ncu_reduce_big.py > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_reduce_launches.csv python tools/ncu_reduce_big.py > gpurun_out/ncu3b.log 2>&1
tail -12 gpurun_out/r02_reduce_launches.csv | cut -c1-200
