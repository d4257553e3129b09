#!/bin/bash
# Multi-GPU measurements on ONE box:  gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh N'
# 1. bench.py at N GPUs (weak scaling headline + strong block + e2e), copy-engine gather of B
# 2. the same with the NCCL all-gather (kernel-only, for the A/B of the gather transport)
# 3. tools/bench_multi.py: C5 (SpGEMM), C3-large (elemwise + reductions), C4 (SDDMM) over N row blocks
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "1" ]; then RUN1="python"; else RUN1="$RUN --master-port 29501"; fi
timeout 900 $RUN1 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-cpu > gpurun_out/bench_n${N}_ce.json 2> gpurun_out/bench_n${N}_ce.err
echo "bench ce rc=$?"; tail -c 3000 gpurun_out/bench_n${N}_ce.json; tail -3 gpurun_out/bench_n${N}_ce.err
if [ "$N" != "1" ]; then
  timeout 600 $RUN --master-port 29502 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl --no-e2e --no-strong > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
  echo "bench nccl rc=$?"; tail -c 1500 gpurun_out/bench_n${N}_nccl.json; tail -3 gpurun_out/bench_n${N}_nccl.err
fi
if [ "$N" = "1" ]; then RUN3="python"; else RUN3="$RUN --master-port 29503"; fi
timeout 900 $RUN3 tools/bench_multi.py c5 c5k c3 c4 > gpurun_out/multi_${N}.log 2>&1
echo "bench_multi rc=$?"; tail -12 gpurun_out/multi_${N}.log | cut -c1-900
