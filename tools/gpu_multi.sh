#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded bench at N ranks.   usage: tools/gpu_multi.sh N [steps]
set -u
N=${1:-2}; STEPS=${2:-300}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps $STEPS --warmup 10 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -1 gpurun_out/bench_n$N.log | cut -c1-2500; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -15
