#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): multi-GPU tests, then the sharded bench at N ranks at the driver's settings.
#   usage: tools/gpu_multi.sh N [extra bench args...]
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
echo "== pytest multi" ; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
run() {  # tag, args...
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n${N}_$tag.log 2> gpurun_out/bench_n${N}_$tag.err
  echo "== $tag rc=$?"; tail -1 gpurun_out/bench_n${N}_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print({k:d.get(k) for k in ('value','ms_per_step','reps','timed_region_s','parity_vs_oracle_top10')}, 'e2e', round(d['e2e']['value']), 'lstm', d['roofline']['ms'], 'other', d['roofline_other']['ms'], 'nccl', d.get('nccl_exchange'), 'rows', d.get('rows'))
except Exception as e:
    print('no line', e)
"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n${N}_$tag.err | tail -8
}
run default "$@"
