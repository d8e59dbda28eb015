#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, short bench, ncu launch list, ncu full capture of the online kernels.
# Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [ncu] [full]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -25 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 400 python bench.py --steps 500 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -2 gpurun_out/bench.log ; tail -5 gpurun_out/bench.err
echo "== bench reference" ; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err ; echo "ref rc=$?" ; tail -1 gpurun_out/bench_ref.log
for a in "$@"; do
  if [ "$a" = "ncu" ]; then
    echo "== ncu launches"
    timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
       python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
  fi
  if [ "$a" = "full" ]; then
    echo "== ncu full (online kernels)"
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lstm|retrieve|tokenize" --launch-skip 25 --launch-count 10 \
       -f -o gpurun_out/online_full python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
  fi
done
