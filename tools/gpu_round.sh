#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, DSMEM micro-benchmark, bench at the driver's settings (+ a long run), reference arm,
# optional ncu launch list / full capture.  Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [ncu] [full] [rows]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
echo "== dsmem" ; timeout 120 tools/bin/dsmem_bench > gpurun_out/dsmem_bench.json 2> gpurun_out/dsmem_bench.err ; echo "dsmem rc=$?" ; cut -c1-400 gpurun_out/dsmem_bench.json ; tail -3 gpurun_out/dsmem_bench.err
echo "== bench (driver settings)" ; timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" ; tail -1 gpurun_out/bench.log | cut -c1-6000 ; tail -5 gpurun_out/bench.err
echo "== bench reference" ; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err ; echo "ref rc=$?" ; tail -1 gpurun_out/bench_ref.log | cut -c1-1500
for a in "$@"; do
  if [ "$a" = "rows" ]; then
    echo "== rows" ; timeout 400 python tools/bench_rows.py > gpurun_out/rows.jsonl 2> gpurun_out/rows.err ; echo "rows rc=$?" ; cat gpurun_out/rows.jsonl | cut -c1-900
  fi
  if [ "$a" = "ncu" ]; then
    echo "== ncu launches"
    timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
       python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_bench.log 2>&1 ; echo "ncu rc=$?"
  fi
  if [ "$a" = "full" ]; then
    echo "== ncu full (online kernels)"
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:lstm|retrieve|tokenize" --launch-skip 25 --launch-count 10 \
       -f -o gpurun_out/online_full python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_full.log 2>&1 ; echo "ncu full rc=$?"
  fi
done
