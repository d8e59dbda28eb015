# knock-out / geometry sweep of the scan kernel at the sharded shape (ncu durations; diagnostic).  The knock-outs exist only in the
# trace build: run tools/make_scan_trace.py first.
export T2P_DIAG_LIB=tools/bin/libexp_SCT.so
for cap in ${CAP_LIST:-40}; do
for d in ${KO_LIST:-0 1 2 3}; do
  T2P_SCAN_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"retrieve_scan" --launch-skip 20 -c 2 --csv --log-file gpurun_out/ko_$d.csv python tools/diag_shard_shape.py 8:12:4:$cap:::n:2 > gpurun_out/ko_$d.log 2>&1
  echo "cap=$cap debug=$d" $(grep -o "gpu__time_duration.sum\",\"[a-z]*\",\"[0-9.]*\|grid_size\",\"\",\"[0-9]*" gpurun_out/ko_$d.csv | tr '\n' ' ')
done
done
