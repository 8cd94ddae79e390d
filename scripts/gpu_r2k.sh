#!/bin/bash
# round 2, call K: hunt the rare slow decode_files call (device outputs, 8 CPUs) with the trace counters
OUT=gpurun_out/r2k
mkdir -p $OUT
for rep in 1 2 3; do
  B200JPG_TRACE=1 taskset -c 0-7 timeout 300 python scripts/files_bench.py --dev-out --threads 8 --reps 30 --tag "hunt-$rep" 2>$OUT/trace_$rep.err | cut -c1-420
  grep decode_files $OUT/trace_$rep.err | awk '{ if ($8+0 > 40) print }' | cut -c1-400
done
B200JPG_TRACE=1 timeout 300 python scripts/files_bench.py --dev-out --reps 30 --tag "hunt-all" 2>$OUT/trace_all.err | cut -c1-420
grep decode_files $OUT/trace_all.err | awk '{ if ($8+0 > 40) print }' | cut -c1-400
