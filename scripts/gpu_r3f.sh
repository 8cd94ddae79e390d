#!/bin/bash
# round 2, call 3F: larger groups when the download queue is deep (host outputs) -- A/B on one box
OUT=gpurun_out/r3f
mkdir -p $OUT
for r in 1 2 3; do
B200JPG_TRACE=1 python scripts/files_bench.py --reps 8 --tag relaxed 2>$OUT/trace_relaxed.err | cut -c1-200 | tee -a $OUT/ab.jsonl
B200JPG_GROUP_RELAX=0 B200JPG_TRACE=1 python scripts/files_bench.py --reps 8 --tag eager 2>$OUT/trace_eager.err | cut -c1-200 | tee -a $OUT/ab.jsonl
done
tail -1 $OUT/trace_relaxed.err | cut -c1-250; tail -1 $OUT/trace_eager.err | cut -c1-250
taskset -c 0-3 python scripts/files_bench.py --reps 6 --threads 4 --tag relaxed-4cpu | cut -c1-200 | tee -a $OUT/ab.jsonl
B200JPG_GROUP_RELAX=0 taskset -c 0-3 python scripts/files_bench.py --reps 6 --threads 4 --tag eager-4cpu | cut -c1-200 | tee -a $OUT/ab.jsonl
