#!/bin/bash
# round 2, call W: rounds kernel with staged tables but scan words through L1 (19 KB per waiting CTA instead of 54); job-ordered groups
OUT=gpurun_out/r2w
mkdir -p $OUT
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 16 --tag "$tag" | cut -c1-200 | tee -a $OUT/ab.jsonl; }
for r in 1 2 3; do
run base
run split B200JPG_SO=libb200jpg_split.so
done
B200JPG_TRACE=1 python scripts/files_bench.py --reps 8 --tag hostout 2>$OUT/trace_hostout.err | cut -c1-300 | tee -a $OUT/ab.jsonl
tail -2 $OUT/trace_hostout.err | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -2
