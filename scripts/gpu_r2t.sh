#!/bin/bash
# round 2, call T: same-box A/B of the synchronisation variants (0: one staged kernel, 1: full pass + unstaged rounds, 2: full pass + staged rounds)
OUT=gpurun_out/r2t
mkdir -p $OUT
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 16 --tag "$tag" | cut -c1-200 | tee -a $OUT/ab.jsonl; }
for r in 1 2 3; do
run sync1
run sync0 B200JPG_SO=libb200jpg_sync0.so
run sync2 B200JPG_SO=libb200jpg_sync2.so
done
