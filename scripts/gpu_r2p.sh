#!/bin/bash
# round 2, call P: one upload per group (submitter gathers into a staging buffer) against per-image uploads
OUT=gpurun_out/r2p
mkdir -p $OUT
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --reps 8 --tag "$tag" | cut -c1-330 | tee -a $OUT/ab.jsonl; }
run hostout
run hostout-gather B200JPG_GATHER=1
run hostout
run hostout-gather B200JPG_GATHER=1
B200JPG_GATHER=1 B200JPG_TIMELINE=1 B200JPG_TRACE=1 python scripts/files_bench.py --reps 3 --tag hostout-gather-timeline 2>$OUT/timeline_gather.err | cut -c1-300
python scripts/files_bench.py --dev-out --reps 10 --tag devout | cut -c1-300 | tee -a $OUT/ab.jsonl
B200JPG_GATHER=1 python scripts/files_bench.py --dev-out --reps 10 --tag devout-gather | cut -c1-300 | tee -a $OUT/ab.jsonl
