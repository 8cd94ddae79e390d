#!/bin/bash
# round 2, call N: per-group device timeline of decode_files (where are the bubbles in the download stream?)
OUT=gpurun_out/r2n
mkdir -p $OUT
B200JPG_TIMELINE=1 B200JPG_TRACE=1 python scripts/files_bench.py --reps 3 --tag hostout-timeline 2>$OUT/timeline_hostout.err | cut -c1-300
B200JPG_TIMELINE=1 B200JPG_TRACE=1 python scripts/files_bench.py --dev-out --reps 3 --tag devout-timeline 2>$OUT/timeline_devout.err | cut -c1-300
python scripts/files_bench.py --dev-out --reps 10 --tag devout-4streams | cut -c1-300
python scripts/files_bench.py --reps 6 --tag hostout-8slots | cut -c1-300
