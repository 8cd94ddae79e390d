#!/bin/bash
# round 2, call M: compute streams / slots A/B for device-output groups; 2048-bit subsequences
OUT=gpurun_out/r2m
mkdir -p $OUT
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 10 --tag "$tag" | cut -c1-330 | tee -a $OUT/ab.jsonl; }
run base
run streams3 B200JPG_COMP_STREAMS=3 B200JPG_SLOTS=6
run streams4 B200JPG_COMP_STREAMS=4 B200JPG_SLOTS=8
run streams1 B200JPG_COMP_STREAMS=1
run sub2048 B200JPG_SO=libb200jpg_sub2048.so
run sub2048s3 B200JPG_SO=libb200jpg_sub2048.so B200JPG_COMP_STREAMS=3 B200JPG_SLOTS=6
timeout 600 env B200JPG_SO=libb200jpg_sub2048.so ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_2048.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
