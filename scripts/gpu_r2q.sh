#!/bin/bash
# round 2, call Q: warp-per-subsequence synchronisation rounds (ENT_COOP_MAX 0 / 4 / 8 / 16)
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -3 | tee $OUT/pytest_ent.txt
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 12 --tag "$tag" | cut -c1-330 | tee -a $OUT/ab.jsonl; }
run coop8
run coop0 B200JPG_SO=libb200jpg_coop0.so
run coop4 B200JPG_SO=libb200jpg_coop4.so
run coop16 B200JPG_SO=libb200jpg_coop16.so
run coop8
run coop0 B200JPG_SO=libb200jpg_coop0.so
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $OUT/launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
