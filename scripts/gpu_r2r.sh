#!/bin/bash
# round 2, call R: lean decode loop (total-bits table entries, transposed scan tile)
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -3 | tee $OUT/pytest_ent.txt
for r in 1 2; do python scripts/files_bench.py --dev-out --reps 12 --tag lean | cut -c1-330 | tee -a $OUT/ab.jsonl; done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $OUT/launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
