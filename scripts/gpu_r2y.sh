#!/bin/bash
# round 2, call Y: DC prefix in one kernel -- tests, launch list, throughput
OUT=gpurun_out/r2y
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
for r in 1 2 3; do python scripts/files_bench.py --dev-out --reps 16 --tag dcscan | cut -c1-200 | tee -a $OUT/ab.jsonl; done
python scripts/files_bench.py --dev-out --reps 16 --config cfg3 --n 128 --tag dcscan-cfg3 | cut -c1-200 | tee -a $OUT/ab.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $OUT/entropy_launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
