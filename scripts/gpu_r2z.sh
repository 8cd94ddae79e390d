#!/bin/bash
# round 2, call Z: DC prefix kernel with independent loads
OUT=gpurun_out/r2z
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -2 | tee $OUT/pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 -k regex:'ent_dc|ent_prefix' --csv --log-file $OUT/dc_launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
grep -o 'ent_[a-z_]*.*' $OUT/dc_launches.csv | awk -F'","' '{print $1, $5, $NF}' | head -12
for r in 1 2; do python scripts/files_bench.py --dev-out --reps 16 --tag dcscan2 | cut -c1-200 | tee -a $OUT/ab.jsonl; done
