#!/bin/bash
# round 2, call V: shift-in quad window in the write pass, staged values in K0, 256 / 384 / 512 subsequences per CTA
OUT=gpurun_out/r2v
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_entropy.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
timeout 600 env B200JPG_SO=libb200jpg_t384.so python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -2
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 16 --tag "$tag" | cut -c1-200 | tee -a $OUT/ab.jsonl; }
for r in 1 2 3; do
run t256
run t384 B200JPG_SO=libb200jpg_t384.so
run t512 B200JPG_SO=libb200jpg_t512.so
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 100 --csv --log-file $OUT/launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
