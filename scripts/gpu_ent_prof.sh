#!/bin/bash
# per-kernel durations of the whole-file path with device entropy decoding (ncu launch list, serialised launches),
# then --set full captures of the entropy kernels
OUT=gpurun_out/${1:-entprof}; mkdir -p $OUT
echo "== entropy tests" ; timeout 900 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -5 | tee $OUT/pytest_ent.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -3 $OUT/run.log
# skip the first (small) group: launch-skip counts matching kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ent_pass' -s 2 -c 2 -o $OUT/ent_full python scripts/files_run.py 64 > $OUT/run_full.log 2>&1
tail -3 $OUT/run_full.log
ls -la $OUT
