#!/bin/bash
OUT=gpurun_out/r2h
mkdir -p $OUT
echo "== new gpu tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_formats.py -x -q -m gpu -k "hand_derived or whole_files or formats or scaled" 2>&1 | tail -5
echo "== dev-out, 24 reps, trace"; B200JPG_TRACE=1 timeout 300 python scripts/files_bench.py --dev-out --reps 24 --tag trace 2>$OUT/trace_devout.err | tee $OUT/devout.json
grep decode_files $OUT/trace_devout.err | awk '{print $7, $8, $10, $11, $12, $13, $14, $18, $19, $21, $22}' | tail -26
echo "== host-out, 12 reps, trace"; B200JPG_TRACE=1 timeout 300 python scripts/files_bench.py --reps 12 --tag trace 2>$OUT/trace_hostout.err | tee $OUT/hostout.json
grep decode_files $OUT/trace_hostout.err | awk '{print $7, $8, $10, $11, $12, $13, $14, $18, $19, $21, $22}' | tail -13
echo "== entropy gpu tests"; timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q -m gpu 2>&1 | tail -3
echo "== variants bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-200
tail -3 $OUT/bench.err
