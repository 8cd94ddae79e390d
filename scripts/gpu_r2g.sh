#!/bin/bash
# round 2, call G (1 GPU): GPU suite with the new K1/K2/K3 variants, SSSE3-mode bench, default bench
OUT=gpurun_out/r2g
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench ssse3 mode"; timeout 600 python bench.py --arith ssse3 --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline 2>$OUT/bench_ssse3.err | tee $OUT/bench_ssse3.json | cut -c1-300
tail -3 $OUT/bench_ssse3.err
echo "== bench default"; timeout 1200 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300
tail -3 $OUT/bench.err
