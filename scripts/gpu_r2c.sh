#!/bin/bash
# round 2, call C: whole GPU suite + the new bench line (cfg2 + cfg3 + cfg4 + CPU baseline) + the reference arm
OUT=gpurun_out/r2c
mkdir -p $OUT
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)"; B200JPG_TRACE=1 timeout 1200 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-3000
tail -30 $OUT/bench.err
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/bench_ref.err | tee $OUT/bench_reference.json | cut -c1-2500
