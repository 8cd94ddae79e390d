#!/bin/bash
# files pipeline trace only
OUT=gpurun_out/${1:-files}; mkdir -p $OUT
B200JPG_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-batch 64 2>$OUT/bench.err > $OUT/bench.json
grep "b200jpg\]" $OUT/bench.err | tail -40
