#!/bin/bash
# One `ncu --set full` capture each of the two 4:2:0 K2 kernels (bulk-copy fed and load/store) on a 256-image batch.
TAG=${1:-ncu_k2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for mode in ${MODES:-0 1}; do
  B200JPG_K2_MODE=$mode SWEEP_BATCH=256 SWEEP_ONLY_AUTO=1 timeout 600 ncu --set full --import-source on --clock-control none \
    -k regex:k2_ycbcr420 -s 6 -c 1 -o $OUT/k2_mode$mode -f python scripts/sweep_kernels.py > $OUT/ncu_mode$mode.log 2>&1
  tail -3 $OUT/ncu_mode$mode.log
done
ls -la $OUT
