#!/bin/bash
# round 2, call A: fused-kernel parity + kernel sweep (K1 butterfly vs direct form, K1+K2 vs fused) on cfg2 and cfg3
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== fused tests"; timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -25 | tee $OUT/pytest_fused.txt
echo "== sweep cfg2"; SWEEP_BATCH=512 SWEEP_SKIP_GENERIC=1 timeout 600 python scripts/sweep_kernels.py 2>&1 | grep -v generic | tee $OUT/sweep_cfg2.jsonl
echo "== sweep cfg3"; SWEEP_BATCH=64 SWEEP_SKIP_GENERIC=1 SWEEP_CONFIG=cfg3 timeout 600 python scripts/sweep_kernels.py 2>&1 | grep -v generic | tee $OUT/sweep_cfg3.jsonl
echo "== all gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
