#!/bin/bash
# round 2, call U: subsequences per CTA (64 / 128 / 256 threads: 8 / 6 / 4 CTAs and 16 / 24 / 32 warps per SM)
OUT=gpurun_out/r2u
mkdir -p $OUT
timeout 600 env B200JPG_SO=libb200jpg_t256.so python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -2
run() { tag=$1; shift; env "$@" python scripts/files_bench.py --dev-out --reps 16 --tag "$tag" | cut -c1-200 | tee -a $OUT/ab.jsonl; }
for r in 1 2 3; do
run t128
run t256 B200JPG_SO=libb200jpg_t256.so
run t64 B200JPG_SO=libb200jpg_t64.so
done
