#!/bin/bash
# round 2, call J: whole-file path with few CPUs per GPU (what an 8-GPU box gives each rank) after the vectorised unstuffing
OUT=gpurun_out/r2j
mkdir -p $OUT
echo "== entropy gpu tests"; timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q -m gpu 2>&1 | tail -3
for t in 2 4 8; do
  taskset -c 0-$((t-1)) timeout 300 python scripts/files_bench.py --dev-out --threads $t --reps 8 --tag "devout-${t}cpu" 2>/dev/null | tee -a $OUT/few_cpus.jsonl | cut -c1-300
done
taskset -c 0-3 timeout 300 python scripts/files_bench.py --threads 4 --reps 6 --tag "hostout-4cpu" 2>/dev/null | tee -a $OUT/few_cpus.jsonl | cut -c1-300
timeout 300 python scripts/files_bench.py --dev-out --reps 8 --tag "devout-allcpu" 2>/dev/null | tee -a $OUT/few_cpus.jsonl | cut -c1-300
