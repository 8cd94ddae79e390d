#!/bin/bash
# One gpurun call: parity tests, smoke, bench (+ variants), ncu launch list + full capture of the top kernels.
# Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt; free -g | head -2 >> $OUT/host.txt; lscpu | grep -iE "numa|socket|thread" >> $OUT/host.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
echo "== bench variants"
for v in "--k1 generic --k2 generic" "--arith ssse3"; do
  echo "-- $v"; timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 0.5 --e2e-batch 16 $v 2>>$OUT/bench.err | tee -a $OUT/bench_variants.json
done
echo "== bench cfg3" ; timeout 900 python bench.py --config cfg3 --batch 128 --unique 2 --e2e-batch 32 --steps 5 --warmup 3 --cpu-seconds 3 2>>$OUT/bench.err | tee $OUT/bench_cfg3.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/bench.err | tee $OUT/bench_reference.json
echo "== sweep" ; timeout 300 python scripts/sweep_kernels.py 2>&1 | tee $OUT/sweep.jsonl
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --batch 256 --e2e-batch 16 --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
grep -E "k1_|k2_" $OUT/launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -k2 | head -30
echo "== ncu full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_idct8_tma|k2_ycbcr420" -s 8 -c 2 -o $OUT/prof \
    python bench.py --steps 2 --warmup 3 --batch 256 --e2e-batch 16 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
