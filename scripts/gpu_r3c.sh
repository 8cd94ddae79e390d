#!/bin/bash
# round 2, call 3C (2 GPUs): full GPU suite + smoke + default bench + reference arm on one GPU, then the 2-GPU bench
OUT=gpurun_out/r3c
mkdir -p $OUT
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench default"; CUDA_VISIBLE_DEVICES=0 timeout 1200 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300
tail -3 $OUT/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/bench.err | tee $OUT/bench_reference.json | cut -c1-300
echo "== N=2 bench"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-300
tail -3 $OUT/bench_n2.err
