#!/bin/bash
# round 2: the scaling line at 8 GPUs (as the driver launches it)
OUT=gpurun_out/r2n8
mkdir -p $OUT
nproc > $OUT/host.txt; nvidia-smi topo -m >> $OUT/host.txt 2>&1; lscpu | grep -iE "numa|socket|model name|thread" >> $OUT/host.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 2>$OUT/bench_n8.err | tee $OUT/bench_n8.json | cut -c1-300
tail -5 $OUT/bench_n8.err
