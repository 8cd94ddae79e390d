#!/bin/bash
# round 2, 8 GPUs: the default bench line at N=8 with the final engine (four compute streams, gather-kernel uploads, relaxed grouping)
OUT=gpurun_out/r3n8
mkdir -p $OUT
nproc > $OUT/host.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra-configs 2>$OUT/bench_n8.err | tee $OUT/bench_n8.json | cut -c1-200
tail -3 $OUT/bench_n8.err | cut -c1-200
