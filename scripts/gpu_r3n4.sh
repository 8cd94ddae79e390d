#!/bin/bash
# round 2, 4 GPUs: the default bench line at N=4 with the final engine
OUT=gpurun_out/r3n4
mkdir -p $OUT
nproc > $OUT/host.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --no-extra-configs 2>$OUT/bench_n4.err | tee $OUT/bench_n4.json | cut -c1-200
tail -3 $OUT/bench_n4.err | cut -c1-200
