#!/bin/bash
# round 2, call 3G (2 GPUs): bench line with 512 files per call and rank at N > 1
OUT=gpurun_out/r3g
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra-configs 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-200
tail -2 $OUT/bench_n2.err | cut -c1-200
