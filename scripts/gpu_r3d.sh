#!/bin/bash
# round 2, call 3D (2 GPUs): 2-GPU bench after the first-use reservation fix, with the engine's trace lines
OUT=gpurun_out/r3d
mkdir -p $OUT
B200JPG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra-configs 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-200
grep "decode_files" $OUT/bench_n2.err | tail -40 | cut -c1-330 > $OUT/trace_tail.txt
wc -l $OUT/trace_tail.txt
