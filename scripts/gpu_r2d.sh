#!/bin/bash
# round 2, call D (2 GPUs): end-to-end scaling of the files path and of the worker boundary, concurrent PCIe probe
OUT=gpurun_out/r2d
mkdir -p $OUT
nproc > $OUT/host.txt; nvidia-smi topo -m >> $OUT/host.txt 2>&1; lscpu | grep -iE "numa|socket|model name|thread" >> $OUT/host.txt
echo "== N=1"; B200JPG_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-extra-configs --no-cpu-baseline 2>$OUT/bench_n1.err | tee $OUT/bench_n1.json | cut -c1-400
echo "== N=2"; B200JPG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-400
grep -c . $OUT/bench_n2.err; tail -5 $OUT/bench_n2.err
