#!/bin/bash
# round 2, call 3E (2 GPUs): uploads by gather kernel -- GPU suite, whole-file throughput on one GPU, 2-GPU bench with trace
OUT=gpurun_out/r3e
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
export CUDA_VISIBLE_DEVICES=0
B200JPG_TRACE=1 python scripts/files_bench.py --reps 8 --tag hostout-gatherkernel 2>$OUT/trace_hostout.err | cut -c1-260 | tee -a $OUT/ab.jsonl
tail -1 $OUT/trace_hostout.err | cut -c1-330
python scripts/files_bench.py --dev-out --reps 16 --tag devout-gatherkernel | cut -c1-260 | tee -a $OUT/ab.jsonl
python scripts/files_bench.py --dev-out --reps 16 --tag devout-gatherkernel | cut -c1-260 | tee -a $OUT/ab.jsonl
unset CUDA_VISIBLE_DEVICES
B200JPG_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra-configs 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-200
grep "decode_files" $OUT/bench_n2.err | sed -n '4,6p' | cut -c1-330
