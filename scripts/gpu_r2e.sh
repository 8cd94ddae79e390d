#!/bin/bash
# round 2, call E (2 GPUs): GPU suite with the new K2 paths, sanitizer evidence, 2-GPU bench after the allocator change
OUT=gpurun_out/r2e
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== memcheck (fused kernel, K1/K2 odd sizes, device entropy)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck.log \
    python -m pytest -q -x -m gpu tests/test_gpu_fused.py -k "many_small or not_taken or full_size" tests/test_gpu_parity.py -k "k2_upsample_ycbcr_bit_exact and scalar-auto or colour_transforms or known_answer" tests/test_gpu_entropy.py -k "fixtures or restart" 2>&1 | tail -6 | tee $OUT/sanitizer_memcheck_pytest.txt
tail -5 $OUT/sanitizer_memcheck.log
echo "== racecheck (shared-memory hazards: fused kernel, TMA rings)"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $OUT/sanitizer_racecheck.log \
    python -m pytest -q -x -m gpu tests/test_gpu_fused.py -k "many_small or not_taken" tests/test_gpu_parity.py -k "bulk_copy_kernel_geometries and 0 or known_answer" 2>&1 | tail -6 | tee $OUT/sanitizer_racecheck_pytest.txt
tail -8 $OUT/sanitizer_racecheck.log
echo "== N=2 bench"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>$OUT/bench_n2.err | tee $OUT/bench_n2.json | cut -c1-300
