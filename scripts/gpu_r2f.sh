#!/bin/bash
# round 2, call F (2 GPUs): entropy subsequence size A/B, files path side by side on two GPUs, sanitizer with a proper selection
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== subsequence size A/B (1 GPU, device outputs and host outputs)"
for so in libb200jpg.so libb200jpg_sub512.so libb200jpg_sub256.so; do
  B200JPG_SO=$so timeout 300 python scripts/files_bench.py --dev-out --tag sub 2>>$OUT/err.txt | tee -a $OUT/files_ab.jsonl
  B200JPG_SO=$so timeout 300 python scripts/files_bench.py --tag sub 2>>$OUT/err.txt | tee -a $OUT/files_ab.jsonl
done
echo "== entropy tests with the 512-bit build"; B200JPG_SO=libb200jpg_sub512.so timeout 600 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -3
echo "== two GPUs side by side, device outputs, threads per process 4 / 8 / 12"
for t in 4 8 12; do
  (CUDA_VISIBLE_DEVICES=0 taskset -c 0-11 timeout 300 python scripts/files_bench.py --dev-out --threads $t --tag "2gpu-a-t$t" 2>>$OUT/err.txt >> $OUT/files_2gpu.jsonl &)
  CUDA_VISIBLE_DEVICES=1 taskset -c 12-23 timeout 300 python scripts/files_bench.py --dev-out --threads $t --tag "2gpu-b-t$t" 2>>$OUT/err.txt >> $OUT/files_2gpu.jsonl
  sleep 3
done
echo "-- one GPU alone on 12 cpus"; CUDA_VISIBLE_DEVICES=0 taskset -c 0-11 timeout 300 python scripts/files_bench.py --dev-out --threads 8 --tag "1gpu-12cpu-t8" 2>>$OUT/err.txt >> $OUT/files_2gpu.jsonl
echo "-- host outputs, two GPUs"
(CUDA_VISIBLE_DEVICES=0 taskset -c 0-11 timeout 300 python scripts/files_bench.py --threads 8 --tag "2gpu-a-host" 2>>$OUT/err.txt >> $OUT/files_2gpu.jsonl &)
CUDA_VISIBLE_DEVICES=1 taskset -c 12-23 timeout 300 python scripts/files_bench.py --threads 8 --tag "2gpu-b-host" 2>>$OUT/err.txt >> $OUT/files_2gpu.jsonl
sleep 3
cat $OUT/files_2gpu.jsonl
SEL="many_small or not_taken or full_size or colour_transforms or known_answer or (k2_upsample_ycbcr_bit_exact and scalar-auto) or fixtures or restart or bulk_copy_kernel_geometries"
echo "== memcheck"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck.log \
    python -m pytest -q -x -m gpu tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_entropy.py -k "$SEL" 2>&1 | tail -4 | tee $OUT/sanitizer_memcheck_pytest.txt
tail -3 $OUT/sanitizer_memcheck.log
echo "== racecheck"
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $OUT/sanitizer_racecheck.log \
    python -m pytest -q -x -m gpu tests/test_gpu_fused.py tests/test_gpu_parity.py tests/test_gpu_entropy.py -k "$SEL" 2>&1 | tail -4 | tee $OUT/sanitizer_racecheck_pytest.txt
tail -3 $OUT/sanitizer_racecheck.log
tail -5 $OUT/err.txt
