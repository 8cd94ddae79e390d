#!/bin/bash
# round 2, call 3A: memcheck + racecheck of the entropy kernels as they stand at the end of round 2 (transposed tile in dynamic
# shared memory, 256-thread CTAs, shift-in write window) on fixtures, restart intervals, corrupted scans
OUT=gpurun_out/r3a
mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/sanitizer_memcheck_entropy.log \
    python -m pytest -q -x -m gpu tests/test_gpu_entropy.py 2>&1 | tail -4 | tee $OUT/memcheck_pytest.txt
tail -4 $OUT/sanitizer_memcheck_entropy.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $OUT/sanitizer_racecheck_entropy.log \
    python -m pytest -q -x -m gpu tests/test_gpu_entropy.py -k "fixtures or restart" 2>&1 | tail -4 | tee $OUT/racecheck_pytest.txt
tail -6 $OUT/sanitizer_racecheck_entropy.log
