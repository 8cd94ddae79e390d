#!/bin/bash
# round 2, call O: job-sorted groups with merged copies -- timeline, throughput, GPU tests
OUT=gpurun_out/r2o
mkdir -p $OUT
B200JPG_TIMELINE=1 B200JPG_TRACE=1 python scripts/files_bench.py --reps 3 --tag hostout-timeline 2>$OUT/timeline_hostout.err | cut -c1-300
python scripts/files_bench.py --reps 8 --tag hostout | cut -c1-300
python scripts/files_bench.py --reps 8 --threads 8 --tag hostout-8thr | cut -c1-300
python scripts/files_bench.py --reps 8 --threads 4 --tag hostout-4thr | cut -c1-300
python scripts/files_bench.py --dev-out --reps 10 --tag devout | cut -c1-300
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
