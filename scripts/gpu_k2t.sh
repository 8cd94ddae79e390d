#!/bin/bash
# gpurun call for a kernel experiment: the K2 tests first (short timeout: a hung ring must not hang the box), then the sweep.
TAG=${1:-k2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== k2 tests" ; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bulk_copy or k2_ or quant_tables or full_size" 2>&1 | tail -15 | tee $OUT/pytest_k2.txt
if grep -q "failed\|Timeout\|error" $OUT/pytest_k2.txt; then echo "k2 tests failed; skipping the rest"; exit 1; fi
echo "== sweep" ; timeout 600 python scripts/sweep_kernels.py 2>&1 | tail -12 | tee $OUT/sweep.jsonl
[ -n "$SHORT" ] && exit 0
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench (default)" ; timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
