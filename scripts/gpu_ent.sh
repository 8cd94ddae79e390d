#!/bin/bash
# gpurun call for the device entropy decoder: its tests first (under compute-sanitizer for a small case), then everything.
TAG=${1:-ent}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== entropy tests" ; timeout 900 python -m pytest tests/test_gpu_entropy.py -x -q 2>&1 | tail -30 | tee $OUT/pytest_ent.txt
echo "== sanitizer (small)" ; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_entropy.py -x -q -k "fixtures or crashtest" 2>&1 | tail -15 | tee $OUT/sanitizer.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; B200JPG_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json
grep "b200jpg\]" $OUT/bench.err | tail -12
