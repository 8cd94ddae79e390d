#!/bin/bash
# Short gpurun call: GPU parity tests + smoke + the default bench line.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; B200JPG_TRACE=1 timeout 900 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json
tail -40 $OUT/bench.err
