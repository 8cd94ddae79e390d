#!/bin/bash
# sparse-stream path: new parity tests first, then the whole GPU suite, smoke, and the bench line with the files trace
TAG=${1:-sbs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/host.txt; grep -m1 "model name" /proc/cpuinfo >> $OUT/host.txt; lscpu | grep -iE "numa|socket|thread|flags" | cut -c1-2000 >> $OUT/host.txt
echo "== new tests" ; timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 100 -k "compaction or streaming_engine" 2>&1 | tail -25 | tee $OUT/pytest_new.txt
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -x -q -m gpu --timeout 100 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; B200JPG_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json
grep "b200jpg\]" $OUT/bench.err | tail -12
tail -5 $OUT/bench.err
