#!/bin/bash
# round 2, call 3H: worker route with buffers and plan kept across images -- parity tests, latency before / after
OUT=gpurun_out/r3h
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
B200JPG_SO=libb200jpg_oldworker.so python scripts/worker_latency.py 30 | tee $OUT/worker_old.json
python scripts/worker_latency.py 30 | tee $OUT/worker_new.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
