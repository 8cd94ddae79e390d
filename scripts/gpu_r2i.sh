#!/bin/bash
# round 2, call I: full GPU suite + smoke + files path with direct device outputs + default bench
OUT=gpurun_out/r2i
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== files dev-out / host-out"; timeout 300 python scripts/files_bench.py --dev-out --reps 12 --tag direct 2>/dev/null | tee $OUT/files.jsonl | cut -c1-330
timeout 300 python scripts/files_bench.py --reps 6 --tag host 2>/dev/null | tee -a $OUT/files.jsonl | cut -c1-330
echo "== bench default"; timeout 1200 python bench.py --steps 10 --warmup 3 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300
tail -3 $OUT/bench.err
