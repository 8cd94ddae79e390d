#!/bin/bash
# round 2, call B: fused kernel -- parity, CTA-shape sweep, ncu profile
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== fused tests (8 warps)"; timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -8 | tee $OUT/pytest_fused_w8.txt
echo "== fused tests (4 warps)"; B200JPG_KF_WARPS=4 timeout 900 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -8 | tee $OUT/pytest_fused_w4.txt
for cfg in "cfg2 512" "cfg3 64"; do set -- $cfg
  for w in 8 4; do for strip in 1920 960 480; do
    echo "== sweep $1 warps=$w strip=$strip"
    B200JPG_KF_WARPS=$w B200JPG_KF_STRIP=$strip SWEEP_ONLY_AUTO=1 SWEEP_BATCH=$2 SWEEP_CONFIG=$1 timeout 300 python scripts/sweep_kernels.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l, end=''); continue
    print(json.dumps({'cfg': d['config'], 'warps': $w, 'strip': $strip, 'k1k2_ms': round(d['k1k2_ms'], 4), 'fused_ms': d['fused_ms'] and round(d['fused_ms'], 4), 'fused_mps': d['fused_mps'], 'same': d['fused_same_pixels']}))
" | tee -a $OUT/kf_sweep.jsonl
  done; done
done
echo "== ncu kf"
SWEEP_ONLY_AUTO=1 SWEEP_BATCH=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:kf_fused -s 3 -c 1 -o $OUT/kf_cfg2 python scripts/sweep_kernels.py > $OUT/ncu_kf.log 2>&1
tail -3 $OUT/ncu_kf.log
ls -la $OUT
