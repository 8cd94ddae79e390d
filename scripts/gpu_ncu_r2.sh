#!/bin/bash
# round 2: ncu --set full of the shipped kernels (cfg2 x1024: K1, K2 bulk-copy; cfg3 x128: KF, K1, K2 4:4:4) + launch list of bench.py
OUT=gpurun_out/r2ncu
mkdir -p $OUT
SWEEP_ONLY_AUTO=1 SWEEP_BATCH=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_idct8_tma|k2_ycbcr420_tma" -s 12 -c 2 -o $OUT/cfg2 python scripts/sweep_kernels.py > $OUT/ncu_cfg2.log 2>&1
tail -2 $OUT/ncu_cfg2.log
SWEEP_ONLY_AUTO=1 SWEEP_BATCH=128 SWEEP_CONFIG=cfg3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_idct8_tma|k2_ycbcr444" -s 12 -c 2 -o $OUT/cfg3 python scripts/sweep_kernels.py > $OUT/ncu_cfg3.log 2>&1
tail -2 $OUT/ncu_cfg3.log
SWEEP_ONLY_AUTO=1 SWEEP_BATCH=128 SWEEP_CONFIG=cfg3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kf_fused" -s 2 -c 1 -o $OUT/cfg3_kf python scripts/sweep_kernels.py > $OUT/ncu_cfg3kf.log 2>&1
tail -2 $OUT/ncu_cfg3kf.log
SWEEP_ONLY_AUTO=1 SWEEP_BATCH=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kf_fused" -s 2 -c 1 -o $OUT/cfg2_kf python scripts/sweep_kernels.py > $OUT/ncu_cfg2kf.log 2>&1
tail -2 $OUT/ncu_cfg2kf.log
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $OUT/ncu_bench.log 2>&1
grep -E "k1_|k2_|kf_|k0_|ent_" $OUT/launches.csv | awk -F'","' '{print $5}' | sed 's/(.*//' | sort | uniq -c | sort -rn | head -20
ls -la $OUT
echo "== group size A/B (device outputs)"
for g in 1 24 48 96; do B200JPG_GROUP_MIN=$g timeout 300 python scripts/files_bench.py --dev-out --reps 10 --tag "groupmin-$g" 2>/dev/null | tee -a $OUT/group_ab.jsonl | cut -c1-330; done
B200JPG_GROUP_MIN=48 timeout 300 python scripts/files_bench.py --reps 6 --tag "hostout" 2>/dev/null | tee -a $OUT/group_ab.jsonl | cut -c1-330
