#!/bin/bash
# --set full capture of the remaining entropy-path kernels of the first full group: ent_sync (10 launches), ent_dc_*, k0_expand_blocks
OUT=gpurun_out/${1:-entprof_b}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ent_sync|ent_dc|k0_expand_blocks' -s 14 -c 14 -o $OUT/ent_full2 python scripts/files_run.py 64 > $OUT/run_full2.log 2>&1
tail -2 $OUT/run_full2.log
