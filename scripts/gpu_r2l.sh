#!/bin/bash
# round 2, call L: where the device entropy chain stands after the lane gathering -- launch list and instruction counts
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -2 $OUT/run.log
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:'ent_|k0_' -s 22 -c 24 --csv --log-file $OUT/ent_metrics.csv python scripts/files_run.py 64 > $OUT/run2.log 2>&1
tail -2 $OUT/run2.log
python scripts/files_bench.py --dev-out --reps 12 --tag devout | cut -c1-300
python scripts/files_bench.py --reps 8 --tag hostout | cut -c1-300
