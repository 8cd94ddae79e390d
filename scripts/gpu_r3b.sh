#!/bin/bash
# round 2, call 3B: 4:4:0 fast kernel tests; counters of the entropy-path kernels as shipped (one 27-image group)
OUT=gpurun_out/r3b
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_formats.py -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
M=gpu__time_duration.sum,launch__registers_per_thread,launch__shared_mem_per_block_static,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
timeout 900 ncu --metrics $M --clock-control none -k regex:'ent_|k0_|k1_|k2_' -s 16 -c 36 --csv --log-file $OUT/ent_metrics.csv python scripts/files_run.py 64 > $OUT/run.log 2>&1
tail -1 $OUT/run.log
python scripts/ent_ncu_summary.py $OUT/ent_metrics.csv $OUT/entropy_ncu_summary.csv
# racecheck of the entropy kernels alone (the unfiltered run's display cap of 100 reports is used up by k2_ycbcr420_tma's known
# mbarrier-ordered "potential WAR" reports, profiles/r02_sanitizer.md)
timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --kernel-regex kns=ent_ --error-exitcode 9 --log-file $OUT/sanitizer_racecheck_ent_only.log \
    python -m pytest -q -x -m gpu tests/test_gpu_entropy.py -k "restart" 2>&1 | tail -3 | tee $OUT/racecheck_ent_pytest.txt
tail -3 $OUT/sanitizer_racecheck_ent_only.log
