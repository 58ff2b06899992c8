#!/bin/bash
# ncu over the accumulation kernels of the SECOND 2^22-point MSM of tools/profile_run.py (bench configuration).
# usage: tools/ncu_tree.sh out.csv [extra env...]
OUT=${1:-gpurun_out/ncu_tree.csv}
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,launch__registers_per_thread,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:"tree_|aff_|msm_accumulate|msm_scatter|msm_hist|msm_reduce" --launch-skip ${SKIP:-19} --launch-count ${COUNT:-19} --csv --log-file $OUT python tools/profile_run.py > ${OUT%.csv}.log 2>&1
