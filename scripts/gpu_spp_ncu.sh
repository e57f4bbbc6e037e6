#!/bin/bash
# instructions / lanes / issue utilisation of render_kernel against spp
TAG=${1:-sppncu}; shift; OUT=gpurun_out; mkdir -p $OUT
for SPP in ${@:-4 32 256}; do
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum \
  --clock-control none -k regex:render_kernel -c 1 --csv --log-file $OUT/spp_ncu_${TAG}_$SPP.csv python bench.py --steps 1 --warmup 0 --spp $SPP --no-cpu-baseline --no-other-configs > /dev/null 2>&1
done
ls $OUT/spp_ncu_${TAG}_*.csv
