#!/bin/bash
# lanes / instruction count / issue / stalls of the C2 GJK kernel (capsule-box launch) for library variants given as args
for v in "$@"; do
  [ "$v" = base ] && lib=$PWD/mind-fcl_b200/libfclb200.so || lib=$PWD/mind-fcl_b200/libfclb200_$v.so
  FCLB_LIB=$lib timeout 300 ncu --clock-control none -k regex:distanceGjkBinned -s 4 -c 1 \
    --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__warps_active.avg.per_cycle_active \
    --csv --log-file gpurun_out/gjk_$v.csv python bench.py --no-cpu-baseline --no-workloads --steps 2 --warmup 1 > /dev/null 2>&1
  echo "variant $v"; grep -v "^==" gpurun_out/gjk_$v.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tail -17
done
