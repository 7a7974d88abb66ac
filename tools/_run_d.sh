cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_prof_d.log
CMD="python tools/profile_run.py --utts 1036 --frames 60 --reps 3"
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
for k in v2 v1; do
  echo "== $k plain" >> $L; GTTS_KERNEL=$k $CMD >> $L 2>&1 || exit 1
  echo "== $k ncu metrics" >> $L
  GTTS_KERNEL=$k ncu --metrics $M --clock-control none -k regex:tube_kernel_v -s 2 -c 1 --csv --log-file gpurun_out/r02_ncu_d_$k.csv $CMD >> $L 2>&1
done
GTTS_KERNEL=v2 ncu --set full --clock-control none --import-source on -k regex:tube_kernel_v2 -s 2 -c 1 -o gpurun_out/prof_r02_v2d $CMD >> $L 2>&1
cat $L | tail -30
