cd $GRAFT_REPO_ROOT
L=gpurun_out/r02_prof_e.log
CMD="python tools/profile_run.py --utts 1036 --frames 200 --reps 4"
for k in v2 v1; do echo "== $k" >> $L; GTTS_KERNEL=$k $CMD >> $L 2>&1; done
echo "== v2 role profile" >> $L; GTTS_PROFILE=1 python tools/profile_run.py --utts 1036 --frames 200 --reps 2 --lib ab/prof.so >> $L 2>&1
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__icc_request_hit_rate.pct,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
GTTS_KERNEL=v2 ncu --metrics $M --clock-control none -k regex:tube_kernel_v -s 2 -c 1 --csv --log-file gpurun_out/r02_ncu_e_v2.csv python tools/profile_run.py --utts 1036 --frames 60 --reps 3 >> $L 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or ragged or stress or fresh or pcm16 or loud or mixed" >> $L 2>&1
cat $L
