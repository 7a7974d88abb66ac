cd $GRAFT_REPO_ROOT
CMD="python tools/profile_run.py --utts 1036 --frames 60 --reps 3"
GTTS_KERNEL=v2 $CMD > gpurun_out/r02_prof_f.log 2>&1 || exit 1
GTTS_KERNEL=v2 ncu --set full --clock-control none --import-source on -k regex:tube_kernel_v2 -s 2 -c 1 -o gpurun_out/prof_r02_v2f $CMD >> gpurun_out/r02_prof_f.log 2>&1
tail -3 gpurun_out/r02_prof_f.log
