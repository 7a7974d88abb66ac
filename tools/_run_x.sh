cd $GRAFT_REPO_ROOT
python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 3 | tail -3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tube_kernel_v2 -c 1 -s 1 -o gpurun_out/prof_r02_v2mixed -f python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 2 > gpurun_out/r02_prof_v2mixed.log 2>&1
tail -2 gpurun_out/r02_prof_v2mixed.log
