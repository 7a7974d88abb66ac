cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_p.log 2>&1; tail -4 gpurun_out/r02_tests_p.log
python tools/profile_run.py --utts 1036 --frames 200 --reps 3 | tail -2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tube_kernel_v2 -c 1 -s 1 -o gpurun_out/prof_r02_p -f python tools/profile_run.py --utts 1036 --frames 200 --reps 2 > gpurun_out/r02_prof_p.log 2>&1
tail -3 gpurun_out/r02_prof_p.log
