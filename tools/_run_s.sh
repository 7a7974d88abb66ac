cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_s.log 2>&1; tail -5 gpurun_out/r02_tests_s.log
python tools/profile_run.py --model5 --utts 1776 --frames 60 --reps 3 | tail -3
python tools/profile_run.py --model5 --utts 1 --frames 332 --reps 3 | tail -2
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tube5_kernel -c 1 -s 1 -o gpurun_out/prof_r02_m5a -f python tools/profile_run.py --model5 --utts 1776 --frames 60 --reps 2 > gpurun_out/r02_prof_m5a.log 2>&1
tail -2 gpurun_out/r02_prof_m5a.log
