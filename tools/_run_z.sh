cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_z.log 2>&1; tail -3 gpurun_out/r02_tests_z.log
echo "== mixed"; python tools/profile_run.py --mixed --utts 4144 --frames 100 --reps 3 | tail -3 | head -2
echo "== uniform"; python tools/profile_run.py --utts 1036 --frames 200 --reps 3 | tail -3 | head -2
python tools/v3_probe.py --parity 0 --kernels v2 2>&1 | tail -1
