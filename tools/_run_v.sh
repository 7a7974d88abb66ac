cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_tests_v.log 2>&1; tail -3 gpurun_out/r02_tests_v.log
for v in old t2c1 t2c4; do echo "== $v"; python tools/profile_run.py --utts 1036 --frames 200 --reps 4 --lib ab/$v.so | tail -3 | head -2; done
echo "== default (chunk 2)"; python tools/profile_run.py --utts 1036 --frames 200 --reps 4 | tail -3 | head -2
