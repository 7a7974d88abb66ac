cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q -k "stream or golden" > gpurun_out/r02_tests_n.log 2>&1; tail -5 gpurun_out/r02_tests_n.log
timeout 600 python tools/bench_configs.py --utts 64 --stream-frames 5000 2>&1 | tail -30
