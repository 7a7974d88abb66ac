cd $GRAFT_REPO_ROOT
tools/microbench/stream_push tools/microbench/stream_frames_20000.f32 20000 | tee gpurun_out/r02_stream_push.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_o.log 2>&1; tail -4 gpurun_out/r02_tests_o.log
python tools/profile_run.py --utts 1036 --frames 200 --reps 3 | tail -2
