cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_tests_q.log 2>&1; tail -6 gpurun_out/r02_tests_q.log
timeout 900 python bench.py > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err; tail -3 gpurun_out/r02_bench_q.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config3 --no-model5 > gpurun_out/r02_launches_q.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tube_kernel_v2 -s 3 -c 1 --csv --log-file gpurun_out/r02_traffic_q.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-model5 > gpurun_out/r02_traffic_q.log 2>&1
tail -2 gpurun_out/r02_traffic_q.csv | cut -c1-400
