cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 > gpurun_out/r02_bench_u.json 2> gpurun_out/r02_bench_u.err; tail -2 gpurun_out/r02_bench_u.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_u_ref.json 2>> gpurun_out/r02_bench_u.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_u.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_u.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tube_kernel_v2 -s 3 -c 1 --csv --log-file gpurun_out/r02_traffic_u.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-model5 > gpurun_out/r02_traffic_u.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tube_kernel_v2 -c 1 -s 1 -o gpurun_out/prof_r02_v2u -f python tools/profile_run.py --utts 1036 --frames 60 --reps 2 > gpurun_out/r02_prof_v2u.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tube5_kernel -c 1 -s 1 -o gpurun_out/prof_r02_m5u -f python tools/profile_run.py --model5 --utts 1776 --frames 60 --reps 2 > gpurun_out/r02_prof_m5u.log 2>&1
tail -2 gpurun_out/r02_prof_m5u.log; tail -1 gpurun_out/r02_traffic_u.csv | cut -c1-300
