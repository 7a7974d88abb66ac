set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r02_final_1gpu.json 2> gpurun_out/bench_r02_final_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_final_reference_arm.json 2> gpurun_out/bench_r02_final_reference_arm.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ncu_r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-model5 > gpurun_out/plain_b.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tube_kernel_v1 -s 3 -c 1 --csv --log-file gpurun_out/ncu_r02_traffic_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config3 --no-model5 > gpurun_out/ncu_b.log 2>&1
python tools/profile_run.py --utts 1036 --frames 60 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tube_kernel_v1 -s 2 -c 1 -o gpurun_out/v1_final -f python tools/profile_run.py --utts 1036 --frames 60 > gpurun_out/ncu_c.log 2>&1
tail -n 2 gpurun_out/plain_c.log; tail -n 2 gpurun_out/ncu_c.log
python tools/events_profile_run.py --reps 3 > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:events -s 4 -c 2 -o gpurun_out/events_final -f python tools/events_profile_run.py --reps 3 > gpurun_out/ncu_d.log 2>&1
tail -n 1 gpurun_out/plain_d.log
