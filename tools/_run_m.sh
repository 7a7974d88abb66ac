cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r02_2gpu_tests.log
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or sharded" >> gpurun_out/r02_2gpu_tests.log 2>&1; tail -4 gpurun_out/r02_2gpu_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -3 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_2gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2))
e=d["e2e"]; print("e2e", round(e["value"]), round(e["ms_per_step"],2), {k:(round(v["value"]),round(v["ms_per_step"],2)) for k,v in e["variants"].items()}, e["ceiling"]["d2h_gbs_per_gpu"], e["ceiling"]["frac_of_ceiling"])
print("config3", d.get("config3"))
PY
