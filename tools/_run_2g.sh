cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu_final.json 2> gpurun_out/r02_bench_2gpu_final.err; tail -3 gpurun_out/r02_bench_2gpu_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_2gpu_ref.json 2>> gpurun_out/r02_bench_2gpu_final.err
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or sharded" 2>&1 | tail -2
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_2gpu_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d["e2e"]["ceiling"]["frac_of_ceiling"])
print("config3", d.get("config3")); print("model5", d.get("model5", {}).get("value"))
print(open("gpurun_out/r02_bench_2gpu_ref.json").read()[:300])
PY
