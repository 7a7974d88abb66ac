cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_gpu or pcm16 or golden" > gpurun_out/r02_tests_i.log 2>&1; tail -5 gpurun_out/r02_tests_i.log
timeout 900 python bench.py > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err; tail -3 gpurun_out/r02_bench_i.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_i.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4))
print("e2e", json.dumps(d["e2e"], indent=None)[:1500])
print("config3", d.get("config3"))
print("cpu", d.get("cpu_baseline"))
PY
