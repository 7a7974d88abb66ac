cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --no-cpu-baseline --no-config3 --steps 20 > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err; tail -3 gpurun_out/r02_bench_k.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_k.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4))
e=d["e2e"]; print("e2e", round(e["value"]), round(e["ms_per_step"],2), {k:(round(v["value"]),round(v["ms_per_step"],2)) for k,v in e["variants"].items()})
PY
