cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_j.json 2> gpurun_out/r02_bench_j.err; tail -3 gpurun_out/r02_bench_j.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_j.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4))
e=d["e2e"]; print("e2e", round(e["value"]), round(e["ms_per_step"],2), {k:(round(v["value"]),round(v["ms_per_step"],2)) for k,v in e["variants"].items()}, e["ceiling"])
c=d.get("config3"); print("config3", round(c["value"]), round(c["ms"],1), round(c["roofline_frac"],4))
PY
