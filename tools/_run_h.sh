cd $GRAFT_REPO_ROOT
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_h_v2.json 2> gpurun_out/r02_bench_h_v2.err
GTTS_KERNEL=v1 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_h_v1.json 2> gpurun_out/r02_bench_h_v1.err
python - <<'PY'
import json
for k in ("v2","v1"):
    d=json.load(open("gpurun_out/r02_bench_h_%s.json"%k))
    print(k, round(d["value"]), round(d["ms_per_step"],2), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), round(d["roofline"]["frac"],4))
PY
