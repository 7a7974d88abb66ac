#!/usr/bin/env python
"""Copies the files tools/capture_final.sh left in gpurun_out/ into profiles/ and writes the text summaries of the two
`--set full` captures (needs ncu for reading the reports; no GPU)."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def run(args, out):
    with open(os.path.join(P, out), "w") as f:
        subprocess.run([sys.executable] + args, stdout=f, stderr=subprocess.STDOUT, cwd=ROOT, check=False)


def events_lines(rep, out, frames):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:events_kernel"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"] + [len(rows)]
    with open(os.path.join(P, out), "w") as f:
        f.write("events_kernel (frame pass): warp instructions per frame and stall samples by source line of events_kernel.cuh; %d frames\n" % frames)
        for a, b in zip(his[:-1], his[1:]):
            hdr = rows[a]
            i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
            lines = [(int(r[0]), r[1].strip()[:120], num(r[i_samp]), num(r[i_inst])) for r in rows[a + 1:b] if r and r[0].isdigit()]
            ti, ts = sum(l[3] for l in lines), sum(l[2] for l in lines)
            if ti < 1e8:
                continue
            f.write("total: %.1f warp instructions per frame, %d stall samples\n--- by instructions\n" % (ti / frames, ts))
            for l in sorted(lines, key=lambda x: -x[3])[:30]:
                f.write("%4d %6.2f instr/frame  samples %4.1f%%  %s\n" % (l[0], l[3] / frames, 100 * l[2] / max(ts, 1), l[1]))
            f.write("--- by stall samples\n")
            for l in sorted(lines, key=lambda x: -x[2])[:20]:
                f.write("%4d %6.2f instr/frame  samples %4.1f%%  %s\n" % (l[0], l[3] / frames, 100 * l[2] / max(ts, 1), l[1]))


def main():
    for a, b in (("ncu_r02_traffic_bench.csv", "ncu_r02_traffic_bench.csv"), ("ncu_r02_launches_bench.csv", "ncu_r02_launches_bench.csv"),
                 ("bench_r02_final_1gpu.json", "bench_r02_1gpu.json"), ("bench_r02_final_reference_arm.json", "bench_r02_1gpu_reference_arm.json")):
        shutil.copy(os.path.join(G, a), os.path.join(P, b))
    rows = [r for r in csv.reader(open(os.path.join(G, "ncu_r02_traffic_bench.csv"))) if len(r) > 14 and r[12].startswith("dram__bytes")]
    vals = {r[12]: float(r[14].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[13]] for r in rows}
    json.dump({"source": "profiles/ncu_r02_traffic_bench.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on one %s launch of `python bench.py --steps 1 "
                         "--warmup 3 --no-cpu-baseline --no-config3 --no-model5` (BASELINE config 2), same build as profiles/ncu_r02_launches_bench.csv and "
                         "profiles/bench_r02_1gpu.json (tools/capture_final.sh)" % rows[0][4].split("(")[0],
               "dram_bytes_read": int(vals["dram__bytes_read.sum"]), "dram_bytes_write": int(vals["dram__bytes_write.sum"])},
              open(os.path.join(P, "ncu_r02_traffic_bench.json"), "w"), indent=1)
    v1, ev = os.path.join(G, "v1_final.ncu-rep"), os.path.join(G, "events_final.ncu-rep")
    run(["tools/ncu_summary.py", v1, "tube_kernel_v1, final round-2 build, tools/profile_run.py --utts 1036 --frames 60, ncu --set full --clock-control none"], "ncu_r02_v1_summary.txt")
    run(["tools/ncu_roles.py", v1], "ncu_r02_v1_roles.txt")
    run(["tools/ncu_lines.py", v1, "30"], "ncu_r02_v1_lines.txt")
    run(["tools/ncu_smem.py", v1], "ncu_r02_v1_smem.txt")
    run(["tools/ncu_summary.py", ev, "events_drift_kernel + events_kernel, final round-2 build, tools/events_profile_run.py (37,888 chunks of 16 postures: 2.42 M events -> 20.3 M frames), ncu --set full"],
        "ncu_r02_events_summary.txt")
    events_lines(ev, "ncu_r02_events_lines.txt", 20293464)
    # launch list digest
    rows = [r for r in csv.reader(open(os.path.join(P, "ncu_r02_launches_bench.csv"))) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        ms = float(r[14].replace(",", "")) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3}[r[13]]
        k = r[4].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += ms
    with open(os.path.join(P, "ncu_r02_launches_digest.txt"), "w") as f:
        f.write("profiles/ncu_r02_launches_bench.csv by kernel (python bench.py --steps 2 --warmup 3 --no-cpu-baseline under ncu: cold-cache, serialised times)\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-70s %5d launches %12.3f ms\n" % (k[:70], n, ms))
    print(open(os.path.join(P, "ncu_r02_launches_digest.txt")).read())


if __name__ == "__main__":
    main()
