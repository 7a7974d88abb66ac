#!/usr/bin/env python
"""Per-function summary of an ncu report for tube_kernel_v1: share of stall samples / instructions per
device function (by source line range read from the .cuh).  python tools/ncu_roles.py rep.ncu-rep"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def func_ranges(path):
    out, cur = [], None
    for i, line in enumerate(open(path), 1):
        m = re.match(r"^(?:GTTS_DEV|__global__|inline|template).*?\b([A-Za-z_0-9]+)\s*\(", line)
        if m and not line.startswith(" "):
            if cur:
                out.append((cur[0], cur[1], i - 1))
            cur = (m.group(1), i)
    if cur:
        out.append((cur[0], cur[1], 10 ** 9))
    return out


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    ranges = {f: func_ranges(os.path.join(ROOT, "gama_tts_b200", "csrc", f)) for f in ("tube_kernel_v2.cuh", "tube_kernel_v1.cuh", "tube_kernel.cuh")}
    agg = collections.defaultdict(lambda: [0, 0])
    cur, hdr = None, None

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif r and r[0].isdigit() and hdr:
            line = int(r[0])
            key = cur
            for name, a, b in ranges.get(cur, []):
                if a <= line <= b:
                    key = cur.replace("tube_kernel", "k").replace(".cuh", "") + ":" + name
            agg[key][0] += num(r[hdr.index("# Samples")])
            agg[key][1] += num(r[hdr.index("Instructions Executed")])
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print("total samples %d, instructions %d" % (ts, ti))
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-40s samples %5.1f%%  inst %5.1f%%" % (k, 100.0 * v[0] / ts, 100.0 * v[1] / ti))


if __name__ == "__main__":
    main()
