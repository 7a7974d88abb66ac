#!/usr/bin/env python
"""Per-source-line summary of an ncu report: `python tools/ncu_lines.py report.ncu-rep [topN]`.
Shares of executed warp instructions and of stall samples per CUDA source line."""
import csv
import subprocess
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    i_samp, i_inst = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = []
    for r in rows[hi + 1:]:
        if r and r[0].isdigit():
            lines.append((int(r[0]), r[1].strip()[:120], num(r[i_samp]), num(r[i_inst])))
    ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
    print("total stall samples %d, total warp instructions %d" % (ts, ti))
    print("--- top %d lines by executed instructions" % top)
    for l in sorted(lines, key=lambda x: -x[3])[:top]:
        print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (l[0], 100.0 * l[3] / ti, 100.0 * l[2] / ts, l[1]))
    print("--- top %d lines by stall samples" % top)
    for l in sorted(lines, key=lambda x: -x[2])[:top]:
        print("%5d  inst %5.1f%%  samples %5.1f%%  %s" % (l[0], 100.0 * l[3] / ti, 100.0 * l[2] / ts, l[1]))


if __name__ == "__main__":
    main()
