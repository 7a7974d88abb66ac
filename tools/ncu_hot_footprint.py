#!/usr/bin/env python
"""Executed instruction-cache footprint of a kernel from an ncu report with source (`--import-source on`):
SASS instructions by how often they were executed, grouped into 128-byte lines and by source function.
python tools/ncu_hot_footprint.py rep.ncu-rep [min_executions_per_cta_iteration]"""
import collections
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_static import func_ranges, ROOT  # noqa: E402


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    ranges = {f: func_ranges(os.path.join(ROOT, "gama_tts_b200", "csrc", f)) for f in ("tube_kernel_v2.cuh", "tube_kernel_v1.cuh", "tube_kernel.cuh")}
    # cuda,sass view: per source line, the SASS rows under it
    per_addr = {}
    cur, curline, hdr = None, None, None
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
            i_exec = hdr.index("Instructions Executed")
        elif hdr:
            if r[0].isdigit():
                curline = int(r[0])
            elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
                try:
                    ex = int(float(r[i_exec]))
                except (ValueError, IndexError):
                    ex = 0
                per_addr.setdefault(int(r[2], 16), (cur, curline, ex))
    addrs = sorted(per_addr)
    base = addrs[0]
    total = len(addrs)
    execd = [a for a in addrs if per_addr[a][2] > 0]
    mx = max(per_addr[a][2] for a in addrs)
    print("SASS instructions %d (%.1f KB), executed at least once %d (%.1f KB), max executions %d" %
          (total, total / 64.0, len(execd), len(execd) / 64.0, mx))
    for frac in (1e-4, 1e-3, 1e-2):
        hot = [a for a in addrs if per_addr[a][2] > mx * frac]
        lines = {(a - base) // 128 for a in hot}
        print("  executed > %.0e of max: %5d instructions, %4d 128-byte lines = %.1f KB" % (frac, len(hot), len(lines), len(lines) / 8.0))
    thr = mx * 1e-3
    by = collections.Counter()
    for a in addrs:
        f, l, ex = per_addr[a]
        if ex <= thr:
            continue
        key = f
        for name, lo, hi in ranges.get(f, []):
            if lo <= (l or 0) <= hi:
                key = f.replace("tube_kernel", "k").replace(".cuh", "") + ":" + name
        by[key] += 1
    print("--- hot instructions (> 1e-3 of max) by source function")
    for k, v in by.most_common(30):
        print("  %-36s %5d  %.1f KB" % (k, v, v / 64.0))


if __name__ == "__main__":
    main()
