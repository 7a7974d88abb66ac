#!/usr/bin/env python
"""Shared-memory wavefronts (the LSU pipe every LDS / STS / shuffle queues for) by source function and by
source line, with the excess over the ideal (bank conflicts).  python tools/ncu_smem.py rep.ncu-rep"""
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
    by_fn, by_line, ideal_fn = collections.Counter(), collections.Counter(), collections.Counter()
    shfl = collections.Counter()
    cur = curline = hdr = None
    seen = set()
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
            i_w, i_i, i_x = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
        elif hdr:
            if r[0].isdigit():
                curline = int(r[0])
            elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
                if r[2] in seen:
                    continue
                seen.add(r[2])

                def num(x):
                    try:
                        return int(float(x))
                    except ValueError:
                        return 0
                w, wi = num(r[i_w]), num(r[i_i])
                key = cur
                for name, lo, hi in ranges.get(cur, []):
                    if lo <= (curline or 0) <= hi:
                        key = cur.replace("tube_kernel", "k").replace(".cuh", "") + ":" + name
                if "SHFL" in r[3]:
                    shfl[key] += num(r[i_x])
                if w:
                    by_fn[key] += w
                    ideal_fn[key] += wi
                    by_line[(cur, curline)] += w
    tot = sum(by_fn.values())
    print("shared-memory wavefronts %d (ideal %d); SHFL instructions %d" % (tot, sum(ideal_fn.values()), sum(shfl.values())))
    for k, v in by_fn.most_common(16):
        print("  %-34s %5.1f%%  (ideal %5.1f%%)  shfl %d" % (k, 100.0 * v / tot, 100.0 * ideal_fn[k] / tot, shfl[k]))
    src = {}
    print("--- top lines")
    for (f, l), v in by_line.most_common(24):
        path = os.path.join(ROOT, "gama_tts_b200", "csrc", f)
        if f not in src:
            src[f] = open(path).read().splitlines() if os.path.exists(path) else []
        text = src[f][l - 1].strip() if l and 0 < l <= len(src[f]) else ""
        print("  %5.1f%%  %s:%s  %s" % (100.0 * v / tot, f.replace("tube_kernel", "k"), l, text[:100]))


if __name__ == "__main__":
    main()
