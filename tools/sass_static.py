#!/usr/bin/env python
"""Static SASS size per source function of tube_kernel_v1 (instruction-cache footprint), from
`nvdisasm -g -c` of the cubin inside libgtts_b200.so.  No GPU needed.  python tools/sass_static.py"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = os.environ.get("KERNEL", "tube_kernel_v2")


def func_ranges(path):
    out, cur = [], None
    for i, line in enumerate(open(path), 1):
        m = re.match(r"^(?:GTTS_DEV_NOINLINE|GTTS_DEV|__global__|inline).*?\b([A-Za-z_0-9]+)\s*\(", line)
        if m and not line.startswith(" "):
            if cur:
                out.append((cur[0], cur[1], i - 1))
            cur = (m.group(1), i)
    if cur:
        out.append((cur[0], cur[1], 10 ** 9))
    return out


def main():
    tmp = tempfile.mkdtemp()
    so = os.path.join(ROOT, "gama_tts_b200", "csrc", "libgtts_b200.so")
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    ranges = {f: func_ranges(os.path.join(ROOT, "gama_tts_b200", "csrc", f)) for f in ("tube_kernel_v2.cuh", "tube_kernel_v1.cuh", "tube_kernel.cuh")}
    counts = collections.Counter()
    lines = collections.Counter()
    lkey = None
    section, key = None, "?"
    for line in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+)", line)
        if m:
            section = m.group(1)
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            f, l = os.path.basename(m.group(1)), int(m.group(2))
            key = f
            lkey = (f, l)
            for name, a, b in ranges.get(f, []):
                if a <= l <= b:
                    key = f.replace("tube_kernel", "k").replace(".cuh", "") + ":" + name
            continue
        if section and KERNEL in section and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            counts[key] += 1
            lines[lkey] += 1
    total = sum(counts.values())
    print("%s: %d SASS instructions, %.1f KB" % (KERNEL, total, total * 16 / 1024.0))
    for k, v in counts.most_common(24):
        print("  %-36s %5d" % (k, v))
    if "--lines" in sys.argv:
        src = {}
        print("--- top source lines by static SASS instructions")
        for (f, l), v in lines.most_common(60):
            path = os.path.join(ROOT, "gama_tts_b200", "csrc", f)
            if f not in src:
                src[f] = open(path).read().splitlines() if os.path.exists(path) else []
            text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ""
            print("  %4d  %s:%d  %s" % (v, f.replace("tube_kernel", "k"), l, text[:110]))


if __name__ == "__main__":
    main()
