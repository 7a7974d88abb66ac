#!/usr/bin/env python
"""Opcode mix of the SASS attributed to a source line range of tube_kernel_v1.cuh (offline, nvdisasm).
python tools/sass_loop_ops.py '<text marking first line>' '<text marking last line>'"""
import collections, glob, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "gama_tts_b200/csrc/tube_kernel_v1.cuh")).read().splitlines()
l0 = [i + 1 for i, t in enumerate(src) if sys.argv[1] in t][0]
l1 = [i + 1 for i, t in enumerate(src) if sys.argv[2] in t][0]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "gama_tts_b200/csrc/libgtts_b200.so")], cwd=tmp, capture_output=True)
cubin = max(glob.glob(os.path.join(tmp, "*.cubin")), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
ops, n, section, cur = collections.Counter(), 0, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        section = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if section and "tube_kernel_v1" in section and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", ln) and cur and cur[0] == "tube_kernel_v1.cuh" and l0 <= cur[1] <= l1:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            ops[m.group(1).split(".")[0]] += 1; n += 1
print("lines %d-%d: %d SASS instructions" % (l0, l1, n))
print(ops.most_common())
