#!/usr/bin/env python
"""SASS instructions of one device function with their stall samples, in program order.
python tools/ncu_sass.py rep.ncu-rep <file.cuh> <first_line> <last_line> [min_samples]"""
import csv
import subprocess
import sys


def main():
    rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    min_s = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi_row = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi_row]
    # need the line mapping: second pass with cuda,sass
    txt2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    addr_line = {}
    cur, curline, h2 = None, None, None
    for r in csv.reader(txt2.splitlines()):
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            h2 = r
        elif r and h2:
            if r[0].isdigit():
                curline = int(r[0])
            elif r[0] == "" and len(r) > 3 and r[2].startswith("0x"):
                addr_line.setdefault(r[2], (cur, curline))

    def num(x):
        try:
            return int(float(x))
        except ValueError:
            return 0
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
    for r in rows[hi_row + 1:]:
        if not r or not r[0].startswith("0x"):
            continue
        f, l = addr_line.get(r[0], (None, None))
        if f != fname or l is None or not (lo <= l <= hi):
            continue
        s = num(r[i_s])
        if s < min_s:
            continue
        top = sorted(((num(r[i]), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        print("%4d %-64s samp %5d exec %8d  %s" % (l, r[1].strip()[:64], s, num(r[i_i]),
                                                    " ".join("%s:%d" % (h, n) for n, h in top if n)))


if __name__ == "__main__":
    main()
