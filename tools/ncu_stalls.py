#!/usr/bin/env python
"""Stall-reason breakdown of the SASS instructions that map to a source line range of one file.
python tools/ncu_stalls.py rep.ncu-rep <file.cuh> <first_line> <last_line>"""
import collections
import csv
import subprocess
import sys


def main():
    rep, fname, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi_row = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi_row]
    txt2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    addr_line = {}
    cur = curline = h2 = None
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
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(" not in h]
    i_i, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg = collections.Counter()
    ops = collections.Counter()
    tot = inst = n = 0
    for r in rows[hi_row + 1:]:
        if not r or not r[0].startswith("0x"):
            continue
        f, l = addr_line.get(r[0], (None, None))
        if f != fname or l is None or not (lo <= l <= hi):
            continue
        n += 1
        for i, h in stall_cols:
            agg[h] += num(r[i])
        tot += num(r[i_s])
        inst += num(r[i_i])
        ops[r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]] += num(r[i_i])
    print("%s:%d-%d: %d SASS instructions, %d executed, %d stall samples" % (fname, lo, hi, n, inst, tot))
    for h, v in agg.most_common(10):
        print("  %-28s %6d  %5.1f%%" % (h, v, 100.0 * v / max(tot, 1)))
    print("  executed by opcode: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(inst, 1)) for k, v in ops.most_common(12)))


if __name__ == "__main__":
    main()
