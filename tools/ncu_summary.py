#!/usr/bin/env python
"""Key metrics of one `ncu --set full` capture as text: python tools/ncu_summary.py report.ncu-rep "header line" > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__icc_request_hit_rate.pct",
    "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def main():
    rep = sys.argv[1]
    print(sys.argv[2] if len(sys.argv) > 2 else rep)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:           # one row per captured kernel launch
        m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        if len(rows) > 3 and "Kernel Name" in m:
            print("---- %s" % m["Kernel Name"][0])
        for k in KEYS:
            if k in m:
                print("%s = %s %s" % (k, m[k][0], m[k][1]))
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                print("%s = %s %s" % (h, m[h][0], m[h][1]))


if __name__ == "__main__":
    main()
