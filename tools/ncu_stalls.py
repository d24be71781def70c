#!/usr/bin/env python
"""Key per-kernel metrics and warp-stall breakdown of an ncu report: python tools/ncu_stalls.py rep.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
def col(name):
    return [r[h.index(name)] for r in rr[2:]] if name in h else None
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]
for k in keys:
    v = col(k)
    if v: print(k, v)
for i, name in enumerate(h):
    if "issue_stalled" in name and name.endswith("per_issue_active.ratio") and "not_issued" not in name:
        vals = [float(r[i]) for r in rr[2:]]
        if max(vals) >= 0.05:
            print("  %-28s" % name.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                  " ".join("%6.2f" % v for v in vals))
