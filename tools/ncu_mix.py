#!/usr/bin/env python
"""Executed-instruction mix per cell from an ncu report (source page): python tools/ncu_mix.py rep cells"""
import collections, csv, subprocess, sys
rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = None; hdr = None; data = collections.defaultdict(list)
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name": kern = r[1]; continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and hdr and len(r) == len(hdr): data[kern].append(r)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
def col(name):
    return [r[h.index(name)] for r in rr[2:]] if name in h else None
for name in ("Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
             "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct"):
    print(name, col(name))
for k, rs in data.items():
    iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples")
    ops = collections.Counter(); samp = collections.Counter(); tot = 0
    for r in rs:
        t = r[iS].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        n = int(r[iE]); ops[op] += n; tot += n; samp[op] += int(r[iSm])
    inst = float(col("smsp__inst_executed.sum")[list(data).index(k)])
    scale = inst / tot     # source-page counts are per-issue replicated; normalise to smsp__inst_executed
    wc = cells / 32
    print(k, "warp-instructions per cell:", inst / wc)
    for op, n in ops.most_common(28):
        print(f"  {op:10s} {n * scale / wc:8.1f}   stall samples {samp[op]}")
