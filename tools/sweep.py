#!/usr/bin/env python
"""BASELINE config C5: vertical-resolution sweep 16..1024 layers at a fixed cell count, plus the
Richards (C3-style) and general-van-Genuchten variants.  Prints one JSON object; run under gpurun."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as graft
lh = graft.load_package()
import workloads as w

PEAK = 6458.4
cells_log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 25
steps, warm = 10, 3
out = {"cells": 1 << cells_log2, "steps": steps, "rows": []}

def run(wl, label, bytes_per_cell_step, flags=0):
    ctx = lh.SoilContext(lh.cuda_library(), wl.config(flags=flags))
    wl.upload(ctx)
    ctx.step(0.0, wl.dt, warm); ctx.sync()
    ctx.step(0.0, wl.dt, steps)
    ms, n = ctx.last_step_timing()
    v = wl.cells * steps / (ms * 1e-3)
    row = {"case": label, "ncol": wl.ncol, "nlayer": wl.nlayer, "ms_per_step": ms / steps, "cell_steps_per_s": v,
           "hbm_frac": v * bytes_per_cell_step / 1e9 / PEAK}
    out["rows"].append(row)
    print(json.dumps(row), file=sys.stderr)
    ctx.close()

for nlayer in (16, 32, 64, 128, 256, 512, 1024):
    ncol = (1 << cells_log2) // nlayer
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, zlim=(-2.0 * nlayer / 64, 0.0))
    run(wl, f"coupled n=2 {nlayer} layers", 152)
for nlayer in (16, 64, 256, 1024):
    ncol = (1 << cells_log2) // nlayer
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, zlim=(-2.0 * nlayer / 64, 0.0))
    run(wl, f"coupled general-vg {nlayer} layers", 152, flags=lh._abi.LH_FLAG_GENERAL_VG)
for nlayer in (16, 100, 1024):
    ncol = (1 << cells_log2) // nlayer
    wl = w.richards_workload(ncol=ncol, nlayer=nlayer, zlim=(-1.5 * nlayer / 100, 0.0))
    run(wl, f"richards sand {nlayer} layers", 88)
wl = w.richards_workload(ncol=1024, nlayer=100)
run(wl, "C3 HybridBox 32x32x100 richards", 88)
wl = w.richards_workload(ncol=1, nlayer=150)
run(wl, "C1 single column richards n=150", 88)
wl = w.coupled_workload(ncol=1, nlayer=64)
run(wl, "C2 single column coupled n=64", 152)
print(json.dumps(out))
