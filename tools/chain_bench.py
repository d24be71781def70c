#!/usr/bin/env python
"""Chained vs whole-grid stage launches (LH_FLAG_NO_CHAIN) on one GPU: python tools/chain_bench.py [steps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as graft
lh = graft.load_package()
import workloads as w
A = lh._abi
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
out = {}
for name, mk in {
    "coupled_1M_x64": lambda: w.coupled_workload(ncol=1 << 20, nlayer=64),
    "coupled_131072_x64": lambda: w.coupled_workload(ncol=131072, nlayer=64),
    "coupled_262144_x64": lambda: w.coupled_workload(ncol=262144, nlayer=64),
    "richards_655360_x100": lambda: w.richards_workload(ncol=655360, nlayer=100),
    "richards_81920_x100": lambda: w.richards_workload(ncol=81920, nlayer=100),
}.items():
    wl = mk()
    row = {}
    for tag, fl in (("chained", A.LH_FLAG_STAGE_LAUNCHES), ("whole_grid", A.LH_FLAG_STAGE_LAUNCHES | A.LH_FLAG_NO_CHAIN)):
        ctx = lh.SoilContext(lh.cuda_library(), wl.config(flags=fl))
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 5); ctx.sync()
        ms = []
        for _ in range(5):
            ctx.step(0.0, wl.dt, steps)
            ms.append(ctx.last_step_timing()[0])
        row[tag] = {"ms_per_step": float(np.median(ms)) / steps, "cell_steps_per_s": wl.cells * steps / (float(np.median(ms)) * 1e-3)}
        ctx.close()
    row["speedup"] = row["whole_grid"]["ms_per_step"] / row["chained"]["ms_per_step"]
    out[name] = row
    print(name, json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps(out))
