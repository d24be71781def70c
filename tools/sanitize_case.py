#!/usr/bin/env python
"""Small invocations of every stage-kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
Covers: tendency, SSPRK33 per-stage launches, the persistent launch, a Shu-Osher and a 2N stepper, multi-chunk
columns (chunk-face exchange through shared memory), ice, Richards and heat models, uploads/downloads, budgets."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import workloads as w
lh, abi = w.lh, w.abi
cuda = lh.cuda_library()

def table(method):
    t = abi.lh_soil_stepper(); assert cuda.soil_stepper_named(method, t) == 0; return t

cases = [
    (w.coupled_workload(ncol=96, nlayer=64, seed=1), 0),
    (w.coupled_workload(ncol=64, nlayer=40, seed=2, ice=True, viscosity=lh.TemperatureDependentViscosity(), impedance=lh.IceImpedance()), 0),
    (w.coupled_workload(ncol=33, nlayer=200, seed=3, zlim=(-6.0, 0.0)), abi.LH_FLAG_GENERAL_VG),
    (w.richards_workload(ncol=70, nlayer=100, seed=4), 0),
    (w.heat_workload(ncol=40, nlayer=37, seed=5), 0),
]
for wl, flags in cases:
    for launch in (abi.LH_FLAG_STAGE_LAUNCHES, abi.LH_FLAG_PERSISTENT):
        ctx = lh.SoilContext(cuda, wl.config(flags=flags | launch))
        wl.upload(ctx)
        ctx.rhs(0.0)
        ctx.get_tendency(0 if wl.model != abi.LH_MODEL_HEAT else 2)
        ctx.step(0.0, wl.dt, 2)
        ctx.step_with(table(abi.LH_METHOD_SSPRK43), 0.0, wl.dt, 1)
        ctx.step_with(table(abi.LH_METHOD_CK2N54), 0.0, wl.dt, 1)
        b = ctx.budgets()
        s = ctx.get_state(0 if wl.model != abi.LH_MODEL_HEAT else 2)
        assert np.all(np.isfinite(s)) and np.all(np.isfinite(b)), wl.name
        ctx.close()
print("sanitize_case: ok")
