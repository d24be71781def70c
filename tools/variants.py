#!/usr/bin/env python
"""Variant table of the fused RHS+stage path on one B200: the kernel specialisations real soils
hit — general van Genuchten n, ice + impedance + viscosity, Richards, heterogeneous columns — and the small / tall
BASELINE configs, each timed in the same process with CUDA events on the ctx stream.

    python tools/variants.py [--quick] [--only name,name] [--steps K] [--reps R]

Prints one JSON object (rows: cell-steps/s, ms/step, contract and on-wire roofline fractions, launched kernel flags).
bench.py imports `measure` / `variant_specs` for its `extra.variants` table.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


# Contract bytes per cell-step (θ_i read in every stage, never written) and what a launched variant
# really moves: the !ICE kernels do not read θ_i (-8 B per stage); Richards reads the prescribed T only when the
# viscosity factor needs it (+8 B per stage, outside the contract figure).
def bytes_per_cell_step(model: str, ice: bool, reads_T: bool = False, cell_params: bool = False):
    contract = {"coupled": 152, "richards": 88, "heat": 88}[model]
    wire = contract - (0 if ice else 24) + (24 if reads_T else 0)
    if cell_params:
        wire += 3 * 72                  # nine per-cell parameter fields read in every stage, outside the contract figure
    return contract, wire


def variant_specs(lh, w, quick=False):
    """name -> (builder, model, extra).  Builders return a Workload; `post(ctx, wl)` applies per-column parameters."""
    A = lh._abi
    C4 = (1 << 20, 64) if not quick else (1 << 16, 64)
    RS = (655360, 100) if not quick else (40960, 100)
    cells = C4[0] * C4[1]

    def het(ctx, wl):
        rng = np.random.default_rng(11)
        n_ = wl.ncol
        ctx.set_column_params(nu=wl.params.nu * rng.uniform(0.98, 1.15, n_), theta_r=rng.uniform(0.0, 0.02, n_),
                              vg_n=rng.uniform(1.5, 3.5, n_), vg_alpha=wl.params.vg_alpha * rng.uniform(0.5, 2.0, n_),
                              Ksat=wl.params.Ksat * 10.0 ** rng.uniform(-1.0, 1.0, n_))

    visc = lambda: lh.TemperatureDependentViscosity()
    imp = lambda: lh.IceImpedance()
    S = {}
    S["coupled_n2"] = dict(make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1]), model="coupled")
    S["coupled_general_n"] = dict(make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1]), model="coupled",
                                  flags=A.LH_FLAG_GENERAL_VG)
    S["coupled_ice_n2"] = dict(make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1], ice=True), model="coupled", ice=True)
    S["coupled_ice_impedance_viscosity"] = dict(
        make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1], ice=True, viscosity=visc(), impedance=imp()),
        model="coupled", ice=True, flags=A.LH_FLAG_GENERAL_VG)
    S["coupled_het"] = dict(make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1]), model="coupled", post=het)
    S["richards_sand"] = dict(make=lambda: w.richards_workload(ncol=RS[0], nlayer=RS[1]), model="richards")
    S["richards_ice_impedance_viscosity"] = dict(
        make=lambda: w.richards_workload(ncol=RS[0], nlayer=RS[1], ice=True, viscosity=visc(), impedance=imp()),
        model="richards", ice=True, reads_T=True)
    S["richards_het"] = dict(make=lambda: w.richards_workload(ncol=RS[0], nlayer=RS[1]), model="richards", post=het)

    # N4 of SURVEY §8(f): per-column HEAT parameters (HETH variants) and per-cell (layered) hydraulic parameters (CELLP variants).
    # A quarter of the C4 columns: the per-cell derivation of nine fields runs on the host, once, at 0.2 us per cell.
    Q4 = (C4[0] // 4, C4[1])

    def het_heat(ctx, wl):
        het(ctx, wl)
        rng = np.random.default_rng(12)
        n_ = wl.ncol
        ctx.set_column_heat_params(rho_c_ds=wl.params.rho_c_ds * rng.uniform(0.8, 1.2, n_),
                                   kappa_sat_unfrozen=wl.params.kappa_sat_unfrozen * rng.uniform(0.8, 1.2, n_),
                                   nu_ss_om=rng.uniform(0.0, 0.2, n_))

    def cell_params(ctx, wl):
        rng = np.random.default_rng(13)
        shape = (wl.ncol, wl.nlayer)
        ctx.set_cell_params(vg_n=rng.uniform(1.5, 3.5, shape), Ksat=wl.params.Ksat * 10.0 ** rng.uniform(-1.0, 1.0, shape))

    def atmos(ctx, wl):
        # PrescribedAtmosForcing at the top face (N3): one Monin-Obukhov solve per column before every stage launch
        ep = lh.EarthParameterSet()
        a = A.lh_soil_atmos()
        a.u_atm, a.theta_atm, a.z_atm, a.theta_scale, a.rho_a_sfc, a.q_atm = 2.0, 288.0, 2.0, 290.0, 1.17, 0.006
        a.R_v, a.R_d, a.grav, a.cp_d, a.cp_v, a.LH_v0 = ep.R_v, ep.R_d, ep.grav, ep.cp_d, ep.cp_v, ep.LH_v0
        a.press_triple, a.T_triple, a.von_karman = ep.press_triple, ep.T_triple, ep.von_karman_const
        a.Pr_0, a.a_m, a.a_h = ep.Pr_0, ep.a_m, ep.a_h
        ctx.set_atmos_forcing(a)

    S["coupled_atmos_forcing"] = dict(make=lambda: w.coupled_workload(ncol=C4[0], nlayer=C4[1]), model="coupled", post=atmos)
    S["coupled_het_heat_params"] = dict(make=lambda: w.coupled_workload(ncol=Q4[0], nlayer=Q4[1]), model="coupled", post=het_heat)
    S["coupled_cell_params"] = dict(make=lambda: w.coupled_workload(ncol=Q4[0], nlayer=Q4[1]), model="coupled", post=cell_params,
                                    cell_params=True)
    S["richards_cell_params"] = dict(make=lambda: w.richards_workload(ncol=Q4[0], nlayer=Q4[1], zlim=(-0.96, 0.0)), model="richards",
                                     post=cell_params, cell_params=True)
    S["C5_coupled_16_layers"] = dict(make=lambda: w.coupled_workload(ncol=cells // 16, nlayer=16, zlim=(-0.5, 0.0)), model="coupled")
    S["C5_coupled_1024_layers"] = dict(make=lambda: w.coupled_workload(ncol=cells // 1024, nlayer=1024, zlim=(-32.0, 0.0)), model="coupled")
    S["C5_richards_16_layers"] = dict(make=lambda: w.richards_workload(ncol=cells // 16, nlayer=16, zlim=(-0.24, 0.0)), model="richards")
    S["C3_hybridbox_32x32x100_richards"] = dict(make=lambda: w.richards_workload(ncol=1024, nlayer=100), model="richards", small=True)
    S["C1_single_column_richards_150"] = dict(make=lambda: w.richards_workload(ncol=1, nlayer=150), model="richards", small=True)
    S["C2_single_column_coupled_64"] = dict(make=lambda: w.coupled_workload(ncol=1, nlayer=64), model="coupled", small=True)
    return S


def measure(lh, spec, steps=20, warmup=3, reps=3, device=0, min_seconds=1.0):
    """Median of the timed blocks of `steps` SSPRK33 steps (CUDA events on the ctx stream): at least `reps` blocks, repeated back
    to back until `min_seconds` of device time have been measured (the board settles at its power-capped clock within a few
    hundred milliseconds of fp64 load: a handful of 20-step blocks alone would be a burst-clock figure), at most 400 blocks."""
    wl = spec["make"]()
    wl.device = device
    ctx = lh.SoilContext(lh.cuda_library(), wl.config(flags=spec.get("flags", 0)))
    try:
        if spec.get("post"):
            spec["post"](ctx, wl)
        wl.upload(ctx)
        k = steps * (50 if spec.get("small") else 1)
        ctx.step(0.0, wl.dt, warmup * (50 if spec.get("small") else 1))
        ctx.sync()
        ms_all = []
        launches = 0
        while len(ms_all) < reps or (sum(ms_all) < min_seconds * 1e3 and len(ms_all) < 400):
            ctx.step(0.0, wl.dt, k)
            ms, launches = ctx.last_step_timing()
            ms_all.append(ms)
        ms = float(np.median(ms_all))
        bud = ctx.budgets()
        info = ctx.kernel_info() if hasattr(ctx, "kernel_info") else None
    finally:
        ctx.close()
    contract, wire = bytes_per_cell_step(spec["model"], spec.get("ice", False), spec.get("reads_T", False), spec.get("cell_params", False))
    v = wl.cells * k / (ms * 1e-3)
    peak = peak_gbs()
    row = {"ncol": wl.ncol, "nlayer": wl.nlayer, "cell_steps_per_s": v, "ms_per_step": ms / k,
           "frac_contract": v * contract / 1e9 / peak, "frac_on_wire": v * wire / 1e9 / peak,
           "bytes_contract": contract, "bytes_on_wire": wire, "launches_per_block": int(launches), "steps_per_block": k,
           "blocks": len(ms_all), "device_seconds": float(sum(ms_all)) * 1e-3,
           "cell_steps_per_s_first_block": wl.cells * k / (ms_all[0] * 1e-3),
           "finite": bool(np.all(np.isfinite(bud)))}
    if info:
        row["kernel"] = info
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import __graft_entry__ as graft

    lh = graft.load_package()
    import workloads as w

    specs = variant_specs(lh, w, quick=args.quick)
    only = [s for s in args.only.split(",") if s]
    out = {"peak_gbs": peak_gbs(), "rows": {}}
    for name, spec in specs.items():
        if only and name not in only:
            continue
        row = measure(lh, spec, steps=args.steps, reps=args.reps)
        out["rows"][name] = row
        print(name, json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
