#!/usr/bin/env python
"""Generates tests/golden/soil_golden.json — the committed golden vectors of the soil RHS + SSPRK33 path.

The reference (Julia) cannot run in this image and keeps no stored arrays of its own, so the vectors come
from two sources, both recorded in the file:

  * ``known_answers``: the literal expectations of the reference's own tests for this path (file:line
    quoted per entry) — closure values, `k_dry`, and the one-RHS known answer of test/SoilModel/coupled.jl.
    They are typed in here from the reference tests, not computed by any restatement.
  * ``cases``: dense input -> tendency / face flux / state-after-3-steps vectors produced by
    oracle/np_soil.py, the pure-Python restatement that is written line by line next to the Julia source
    (independent of the C oracle and of the CUDA kernels, which are both checked against this file).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import workloads as w  # noqa: E402
import np_soil  # noqa: E402

lh, abi = w.lh, w.abi
D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
NSTEPS = 3

PARAM_FIELDS = [name for name, _ in abi.lh_soil_params._fields_]


def params_dict(p):
    return {k: getattr(p, k) for k in PARAM_FIELDS}


def case_of(name, wl, ncol):
    bcv = [wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]
    aux_T = wl.aux_T if wl.aux_T is not None else np.full(wl.nlayer, 288.0)
    top, bottom = (wl.top[0], wl.top[2]), (wl.bottom[0], wl.bottom[2])
    rec = {"name": name, "model": wl.model, "nlayer": wl.nlayer, "ncol": ncol, "zmin": wl.zmin, "zmax": wl.zmax,
           "dt": wl.dt, "nsteps": NSTEPS, "params": params_dict(wl.params), "top": list(wl.top), "bottom": list(wl.bottom),
           "aux_T": None if wl.aux_T is None else [float(x) for x in wl.aux_T], "columns": []}
    for c in range(ncol):
        th, ti = wl.fields[0][c], wl.fields[1][c]
        re = wl.fields.get(2, np.zeros_like(wl.fields[0]))[c]

        def rhs(u, stage):
            dth, dti, dre, _, _ = np_soil.column_rhs(wl.params, wl.model, wl.zmin, wl.zmax, top, bottom, bcv,
                                                     list(u[0]), list(u[1]), list(u[2]), list(aux_T))
            return dth, dti, dre

        dth, dti, dre, Fw, Fe = np_soil.column_rhs(wl.params, wl.model, wl.zmin, wl.zmax, top, bottom, bcv,
                                                   list(th), list(ti), list(re), list(aux_T))
        u = (th.copy(), ti.copy(), re.copy())
        for _ in range(NSTEPS):
            u = np_soil.ssprk33_step(rhs, u, wl.dt)
        rec["columns"].append({
            "theta_l": th.tolist(), "theta_i": ti.tolist(), "rho_e_int": re.tolist(),
            "d_theta_l": dth.tolist(), "d_rho_e_int": dre.tolist(), "Fw": Fw.tolist(), "Fe": Fe.tolist(),
            "theta_l_after": np.asarray(u[0]).tolist(), "rho_e_int_after": np.asarray(u[2]).tolist(),
        })
    return rec


def cases():
    visc, imp = lh.TemperatureDependentViscosity(), lh.IceImpedance()
    out = []
    out.append(case_of("coupled_dirichlet_top_free_drainage", w.coupled_workload(ncol=3, nlayer=20, seed=101), 3))
    out.append(case_of("coupled_flux_top_dirichlet_bottom",
                       w.coupled_workload(ncol=2, nlayer=16, seed=102, top=(F, 2.0, F, -1e-8), bottom=(D, 281.0, D, 0.3)), 2))
    out.append(case_of("coupled_ice_viscosity_impedance",
                       w.coupled_workload(ncol=2, nlayer=12, seed=103, ice=True, viscosity=visc, impedance=imp), 2))
    out.append(case_of("coupled_64_layers", w.coupled_workload(ncol=1, nlayer=64, seed=104), 1))
    out.append(case_of("richards_sand_dirichlet_top_free_drainage", w.richards_workload(ncol=3, nlayer=30, seed=105), 3))
    out.append(case_of("richards_flux_top_dirichlet_bottom",
                       w.richards_workload(ncol=2, nlayer=25, seed=106, top=(N, 0.0, F, 1e-7), bottom=(N, 0.0, D, 0.12)), 2))
    out.append(case_of("richards_ice_viscosity_impedance",
                       w.richards_workload(ncol=2, nlayer=12, seed=107, ice=True, viscosity=visc, impedance=imp), 2))
    out.append(case_of("heat_dirichlet_both", w.heat_workload(ncol=2, nlayer=24, seed=108), 2))
    out.append(case_of("heat_flux_both_ice",
                       w.heat_workload(ncol=2, nlayer=10, seed=109, ice=True, top=(F, 4.0, N, 0.0), bottom=(F, -4.0, N, 0.0)), 2))
    return out


# Coefficient tables of the explicit low-storage steppers, typed here from the literature (NOT read from the libraries
# under test): Shu & Osher 1988 (SSPRK22, SSPRK33), Spiteri & Ruuth 2002 (SSPRK43), Carpenter & Kennedy 1994 (2N54).
STEPPERS = {
    "Euler": ("shu_osher", [0.0], [1.0], [1.0]),
    "SSPRK22": ("shu_osher", [0.0, 0.5], [1.0, 0.5], [1.0, 0.5]),
    "SSPRK33": ("shu_osher", [0.0, 0.75, 1.0 / 3.0], [1.0, 0.25, 2.0 / 3.0], [1.0, 0.25, 2.0 / 3.0]),
    "SSPRK43": ("shu_osher", [0.0, 0.0, 2.0 / 3.0, 0.0], [1.0, 1.0, 1.0 / 3.0, 1.0], [0.5, 0.5, 1.0 / 6.0, 0.5]),
    "CarpenterKennedy2N54": ("2n",
                             [0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                              -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0],
                             [1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                              1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                              2277821191437.0 / 14882151754819.0], None),
}


def stepper_cases():
    """State of one coupled column after 2 steps of every stepper (recurrences written out here over np_soil's RHS)."""
    wl = w.coupled_workload(ncol=1, nlayer=16, seed=111)
    bcv = [wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]
    top, bottom = (wl.top[0], wl.top[2]), (wl.bottom[0], wl.bottom[2])

    def rhs(u):
        dth, dti, dre, _, _ = np_soil.column_rhs(wl.params, wl.model, wl.zmin, wl.zmax, top, bottom, bcv,
                                                 list(u[0]), list(u[1]), list(u[2]), [288.0] * wl.nlayer)
        return [dth, dti, dre]

    out = {"nlayer": wl.nlayer, "zmin": wl.zmin, "zmax": wl.zmax, "dt": wl.dt, "nsteps": 2, "params": params_dict(wl.params),
           "top": list(wl.top), "bottom": list(wl.bottom), "theta_l": wl.fields[0][0].tolist(), "theta_i": wl.fields[1][0].tolist(),
           "rho_e_int": wl.fields[2][0].tolist(), "methods": {}}
    for name, (kind, a, b, g) in STEPPERS.items():
        u = [wl.fields[k][0].copy() for k in range(3)]
        for _ in range(2):
            if kind == "shu_osher":
                u0, v = [x.copy() for x in u], [x.copy() for x in u]
                for i in range(len(a)):
                    k = rhs(v)
                    v = [a[i] * x0 + b[i] * vi + (g[i] * wl.dt) * ki for x0, vi, ki in zip(u0, v, k)]
                u = v
            else:
                r = [np.zeros_like(x) for x in u]
                for i in range(len(a)):
                    k = rhs(u)
                    r = [wl.dt * ki if i == 0 else a[i] * ri + wl.dt * ki for ri, ki in zip(r, k)]
                    u = [ui + b[i] * ri for ui, ri in zip(u, r)]
        out["methods"][name] = {"theta_l_after": np.asarray(u[0]).tolist(), "rho_e_int_after": np.asarray(u[2]).tolist()}
    return out


def column_param_case():
    """Three Richards columns with different (nu, theta_r, n, alpha, Ksat): each is an independent np_soil column."""
    wl = w.richards_workload(ncol=3, nlayer=20, seed=112)
    cp = {"nu": [0.287, 0.35, 0.42], "theta_r": [0.075, 0.02, 0.0], "vg_n": [3.96, 1.8, 2.4], "vg_alpha": [2.7, 1.1, 3.6],
          "Ksat": [34 / 3600 / 100, 2.0e-6, 5.0e-5]}
    p0 = wl.params
    S = (wl.fields[0] - p0.theta_r) / (p0.nu - p0.theta_r)
    rec = {"nlayer": wl.nlayer, "zmin": wl.zmin, "zmax": wl.zmax, "dt": wl.dt, "nsteps": NSTEPS, "params": params_dict(p0),
           "top": list(wl.top), "bottom": list(wl.bottom), "column_params": cp, "columns": []}
    bcv = [wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]
    top, bottom = (wl.top[0], wl.top[2]), (wl.bottom[0], wl.bottom[2])
    import copy
    for c in range(3):
        p = copy.copy(p0)
        p.nu, p.theta_r, p.vg_n, p.vg_alpha, p.Ksat = (cp[k][c] for k in ("nu", "theta_r", "vg_n", "vg_alpha", "Ksat"))
        p.vg_m = 1.0 - 1.0 / p.vg_n
        th = p.theta_r + S[c] * (p.nu - p.theta_r)
        ti, re = np.zeros(wl.nlayer), np.zeros(wl.nlayer)

        def rhs(u, stage, p=p):
            dth, dti, dre, _, _ = np_soil.column_rhs(p, wl.model, wl.zmin, wl.zmax, top, bottom, bcv,
                                                     list(u[0]), list(u[1]), list(u[2]), [288.0] * wl.nlayer)
            return dth, dti, dre

        dth, _, _, Fw, _ = np_soil.column_rhs(p, wl.model, wl.zmin, wl.zmax, top, bottom, bcv, list(th), list(ti), list(re),
                                              [288.0] * wl.nlayer)
        u = (th.copy(), ti.copy(), re.copy())
        for _ in range(NSTEPS):
            u = np_soil.ssprk33_step(rhs, u, wl.dt)
        rec["columns"].append({"theta_l": th.tolist(), "d_theta_l": dth.tolist(), "Fw": Fw.tolist(),
                               "theta_l_after": np.asarray(u[0]).tolist()})
    return rec


# The reference's own expectations for this path, typed from its tests (paths relative to /root/reference).
KNOWN_ANSWERS = [
    {"what": "effective_saturation(0.5, [0.25, 0.5, 0.75], 0.0)", "ref": "test/SoilModel/test_water_parameterizations.jl:10-13",
     "fn": "effective_saturation", "args": [[0.5, 0.25, 0.0], [0.5, 0.5, 0.0], [0.5, 0.75, 0.0]], "expect": [0.5, 1.0, 1.5], "rtol": 0.0},
    {"what": "pressure_head saturated branch: (0.5 - 0.4) / 1e-2", "ref": "test/SoilModel/test_water_parameterizations.jl:24-26",
     "fn": "pressure_head_saturated", "args": [[0.5, 0.4, 1e-2]], "expect": [10.0], "rtol": 1e-6},
    {"what": "IceImpedance(Omega = 7): impedance_factor(f_i = 1) = 10^-7", "ref": "test/SoilModel/test_water_parameterizations.jl:40-41",
     "fn": "impedance_factor", "args": [[7.0, 1.0]], "expect": [1e-7], "rtol": 1e-6},
    {"what": "hydraulic_conductivity clamps to Ksat for S >= 1", "ref": "test/SoilModel/test_water_parameterizations.jl:30-36",
     "fn": "hydraulic_conductivity_over_Ksat", "args": [[1.0], [1.5]], "expect": [1.0, 1.0], "rtol": 0.0},
    {"what": "k_dry(nu = 0.495, kappa_solid = 8, rho_p = 2700, kappa_dry_parameter = 0.053, K_therm = 0.024)",
     "ref": "test/SoilModel/heat_test_interface.jl:7", "fn": "k_dry", "args": [[0.495, 8.0, 2700.0, 0.053, 0.024]],
     "expect": [0.43314518988433487], "rtol": 0.0},
    {"what": "one RHS of the coupled model at default ICs: n = 20, z in [-2, 0], theta_l = 0.25 (S = 0.5), zero fluxes: "
             "d theta_l = +K/dz at the bottom cell, -K/dz at the top cell, 0 inside; d rho_e_int = 0; d theta_i = 0",
     "ref": "test/SoilModel/coupled.jl:196-234", "fn": "coupled_rhs_known_answer", "args": [],
     "expect": {"K": 1.5618205801102845e-09, "dz": 0.1}, "rtol": 1e-12},
]


def main():
    doc = {
        "about": "golden vectors of the soil RHS + SSPRK33 path; see tests/golden/make_golden.py",
        "generator": "oracle/np_soil.py (pure-Python restatement of the reference), numpy " + np.__version__,
        "known_answers": KNOWN_ANSWERS,
        "cases": cases(),
        "steppers": stepper_cases(),
        "column_params": column_param_case(),
    }
    path = os.path.join(HERE, "soil_golden.json")
    with open(path, "w") as f:
        json.dump(doc, f, indent=0, separators=(",", ":"))
    print(path, os.path.getsize(path), "bytes,", len(doc["cases"]), "cases")


if __name__ == "__main__":
    main()
