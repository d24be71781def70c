"""Committed golden vectors (tests/golden/soil_golden.json, made by tests/golden/make_golden.py).

CPU (`-m "not gpu"`): the reference's own known answers hold for BOTH restatements, and the C oracle
reproduces every dense case of the file.  GPU (`-m gpu`): the CUDA path, through the C ABI, reproduces
the same cases within the north-star tolerances (1e-12 scaled per tendency evaluation, 1e-10 on the state
after the file's number of SSPRK33 steps)."""
import json
import os
import sys
import types

import numpy as np
import pytest

import workloads as w

sys.path.insert(0, os.path.join(w.ROOT, "oracle"))
import np_soil  # noqa: E402

lh, abi = w.lh, w.abi
EPS = np.finfo(np.float64).eps
GOLDEN = json.load(open(os.path.join(w.ROOT, "tests", "golden", "soil_golden.json")))
CASES = {c["name"]: c for c in GOLDEN["cases"]}


def _params(rec):
    p = abi.lh_soil_params()
    for k, v in rec["params"].items():
        setattr(p, k, v)
    return p


def _workload(rec):
    ncol, n = rec["ncol"], rec["nlayer"]
    fields = {0: np.array([c["theta_l"] for c in rec["columns"]]), 1: np.array([c["theta_i"] for c in rec["columns"]])}
    if rec["model"] != abi.LH_MODEL_RICHARDS:
        fields[2] = np.array([c["rho_e_int"] for c in rec["columns"]])
    return w.Workload(model=rec["model"], ncol=ncol, nlayer=n, zmin=rec["zmin"], zmax=rec["zmax"], params=_params(rec),
                      top=tuple(rec["top"]), bottom=tuple(rec["bottom"]), dt=rec["dt"], fields=fields,
                      aux_T=None if rec["aux_T"] is None else np.array(rec["aux_T"]), name=rec["name"])


def _check_case(lib, rec, tend_tol, state_tol):
    wl = _workload(rec)
    ctx = lh.SoilContext(lib, wl.config())
    wl.upload(ctx)
    ctx.rhs(0.0)
    dz = (rec["zmax"] - rec["zmin"]) / rec["nlayer"]
    for fid, key, fkey in ((0, "d_theta_l", "Fw"), (2, "d_rho_e_int", "Fe")):
        if fid == 2 and rec["model"] == abi.LH_MODEL_RICHARDS or fid == 0 and rec["model"] == abi.LH_MODEL_HEAT:
            continue
        got = ctx.get_tendency(fid)
        for c, col in enumerate(rec["columns"]):
            ref = np.array(col[key])
            scale = max(np.max(np.abs(ref)), np.max(np.abs(col[fkey])) / dz, 1e-300)
            err = np.max(np.abs(got[c] - ref)) / scale
            assert err <= tend_tol, f"{rec['name']} field {fid} column {c}: scaled tendency error {err:.3e}"
    assert np.all(ctx.get_tendency(1) == 0.0)                       # d theta_i = 0, right_hand_side.jl:182,359
    ctx.step(0.0, rec["dt"], rec["nsteps"])
    for fid, key in ((0, "theta_l_after"), (2, "rho_e_int_after")):
        if fid == 2 and rec["model"] == abi.LH_MODEL_RICHARDS or fid == 0 and rec["model"] == abi.LH_MODEL_HEAT:
            continue
        got = ctx.get_state(fid)
        ref = np.array([col[key] for col in rec["columns"]])
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        assert err <= state_tol, f"{rec['name']} field {fid}: state error {err:.3e} after {rec['nsteps']} steps"
    ctx.close()


# ---- the reference's own known answers, against both restatements ---------------------------------------------
def _eval_known(entry, oracle):
    import ctypes as C

    fn, args = entry["fn"], entry["args"]
    dbl = C.c_double

    def ofn(name, *argtypes):
        f = oracle.raw(name)
        f.restype = dbl
        f.argtypes = list(argtypes)
        return f

    p = abi.lh_soil_params()
    pp = C.POINTER(abi.lh_soil_params)
    if fn == "effective_saturation":
        f = ofn("lho_effective_saturation", dbl, dbl, dbl)
        return [[np_soil.effective_saturation(*a) for a in args], [f(*a) for a in args]]
    if fn == "pressure_head_saturated":
        out = [[], []]
        f = ofn("lho_pressure_head", pp, dbl, dbl, dbl)
        for th, nu_eff, S_s in args:
            q = types.SimpleNamespace(vg_n=2.0, vg_alpha=2.6, vg_m=0.5, theta_r=0.0)
            p.vg_n, p.vg_alpha, p.vg_m, p.theta_r = 2.0, 2.6, 0.5, 0.0
            out[0].append(np_soil.pressure_head(q, th, nu_eff, S_s))
            out[1].append(f(C.byref(p), th, nu_eff, S_s))
        return out
    if fn == "impedance_factor":
        out = [[], []]
        f = ofn("lho_impedance_factor", pp, dbl)
        for Omega, f_i in args:
            q = types.SimpleNamespace(impedance_factor=1, imp_Omega=Omega)
            p.impedance_factor, p.imp_Omega = 1, Omega
            out[0].append(np_soil.impedance_factor(q, f_i))
            out[1].append(f(C.byref(p), f_i))
        return out
    if fn == "hydraulic_conductivity_over_Ksat":
        out = [[], []]
        f = ofn("lho_hydraulic_conductivity", pp, dbl, dbl, dbl)
        for (S,) in args:
            q = types.SimpleNamespace(vg_m=0.5, Ksat=1.0)
            p.vg_m, p.Ksat = 0.5, 1.0
            out[0].append(np_soil.hydraulic_conductivity(q, S, 1.0, 1.0))
            out[1].append(f(C.byref(p), S, 1.0, 1.0))
        return out
    if fn == "k_dry":
        out = [[], []]
        f = ofn("lho_k_dry", pp)
        for nu, ks, rho_p, kdp, K_therm in args:
            q = types.SimpleNamespace(nu=nu, kappa_solid=ks, rho_p=rho_p, kappa_dry_parameter=kdp, K_therm=K_therm)
            p.nu, p.kappa_solid, p.rho_p, p.kappa_dry_parameter, p.K_therm = nu, ks, rho_p, kdp, K_therm
            out[0].append(np_soil.k_dry(q))
            out[1].append(f(C.byref(p)))
        return out
    raise KeyError(fn)


@pytest.mark.parametrize("entry", [e for e in GOLDEN["known_answers"] if e["fn"] != "coupled_rhs_known_answer"],
                         ids=lambda e: e["fn"])
def test_reference_known_answers(entry, oracle):
    for got in _eval_known(entry, oracle):
        for g, e in zip(got, entry["expect"]):
            assert abs(g - e) <= entry["rtol"] * abs(e), (entry["what"], g, e)


def _known_rhs_workload():
    """test/SoilModel/coupled.jl:123-235: default ICs of the coupled model (theta_l = 0.25 = S 0.5, T = T_0), zero fluxes."""
    F = abi.LH_BC_FLUX
    wl = w.coupled_workload(ncol=1, nlayer=20, seed=1, top=(F, 0.0, F, 0.0), bottom=(F, 0.0, F, 0.0))
    th = np.full((1, 20), 0.25)
    ti = np.zeros((1, 20))
    T = np.full((1, 20), wl.params.T_0)
    wl.fields = {0: th, 1: ti, 2: w.rho_e_int_from_T(wl.params, th, ti, T)}
    return wl


def _check_known_rhs(lib, tol):
    entry = next(e for e in GOLDEN["known_answers"] if e["fn"] == "coupled_rhs_known_answer")
    K, dz = entry["expect"]["K"], entry["expect"]["dz"]
    wl = _known_rhs_workload()
    ctx = lh.SoilContext(lib, wl.config())
    wl.upload(ctx)
    ctx.rhs(0.0)
    d = ctx.get_tendency(0)[0]
    expect = np.zeros(20)
    expect[0], expect[-1] = K / dz, -K / dz          # coupled.jl:223-234
    assert np.max(np.abs(d - expect)) <= tol * K / dz
    assert np.max(np.abs(ctx.get_tendency(2)[0])) <= tol * 4.2e6 * 273.16 * K / dz    # coupled.jl:222 (≈ 0)
    assert np.all(ctx.get_tendency(1) == 0.0)                                          # coupled.jl:221
    ctx.close()


def test_known_rhs_oracle(oracle):
    _check_known_rhs(oracle, 1e-12)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    # two independent restatements on the same libm: a few ulp of the column scale
    _check_case(oracle, CASES[name], 8 * EPS, 64 * EPS)


# ---- the CUDA path -----------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_known_rhs_cuda(cuda):
    _check_known_rhs(cuda, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_golden(cuda, name):
    _check_case(cuda, CASES[name], 1e-12, 1e-10)


# ---- steppers and per-column parameters (fixtures generated with literature coefficient tables / per-column np_soil) ----
_METHOD_IDS = {"Euler": abi.LH_METHOD_EULER, "SSPRK22": abi.LH_METHOD_SSPRK22, "SSPRK33": abi.LH_METHOD_SSPRK33,
               "SSPRK43": abi.LH_METHOD_SSPRK43, "CarpenterKennedy2N54": abi.LH_METHOD_CK2N54}


def _check_steppers(lib, tol):
    rec = GOLDEN["steppers"]
    fields = {0: np.array([rec["theta_l"]]), 1: np.array([rec["theta_i"]]), 2: np.array([rec["rho_e_int"]])}
    wl = w.Workload(model=abi.LH_MODEL_COUPLED, ncol=1, nlayer=rec["nlayer"], zmin=rec["zmin"], zmax=rec["zmax"], params=_params(rec),
                    top=tuple(rec["top"]), bottom=tuple(rec["bottom"]), dt=rec["dt"], fields=fields)
    for name, ref in rec["methods"].items():
        tab = abi.lh_soil_stepper()
        assert lib.soil_stepper_named(_METHOD_IDS[name], tab) == abi.LH_OK
        ctx = lh.SoilContext(lib, wl.config())
        wl.upload(ctx)
        ctx.step_with(tab, 0.0, rec["dt"], rec["nsteps"])
        for fid, key in ((0, "theta_l_after"), (2, "rho_e_int_after")):
            r = np.array(ref[key])
            err = np.max(np.abs(ctx.get_state(fid)[0] - r)) / np.max(np.abs(r))
            assert err <= tol, (name, fid, err)
        ctx.close()


def _check_column_params(lib, tend_tol, state_tol):
    rec = GOLDEN["column_params"]
    cols = rec["columns"]
    n = rec["nlayer"]
    fields = {0: np.array([c["theta_l"] for c in cols]), 1: np.zeros((len(cols), n))}
    wl = w.Workload(model=abi.LH_MODEL_RICHARDS, ncol=len(cols), nlayer=n, zmin=rec["zmin"], zmax=rec["zmax"], params=_params(rec),
                    top=tuple(rec["top"]), bottom=tuple(rec["bottom"]), dt=rec["dt"], fields=fields)
    ctx = lh.SoilContext(lib, wl.config())
    ctx.set_column_params(**{k: np.array(v) for k, v in rec["column_params"].items()})
    wl.upload(ctx)
    ctx.rhs(0.0)
    got = ctx.get_tendency(0)
    dz = (rec["zmax"] - rec["zmin"]) / n
    for c, col in enumerate(cols):
        r = np.array(col["d_theta_l"])
        scale = max(np.max(np.abs(r)), np.max(np.abs(col["Fw"])) / dz)
        assert np.max(np.abs(got[c] - r)) <= tend_tol * scale, c
    ctx.step(0.0, rec["dt"], rec["nsteps"])
    ref = np.array([c["theta_l_after"] for c in cols])
    assert np.max(np.abs(ctx.get_state(0) - ref)) <= state_tol * np.max(np.abs(ref))
    ctx.close()


def test_oracle_reproduces_golden_steppers_and_column_params(oracle):
    _check_steppers(oracle, 64 * EPS)
    _check_column_params(oracle, 8 * EPS, 64 * EPS)


@pytest.mark.gpu
def test_cuda_reproduces_golden_steppers_and_column_params(cuda):
    _check_steppers(cuda, 1e-10)
    _check_column_params(cuda, 1e-12, 1e-10)
