"""Two restatements written separately must agree: the C oracle (oracle/lho_soil.c) against the
pure-Python one (oracle/np_soil.py) on single columns, every model kind and BC kind.  Both use glibc's
libm and evaluate in source order without FMA, so agreement is expected to the last bits; the assertion
allows 4 ulp of the column's scale.  CPU only."""
import os
import sys

import numpy as np
import pytest

import workloads as w

sys.path.insert(0, os.path.join(w.ROOT, "oracle"))
import np_soil  # noqa: E402

lh = w.lh
abi = w.abi
D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
EPS = np.finfo(np.float64).eps


def _check(oracle, wl, ncheck=4):
    ctx = lh.SoilContext(oracle, wl.config())
    wl.upload(ctx)
    ctx.rhs(0.0)
    d = [ctx.get_tendency(f) for f in range(3)]
    bcv = [wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]
    aux_T = wl.aux_T if wl.aux_T is not None else np.full(wl.nlayer, 288.0)
    for c in range(min(ncheck, wl.ncol)):
        th = wl.fields[0][c]
        ti = wl.fields[1][c]
        re = wl.fields.get(2, np.zeros_like(wl.fields[0]))[c]
        dth, dti, dre, Fw, Fe = np_soil.column_rhs(
            wl.params, wl.model, wl.zmin, wl.zmax, (wl.top[0], wl.top[2]), (wl.bottom[0], wl.bottom[2]), bcv,
            list(th), list(ti), list(re), list(aux_T))
        dz = (wl.zmax - wl.zmin) / wl.nlayer
        for a, r, Fl in ((d[0][c], dth, Fw), (d[2][c], dre, Fe)):
            scale = max(np.max(np.abs(r)), np.max(np.abs(Fl)) / dz, 1e-300)
            assert np.max(np.abs(a - r)) <= 4 * EPS * scale
        assert np.all(d[1][c] == 0.0)


@pytest.mark.parametrize("top,bottom", [((D, 288.0, D, 0.4), (F, 0.0, FD, 0.0)), ((F, 2.0, F, -1e-8), (D, 281.0, D, 0.3)),
                                        ((D, 280.0, FD, 0.0), (F, 0.0, F, 0.0))])
@pytest.mark.parametrize("ice", [False, True])
def test_coupled(oracle, top, bottom, ice):
    _check(oracle, w.coupled_workload(ncol=4, nlayer=20, seed=3, ice=ice, top=top, bottom=bottom,
                                      viscosity=lh.TemperatureDependentViscosity() if ice else None,
                                      impedance=lh.IceImpedance() if ice else None))


@pytest.mark.parametrize("top,bottom", [((N, 0.0, D, 0.267), (N, 0.0, FD, 0.0)), ((N, 0.0, F, 1e-7), (N, 0.0, D, 0.12))])
def test_richards(oracle, top, bottom):
    _check(oracle, w.richards_workload(ncol=4, nlayer=30, seed=4, top=top, bottom=bottom))
    _check(oracle, w.richards_workload(ncol=2, nlayer=12, seed=5, top=top, bottom=bottom, ice=True,
                                       viscosity=lh.TemperatureDependentViscosity(), impedance=lh.IceImpedance()))


@pytest.mark.parametrize("top,bottom", [((D, 290.0, N, 0.0), (D, 280.0, N, 0.0)), ((F, 4.0, N, 0.0), (F, -4.0, N, 0.0))])
def test_heat(oracle, top, bottom):
    _check(oracle, w.heat_workload(ncol=4, nlayer=24, seed=6, top=top, bottom=bottom))
    _check(oracle, w.heat_workload(ncol=2, nlayer=10, seed=7, top=top, bottom=bottom, ice=True))


def test_ssprk33_step_matches(oracle):
    """The stage combine of the C oracle against the Python Shu-Osher restatement."""
    wl = w.coupled_workload(ncol=1, nlayer=16, seed=9, top=(F, 0.0, F, 0.0), bottom=(F, 0.0, F, 0.0))
    ctx = lh.SoilContext(oracle, wl.config())
    wl.upload(ctx)
    ctx.step(0.0, wl.dt, 1)
    p = wl.params

    def rhs(u, stage):
        dth, dti, dre, _, _ = np_soil.column_rhs(p, wl.model, wl.zmin, wl.zmax, (F, F), (F, F), [0.0] * 4,
                                                 list(u[0]), list(u[1]), list(u[2]), [288.0] * wl.nlayer)
        return dth, dti, dre

    u0 = (wl.fields[0][0].copy(), wl.fields[1][0].copy(), wl.fields[2][0].copy())
    u = np_soil.ssprk33_step(rhs, u0, wl.dt)
    for f in range(3):
        got = ctx.get_state(f)[0]
        assert np.max(np.abs(got - u[f])) <= 8 * EPS * max(np.max(np.abs(u[f])), 1e-300)


@pytest.mark.parametrize("method", [abi.LH_METHOD_EULER, abi.LH_METHOD_SSPRK22, abi.LH_METHOD_SSPRK33, abi.LH_METHOD_SSPRK43,
                                    abi.LH_METHOD_CK2N54])
def test_generic_steppers_match_python(oracle, method):
    """lho_soil_step (two-register Shu-Osher and Williamson 2N recurrences) against the same recurrences written out in
    Python over np_soil's right-hand side, per-stage boundary values included."""
    wl = w.coupled_workload(ncol=1, nlayer=14, seed=19, top=(D, 288.0, D, 0.4), bottom=(F, 0.0, FD, 0.0))
    tab = abi.lh_soil_stepper()
    assert oracle.soil_stepper_named(method, tab) == abi.LH_OK
    ns, nsteps = tab.nstages, 3
    rng = np.random.default_rng(8)
    base = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
    bct = base * (1.0 + 1e-3 * rng.standard_normal((nsteps, ns, 4)))
    ctx = lh.SoilContext(oracle, wl.config())
    wl.upload(ctx)
    ctx.step_with(tab, 0.0, wl.dt, nsteps, bct)
    p = wl.params

    def rhs(u, bcv):
        dth, dti, dre, _, _ = np_soil.column_rhs(p, wl.model, wl.zmin, wl.zmax, (wl.top[0], wl.top[2]), (wl.bottom[0], wl.bottom[2]),
                                                 list(bcv), list(u[0]), list(u[1]), list(u[2]), [288.0] * wl.nlayer)
        return [dth, dti, dre]

    u = [wl.fields[k][0].copy() for k in range(3)]
    dt = wl.dt
    for s in range(nsteps):
        if tab.kind == abi.LH_STEPPER_SHU_OSHER:
            u0, v = [x.copy() for x in u], [x.copy() for x in u]
            for i in range(ns):
                k = rhs(v, bct[s, i])
                v = [tab.a[i] * a0 + tab.b[i] * vi + (tab.g[i] * dt) * ki for a0, vi, ki in zip(u0, v, k)]
            u = v
        else:
            r = [np.zeros_like(x) for x in u]
            for i in range(ns):
                k = rhs(u, bct[s, i])
                r = [dt * ki if i == 0 else tab.a[i] * ri + dt * ki for ri, ki in zip(r, k)]
                u = [ui + tab.b[i] * ri for ui, ri in zip(u, r)]
    for f in range(3):
        got = ctx.get_state(f)[0]
        assert np.max(np.abs(got - u[f])) <= 16 * EPS * max(np.max(np.abs(u[f])), 1e-300), (method, f)
