"""Random API sequences on the emulated build of the product library vs the oracle (CPU only; tests/hostemu.py).

The `-m gpu` parity tests drive one feature at a time.  The host logic behind the C ABI (csrc/lh_soil_api.cu, 1500 lines) keeps
state between calls — which kernel variant is selected (ice detected on upload, per-column / per-cell parameters set or
cleared), whether the fused per-block budget sums describe the current state, whether the next stage launch may chain to the
previous one, the position in the prescribed-profile tables — and a stale piece of that state only shows when calls are
INTERLEAVED.  Here every seed builds a small random problem (model, shape, boundary kinds, ice, factors, launch flags) and
applies the same random sequence of calls to the emulated product library and to the oracle, comparing after every call."""
import ctypes as C

import numpy as np
import pytest

import hostemu
import workloads as w

lh, abi = w.lh, w.abi
D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
RICHARDS, HEAT, COUPLED = abi.LH_MODEL_RICHARDS, abi.LH_MODEL_HEAT, abi.LH_MODEL_COUPLED
METHODS = [abi.LH_METHOD_EULER, abi.LH_METHOD_SSPRK22, abi.LH_METHOD_SSPRK33, abi.LH_METHOD_SSPRK43, abi.LH_METHOD_CK2N54]


def _table(lib, method):
    t = abi.lh_soil_stepper()
    assert lib.soil_stepper_named(method, C.byref(t)) == abi.LH_OK
    return t


def _problem(rng):
    model = int(rng.choice([RICHARDS, HEAT, COUPLED]))
    ncol, nlayer = int(rng.integers(1, 100)), int(rng.choice([1, 2, 3, 5, 8, 13, 21, 34, 47]))
    ice = bool(rng.random() < 0.4)
    factors = bool(rng.random() < 0.3)
    visc = lh.TemperatureDependentViscosity() if factors else None
    imp = lh.IceImpedance() if factors else None
    seed = int(rng.integers(1, 10 ** 6))
    # layers of ~0.1 m: with centimetre layers the energy tendency is dominated by ulp(T) / dz^2, which the parity norm does not scale with
    zlim = (-0.1 * nlayer - 0.05, 0.0)
    if model == COUPLED:
        top = [(D, 285.0 + 6 * rng.random(), D, 0.25 + 0.2 * rng.random()), (F, -3.0 * rng.random(), F, -1e-7 * rng.random()),
               (D, 290.0, F, 0.0), (F, 0.0, D, 0.35)][int(rng.integers(0, 4))]
        bottom = [(F, 0.0, FD, 0.0), (D, 283.0, D, 0.3), (F, 0.5, F, 0.0)][int(rng.integers(0, 3))]
        wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=seed, ice=ice, zlim=zlim, viscosity=visc, impedance=imp, top=top, bottom=bottom)
    elif model == RICHARDS:
        top = [(N, 0.0, D, 0.2 + 0.06 * rng.random()), (N, 0.0, F, -1e-6 * rng.random())][int(rng.integers(0, 2))]
        bottom = [(N, 0.0, FD, 0.0), (N, 0.0, D, 0.15), (N, 0.0, F, 0.0)][int(rng.integers(0, 3))]
        wl = w.richards_workload(ncol=ncol, nlayer=nlayer, seed=seed, ice=ice, zlim=zlim, viscosity=visc, impedance=imp, top=top, bottom=bottom)
    else:
        top = [(D, 290.0, N, 0.0), (F, -2.0, N, 0.0)][int(rng.integers(0, 2))]
        bottom = [(D, 280.0, N, 0.0), (F, 0.0, N, 0.0)][int(rng.integers(0, 2))]
        wl = w.heat_workload(ncol=ncol, nlayer=nlayer, seed=seed, ice=ice, zlim=(0.0, 0.08 * nlayer + 0.05), top=top, bottom=bottom)
    flags = int(rng.choice([0, abi.LH_FLAG_STAGE_LAUNCHES, abi.LH_FLAG_PERSISTENT, abi.LH_FLAG_STAGE_LAUNCHES | abi.LH_FLAG_NO_CHAIN]))
    if rng.random() < 0.3:
        flags |= abi.LH_FLAG_GENERAL_VG
    # small steps: the sequences below take up to ~25 steps and the comparison must stay a round-off comparison
    wl.dt = wl.dt * 0.2
    return wl, flags


def _prognostic(wl):
    return {RICHARDS: (0,), HEAT: (2,), COUPLED: (0, 2)}[wl.model]


class LeftStableRegime(Exception):
    """The ORACLE's state is no longer finite: the random problem left the explicit scheme's stability range (a forward-Euler
    step on a freshly oversaturated cell); nothing to compare any more."""


def _compare(tag, g, o, wl, rtol=2e-10):
    for f in _prognostic(wl):
        a, r = g.get_state(f), o.get_state(f)
        if not np.all(np.isfinite(r)):
            raise LeftStableRegime(tag)
        assert np.max(np.abs(a - r)) <= rtol * np.max(np.abs(r)), (tag, "state", f, np.max(np.abs(a - r)) / np.max(np.abs(r)))


def _sequence(rng, wl, g, o, nops):
    both = (g, o)
    t = 0.0
    has_flux_face = {k: kind == F for k, kind in zip(("top_energy", "top_hydrology", "bottom_energy", "bottom_hydrology"),
                                                     (wl.top[0], wl.top[2], wl.bottom[0], wl.bottom[2]))}
    het = False
    atmos_on = [False]
    for i in range(nops):
        ops = ["step", "step", "stages", "rhs", "budgets", "upload", "stepper", "run", "checkpoint", "diag", "bc"]
        if wl.model != HEAT:
            ops += ["colp", "cellp"]
        if wl.model != RICHARDS:
            ops += ["heatp"]
        if any(has_flux_face.values()):
            ops += ["fluxcols"]
        if wl.model != COUPLED:
            ops += ["auxtab"]
        elif wl.params.z_0m > 0.0:
            ops += ["atmos"]
        op = str(rng.choice(ops))
        tag = (i, op)
        if op == "step":
            k = int(rng.integers(0, 4))
            table = None
            if rng.random() < 0.5 and k > 0:
                vals = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
                table = np.broadcast_to(vals, (k, 3, 4)) * (1.0 + 1e-3 * rng.standard_normal((k, 3, 4)))
            for c in both:
                c.step(t, wl.dt, k, table)
            t += k * wl.dt
        elif op == "stages":
            for c in both:
                for s in (1, 2, 3):
                    c.stage(s, wl.dt)
            t += wl.dt
        elif op == "rhs":
            for c in both:
                c.rhs(t)
            dz = (wl.zmax - wl.zmin) / wl.nlayer
            for f in _prognostic(wl):
                a, r = g.get_tendency(f), o.get_tendency(f)
                scale = w.tendency_scale(o, f)
                # energy: the oracle rounds T = T_0 + (T - T_0) to ulp(T) = 5.7e-14 K before differencing it (the device keeps the
                # quotient); a conductive flux divergence sees that as kappa ulp(T) / dz^2, whatever the column's flux scale is
                floor = 4.0 * np.max(o.diagnostic(abi.LH_DIAG_KAPPA)) * 5.7e-14 / dz ** 2 if f == 2 else 0.0
                assert np.max((np.abs(a - r) - floor) / scale[:, None]) <= 5e-12, (tag, f, g.kernel_info())
        elif op == "budgets":
            if rng.random() < 0.5:
                a = g.budgets_wait(g.budgets_async())
                r = o.budgets()
            else:
                a, r = g.budgets(), o.budgets()
            assert np.allclose(a, r, rtol=1e-11, atol=0.0), (tag, a, r)
        elif op == "upload":
            f = int(rng.choice(list(wl.fields)))
            cur = o.get_state(f)
            if f == 1:
                # ice appears, changes or disappears: the ICE kernel variants are (de)selected on the upload
                new = np.zeros_like(cur) if rng.random() < 0.4 else rng.uniform(0.0, 0.03, cur.shape)
                th = o.get_state(0)
                room = wl.params.nu - new
                for c in both:
                    c.set_state(0, np.minimum(th, 0.98 * room))
            else:
                new = cur * (1.0 + 1e-3 * rng.standard_normal(cur.shape))
                if f == 0:
                    new = np.minimum(new, 0.98 * (wl.params.nu - o.get_state(1))) if not het else cur
                    if not het and wl.model == COUPLED and rng.random() < 0.4:
                        # a few exactly saturated, oversaturated and almost dry cells: the saturated branches of the pressure
                        # head and the conductivity, the residual-water clamp
                        room = wl.params.nu - o.get_state(1)
                        pick = rng.random(cur.shape)
                        new = np.where(pick < 0.03, room, new)
                        new = np.where((pick >= 0.03) & (pick < 0.06), room * (1.0 + 2e-4), new)
                        new = np.where((pick >= 0.06) & (pick < 0.09), wl.params.theta_r + 0.02 * room, new)
            for c in both:
                c.set_state(f, new)
        elif op == "stepper":
            m = int(rng.choice(METHODS))
            k = int(rng.integers(1, 3))
            for c in both:
                c.step_with(_table(c.lib, m), t, wl.dt, k)
            t += k * wl.dt
        elif op == "run":
            k = int(rng.integers(1, 6))
            be = int(rng.choice([0, 1, 2]))
            se = int(rng.choice([0, 1, 3]))
            first = bool(rng.random() < 0.5)
            fields = list(_prognostic(wl))
            table = None
            if rng.random() < 0.5:                    # a boundary-value table: read from device memory by the persistent launches
                vals = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
                table = np.broadcast_to(vals, (k, 3, 4)) * (1.0 + 1e-3 * rng.standard_normal((k, 3, 4)))
            kw = dict(bc_table=table, budget_every=be, save_every=se, save_first=first, save_fields=fields if (se or first) else ())
            ns = (k // se if se else 0) + (1 if first else 0)
            pin = None
            if ns and rng.random() < 0.5:
                # snapshots into pinned memory from the library: D2H copies that really are asynchronous (a pageable numpy
                # destination turns each of them into a synchronisation point)
                pin = C.c_void_p()
                shape = (ns, len(fields), wl.ncol, wl.nlayer)
                assert g.lib.soil_alloc_host(int(np.prod(shape)) * 8, C.byref(pin)) == abi.LH_OK
                dest = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_double)), shape=shape)
                bg, sg = g.run(t, wl.dt, k, save_out=dest, **kw)
                sg = sg.copy()
                assert g.lib.soil_free_host(pin) == abi.LH_OK
                out = [(bg, sg), o.run(t, wl.dt, k, **kw)]
            else:
                out = [c.run(t, wl.dt, k, **kw) for c in both]
            t += k * wl.dt
            (bg, sg), (bo, so) = out
            if be:
                assert np.allclose(bg, bo, rtol=1e-11, atol=0.0), tag
            if so is not None and so.size:
                assert sg.shape == so.shape
                for j in range(so.shape[1]):
                    assert np.max(np.abs(sg[:, j] - so[:, j])) <= 2e-10 * np.max(np.abs(so[:, j])), (tag, "snapshot field", j)
        elif op == "checkpoint":
            saved = [c.checkpoint() for c in both]
            before = [g.get_state(f) for f in _prognostic(wl)]
            for c in both:
                c.step(t, wl.dt, 2)
            for c, buf in zip(both, saved):
                c.restore(buf)
            for f, b in zip(_prognostic(wl), before):
                assert np.array_equal(g.get_state(f), b), (tag, "restore is not bit-exact", f)
        elif op == "diag":
            which = int(rng.choice([abi.LH_DIAG_K, abi.LH_DIAG_PSI, abi.LH_DIAG_KAPPA, abi.LH_DIAG_T]))
            a, r = g.diagnostic(which), o.diagnostic(which)
            ok = np.isfinite(r)
            assert np.array_equal(np.isfinite(a), ok), (tag, which)
            # K: the oracle's literal 1 - (1 - S^(1/m))^m cancels in nearly dry cells (S = 0.02: 5000 ulp); the device form does not
            tol = 2e-11 if which == abi.LH_DIAG_K else 5e-13
            assert np.max(np.abs(a[ok] - r[ok]) / np.maximum(np.abs(r[ok]), 1e-300), initial=0.0) <= tol, (tag, which, g.kernel_info())
        elif op == "bc":
            vals = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]]) * (1.0 + 1e-3 * rng.standard_normal(4))
            for c in both:
                c.set_bc_values(vals)
        elif op == "colp":
            if het and rng.random() < 0.5:
                for c in both:
                    c.set_column_params()
                    c.set_cell_params()
                het = False
            else:
                n = wl.ncol
                th, ti = o.get_state(0), o.get_state(1)
                # only parameters that keep every cell below saturation and above the residual water content
                cp = {"vg_n": rng.uniform(1.3, 4.0, n), "vg_alpha": wl.params.vg_alpha * rng.uniform(0.5, 2.0, n),
                      "Ksat": wl.params.Ksat * 10.0 ** rng.uniform(-1.0, 1.0, n)}
                if rng.random() < 0.5:
                    cp["nu"] = np.maximum(wl.params.nu * rng.uniform(1.0, 1.15, n), (th + ti).max(axis=1) * 1.02)
                for c in both:
                    c.set_column_params(**cp)
                het = True
        elif op == "cellp":
            shape = (wl.ncol, wl.nlayer)
            cp = {"vg_n": rng.uniform(1.4, 3.5, shape), "Ksat": wl.params.Ksat * 10.0 ** rng.uniform(-1.0, 1.0, shape)}
            for c in both:
                c.set_cell_params(**cp)
            het = True
        elif op == "heatp":
            n = wl.ncol
            if rng.random() < 0.3:
                for c in both:
                    c.set_column_heat_params()
            else:
                hp = {"rho_c_ds": wl.params.rho_c_ds * rng.uniform(0.8, 1.2, n), "kappa_sat_unfrozen": wl.params.kappa_sat_unfrozen * rng.uniform(0.8, 1.2, n)}
                if rng.random() < 0.5:
                    hp["nu_ss_om"] = rng.uniform(0.0, 0.2, n)
                for c in both:
                    c.set_column_heat_params(**hp)
                if wl.model == COUPLED:
                    het = True
        elif op == "fluxcols":
            kw = {}
            for name, on in has_flux_face.items():
                if on and rng.random() < 0.6:
                    scale = 2.0 if name.endswith("energy") else 1e-7
                    kw[name] = -scale * rng.random(wl.ncol)
            for c in both:
                c.set_column_fluxes(**kw)
        elif op == "auxtab":
            # make_update_aux (right_hand_side.jl:54-81): a time-dependent prescribed profile as a device table, one row per stage
            k = int(rng.integers(1, 4))
            if wl.model == RICHARDS:
                field = abi.LH_FIELD_T
                rows = 288.0 + 6.0 * rng.random((3 * k, 1)) + np.linspace(0.0, 2.0, wl.nlayer)[None, :]
            else:
                field = abi.LH_FIELD_THETA_L
                rows = wl.params.nu * (0.4 + 0.3 * rng.random((3 * k, 1))) * np.ones((1, wl.nlayer))
            for c in both:
                c.set_aux_table(field, rows)
                c.step(t, wl.dt, k)
                c.set_aux_table(field, None)
            t += k * wl.dt
        elif op == "atmos":
            if atmos_on[0]:
                for c in both:
                    c.set_atmos_forcing(None)
                atmos_on[0] = False
            else:
                a = abi.lh_soil_atmos()
                ep = lh.EarthParameterSet()
                a.u_atm, a.theta_atm, a.z_atm, a.theta_scale, a.rho_a_sfc = 0.3 + 3 * rng.random(), 284.0 + 10 * rng.random(), 2.0, 290.0, 1.17
                a.q_atm = 0.004 + 0.006 * rng.random()
                a.R_v, a.R_d, a.grav, a.cp_d, a.cp_v = ep.R_v, ep.R_d, ep.grav, ep.cp_d, ep.cp_v
                a.LH_v0, a.press_triple, a.T_triple, a.von_karman = ep.LH_v0, ep.press_triple, ep.T_triple, ep.von_karman_const
                a.Pr_0, a.a_m, a.a_h = ep.Pr_0, ep.a_m, ep.a_h
                for c in both:
                    c.set_atmos_forcing(a)
                atmos_on[0] = True
        _compare(tag, g, o, wl, rtol=2e-9 if atmos_on[0] else 2e-10)
    return t


# async: 0 = every operation completes inside the call that enqueues it; 1 = streams are queues executed only when a
# synchronisation needs them and only as far as it needs them; 2 = seeded random interleaving of all streams' runnable operations
# (tests/support/hostemu/hostemu.cpp).  All three are legal executions of the same CUDA program.
@pytest.mark.parametrize("seed,mode", [(s, 0) for s in range(24)] + [(s, 1) for s in range(24, 36)] + [(s, 2) for s in range(36, 48)])
def test_random_api_sequences_match_oracle(oracle, seed, mode):
    emu = hostemu.library(lh)
    knobs = hostemu.controls()
    knobs.lh_emu_set_async(mode, seed + 1)
    try:
        _one_sequence(emu, oracle, seed)
    finally:
        knobs.lh_emu_set_async(0, 1)


def _one_sequence(emu, oracle, seed):
    rng = np.random.default_rng(9000 + seed)
    wl, flags = _problem(rng)
    g, o = lh.SoilContext(emu, wl.config(flags=flags)), lh.SoilContext(oracle, wl.config(flags=flags & abi.LH_FLAG_GENERAL_VG))
    for c in (g, o):
        wl.upload(c)
    try:
        _sequence(rng, wl, g, o, nops=14)
    except LeftStableRegime:
        pass
    finally:
        g.close()
        o.close()
