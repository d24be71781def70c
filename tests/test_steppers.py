"""Explicit low-storage steppers on the fused RHS+stage kernel (include/lh_soil.h "other explicit steppers").

The reference's test driver imports SSPRK33, SSPRK73 and CarpenterKennedy2N54 (test/runtests.jl:5-10) and
only ever uses SSPRK33.  CPU: the built-in coefficient tables satisfy the Runge-Kutta order conditions, the
product's tables equal the oracle's, the oracle's generic SSPRK33 equals its specialised SSPRK33, and the
observed convergence order of each method on the soil heat problem is the nominal one.  GPU: the CUDA path
reproduces the oracle state for every method (1e-10), through lh_soil_step and through Simulation."""
import ctypes as C
import itertools

import numpy as np
import pytest

import workloads as w

lh, abi = w.lh, w.abi
F, D, FD, N = abi.LH_BC_FLUX, abi.LH_BC_DIRICHLET, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
METHODS = {"Euler": (abi.LH_METHOD_EULER, 1), "SSPRK22": (abi.LH_METHOD_SSPRK22, 2), "SSPRK33": (abi.LH_METHOD_SSPRK33, 3),
           "SSPRK43": (abi.LH_METHOD_SSPRK43, 3), "CK2N54": (abi.LH_METHOD_CK2N54, 4)}


def table(lib, method):
    t = abi.lh_soil_stepper()
    assert lib.soil_stepper_named(method, t) == abi.LH_OK
    return t


def butcher(t):
    """Butcher tableau (A, b, c) of a Shu-Osher / 2N table, by propagating the stage recurrences symbolically:
    every register is a row vector of weights on (u^n, dt k_0, ..., dt k_{s-1})."""
    s = t.nstages
    e = lambda j: np.eye(s + 1)[j]
    Y = []                                   # stage values the RHS is evaluated at
    if t.kind == abi.LH_STEPPER_SHU_OSHER:
        u = e(0)
        for i in range(s):
            Y.append(u.copy())
            u = t.a[i] * e(0) + t.b[i] * u + t.g[i] * e(i + 1)
    else:
        u, r = e(0), np.zeros(s + 1)
        for i in range(s):
            Y.append(u.copy())
            r = t.a[i] * r + e(i + 1)
            u = u + t.b[i] * r
    assert abs(u[0] - 1) < 1e-14 and all(abs(y[0] - 1) < 1e-14 for y in Y)      # consistency
    A = np.array([y[1:] for y in Y])
    return A, u[1:], A.sum(axis=1)


def order_residuals(A, b, c):
    """Residuals of the rooted-tree order conditions up to order 4."""
    r = {1: [b.sum() - 1], 2: [b @ c - 1 / 2], 3: [b @ c**2 - 1 / 3, b @ A @ c - 1 / 6],
         4: [b @ c**3 - 1 / 4, (b * c) @ A @ c - 1 / 8, b @ A @ c**2 - 1 / 12, b @ A @ A @ c - 1 / 24]}
    return {k: max(abs(x) for x in v) for k, v in r.items()}


@pytest.mark.parametrize("name", sorted(METHODS))
def test_order_conditions_and_stage_times(oracle, name):
    method, order = METHODS[name]
    t = table(oracle, method)
    A, b, c = butcher(t)
    res = order_residuals(A, b, c)
    for k in range(1, order + 1):
        assert res[k] < 5e-14, (name, k, res[k])
    if order < 4:
        assert res[order + 1] > 1e-3                                  # and no better than nominal
    assert np.allclose(c, [t.c[i] for i in range(t.nstages)], atol=1e-14)   # c[] is what the host evaluates BCs at
    assert np.all(np.tril(A, -1) == A)                                # explicit
    host = {"Euler": lh.Euler, "SSPRK22": lh.SSPRK22, "SSPRK33": lh.SSPRK33, "SSPRK43": lh.SSPRK43,
            "CK2N54": lh.CarpenterKennedy2N54}[name]()
    assert np.allclose(host.c, c, atol=1e-14) and host.stages == t.nstages


def test_product_tables_equal_oracle_tables(oracle):
    """lh_soil_stepper_named needs no device: the CUDA library's tables against the oracle's, bit for bit."""
    cuda = lh.cuda_library()
    for method, _ in METHODS.values():
        a, b = table(cuda, method), table(oracle, method)
        assert bytes(a) == bytes(b)
    bad = abi.lh_soil_stepper()
    assert cuda.soil_stepper_named(99, bad) == abi.LH_ERR_INVALID_ARG
    assert oracle.soil_stepper_named(99, bad) == abi.LH_ERR_INVALID_ARG


def test_generic_ssprk33_equals_specialised(oracle):
    wl = w.coupled_workload(ncol=3, nlayer=24, seed=31)
    a, b = lh.SoilContext(oracle, wl.config()), lh.SoilContext(oracle, wl.config())
    for ctx in (a, b):
        wl.upload(ctx)
    a.step(0.0, wl.dt, 4)
    b.step_with(table(oracle, abi.LH_METHOD_SSPRK33), 0.0, wl.dt, 4)
    for f in (0, 2):
        ra, rb = a.get_state(f), b.get_state(f)
        assert np.max(np.abs(ra - rb)) <= 1e-14 * np.max(np.abs(ra))     # same scheme, different rounding of the combine


def _heat_problem(lib, method, nsteps, T=2.0e4):
    wl = w.heat_workload(ncol=2, nlayer=12, seed=41, top=(F, 0.0, N, 0.0), bottom=(F, 0.0, N, 0.0))
    ctx = lh.SoilContext(lib, wl.config())
    wl.upload(ctx)
    ctx.step_with(table(lib, method), 0.0, T / nsteps, nsteps)
    out = ctx.get_state(2)
    ctx.close()
    return out


@pytest.mark.parametrize("name", sorted(METHODS))
def test_observed_order(oracle, name):
    """Self-convergence on heat diffusion with insulated ends: error(dt) / error(dt/2) -> 2^order."""
    method, order = METHODS[name]
    ref = _heat_problem(oracle, abi.LH_METHOD_CK2N54, 512)
    e1 = np.max(np.abs(_heat_problem(oracle, method, 16) - ref))
    e2 = np.max(np.abs(_heat_problem(oracle, method, 32) - ref))
    observed = np.log2(e1 / e2)
    assert abs(observed - order) < 0.35, (name, observed, e1, e2)


def test_argument_validation(oracle):
    wl = w.richards_workload(ncol=1, nlayer=8, seed=2)
    ctx = lh.SoilContext(oracle, wl.config())
    t = table(oracle, abi.LH_METHOD_CK2N54)
    t.a[0] = 0.5
    with pytest.raises(lh._abi.SoilError):
        ctx.step_with(t, 0.0, 0.1, 1)
    t = table(oracle, abi.LH_METHOD_SSPRK22)
    t.nstages = 0
    with pytest.raises(lh._abi.SoilError):
        ctx.step_with(t, 0.0, 0.1, 1)
    t = table(oracle, abi.LH_METHOD_SSPRK22)
    with pytest.raises(ValueError):
        ctx.step_with(t, 0.0, 0.1, 2, np.zeros(5))


# ---- through the host mirror: Simulation(model, method) ------------------------------------------------------
@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, oracle):
    lib = oracle if request.param == "oracle" else lh.cuda_library()
    with lh.use_library(lib):
        yield lib


@pytest.mark.parametrize("method", [lh.Euler(), lh.SSPRK22(), lh.SSPRK43(), lh.CarpenterKennedy2N54(),
                                    lh.ShuOsherRK([0.0, 0.5], [1.0, 0.5], [1.0, 0.5], [0.0, 1.0])], ids=repr)
def test_simulation_with_other_methods(backend, oracle, method):
    """Richards column with a time-dependent Dirichlet top: Simulation must evaluate the closure at t + c_i dt."""
    sp = w.sand_soil_params()
    model = lh.SoilModel(
        np.float64, domain=lh.Column(np.float64, zlim=(-1.0, 0.0), nelements=16), energy_model=lh.PrescribedTemperatureModel(),
        hydrology_model=lh.SoilHydrologyModel(hydraulic_model=w.sand_vg()),
        boundary_conditions=lh.SoilColumnBC(
            top=lh.SoilComponentBC(hydrology=lh.Dirichlet(lambda t: 0.2 + 0.05 * np.sin(t / 3.0))),
            bottom=lh.SoilComponentBC(hydrology=lh.FreeDrainage())),
        soil_param_set=sp, earth_param_set=lh.EarthParameterSet())
    Y, Ya = lh.initialize_states(model, lambda z, m: {"ϑ_l": 0.12 + 0.02 * z, "θ_i": 0.0}, 0.0)
    dt, nsteps = 0.05, 12
    sim = lh.Simulation(model, method, Y_init=Y, dt=dt, tspan=(0.0, dt * nsteps), Ya_init=Ya, saveat=4 * dt)
    sol = lh.run_(sim)
    assert np.allclose(sol.t, [0.0, 4 * dt, 8 * dt, 12 * dt])
    got = np.asarray(lh.parent(sol.u[-1].soil.ϑ_l)).reshape(-1)
    # the same integration driven by hand on the oracle, stage times from the table
    tab = method.table(oracle)
    wl = w.richards_workload(ncol=1, nlayer=16, seed=1, zlim=(-1.0, 0.0), top=(N, 0.0, D, 0.2), bottom=(N, 0.0, FD, 0.0))
    ctx = lh.SoilContext(oracle, wl.config())
    y0 = np.asarray(lh.parent(Y.soil.ϑ_l)).reshape(-1).copy()
    ctx.set_state(0, y0[None, :].copy())
    ctx.set_state(1, np.zeros((1, 16)))
    bct = np.zeros((nsteps, tab.nstages, 4))
    for s, i in itertools.product(range(nsteps), range(tab.nstages)):
        bct[s, i, 1] = 0.2 + 0.05 * np.sin((s * dt + tab.c[i] * dt) / 3.0)
    ctx.step_with(tab, 0.0, dt, nsteps, bct)
    ref = ctx.get_state(0)[0]
    assert np.max(np.abs(got - ref)) <= 1e-10 * np.max(np.abs(ref))
    assert np.max(np.abs(got - y0)) > 1e-6          # and it did move


# ---- the CUDA path -----------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(METHODS))
@pytest.mark.parametrize("kind", ["coupled", "coupled_ice", "richards", "heat"])
def test_cuda_matches_oracle(cuda, oracle, name, kind):
    method, _ = METHODS[name]
    if kind == "coupled":
        wl = w.coupled_workload(ncol=96, nlayer=64, seed=51)
    elif kind == "coupled_ice":
        wl = w.coupled_workload(ncol=64, nlayer=20, seed=52, ice=True, viscosity=lh.TemperatureDependentViscosity(),
                                impedance=lh.IceImpedance())
    elif kind == "richards":
        wl = w.richards_workload(ncol=70, nlayer=100, seed=53)
    else:
        wl = w.heat_workload(ncol=40, nlayer=37, seed=54)
    g, o = lh.SoilContext(cuda, wl.config()), lh.SoilContext(oracle, wl.config())
    nsteps = 6
    rng = np.random.default_rng(5)
    tab_c, tab_o = table(cuda, method), table(oracle, method)
    base = np.array([wl.top[1], wl.top[3], wl.bottom[1], wl.bottom[3]])
    bct = base * (1.0 + 1e-3 * rng.standard_normal((nsteps, tab_o.nstages, 4)))      # per-stage boundary values
    for ctx, tab in ((g, tab_c), (o, tab_o)):
        wl.upload(ctx)
        ctx.step_with(tab, 0.0, wl.dt, nsteps, bct)
    for f in ((0, 2) if wl.model == abi.LH_MODEL_COUPLED else (0,) if wl.model == abi.LH_MODEL_RICHARDS else (2,)):
        a, r = g.get_state(f), o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (name, kind, f)
    ms, launches = g.last_step_timing()
    assert launches == nsteps * tab_o.nstages                                  # one fused launch per stage
