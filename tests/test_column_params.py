"""Heterogeneous soils: per-column hydraulic parameters (lh_soil_set_column_params — ν, θr, van Genuchten n and α,
K_sat per column).  CPU: the oracle with constant arrays equals the homogeneous oracle bit for bit and
really uses the arrays.  GPU: the per-lane-parameter kernels against the oracle for every model, all launch paths,
and the diagnostics."""
import numpy as np
import pytest

import workloads as w

lh, abi = w.lh, w.abi


def random_column_params(wl, seed, which=("nu", "theta_r", "vg_n", "vg_alpha", "Ksat")):
    rng = np.random.default_rng(seed)
    p, n = wl.params, wl.ncol
    out = {}
    if "nu" in which:
        out["nu"] = p.nu * rng.uniform(0.85, 1.15, n)
    if "theta_r" in which:
        out["theta_r"] = rng.uniform(0.0, 0.06, n)
    if "vg_n" in which:
        out["vg_n"] = rng.uniform(1.3, 4.0, n)
    if "vg_alpha" in which:
        out["vg_alpha"] = p.vg_alpha * rng.uniform(0.5, 2.0, n)
    if "Ksat" in which:
        out["Ksat"] = p.Ksat * 10.0 ** rng.uniform(-1.5, 1.5, n)
    return out


def rescale_state(wl, cp):
    """Keep the workload's saturation profile under the new per-column porosity / residual water content."""
    p = wl.params
    S = (wl.fields[0] - p.theta_r) / ((p.nu - wl.fields[1]) - p.theta_r)
    nu = cp.get("nu", np.full(wl.ncol, p.nu))[:, None]
    thr = cp.get("theta_r", np.full(wl.ncol, p.theta_r))[:, None]
    wl.fields[0] = thr + S * ((nu - wl.fields[1]) - thr)


def test_oracle_constant_arrays_equal_homogeneous(oracle):
    wl = w.coupled_workload(ncol=6, nlayer=20, seed=81)
    a, b = lh.SoilContext(oracle, wl.config()), lh.SoilContext(oracle, wl.config())
    p = wl.params
    b.set_column_params(nu=np.full(6, p.nu), theta_r=np.full(6, p.theta_r), vg_n=np.full(6, p.vg_n),
                        vg_alpha=np.full(6, p.vg_alpha), Ksat=np.full(6, p.Ksat))
    for ctx in (a, b):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 3)
    for f in (0, 2):
        assert np.array_equal(a.get_state(f), b.get_state(f))
    c = lh.SoilContext(oracle, wl.config())
    c.set_column_params(Ksat=p.Ksat * np.array([1.0, 2.0, 1.0, 1.0, 0.5, 1.0]))
    wl.upload(c)
    c.step(0.0, wl.dt, 3)
    same = [np.array_equal(a.get_state(0)[k], c.get_state(0)[k]) for k in range(6)]
    assert same == [True, False, True, True, False, True]         # columns are independent; only the changed ones move
    with pytest.raises(ValueError):
        c.set_column_params(nu=np.zeros(5))
    c.set_column_params()                                         # all None: homogeneous again
    wl.upload(c)
    c.step(0.0, wl.dt, 3)
    assert np.array_equal(a.get_state(0), c.get_state(0))


CASES = {
    "coupled": lambda: w.coupled_workload(ncol=200, nlayer=64, seed=82),
    "coupled_ice": lambda: w.coupled_workload(ncol=96, nlayer=24, seed=83, ice=True, viscosity=lh.TemperatureDependentViscosity(),
                                              impedance=lh.IceImpedance()),
    "richards": lambda: w.richards_workload(ncol=130, nlayer=100, seed=84),
    "richards_tall": lambda: w.richards_workload(ncol=40, nlayer=300, seed=85, zlim=(-4.5, 0.0)),
    "heat": lambda: w.heat_workload(ncol=64, nlayer=37, seed=86),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("launch", ["stage", "persistent"])
def test_cuda_matches_oracle(cuda, oracle, name, launch):
    wl = CASES[name]()
    cp = random_column_params(wl, seed=7)
    rescale_state(wl, cp)
    if wl.model == abi.LH_MODEL_COUPLED:      # the Dirichlet top value must stay below every column's porosity
        wl.top = (wl.top[0], wl.top[1], wl.top[2], 0.3)
    flags = abi.LH_FLAG_STAGE_LAUNCHES if launch == "stage" else abi.LH_FLAG_PERSISTENT
    g, o = lh.SoilContext(cuda, wl.config(flags=flags)), lh.SoilContext(oracle, wl.config())
    for ctx in (g, o):
        ctx.set_column_params(**cp)
        wl.upload(ctx)
        ctx.rhs(0.0)
    fields = (0, 2) if wl.model == abi.LH_MODEL_COUPLED else (0,) if wl.model == abi.LH_MODEL_RICHARDS else (2,)
    for f in fields:
        a, r = g.get_tendency(f), o.get_tendency(f)
        scale = w.tendency_scale(o, f)
        err = np.max(np.abs(a - r) / scale[:, None])
        assert err <= 1e-12, (name, f, err)
    for which in (abi.LH_DIAG_K, abi.LH_DIAG_PSI, abi.LH_DIAG_KAPPA):
        a, r = g.diagnostic(which), o.diagnostic(which)
        assert np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300)) <= 2e-13, (name, which)
    for ctx in (g, o):
        ctx.step(0.0, wl.dt, 6)
    for f in fields:
        a, r = g.get_state(f), o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (name, f)


@pytest.mark.gpu
def test_cuda_constant_arrays_match_homogeneous(cuda):
    """Constant arrays take the per-lane kernels (general closures); the homogeneous context takes the n = 2 fast path:
    two different evaluation orders of the same closures, agreeing to round-off."""
    wl = w.coupled_workload(ncol=128, nlayer=64, seed=87)
    p = wl.params
    a, b = lh.SoilContext(cuda, wl.config()), lh.SoilContext(cuda, wl.config())
    b.set_column_params(nu=np.full(128, p.nu), vg_n=np.full(128, p.vg_n), Ksat=np.full(128, p.Ksat))
    for ctx in (a, b):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 5)
    for f in (0, 2):
        ra, rb = a.get_state(f), b.get_state(f)
        assert np.max(np.abs(ra - rb)) <= 1e-11 * np.max(np.abs(ra))
    b.set_column_params()                       # back to the homogeneous kernels
    wl.upload(b)
    b.step(0.0, wl.dt, 5)
    assert np.array_equal(a.get_state(0), b.get_state(0))


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, oracle):
    lib = oracle if request.param == "oracle" else lh.cuda_library()
    with lh.use_library(lib):
        yield lib


def test_simulation_with_column_params(backend, oracle):
    """HybridBox of 6 x 4 columns with per-column K_sat and van Genuchten n through Simulation(column_params=...)."""
    box = lh.HybridBox(zlim=(-1.0, 0.0), nelements=(6, 4, 16))
    model = lh.SoilModel(
        np.float64, domain=box, energy_model=lh.PrescribedTemperatureModel(),
        hydrology_model=lh.SoilHydrologyModel(hydraulic_model=w.sand_vg()),
        boundary_conditions=lh.SoilColumnBC(top=lh.SoilComponentBC(hydrology=lh.Dirichlet(lambda t: 0.25)),
                                            bottom=lh.SoilComponentBC(hydrology=lh.FreeDrainage())),
        soil_param_set=w.sand_soil_params(), earth_param_set=lh.EarthParameterSet())
    Y, Ya = lh.initialize_states(model, lambda z, m: {"ϑ_l": 0.12, "θ_i": 0.0}, 0.0)
    rng = np.random.default_rng(3)
    cp = {"Ksat": w.sand_vg().Ksat * 10.0 ** rng.uniform(-1, 1, 24), "n": rng.uniform(2.0, 4.5, 24)}
    sim = lh.Simulation(model, lh.SSPRK33(), Y_init=Y, dt=0.25, tspan=(0.0, 5.0), Ya_init=Ya, column_params=cp)
    sol = lh.run_(sim)
    got = np.asarray(lh.parent(sol.u[-1].soil.ϑ_l))
    assert got.shape[0] == 24 and np.all(np.isfinite(got))
    # columns differ because their parameters do
    assert np.max(np.abs(got - got[0])) > 1e-6
    # column k equals a single Column run with that column's parameters
    k = 5
    vg = lh.vanGenuchten(n=cp["n"][k], α=w.sand_vg().α, Ksat=cp["Ksat"][k], θr=w.sand_vg().θr)
    p = w.make_params(w.sand_soil_params(), vg)
    D, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE
    wl = w.Workload(model=abi.LH_MODEL_RICHARDS, ncol=1, nlayer=16, zmin=-1.0, zmax=0.0, params=p,
                    top=(N, 0.0, D, 0.25), bottom=(N, 0.0, FD, 0.0), dt=0.25)
    ctx = lh.SoilContext(oracle, wl.config())
    ctx.set_state(0, np.full((1, 16), 0.12)); ctx.set_state(1, np.zeros((1, 16)))
    ctx.step(0.0, 0.25, 20)
    ref = ctx.get_state(0)[0]
    assert np.max(np.abs(got[k] - ref)) <= 1e-10 * np.max(np.abs(ref))


# ---- per-column HEAT parameters (lh_soil_set_column_heat_params; reference parameters.jl:11-43) ---------------------------
def random_heat_params(wl, seed):
    rng = np.random.default_rng(seed)
    p, n = wl.params, wl.ncol
    om = rng.uniform(0.0, 0.15, n)
    om[::3] = 0.0                                   # some columns without organic matter (outer Kersten exponents exactly 1)
    return dict(rho_c_ds=p.rho_c_ds * rng.uniform(0.7, 1.3, n), kappa_sat_unfrozen=p.kappa_sat_unfrozen * rng.uniform(0.7, 1.4, n),
                kappa_sat_frozen=p.kappa_sat_frozen * rng.uniform(0.7, 1.4, n), kappa_solid=p.kappa_solid * rng.uniform(0.6, 1.5, n),
                nu_ss_om=om, nu_ss_quartz=rng.uniform(0.2, 0.8, n), nu_ss_gravel=rng.uniform(0.0, 0.15, n))


def test_oracle_heat_columns_are_independent_single_column_runs(oracle):
    """Column k of a run with per-column heat parameters == a homogeneous single-column run with column k's SoilParams."""
    wl = w.coupled_workload(ncol=5, nlayer=18, seed=91, ice=True)
    hp = random_heat_params(wl, 3)
    a = lh.SoilContext(oracle, wl.config())
    a.set_column_heat_params(**hp)
    wl.upload(a)
    a.step(0.0, wl.dt, 4)
    for k in (0, 1, 4):
        cfg = wl.config(ncol=1)
        for name, arr in hp.items():
            setattr(cfg.params, name, float(arr[k]))
        b = lh.SoilContext(oracle, cfg)
        for f, arr in wl.fields.items():
            b.set_state(f, arr[k:k + 1])
        b.step(0.0, wl.dt, 4)
        for f in (0, 2):
            assert np.array_equal(a.get_state(f)[k], b.get_state(f)[0]), (k, f)
    r = lh.SoilContext(oracle, w.richards_workload(ncol=4, nlayer=8).config())
    with pytest.raises(lh._abi.SoilError):
        r.set_column_heat_params(rho_c_ds=np.ones(4))
    a.set_column_heat_params()                      # all None: back to the scalars
    h = lh.SoilContext(oracle, wl.config())
    for ctx in (a, h):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 2)
    assert np.array_equal(a.get_state(2), h.get_state(2))


HEAT_CASES = {
    "coupled": lambda: w.coupled_workload(ncol=200, nlayer=64, seed=92),
    "coupled_ice": lambda: w.coupled_workload(ncol=96, nlayer=24, seed=93, ice=True),
    "heat": lambda: w.heat_workload(ncol=64, nlayer=37, seed=94),
    "heat_ice": lambda: w.heat_workload(ncol=70, nlayer=20, seed=95, ice=True),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(HEAT_CASES))
@pytest.mark.parametrize("with_hydraulic", [False, True])
def test_cuda_heat_params_match_oracle(cuda, oracle, name, with_hydraulic):
    wl = HEAT_CASES[name]()
    hp = random_heat_params(wl, seed=11)
    cp = random_column_params(wl, seed=12) if (with_hydraulic and wl.model == abi.LH_MODEL_COUPLED) else {}
    if cp:
        rescale_state(wl, cp)
        wl.top = (wl.top[0], wl.top[1], wl.top[2], 0.3)
    g, o = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_STAGE_LAUNCHES)), lh.SoilContext(oracle, wl.config())
    for ctx in (g, o):
        if cp:
            ctx.set_column_params(**cp)
        ctx.set_column_heat_params(**hp)
        wl.upload(ctx)
        ctx.rhs(0.0)
    assert "HETH" in g.kernel_info()
    fields = (0, 2) if wl.model == abi.LH_MODEL_COUPLED else (2,)
    for f in fields:
        a, r = g.get_tendency(f), o.get_tendency(f)
        scale = w.tendency_scale(o, f)
        assert np.max(np.abs(a - r) / scale[:, None]) <= 1e-12, (name, f)
    a, r = g.diagnostic(abi.LH_DIAG_KAPPA), o.diagnostic(abi.LH_DIAG_KAPPA)
    assert np.max(np.abs(a - r) / np.abs(r)) <= 2e-13
    a, r = g.diagnostic(abi.LH_DIAG_T), o.diagnostic(abi.LH_DIAG_T)
    assert np.max(np.abs(a - r) / np.abs(r)) <= 2e-13
    for ctx in (g, o):
        ctx.step(0.0, wl.dt, 6)
    for f in fields:
        a, r = g.get_state(f), o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (name, f)
    g.set_column_heat_params()                      # heat scalars again; hydraulic arrays (if any) stay
    assert "HETH" not in g.kernel_info() and (("HET" in g.kernel_info()) == bool(cp))


# ---- per-CELL (layered) hydraulic parameters (lh_soil_set_cell_params) ---------------------------------------------------
def random_cell_params(wl, seed, which=("nu", "theta_r", "vg_n", "vg_alpha", "Ksat")):
    """Soil horizons: piecewise-constant in depth with per-column horizon depths, plus a little cell noise."""
    rng = np.random.default_rng(seed)
    p, nc, n = wl.params, wl.ncol, wl.nlayer
    horizon = (np.arange(n)[None, :] >= rng.integers(1, max(2, n - 1), nc)[:, None]).astype(float)   # 0 below, 1 above the interface
    def two(lo, hi):
        a, b = rng.uniform(lo, hi, (nc, 1)), rng.uniform(lo, hi, (nc, 1))
        return (a + (b - a) * horizon) * rng.uniform(0.99, 1.01, (nc, n))
    out = {}
    if "nu" in which:
        out["nu"] = p.nu * two(0.9, 1.15)
    if "theta_r" in which:
        out["theta_r"] = two(0.0, 0.05)
    if "vg_n" in which:
        out["vg_n"] = two(1.4, 3.5)
    if "vg_alpha" in which:
        out["vg_alpha"] = p.vg_alpha * two(0.6, 1.8)
    if "Ksat" in which:
        out["Ksat"] = p.Ksat * 10.0 ** two(-1.0, 1.0)
    return out


def rescale_state_cells(wl, cp):
    p = wl.params
    S = (wl.fields[0] - p.theta_r) / ((p.nu - wl.fields[1]) - p.theta_r)
    nu = cp.get("nu", np.full((wl.ncol, wl.nlayer), p.nu))
    thr = cp.get("theta_r", np.full((wl.ncol, wl.nlayer), p.theta_r))
    wl.fields[0] = thr + S * ((nu - wl.fields[1]) - thr)


def test_oracle_cell_params_semantics(oracle):
    """Constant fields == the homogeneous model bit for bit; per-cell fields constant in depth == per-column arrays."""
    wl = w.coupled_workload(ncol=5, nlayer=14, seed=96, zlim=(-1.4, 0.0))
    p = wl.params
    ones = np.ones((5, 14))
    a, b = lh.SoilContext(oracle, wl.config()), lh.SoilContext(oracle, wl.config())
    b.set_cell_params(nu=p.nu * ones, theta_r=p.theta_r * ones, vg_n=p.vg_n * ones, vg_alpha=p.vg_alpha * ones, Ksat=p.Ksat * ones)
    for ctx in (a, b):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 3)
    for f in (0, 2):
        assert np.array_equal(a.get_state(f), b.get_state(f))
    ks = p.Ksat * np.array([1.0, 2.0, 0.5, 1.0, 3.0])
    c, d = lh.SoilContext(oracle, wl.config()), lh.SoilContext(oracle, wl.config())
    c.set_column_params(Ksat=ks)
    d.set_cell_params(Ksat=ks[:, None] * ones)
    for ctx in (c, d):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 3)
    assert np.array_equal(c.get_state(0), d.get_state(0)) and not np.array_equal(a.get_state(0), c.get_state(0))
    # a less conductive top horizon slows the infiltration front coming from the Dirichlet top
    e = lh.SoilContext(oracle, wl.config())
    k2 = p.Ksat * ones.copy(); k2[:, 10:] *= 0.01
    e.set_cell_params(Ksat=k2)
    wl.upload(e)
    e.step(0.0, wl.dt, 3)
    assert not np.array_equal(e.get_state(0), a.get_state(0))
    with pytest.raises(ValueError):
        e.set_cell_params(nu=np.ones((5, 13)))
    e.set_cell_params()
    wl.upload(e)
    e.step(0.0, wl.dt, 3)
    assert np.array_equal(e.get_state(0), a.get_state(0))


CELL_CASES = {
    "coupled": lambda: w.coupled_workload(ncol=200, nlayer=64, seed=97),
    "coupled_ice": lambda: w.coupled_workload(ncol=96, nlayer=24, seed=98, ice=True, viscosity=lh.TemperatureDependentViscosity(),
                                              impedance=lh.IceImpedance()),
    "richards": lambda: w.richards_workload(ncol=130, nlayer=100, seed=99),
    "richards_tall": lambda: w.richards_workload(ncol=40, nlayer=300, seed=100, zlim=(-4.5, 0.0)),
    "heat": lambda: w.heat_workload(ncol=64, nlayer=37, seed=101),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CELL_CASES))
@pytest.mark.parametrize("launch", ["stage", "persistent"])
def test_cuda_cell_params_match_oracle(cuda, oracle, name, launch):
    wl = CELL_CASES[name]()
    cp = random_cell_params(wl, seed=17)
    rescale_state_cells(wl, cp)
    if wl.model == abi.LH_MODEL_COUPLED:
        wl.top = (wl.top[0], wl.top[1], wl.top[2], 0.3)
    flags = abi.LH_FLAG_STAGE_LAUNCHES if launch == "stage" else abi.LH_FLAG_PERSISTENT
    g, o = lh.SoilContext(cuda, wl.config(flags=flags)), lh.SoilContext(oracle, wl.config())
    for ctx in (g, o):
        ctx.set_cell_params(**cp)
        if name == "coupled":                       # together with per-column heat parameters
            ctx.set_column_heat_params(**random_heat_params(wl, 18))
        wl.upload(ctx)
        ctx.rhs(0.0)
    assert "CELLP" in g.kernel_info()
    fields = (0, 2) if wl.model == abi.LH_MODEL_COUPLED else (0,) if wl.model == abi.LH_MODEL_RICHARDS else (2,)
    for f in fields:
        a, r = g.get_tendency(f), o.get_tendency(f)
        scale = w.tendency_scale(o, f)
        assert np.max(np.abs(a - r) / scale[:, None]) <= 1e-12, (name, f)
    for which in (abi.LH_DIAG_K, abi.LH_DIAG_PSI, abi.LH_DIAG_KAPPA):
        a, r = g.diagnostic(which), o.diagnostic(which)
        assert np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300)) <= 2e-13, (name, which)
    for ctx in (g, o):
        ctx.step(0.0, wl.dt, 6)
    for f in fields:
        a, r = g.get_state(f), o.get_state(f)
        assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (name, f)
    g.set_cell_params()
    assert "CELLP" not in g.kernel_info()
