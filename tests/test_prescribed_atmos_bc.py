"""PrescribedAtmosForcing (reference src/SoilModel/boundary_conditions.jl:103-131, 516-620), restating the reference's own
test/SoilModel/test_prescribed_atmos_bc.jl through the host mirror and the C ABI.

PARITY UNPINNED: `surface_conditions` (SurfaceFluxes v0.1) and `q_vap_saturation_generic` / `cp_m` (Thermodynamics v0.5) are
not under the reference tree; the reference test compares the function with a re-evaluation through the SAME packages, so it
pins only structure.  That structure is what is tested here — on the oracle (CPU) and on the CUDA path (GPU) — plus CUDA vs
oracle agreement and physical sanity of the restated similarity solution."""
import numpy as np
import pytest

import workloads as w

lh, abi = w.lh, w.abi
param_set = lh.EarthParameterSet()


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, oracle):
    lib = oracle if request.param == "oracle" else lh.cuda_library()
    with lh.use_library(lib):
        yield lib


def reference_setup(T_surf=299.0, q_atm=None, u_atm=0.34, θ_atm=None):
    """test_prescribed_atmos_bc.jl:9-63"""
    ν, vg_n, vg_α, θ_r = 0.55, 1.68, 5.0, 0.084
    msp = lh.SoilParams(ν=ν, ρc_ds=1.0)
    domain = lh.Column(zlim=(-0.55, 0.0), nelements=10)
    hm = lh.vanGenuchten(n=vg_n, α=vg_α, Ksat=0.0, θr=θ_r)
    ρ_a_sfc = 1.17
    if q_atm is None:
        q_atm = lh.q_vap_saturation_liquid(param_set, T_surf, ρ_a_sfc)           # :28
    surface_bc = lh.PrescribedAtmosForcing(u_atm=u_atm, θ_atm=T_surf if θ_atm is None else θ_atm, z_atm=0.05, θ_scale=T_surf,
                                           ρ_a_sfc=ρ_a_sfc, q_atm=q_atm)
    bc = lh.SoilColumnBC(top=surface_bc, bottom=lh.SoilComponentBC(energy=lh.VerticalFlux(0.0), hydrology=lh.VerticalFlux(0.0)))
    model = lh.SoilModel(domain=domain, energy_model=lh.SoilEnergyModel(), hydrology_model=lh.SoilHydrologyModel(hydraulic_model=hm),
                         boundary_conditions=bc, soil_param_set=msp, earth_param_set=param_set)
    return model, hm, ν


def saturated_ic(z, model):
    """:66-80: saturated soil at 299 K"""
    ps, sp = model.earth_param_set, model.soil_param_set
    ρc_s = lh.volumetric_heat_capacity(sp.ν, 0.0, sp.ρc_ds, ps)
    return dict(ϑ_l=sp.ν, θ_i=0.0, ρe_int=lh.volumetric_internal_energy(0.0, ρc_s, 299.0, ps))


def test_equilibrium_gives_exactly_zero_tendency(backend):
    """:75-79: atmosphere at the surface temperature, saturated with respect to the surface: `sum(parent(dY)) == 0.0`."""
    model, _, _ = reference_setup()
    Y, Ya = lh.initialize_states(model, saturated_ic, 0.0)
    dY = lh.similar(Y)
    lh.make_rhs(model)(dY, Y, Ya, 0.0)
    tot = sum(float(np.sum(dY.soil[k])) for k in ("ϑ_l", "θ_i", "ρe_int"))
    assert tot == 0.0
    assert all(np.all(dY.soil[k] == 0.0) for k in ("ϑ_l", "θ_i", "ρe_int"))


def test_flux_structure(backend):
    """:81-156: the four surface states of the reference test."""
    model, hm, ν = reference_setup()
    T_surf = 299.0
    ϑ_l = np.array([ν, ν + 1e-3, ν - 1e-3, ν])
    θ_i = np.array([0.0, 0.0, 0.0, 0.1])
    T = np.array([T_surf, T_surf, 289.5, 289.5])
    heat, water = lh.compute_turbulent_surface_fluxes(model.energy_model, model.hydrology_model, model, ϑ_l, θ_i, T)
    assert heat.dtype == np.float64 and water.dtype == np.float64                       # :158-159 typing
    assert heat[0] == heat[1] and water[0] == water[1]                                  # :155 oversaturated == exactly saturated
    assert heat[0] == 0.0 and water[0] == 0.0                                           # equilibrium state: t* = q* = 0 (:150-153)
    # colder surface under the same air: sensible heat flows DOWN into the soil (negative = along -z) and the air, saturated
    # at 299 K, is supersaturated with respect to the 289.5 K surface: condensation onto the soil (E < 0)
    assert heat[2] < 0.0 and water[2] < 0.0 and heat[3] < 0.0 and water[3] < 0.0
    # the scalar form returns a pair of floats
    h0, w0 = lh.compute_turbulent_surface_fluxes(model.energy_model, model.hydrology_model, model, float(ϑ_l[2]), 0.0, 289.5)
    assert h0 == heat[2] and w0 == water[2]


def test_no_method_for_prescribed_components_or_bottom(backend):
    """:161-194"""
    model, _, _ = reference_setup()
    for en, hy in ((lh.PrescribedTemperatureModel(), lh.PrescribedHydrologyModel()),
                   (lh.SoilEnergyModel(), lh.PrescribedHydrologyModel()),
                   (lh.PrescribedTemperatureModel(), lh.SoilHydrologyModel())):
        with pytest.raises(Exception):
            lh.compute_turbulent_surface_fluxes(en, hy, model, 0.55, 0.0, 299.0)
    with pytest.raises(Exception):
        lh.boundary_fluxes(None, model.boundary_conditions.top, "bottom", model, None, None)
    # a model with a prescribed component cannot be built into a context with this BC
    bad = lh.SoilModel(domain=model.domain, energy_model=lh.PrescribedTemperatureModel(),
                       hydrology_model=lh.SoilHydrologyModel(hydraulic_model=lh.vanGenuchten(n=1.68, α=5.0, Ksat=0.0, θr=0.084)),
                       boundary_conditions=model.boundary_conditions, soil_param_set=model.soil_param_set, earth_param_set=param_set)
    with pytest.raises(Exception):
        lh.SoilEngine(bad, 0.0)


def test_abi_level_errors(oracle):
    wl = w.richards_workload(ncol=4, nlayer=8)
    ctx = lh.SoilContext(oracle, wl.config())
    a = lh.build_atmos(reference_setup()[0])
    assert oracle.soil_set_atmos_forcing(ctx._h, a) == abi.LH_ERR_UNSUPPORTED_BC
    c = lh.SoilContext(oracle, w.coupled_workload(ncol=4, nlayer=8, zlim=(-0.8, 0.0)).config())
    a.struct_size = 8
    assert oracle.soil_set_atmos_forcing(c._h, a) == abi.LH_ERR_INVALID_ARG
    with pytest.raises(lh._abi.SoilError):
        c.atmos_fluxes(0.3, 0.0, 290.0)                      # no forcing set


def test_fluxes_are_a_fixed_point_of_the_similarity_relations(backend):
    """Unstable, stable and neutral cases: recover (u*, θ*, q*) from the returned fluxes and check that they satisfy
    u* = κ Δu / (ln(z/z0) - ψ_m(z/L) + ψ_m(z0/L)) etc. with L from the same u*, θ* (Businger-Dyer functions)."""
    import math

    for θ_atm, q_scale in ((305.0, 0.5), (285.0, 0.5), (299.0, 0.5), (299.5, 1.0)):
        model, hm, ν = reference_setup(θ_atm=θ_atm)
        ep, bc, sp = model.earth_param_set, model.boundary_conditions.top, model.soil_param_set
        model2, _, _ = reference_setup(θ_atm=θ_atm, q_atm=bc.q_atm * q_scale)
        bc = model2.boundary_conditions.top
        T = 299.0
        heat, water = lh.compute_turbulent_surface_fluxes(model2.energy_model, model2.hydrology_model, model2, ν - 0.1, 0.0, T)
        # surface humidity as the reference defines it (:584-593)
        S = (ν - 0.1 - hm.θr) / (ν - hm.θr)
        ψ = lh.matric_potential(hm, S)
        q_s = lh.q_vap_saturation_liquid(ep, T, bc.ρ_a_sfc) * math.exp(ep.grav * ψ / ep.R_v / T)
        E = water * ep.ρ_cloud_liq
        cpm = ep.cp_d + (ep.cp_v - ep.cp_d) * q_s
        h_d = ep.cp_d * (T - ep.T_0) + ep.R_d * ep.T_0
        lhv = ep.cp_v * (T - ep.T_0) + ep.LH_v0
        ust_tst = -(heat - lhv * E + h_d * E) / (cpm * bc.ρ_a_sfc)          # u* θ*
        ust_qst = -E / bc.ρ_a_sfc                                           # u* q*
        dθ, dq, du = bc.θ_atm - T, bc.q_atm - q_s, bc.u_atm
        κ, z, z0m, z0s = ep.von_karman_const, bc.z_atm, sp.z_0m, sp.z_0s

        def psi_m(ζ):
            if ζ >= 0:
                return -ep.a_m * ζ
            X = (1 - 15 * ζ) ** 0.25
            return 2 * math.log((1 + X) / 2) + math.log((1 + X * X) / 2) - 2 * math.atan(X) + math.pi / 2

        def psi_h(ζ):
            if ζ >= 0:
                return -ep.a_h * ζ / ep.Pr_0
            return 2 * math.log((1 + math.sqrt(1 - 9 * ζ)) / 2)

        # solve the 1-D problem independently by bisection on x = 1/L
        def G(x):
            us = κ * du / (math.log(z / z0m) - psi_m(z * x) + psi_m(z0m * x))
            ts = κ * dθ / (ep.Pr_0 * (math.log(z / z0s) - psi_h(z * x) + psi_h(z0s * x)))
            return us, ts, x - κ * ep.grav * ts / (us * us * bc.θ_scale)

        lo, hi = -1000.0 / z, 10.0 / z
        flo = G(lo)[2]
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            fm = G(mid)[2]
            if (fm > 0) == (flo > 0):
                lo, flo = mid, fm
            else:
                hi = mid
        us, ts, _ = G(0.5 * (lo + hi))
        qs = κ * dq / (ep.Pr_0 * (math.log(z / z0s) - psi_h(z * 0.5 * (lo + hi)) + psi_h(z0s * 0.5 * (lo + hi))))
        assert abs(ust_tst - us * ts) <= 1e-9 * max(abs(us * ts), 1e-5), (θ_atm, ust_tst, us * ts)      # u* θ* ~ 1e-2 when Δθ != 0
        assert abs(ust_qst - us * qs) <= 1e-9 * max(abs(us * qs), 1e-8), (θ_atm, ust_qst, us * qs)


def test_run_with_atmospheric_forcing_conserves_what_the_surface_exchanges(backend):
    """A drying, warming column (warm dry air above moist soil): the water the column loses per step equals the
    time-integrated evaporation Ẽ dt within the stepping error, and the state stays finite and physical."""
    ν = 0.55
    msp = lh.SoilParams(ν=ν, ρc_ds=2.0e6)
    hm = lh.vanGenuchten(n=1.68, α=5.0, Ksat=1e-6, θr=0.084)
    bcA = lh.PrescribedAtmosForcing(u_atm=2.0, θ_atm=303.0, z_atm=2.0, θ_scale=300.0, ρ_a_sfc=1.17,
                                    q_atm=0.4 * lh.q_vap_saturation_liquid(param_set, 303.0, 1.17))
    model = lh.SoilModel(domain=lh.Column(zlim=(-1.0, 0.0), nelements=20), energy_model=lh.SoilEnergyModel(),
                         hydrology_model=lh.SoilHydrologyModel(hydraulic_model=hm),
                         boundary_conditions=lh.SoilColumnBC(top=bcA, bottom=lh.SoilComponentBC(energy=lh.VerticalFlux(0.0),
                                                                                                hydrology=lh.VerticalFlux(0.0))),
                         soil_param_set=msp, earth_param_set=param_set)

    def ic(z, m):
        ρc_s = lh.volumetric_heat_capacity(0.4, 0.0, m.soil_param_set.ρc_ds, m.earth_param_set)
        return dict(ϑ_l=0.4, θ_i=0.0, ρe_int=lh.volumetric_internal_energy(0.0, ρc_s, 295.0, m.earth_param_set))

    Y, Ya = lh.initialize_states(model, ic, 0.0)
    dt, nsteps = 10.0, 50
    sim = lh.Simulation(model, lh.SSPRK33(), Y_init=lh.copy(Y), dt=dt, tspan=(0.0, dt * nsteps), Ya_init=Ya, saveat=dt * nsteps)
    lh.run_(sim)
    u = sim.integrator.u
    assert np.all(np.isfinite(u.soil["ϑ_l"])) and np.all(np.isfinite(u.soil["ρe_int"]))
    dz = 1.0 / 20
    lost = (np.sum(Y.soil["ϑ_l"]) - np.sum(u.soil["ϑ_l"])) * dz                 # m of water
    _, E0 = lh.compute_turbulent_surface_fluxes(model.energy_model, model.hydrology_model, model, 0.4, 0.0, 295.0)
    assert E0 > 0.0 and lost > 0.0
    assert 0.3 * E0 * dt * nsteps < lost < 3.0 * E0 * dt * nsteps                # same order: the surface dries and warms as it goes
    assert u.soil["ϑ_l"][-1] < 0.4                                               # the top cell dried


@pytest.mark.gpu
def test_cuda_matches_oracle_fluxes_and_steps(cuda, oracle):
    """PrescribedAtmosForcing through the C ABI: surface fluxes for random states (1e-12 relative), the tendency under the
    forcing (1e-12, cancellation-aware norm) and the state after 10 steps (1e-10), homogeneous and per-column parameters."""
    model, _, _ = reference_setup(θ_atm=303.0)
    atm = lh.build_atmos(model)
    wl = w.coupled_workload(ncol=300, nlayer=40, seed=121, top=(w.F, 0.0, w.F, 0.0))
    wl.params.z_0m, wl.params.z_0s = 0.001, 0.002
    rng = np.random.default_rng(4)
    th = rng.uniform(0.05, 0.55, 500)
    ti = np.where(rng.uniform(size=500) < 0.3, rng.uniform(0.0, 0.04, 500), 0.0)
    T = rng.uniform(270.0, 320.0, 500)
    for het in (False, True):
        g, o = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_STAGE_LAUNCHES)), lh.SoilContext(oracle, wl.config())
        for ctx in (g, o):
            if het:
                cp = dict(nu=wl.params.nu * np.random.default_rng(5).uniform(1.0, 1.1, wl.ncol), Ksat=wl.params.Ksat * np.random.default_rng(6).uniform(0.5, 2.0, wl.ncol))
                ctx.set_column_params(**cp)
            ctx.set_atmos_forcing(atm)
            wl.upload(ctx)
        if not het:
            hg, wg = g.atmos_fluxes(th, ti, T)
            ho, wo = o.atmos_fluxes(th, ti, T)
            assert np.max(np.abs(hg - ho) / np.maximum(np.abs(ho), 1e-3)) <= 1e-11
            assert np.max(np.abs(wg - wo) / np.maximum(np.abs(wo), 1e-12)) <= 1e-11
        for ctx in (g, o):
            ctx.rhs(0.0)
        for f in (0, 2):
            a, r = g.get_tendency(f), o.get_tendency(f)
            scale = w.tendency_scale(o, f)
            assert np.max(np.abs(a - r) / scale[:, None]) <= 1e-12, (het, f)
        for ctx in (g, o):
            ctx.step(0.0, wl.dt, 10)
        for f in (0, 2):
            a, r = g.get_state(f), o.get_state(f)
            assert np.max(np.abs(a - r)) <= 1e-10 * np.max(np.abs(r)), (het, f)
        g.set_atmos_forcing(None)                           # back to the configured (zero-flux) top
        o.set_atmos_forcing(None)


def test_column_flux_fields(backend):
    """Per-column prescribed fluxes (lh_soil_set_column_fluxes): column k == a homogeneous run with that column's flux."""
    lib = backend
    wl = w.coupled_workload(ncol=6, nlayer=16, seed=122, top=(w.F, 0.0, w.F, 0.0), bottom=(w.F, 0.0, w.F, 0.0), zlim=(-1.6, 0.0))
    rng = np.random.default_rng(8)
    fe, fw = rng.uniform(-50.0, 50.0, 6), rng.uniform(-1e-7, 1e-7, 6)
    a = lh.SoilContext(lib, wl.config())
    a.set_column_fluxes(top_energy=fe, top_hydrology=fw)
    wl.upload(a)
    a.step(0.0, wl.dt, 4)
    for k in (0, 3, 5):
        one = w.coupled_workload(ncol=6, nlayer=16, seed=122, top=(w.F, float(fe[k]), w.F, float(fw[k])), bottom=(w.F, 0.0, w.F, 0.0),
                                 zlim=(-1.6, 0.0))
        b = lh.SoilContext(lib, one.config())
        one.upload(b)
        b.step(0.0, one.dt, 4)
        for f in (0, 2):
            assert np.array_equal(a.get_state(f)[k], b.get_state(f)[k]), (k, f)
    a.set_column_fluxes()                                   # all None: the scalars again
    d = lh.SoilContext(lib, wl.config())
    for ctx in (a, d):
        wl.upload(ctx)
        ctx.step(0.0, wl.dt, 2)
    assert np.array_equal(a.get_state(0), d.get_state(0))
    bad = lh.SoilContext(lib, w.coupled_workload(ncol=6, nlayer=16, seed=1).config())     # Dirichlet top
    with pytest.raises(lh._abi.SoilError):
        bad.set_column_fluxes(top_energy=fe)


def test_dirichlet_as_flux_property(backend):
    """The intent of test/SoilModel/dirichlet_bc_as_flux.jl:200-227 (not run by the reference's driver): a Dirichlet condition is
    applied by converting it into a boundary flux, so the tendency under Dirichlet(ϑ_l, T) at the top equals, bit for bit, the
    tendency under VerticalFlux with the fluxes the host closures give for that face (boundary_conditions.jl:371-444)."""
    lib = backend
    wl = w.coupled_workload(ncol=5, nlayer=24, seed=131, zlim=(-2.4, 0.0))
    p = wl.params
    a = lh.SoilContext(lib, wl.config())
    wl.upload(a)
    a.rhs(0.0)
    th, ti, T = wl.fields[0][:, -1], wl.fields[1][:, -1], a.diagnostic(abi.LH_DIAG_T)[:, -1]
    psi_c, = (a.diagnostic(abi.LH_DIAG_PSI)[:, -1],)
    dzb = (wl.zmax - wl.zmin) / wl.nlayer / 2.0
    # face state: Dirichlet ϑ_l and T at the face, θ_i of the cell (boundary_conditions.jl:218-288)
    face = lh.SoilContext(lib, w.coupled_workload(ncol=5, nlayer=1, seed=1, zlim=(-0.1, 0.0)).config())
    ρc_s = p.rho_c_ds + np.minimum(wl.top[3], p.nu - ti) * p.cp_l * p.rho_cloud_liq + ti * p.cp_i * p.rho_cloud_ice
    face.set_state(0, np.full((5, 1), wl.top[3])); face.set_state(1, ti[:, None].copy())
    face.set_state(2, (ρc_s * (wl.top[1] - p.T_0) - ti * p.rho_cloud_ice * p.LH_f0)[:, None])
    K_f, psi_f, kap_f = (face.diagnostic(k)[:, 0] for k in (abi.LH_DIAG_K, abi.LH_DIAG_PSI, abi.LH_DIAG_KAPPA))
    fw = -K_f * (psi_f - psi_c + dzb) / dzb                                   # :371-401 (top face: no sign flip)
    fe = -kap_f * (wl.top[1] - T) / dzb                                        # :416-444
    fl = w.coupled_workload(ncol=5, nlayer=24, seed=131, zlim=(-2.4, 0.0), top=(w.F, 0.0, w.F, 0.0))
    b = lh.SoilContext(lib, fl.config())
    b.set_column_fluxes(top_energy=fe, top_hydrology=fw)
    fl.upload(b)
    b.rhs(0.0)
    for f in (0, 2):
        ra, rb = a.get_tendency(f), b.get_tendency(f)
        assert np.array_equal(ra[:, :-1], rb[:, :-1])                           # interior cells do not see the top face at all
        top_scale = np.max(np.abs(ra[:, -1]))
        assert np.max(np.abs(ra[:, -1] - rb[:, -1])) <= 1e-12 * top_scale       # the face flux, re-derived from the diagnostics
