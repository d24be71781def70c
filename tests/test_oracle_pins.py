"""Pin the oracle against every known-answer test the reference holds for this path.

The reference cannot run here (Julia is absent), so the oracle (oracle/lho_soil.c) is pinned to
the reference's OWN tests: each test below quotes the file:line under /root/reference/test it
restates, with the reference's threshold.  CPU only.
"""
import ctypes as C
import math

import numpy as np
import pytest

import workloads as w

lh = w.lh
abi = w.abi
EPS = np.finfo(np.float64).eps


def _fn(oracle, name, nargs_after_p=None, with_p=True, nargs=None):
    f = oracle.raw(name)
    f.restype = C.c_double
    if with_p:
        f.argtypes = [C.POINTER(abi.lh_soil_params)] + [C.c_double] * nargs_after_p
    else:
        f.argtypes = [C.c_double] * nargs
    return f


@pytest.fixture(scope="module")
def fns(oracle):
    P = lambda name, n: _fn(oracle, name, n)
    S = lambda name, n: _fn(oracle, name, with_p=False, nargs=n)
    return dict(
        vlf=S("lho_volumetric_liquid_fraction", 2),
        eff_sat=S("lho_effective_saturation", 3),
        psi_m=P("lho_matric_potential", 1),
        inv_psi=P("lho_inverse_matric_potential", 1),
        head=P("lho_pressure_head", 3),
        K=P("lho_hydraulic_conductivity", 3),
        visc=P("lho_viscosity_factor", 1),
        imp=P("lho_impedance_factor", 1),
        hydrostatic=P("lho_hydrostatic_profile", 4),
        rho_c_s=P("lho_volumetric_heat_capacity", 3),
        T_from=P("lho_temperature_from_rho_e_int", 3),
        rho_e=P("lho_volumetric_internal_energy", 3),
        k_sat=S("lho_saturated_thermal_conductivity", 4),
        S_r=S("lho_relative_saturation", 3),
        kersten=P("lho_kersten_number", 2),
        kappa=S("lho_thermal_conductivity", 3),
        rho_e_liq=P("lho_volumetric_internal_energy_liq", 1),
        k_solid=S("lho_k_solid", 5),
        ksat_frozen=S("lho_ksat_frozen", 3),
        ksat_unfrozen=S("lho_ksat_unfrozen", 3),
        k_dry=P("lho_k_dry", 0),
    )


def test_water_parameterizations(fns):
    """test/SoilModel/test_water_parameterizations.jl:1-63 (Float32 there, fp64 here)."""
    θr = 0.2
    vg = lh.vanGenuchten(θr=θr)                 # defaults n = 1.56, α = 3.6, Ksat = 2.9e-7
    p = w.make_params(lh.SoilParams(), vg)
    pp = C.byref(p)
    n, α, m, Ksat = vg.n, vg.α, vg.m, vg.Ksat
    assert m == 1.0 - 1.0 / 1.56
    ν, S_s = 0.4, 1e-2
    θ = [0.3, 0.4, 0.5]
    S = [fns["eff_sat"](ν, t, θr) for t in θ]
    assert np.allclose(S, [0.5, 1.0, 1.5], rtol=1e-14)                       # :10-13
    va = -(((S[0] ** (-1.0 / m) - 1.0) * α ** (-n)) ** (1.0 / n))            # :16
    ψ = [fns["psi_m"](pp, s) for s in S[:2]]
    assert np.allclose([fns["inv_psi"](pp, x) for x in ψ], S[:2], rtol=1e-13)  # :18
    assert math.isnan(fns["inv_psi"](pp, 1.0))                               # :19 (throws there)
    assert ψ[0] == va and abs(ψ[1]) < 1e-7                                   # :20  (S≈1 -> ψ≈0)
    head = [fns["head"](pp, t, ν, S_s) for t in θ]
    assert np.allclose(head[:2], ψ, atol=1e-7) and np.isclose(head[2], 10.0, rtol=1e-14)  # :24-26: (0.5-0.4)/1e-2
    k = [fns["K"](pp, s, 1.0, 1.0) for s in S]
    va = (math.sqrt(S[0]) * (1.0 - (1.0 - S[0] ** (1.0 / m)) ** m) ** 2.0) * Ksat
    assert k[0] == va and k[2] == Ksat and np.isclose(k[1], Ksat, rtol=1e-6)      # :30-36
    p_imp = w.make_params(lh.SoilParams(), vg, impedance=lh.IceImpedance())
    assert np.isclose(fns["imp"](C.byref(p_imp), 1.0), 1e-7, rtol=1e-14)      # :40-41
    assert fns["imp"](pp, 1.0) == 1.0                                        # NoEffect
    p_v = w.make_params(lh.SoilParams(), vg, viscosity=lh.TemperatureDependentViscosity())
    for T in (278.0, 288.0, 298.0):                                          # :44-46
        assert fns["visc"](C.byref(p_v), T) == math.exp(2.64e-2 * (T - 288.0))
    assert fns["visc"](pp, 300.0) == 1.0
    z = np.arange(-1.0, 0.0 + 1e-12, 0.1)                                    # :49-54
    θh = [fns["hydrostatic"](pp, zz, -0.5, ν, S_s) for zz in z]
    h = np.array([fns["head"](pp, t, ν, S_s) for t in θh]) + z
    assert np.std(h, ddof=1) < 1e-6
    assert [fns["vlf"](x, 0.5) for x in (0.25, 0.5, 0.75)] == [0.25, 0.5, 0.5]  # :57-58


def test_heat_parameterizations(fns):
    """test/SoilModel/test_heat_parameterizations.jl:5-83: bitwise `==` against the literal formula."""
    ep = lh.EarthParameterSet()
    ρ_l, ρ_i = ep.ρ_cloud_liq, ep.ρ_cloud_ice
    ρcp_l, ρcp_i = ep.cp_l * ρ_l, ep.cp_i * ρ_i
    T_ref, LH = ep.T_0, ep.LH_f0
    κ_air = ep.K_therm
    sp = lh.SoilParams(ν=0.2, S_s=1e-3, ν_ss_om=0.1, ν_ss_gravel=0.1, ν_ss_quartz=0.1, ρc_ds=0.0,
                       κ_solid=0.1, ρp=1.0, κ_sat_unfrozen=0.0, κ_sat_frozen=0.0)
    p = w.make_params(sp, lh.vanGenuchten())
    pp = C.byref(p)
    assert fns["T_from"](pp, 5.4e7, 0.05, 2.1415e6) == T_ref + (5.4e7 + 0.05 * ρ_i * LH) / 2.1415e6   # :22-23
    assert fns["rho_c_s"](pp, 0.25, 0.05, 1e6) == 1e6 + 0.25 * ρcp_l + 0.05 * ρcp_i                   # :25-26
    assert fns["rho_e"](pp, 0.05, 2.1415e6, 300.0) == 2.1415e6 * (300.0 - T_ref) - 0.05 * ρ_i * LH     # :28-29
    assert fns["k_sat"](0.25, 0.05, 0.57, 2.29) == 0.57 ** (0.25 / (0.05 + 0.25)) * 2.29 ** (0.05 / (0.05 + 0.25))  # :31-32
    assert fns["k_sat"](0.0, 0.0, 0.57, 2.29) == 0.0                                                  # :34
    assert fns["S_r"](0.25, 0.05, 0.4) == (0.25 + 0.05) / 0.4                                         # :36
    assert fns["kersten"](pp, 0.0, 0.75) == 0.75 ** ((1.0 + 0.1 - 0.24 * 0.1 - 0.1) / 2.0) * (
        (1.0 + math.exp(-18.1 * 0.75)) ** (-3.0) - ((1.0 - 0.75) / 2.0) ** 3.0
    ) ** (1.0 - 0.1)                                                                                  # :52-58
    assert fns["kersten"](pp, 0.05, 0.75) == 0.75 ** (1.0 + 0.1)                                      # :61-62
    assert fns["kappa"](1.5, 0.7287, 0.7187) == 0.7287 * 0.7187 + (1.0 - 0.7287) * 1.5                # :64-65
    assert fns["rho_e_liq"](pp, 300.0) == ρcp_l * (300.0 - T_ref)                                     # :67-68
    assert fns["k_solid"](0.5, 0.25, 2.0, 3.0, 2.0) == 2.0 ** 0.5 * 2.0 ** 0.25 * 3.0 ** 0.25          # :70-71
    assert fns["ksat_frozen"](0.5, 0.1, 0.4) == 0.5 ** 0.9 * 0.4 ** 0.1                               # :73-74
    assert fns["ksat_unfrozen"](0.5, 0.1, 0.4) == 0.5 ** 0.9 * 0.4 ** 0.1                             # :76-77
    assert fns["k_dry"](pp) == ((0.053 * 0.1 - κ_air) * 0.8 + κ_air * 1.0) / (1.0 - (1.0 - 0.053) * 0.8)  # :79-81


def test_k_therm_pin(fns):
    """test/SoilModel/heat_test_interface.jl:7: ρc_ds = 0.43314518988433487 is k_dry of that soil
    (κ_dry/ρc_ds = 1 exactly), which pins K_therm = 0.024 numerically."""
    sp = lh.SoilParams(ν=0.495, ν_ss_gravel=0.1, ν_ss_om=0.1, ν_ss_quartz=0.1, κ_solid=8.0)
    p = w.make_params(sp, lh.vanGenuchten())
    assert fns["k_dry"](C.byref(p)) == 0.43314518988433487


def test_host_closures_match_oracle(fns):
    """The package's host-side closures (setup helpers) and the oracle are literal twins."""
    rng = np.random.default_rng(5)
    sp, vg, ep = w.coupled_soil_params(), w.coupled_vg(), lh.EarthParameterSet()
    p = w.make_params(sp, vg)
    pp = C.byref(p)
    for _ in range(200):
        S = rng.uniform(0.02, 0.999)
        θl, θi, T = rng.uniform(0.05, 0.45), rng.uniform(0, 0.04), rng.uniform(270, 300)
        assert lh.matric_potential(vg, S) == fns["psi_m"](pp, S)
        assert lh.hydraulic_conductivity(vg, S, 1.0, 1.0) == fns["K"](pp, S, 1.0, 1.0)
        assert lh.pressure_head(vg, θl, sp.ν - θi, sp.S_s) == fns["head"](pp, θl, sp.ν - θi, sp.S_s)
        ρc = lh.volumetric_heat_capacity(θl, θi, sp.ρc_ds, ep)
        assert ρc == fns["rho_c_s"](pp, θl, θi, sp.ρc_ds)
        assert lh.volumetric_internal_energy(θi, ρc, T, ep) == fns["rho_e"](pp, θi, ρc, T)
        Sr = lh.relative_saturation(θl, θi, sp.ν)
        assert lh.kersten_number(θi, Sr, sp) == fns["kersten"](pp, θi, Sr)
        assert lh.kersten_number(0.0, Sr, sp) == fns["kersten"](pp, 0.0, Sr)
        assert lh.saturated_thermal_conductivity(θl, θi, sp.κ_sat_unfrozen, sp.κ_sat_frozen) == fns["k_sat"](
            θl, θi, sp.κ_sat_unfrozen, sp.κ_sat_frozen)
    assert lh.k_dry(ep, sp) == fns["k_dry"](pp)


# --- one RHS evaluation with an analytic expectation ------------------------------------------------
def _coupled_zero_flux_ctx(lib, n=20):
    p = w.make_params(w.coupled_soil_params(), w.coupled_vg())
    F = abi.LH_BC_FLUX
    wl = w.Workload(model=abi.LH_MODEL_COUPLED, ncol=1, nlayer=n, zmin=-2.0, zmax=0.0, params=p,
                    top=(F, 0.0, F, 0.0), bottom=(F, 0.0, F, 0.0), dt=20.0)
    return wl, lh.SoilContext(lib, wl.config())


def test_rhs_known_answer(oracle, fns):
    """test/SoilModel/coupled.jl:196-234: default ICs (ϑ_l = ν/2 = 0.25, θ_i = 0, T = T_0) and one
    RHS evaluation: dθ_i = 0, dρe_int = 0, dϑ_l = -div(-K ∇h) with interior flux -K and zero
    boundary fluxes."""
    wl, ctx = _coupled_zero_flux_ctx(oracle)
    p = wl.params
    assert np.allclose(ctx.zc(), np.arange(-1.95, 0.0, 0.1), rtol=0, atol=1e-14)       # :198
    ep, sp = lh.EarthParameterSet(), w.coupled_soil_params()
    θl = 0.5 * sp.ν
    ρc_s = lh.volumetric_heat_capacity(θl, 0.0, sp.ρc_ds, ep)
    ρe = lh.volumetric_internal_energy(0.0, ρc_s, 273.16, ep)
    ctx.set_state(0, np.full(20, θl)); ctx.set_state(1, np.zeros(20)); ctx.set_state(2, np.full(20, ρe))
    ctx.rhs(0.0)
    dϑ, dθi, dρe = (ctx.get_tendency(f)[0] for f in (0, 1, 2))
    assert np.allclose(dθi, 0.0, atol=1e-8)                                            # :221
    assert np.allclose(dρe, 0.0, atol=1e-8)                                            # :222
    S = fns["eff_sat"](sp.ν, 0.25, 0.0)
    K = fns["K"](C.byref(p), S, 1.0, 1.0)
    assert K == pytest.approx(1.5618205801102845e-09, rel=1e-14)
    expected_flux = np.zeros(21) - K
    expected_flux[-1] = 0.0
    expected_flux[0] = 0.0
    minus_div_flux = -(expected_flux[1:] - expected_flux[:-1]) / 0.1
    assert np.sum(dϑ - minus_div_flux) < EPS                                           # :234
    # stronger than the reference's (signed-sum) check: element-wise, scaled by K/dz
    assert np.max(np.abs(dϑ - minus_div_flux)) <= 1e-12 * K / 0.1


def _steps(ctx, t0, dt, nsteps):
    ctx.step(t0, dt, nsteps)


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def anylib(request, oracle):
    """The reference's own integration tests, at the reference's own thresholds: on the oracle (CPU) and, marked gpu,
    on the product through the C ABI (single columns: the persistent launch, ~25 us per step)."""
    return oracle if request.param == "oracle" else lh.cuda_library()


def test_coupled_equilibrium(anylib):
    oracle = anylib
    """test/SoilModel/coupled.jl:1-120: coupled water+heat, n = 20, zero-flux BCs, SSPRK33
    dt = 20 s for 32 days -> hydrostatic profile with interface at -0.3 and mean T = 284."""
    wl, ctx = _coupled_zero_flux_ctx(oracle)
    ep, sp = lh.EarthParameterSet(), w.coupled_soil_params()
    z = ctx.zc()
    T0 = 289.0 + 5.0 * z
    θl = 0.495
    ρc_s = lh.volumetric_heat_capacity(θl, 0.0, sp.ρc_ds, ep)
    ρe = np.array([lh.volumetric_internal_energy(0.0, ρc_s, T, ep) for T in T0])
    ctx.set_state(0, np.full(20, θl)); ctx.set_state(1, np.zeros(20)); ctx.set_state(2, ρe)
    W0, E0 = ctx.budgets()
    nsteps = int(round(60 * 60 * 24 * 32 / 20.0))
    assert nsteps == 138240
    _steps(ctx, 0.0, 20.0, nsteps)
    vlf = ctx.get_state(0)[0]
    ρeint = ctx.get_state(2)[0]
    ρc = np.array([lh.volumetric_heat_capacity(v, 0.0, sp.ρc_ds, ep) for v in vlf])
    temp = np.array([lh.temperature_from_ρe_int(e, 0.0, c, ep) for e, c in zip(ρeint, ρc)])

    def expected(z, z_interface):
        ν, S_s, α, n, m = 0.5, 1e-3, 2.6, 2.0, 0.5
        if z < z_interface:
            return -S_s * (z - z_interface) + ν
        return ν * (1 + (α * (z - z_interface)) ** n) ** (-m)

    exp = np.array([expected(zz, -0.3) for zz in z])
    assert math.sqrt(np.mean(vlf - exp) ** 2.0) < 1e-3                                 # :117
    assert math.sqrt(np.mean(temp - 284.0) ** 2.0) < 1e-3                              # :118
    # flux-form divergence + zero boundary fluxes: budgets are conserved
    W1, E1 = ctx.budgets()
    assert abs(W1 - W0) <= 1e-11 * abs(W0), (W1 - W0) / W0
    assert abs(E1 - E0) <= 1e-9 * abs(E0), (E1 - E0) / E0


def test_richards_equilibrium(anylib):
    oracle = anylib
    """test/SoilModel/richards_equation.jl:1-95: Richards only, n = 50, z in [-10, 0], zero flux,
    dt = 100 s for 36 days -> hydrostatic with interface at -0.56 (< 1e-4)."""
    vg = lh.vanGenuchten(n=2.0, α=2.6, Ksat=0.0443 / 3600 / 100, θr=0.0)
    p = w.make_params(lh.SoilParams(ν=0.495, S_s=1e-3), vg)
    F, N = abi.LH_BC_FLUX, abi.LH_BC_NONE
    wl = w.Workload(model=abi.LH_MODEL_RICHARDS, ncol=1, nlayer=50, zmin=-10.0, zmax=0.0, params=p,
                    top=(N, 0.0, F, 0.0), bottom=(N, 0.0, F, 0.0), dt=100.0)
    ctx = lh.SoilContext(oracle, wl.config())
    ctx.set_state(0, np.full(50, 0.494)); ctx.set_state(1, np.zeros(50))
    nsteps = int(round(60 * 60 * 24 * 36 / 100.0))
    assert nsteps == 31104
    _steps(ctx, 0.0, 100.0, nsteps)
    z = ctx.zc()
    ϑ = ctx.get_state(0)[0]

    def expected(z, zi):
        ν, S_s, α, n, m = 0.495, 1e-3, 2.6, 2.0, 0.5
        if z < zi:
            return -S_s * (z - zi) + ν
        return ν * (1 + (α * (z - zi)) ** n) ** (-m)

    exp = np.array([expected(zz, -0.56) for zz in z])
    assert math.sqrt(np.mean(ϑ - exp) ** 2.0) < 1e-4                                   # :94


def test_heat_analytic(anylib):
    oracle = anylib
    """test/SoilModel/heat_test_interface.jl:1-100: heat only, n = 60, dry soil with
    κ_dry / ρc_ds = 1, Dirichlet T: top 0, bottom 5 cos(2π t); dt = 1e-4 to t = 2; MSE < 1e-6 against
    the closed-form solution.  Pins the Dirichlet-as-flux half-cell distance and both sign
    conventions for energy."""
    sp = lh.SoilParams(ν=0.495, ν_ss_gravel=0.1, ν_ss_om=0.1, ν_ss_quartz=0.1, ρc_ds=0.43314518988433487,
                       κ_solid=8.0, κ_sat_unfrozen=0.57, κ_sat_frozen=2.29)
    p = w.make_params(sp, lh.vanGenuchten())
    D, N = abi.LH_BC_DIRICHLET, abi.LH_BC_NONE
    A, ω = 5.0, 2 * math.pi
    wl = w.Workload(model=abi.LH_MODEL_HEAT, ncol=1, nlayer=60, zmin=0.0, zmax=1.0, params=p,
                    top=(D, 0.0, N, 0.0), bottom=(D, A, N, 0.0), dt=1e-4)
    ctx = lh.SoilContext(oracle, wl.config())
    ep = lh.EarthParameterSet()
    ρc_s = lh.volumetric_heat_capacity(0.0, 0.0, sp.ρc_ds, ep)
    ρe0 = lh.volumetric_internal_energy(0.0, ρc_s, 0.0, ep)
    ctx.set_state(0, np.zeros(60)); ctx.set_state(1, np.zeros(60)); ctx.set_state(2, np.full(60, ρe0))
    dt, nsteps = 1e-4, 20000
    table = np.zeros((nsteps, 3, 4))
    t = 0.0
    for s in range(nsteps):
        for k, ts in enumerate((t, t + dt, t + dt / 2)):
            table[s, k, abi.LH_BCV_BOTTOM_ENERGY] = A * math.cos(ω * ts)
        t = t + dt
    ctx.step(0.0, dt, nsteps, table)
    tf = 2.0
    z = ctx.zc()
    num = np.exp(np.sqrt(ω / 2) * (1 + 1j) * (1 - z)) - np.exp(-np.sqrt(ω / 2) * (1 + 1j) * (1 - z))
    denom = np.exp(np.sqrt(ω / 2) * (1 + 1j)) - np.exp(-np.sqrt(ω / 2) * (1 + 1j))
    analytic = np.real(num * A * np.exp(1j * ω * tf) / denom)
    ρe = ctx.get_state(2)[0]
    Tfinal = np.array([lh.temperature_from_ρe_int(e, 0.0, ρc_s, ep) for e in ρe])
    MSE = np.mean((analytic - Tfinal) ** 2.0)
    assert MSE < 1e-6                                                                  # :99


def test_bonan_sand_infiltration_runs(anylib):
    oracle = anylib
    """test/SoilModel/richards_equation.jl:98-190: Bonan sand, Dirichlet top 0.267, FreeDrainage
    bottom, ϑ_l0 = 0.1, dt = 0.25 s.  The reference compares with an external CSV that cannot be
    downloaded here, so this pins the physically necessary properties instead: a monotone wetting
    front moving down from the top, bounded by the Dirichlet value, and water balance
    dW = -(F_top - F_bot) dt accumulated over the run."""
    wl = w.richards_workload(ncol=1, nlayer=150)
    ctx = lh.SoilContext(oracle, wl.config())
    ctx.set_state(0, np.full(150, 0.1)); ctx.set_state(1, np.zeros(150))
    W0 = ctx.budgets()[0]
    nsteps = 2880  # 0.2 h of the 0.8 h run
    ctx.step(0.0, 0.25, nsteps)
    ϑ = ctx.get_state(0)[0]
    assert np.all(ϑ <= 0.267 + 1e-9) and np.all(ϑ >= 0.1 - 1e-9)
    assert ϑ[-1] > 0.25 and abs(ϑ[0] - 0.1) < 1e-6
    assert np.all(np.diff(ϑ) >= -1e-9)          # wetter towards the top
    assert ctx.budgets()[0] > W0
