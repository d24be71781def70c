"""GPU parity tests proper: the CUDA library (through the C ABI) against the CPU oracle on the
same seeded inputs.

Gates (BASELINE.json north_star / SURVEY §8d):
  * per tendency evaluation:  max_i |a_i - r_i| <= 1e-12 * max(‖r‖∞ over the column, max_j|F_j|/Δz)
  * state after N SSPRK33 steps: max |u - u_ref| <= 1e-10 * max |u_ref| per field
"""
import numpy as np
import pytest

import workloads as w

pytestmark = pytest.mark.gpu

lh = w.lh
abi = w.abi
D, F, FD, N = abi.LH_BC_DIRICHLET, abi.LH_BC_FLUX, abi.LH_BC_FREE_DRAINAGE, abi.LH_BC_NONE

TOL_TENDENCY = 1e-12
TOL_STATE = 1e-10


def _pair(cuda, oracle, wl, flags=0):
    g = lh.SoilContext(cuda, wl.config(flags=flags))
    o = lh.SoilContext(oracle, wl.config())
    wl.upload(g)
    wl.upload(o)
    return g, o


def _prognostic(model):
    return {abi.LH_MODEL_RICHARDS: (0,), abi.LH_MODEL_HEAT: (2,), abi.LH_MODEL_COUPLED: (0, 2)}[model]


def assert_tendency_parity(g, o, model, tol=TOL_TENDENCY):
    g.rhs(0.0)
    o.rhs(0.0)
    worst = 0.0
    for f in _prognostic(model):
        a, r = g.get_tendency(f), o.get_tendency(f)
        assert np.all(np.isfinite(r)), "oracle produced non-finite tendencies: bad test input"
        scale = w.tendency_scale(o, f)
        err = np.max(np.abs(a - r) / scale[:, None])
        worst = max(worst, err)
        assert err <= tol, f"field {f}: scaled tendency error {err:.3e} > {tol:g}"
    assert np.all(g.get_tendency(1) == 0.0)      # dθ_i ≡ 0 (right_hand_side.jl:182,359)
    return worst


def assert_state_parity(g, o, model, dt, nsteps, table=None, tol=TOL_STATE):
    g.step(0.0, dt, nsteps, table)
    o.step(0.0, dt, nsteps, table)
    for f in _prognostic(model):
        a, r = g.get_state(f), o.get_state(f)
        assert np.all(np.isfinite(r))
        err = np.max(np.abs(a - r)) / np.max(np.abs(r))
        assert err <= tol, f"field {f}: state error {err:.3e} > {tol:g} after {nsteps} steps"
    # dθ_i ≡ 0.  The device never rewrites θ_i (bitwise constant); the reference's generic axpy
    # (u0 + 2 u2 + 2 dt 0)/3 may move a non-zero θ_i by an ulp per step, which the oracle reproduces.
    ti_g, ti_o = g.get_state(1), o.get_state(1)
    assert np.max(np.abs(ti_g - ti_o)) <= 4 * nsteps * np.finfo(np.float64).eps * max(np.max(np.abs(ti_o)), 1e-300)


# ---- the BASELINE configs at oracle-sized shapes ---------------------------------------------------
@pytest.mark.parametrize("ncol,nlayer", [(1, 20), (1, 64), (96, 64), (33, 16), (257, 100), (64, 256)])
def test_coupled_tendency_and_state(cuda, oracle, ncol, nlayer):
    """C2/C4: coupled water+heat, Dirichlet top (ϑ_l, T), FreeDrainage/zero-flux bottom."""
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=100 + ncol + nlayer)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)


@pytest.mark.parametrize("ncol,nlayer", [(1, 150), (1, 50), (1024, 100), (40, 37)])
def test_richards_tendency_and_state(cuda, oracle, ncol, nlayer):
    """C1/C3: Richards only, Bonan sand, Dirichlet top 0.267, FreeDrainage bottom."""
    wl = w.richards_workload(ncol=ncol, nlayer=nlayer, seed=200 + ncol + nlayer)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)


@pytest.mark.parametrize("ncol,nlayer", [(1, 60), (128, 60), (70, 24)])
def test_heat_tendency_and_state(cuda, oracle, ncol, nlayer):
    wl = w.heat_workload(ncol=ncol, nlayer=nlayer, seed=300 + ncol + nlayer)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)


@pytest.mark.parametrize("ice", [False, True])
def test_coupled_general_van_genuchten_path(cuda, oracle, ice):
    """The coupled workload has n = 2, which the library evaluates with the square-root
    specialisation (LH_FLAG_VG2 kernels); LH_FLAG_GENERAL_VG forces the general-n log/exp closures on
    the same inputs, and a non-integer n exercises them without the flag."""
    wl = w.coupled_workload(ncol=96, nlayer=64, seed=411, ice=ice)
    g, o = _pair(cuda, oracle, wl, flags=abi.LH_FLAG_GENERAL_VG)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)
    wl = w.coupled_workload(ncol=96, nlayer=64, seed=412, ice=ice)
    wl.params.vg_n = 1.56
    wl.params.vg_m = 1.0 - 1.0 / 1.56
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)


# ---- every BC kind at both faces ---------------------------------------------------------------------
BC_FACES_COUPLED = [
    (F, 0.0, F, 0.0), (F, 3.0, F, -1e-7), (D, 284.0, D, 0.35), (D, 292.0, F, 2e-8), (F, -2.0, D, 0.45),
    (F, 0.0, FD, 0.0), (D, 280.0, FD, 0.0),
]


@pytest.mark.parametrize("top", BC_FACES_COUPLED)
@pytest.mark.parametrize("bottom", BC_FACES_COUPLED)
def test_coupled_all_bc_kinds(cuda, oracle, top, bottom):
    wl = w.coupled_workload(ncol=40, nlayer=24, seed=17, top=top, bottom=bottom)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


@pytest.mark.parametrize("top", [(N, 0.0, F, 1e-7), (N, 0.0, D, 0.2), (N, 0.0, FD, 0.0), (F, 1.0, D, 0.26)])
@pytest.mark.parametrize("bottom", [(N, 0.0, F, 0.0), (N, 0.0, D, 0.15), (N, 0.0, FD, 0.0)])
def test_richards_all_bc_kinds(cuda, oracle, top, bottom):
    wl = w.richards_workload(ncol=40, nlayer=30, seed=18, top=top, bottom=bottom)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


@pytest.mark.parametrize("top", [(F, 5.0, N, 0.0), (D, 300.0, N, 0.0), (D, 275.0, F, 0.0)])
@pytest.mark.parametrize("bottom", [(F, -5.0, N, 0.0), (D, 281.0, N, 0.0)])
def test_heat_all_bc_kinds(cuda, oracle, top, bottom):
    wl = w.heat_workload(ncol=40, nlayer=30, seed=19, top=top, bottom=bottom)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


# ---- branch coverage of the closures ---------------------------------------------------------------
def test_ice_and_factors(cuda, oracle):
    """θ_i > 0 (Kersten frozen branch, κ_sat mix, S != S_eff), TemperatureDependentViscosity and
    IceImpedance on (SoilWaterParameterizations.jl:76-126)."""
    wl = w.coupled_workload(ncol=64, nlayer=32, seed=23, ice=True,
                            viscosity=lh.TemperatureDependentViscosity(), impedance=lh.IceImpedance())
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)
    wl = w.richards_workload(ncol=64, nlayer=32, seed=24, ice=True,
                             viscosity=lh.TemperatureDependentViscosity(), impedance=lh.IceImpedance())
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 10)
    wl = w.heat_workload(ncol=64, nlayer=32, seed=25, ice=True)
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


def test_saturated_cells(cuda, oracle):
    """ϑ_l > ν_eff: the positive pressure-head branch (SoilWaterParameterizations.jl:236-240) and
    the S >= 1 clamp of K (:276-280), mixed with unsaturated cells in the same columns."""
    wl = w.coupled_workload(ncol=48, nlayer=32, seed=31)
    th = wl.fields[0]
    rng = np.random.default_rng(5)
    mask = rng.uniform(size=th.shape) < 0.3
    th[mask] = wl.params.nu + rng.uniform(0.0, 5e-4, size=mask.sum())      # ψ = (ϑ_l - ν)/S_s up to 0.5 m
    th[0, :4] = wl.params.nu                                              # exactly S = 1
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


@pytest.mark.parametrize("nu,theta_r", [(0.4547846749285817, 0.002831967114546297), (0.44265431030687197, 0.09972099357892111),
                                         (0.287, 0.075)])
def test_exactly_saturated_cells(cuda, oracle, nu, theta_r):
    """ϑ_l == ν bit for bit (a saturated initial condition, a ponded Dirichlet top): the reference's quotient is exactly 1
    there, K = K_sat and ψ = 0.  The first two (ν, θr) pairs are ones for which (ν - θr) * (1 / (ν - θr)) rounds to
    1 - 2^-53, i.e. a device S computed with the reciprocal would fall on the wrong side of the S < 1 branches."""
    for make in (w.richards_workload, w.coupled_workload):
        wl = make(ncol=40, nlayer=24, seed=33, top=(w.F if make is w.coupled_workload else w.N, 0.0, w.D, nu),
                  bottom=(w.F if make is w.coupled_workload else w.N, 0.0, w.FD, 0.0))
        p = wl.params
        S = (wl.fields[0] - p.theta_r) / (p.nu - p.theta_r)
        p.nu, p.theta_r = nu, theta_r
        if make is w.coupled_workload:
            p.vg_n, p.vg_m = 1.9, 1.0 - 1.0 / 1.9          # general-n closures with theta_r != 0
        th = theta_r + S * (nu - theta_r)
        th[:, 8:14] = nu                                     # a saturated band, and the boundary cells
        th[::2, 0] = nu
        th[1::2, -1] = nu
        wl.fields[0] = th
        if 2 in wl.fields:
            wl.fields[2] = w.rho_e_int_from_T(p, th, wl.fields[1], w.temperature_profiles(33, (0, 40), 24, wl.zmin, wl.zmax))
        g, o = _pair(cuda, oracle, wl)
        assert_tendency_parity(g, o, wl.model)
        for which in (abi.LH_DIAG_K, abi.LH_DIAG_PSI):
            a = g.diagnostic(which)
            assert np.all(a[:, 8:14] == (p.Ksat if which == abi.LH_DIAG_K else 0.0))


def test_dry_soil_heat(cuda, oracle):
    """θ_w < eps: κ_sat = 0, κ = κ_dry (SoilHeatParameterizations.jl:121-123) — the analytic heat
    test's regime (heat_test_interface.jl)."""
    wl = w.heat_workload(ncol=16, nlayer=60, seed=41)
    wl.fields[0][...] = 0.0
    wl.fields[1][...] = 0.0
    g, o = _pair(cuda, oracle, wl)
    assert_tendency_parity(g, o, wl.model)


def test_diagnostics_match_oracle(cuda, oracle):
    """K, ψ, κ, T closures evaluated on the device vs the literal reference formulas."""
    for ice in (False, True):
        wl = w.coupled_workload(ncol=64, nlayer=48, seed=51, ice=ice)
        g, o = _pair(cuda, oracle, wl)
        for which, tol in ((abi.LH_DIAG_K, 2e-13), (abi.LH_DIAG_PSI, 1e-13), (abi.LH_DIAG_KAPPA, 1e-13), (abi.LH_DIAG_T, 1e-14)):
            a, r = g.diagnostic(which), o.diagnostic(which)
            rel = np.max(np.abs(a - r) / np.abs(r))
            assert rel <= tol, f"diag {which} ice={ice}: rel err {rel:.3e}"


# ---- time-dependent Dirichlet values through the bc table ------------------------------------------
def test_bc_table(cuda, oracle):
    wl = w.heat_workload(ncol=32, nlayer=40, seed=61)
    g, o = _pair(cuda, oracle, wl)
    nsteps = 12
    table = np.zeros((nsteps, 3, 4))
    t = 0.0
    for s in range(nsteps):
        for k, ts in enumerate((t, t + wl.dt, t + wl.dt / 2)):
            table[s, k, abi.LH_BCV_TOP_ENERGY] = 290.0 + 3.0 * np.sin(ts / 200.0)
            table[s, k, abi.LH_BCV_BOTTOM_ENERGY] = 280.0 + 2.0 * np.cos(ts / 300.0)
        t += wl.dt
    assert_state_parity(g, o, wl.model, wl.dt, nsteps, table)


# ---- transfers / layouts -----------------------------------------------------------------------------
@pytest.mark.parametrize("ncol,nlayer", [(1, 7), (31, 5), (100, 64), (1000, 33)])
def test_state_roundtrip_layouts(cuda, ncol, nlayer):
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=71)
    g = lh.SoilContext(cuda, wl.config())
    rng = np.random.default_rng(1)
    a = rng.standard_normal((ncol, nlayer))
    g.set_state(0, a)                                   # reference layout, dense
    assert np.array_equal(g.get_state(0), a)
    big = rng.standard_normal((ncol, 3, nlayer + 2))    # strided view: field-interleaved host block
    view = big[:, 1, 1:-1]
    g.set_state(2, view)
    out = np.zeros_like(big)
    g.get_state(2, out[:, 1, 1:-1])
    assert np.array_equal(out[:, 1, 1:-1], view) and np.all(out[:, 0] == 0) and np.all(out[:, 2] == 0)
    soa = np.ascontiguousarray(a.T)                     # column-fastest host block
    g.set_state(1, soa.T)
    assert np.array_equal(g.get_state(1), a)
    back = np.empty_like(soa)
    g.get_state(1, back.T)
    assert np.array_equal(back, soa)


def test_budgets_match_oracle(cuda, oracle):
    wl = w.coupled_workload(ncol=777, nlayer=64, seed=81)
    g, o = _pair(cuda, oracle, wl)
    bg, bo = g.budgets(), o.budgets()
    assert np.all(np.abs(bg - bo) <= 1e-13 * np.abs(bo))
    bg2 = g.budgets()
    assert np.array_equal(bg, bg2)                      # fixed-tree reduction: bitwise reproducible


# ---- size-independent properties at the BASELINE's full size ---------------------------------------
def test_full_size_properties(cuda, oracle):
    """C4 shape (2^20 columns x 64 layers, coupled).  The oracle cannot run this in seconds, so:
    (1) columns are independent and deterministic: the first 64 columns evolve bit-identically
        inside the 1M batch and alone;  (2) those 64 columns match the oracle to 1e-10;
    (3) with zero-flux BCs the water and energy budgets are conserved to round-off."""
    ncol, nlayer, nsteps = 1 << 20, 64, 5
    wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=w.BASE_SEED + 3,
                            top=(F, 0.0, F, 0.0), bottom=(F, 0.0, F, 0.0))
    g = lh.SoilContext(cuda, wl.config())
    wl.upload(g)
    W0 = g.budgets()
    g.step(0.0, wl.dt, nsteps)
    W1 = g.budgets()
    assert abs(W1[0] - W0[0]) <= 1e-12 * abs(W0[0])
    assert abs(W1[1] - W0[1]) <= 1e-12 * abs(W0[1])
    sub = 64
    th_big = g.get_state(0)[:sub].copy()
    re_big = g.get_state(2)[:sub].copy()
    g.close()
    small = w.Workload(model=wl.model, ncol=sub, nlayer=nlayer, zmin=wl.zmin, zmax=wl.zmax, params=wl.params,
                       top=wl.top, bottom=wl.bottom, dt=wl.dt,
                       fields={k: v[:sub].copy() for k, v in wl.fields.items()})
    gs, os_ = _pair(cuda, oracle, small)
    gs.step(0.0, wl.dt, nsteps)
    os_.step(0.0, wl.dt, nsteps)
    assert np.array_equal(gs.get_state(0), th_big) and np.array_equal(gs.get_state(2), re_big)
    for f, big in ((0, th_big), (2, re_big)):
        r = os_.get_state(f)
        assert np.max(np.abs(big - r)) <= TOL_STATE * np.max(np.abs(r))


def test_vertical_sweep_shapes(cuda, oracle):
    """C5: layers 16..1024 at a fixed (oracle-sized) cell count: every chunking of the vertical."""
    for nlayer in (16, 32, 64, 128, 256, 512, 1024):
        ncol = max(1, 16384 // nlayer)
        wl = w.coupled_workload(ncol=ncol, nlayer=nlayer, seed=900 + nlayer, zlim=(-2.0 * nlayer / 64, 0.0))
        g, o = _pair(cuda, oracle, wl)
        assert_tendency_parity(g, o, wl.model)
        assert_state_parity(g, o, wl.model, wl.dt / 4, 4)


def test_check_finite_flag(cuda):
    wl = w.richards_workload(ncol=8, nlayer=16, seed=3)
    wl.fields[0][3, 5] = np.nan
    g = lh.SoilContext(cuda, wl.config(flags=abi.LH_FLAG_CHECK_FINITE))
    wl.upload(g)
    with pytest.raises(lh.NonFiniteStateError):
        g.step(0.0, wl.dt, 1)


# ---- the device elementary functions themselves, evaluated on the GPU ------------------------------
def test_device_math_accuracy_on_gpu(cuda):
    """lh_log2 / lh_exp2 / lh_exp2m1 / lh_sqrt / lh_rsqrt / lh_rcp / lh_div (csrc/lh_math.cuh) evaluated ON
    THE B200 through lh_soil_eval_math, against mpmath at 40 digits.  This is what validates the MUFU
    seed accuracy the Newton/cubic refinements rely on (tests/test_device_math.py checks the same
    source on the host with emulated 20-bit seeds)."""
    import mpmath as mp

    mp.mp.dps = 40
    wl = w.coupled_workload(ncol=32, nlayer=4, seed=1)
    ctx = lh.SoilContext(cuda, wl.config())
    rng = np.random.default_rng(12)

    def ulps(y, exact):
        out = []
        for a, e in zip(y, exact):
            ulp = mp.mpf(2) ** (mp.floor(mp.log(abs(e), 2)) - 52)
            out.append(float(abs(mp.mpf(float(a)) - e) / ulp))
        return np.array(out)

    x = np.concatenate([rng.uniform(1e-16, 2.0, 3000), np.exp(rng.uniform(-40, 40, 1500)), 1 + rng.uniform(-1e-3, 1e-3, 500)])
    assert ulps(ctx.eval_math(0, x), [mp.log(mp.mpf(float(v)), 2) for v in x]).max() <= 2.5
    assert ulps(ctx.eval_math(3, x), [mp.sqrt(mp.mpf(float(v))) for v in x]).max() <= 1.5
    assert ulps(ctx.eval_math(4, x), [1 / mp.sqrt(mp.mpf(float(v))) for v in x]).max() <= 2.5
    assert ulps(ctx.eval_math(5, x), [1 / mp.mpf(float(v)) for v in x]).max() <= 2.0
    num = rng.uniform(-1e8, 1e8, x.size)
    assert ulps(ctx.eval_math(6, np.concatenate([num, x])), [mp.mpf(float(a)) / mp.mpf(float(b)) for a, b in zip(num, x)]).max() <= 1.5
    xe = np.concatenate([rng.uniform(-50, 50, 3000), rng.uniform(-1, 1, 1500), rng.uniform(-700, 700, 500)])
    assert ulps(ctx.eval_math(1, xe), [mp.mpf(2) ** mp.mpf(float(v)) for v in xe]).max() <= 2.0
    ym = ctx.eval_math(2, xe)
    exact = np.array([float(mp.expm1(mp.mpf(float(v)) * mp.log(2))) for v in xe])
    assert np.max(np.abs(ym - exact) / np.abs(exact)) <= 6e-15
    xs = rng.uniform(-1 / 128, 1 / 128, 1000)     # k == 0: the result is the polynomial itself
    ys = ctx.eval_math(2, xs)
    assert ulps(ys, [mp.expm1(mp.mpf(float(v)) * mp.log(2)) for v in xs]).max() <= 8.0
    assert ctx.eval_math(3, np.array([0.0]))[0] == 0.0 and np.isnan(ctx.eval_math(0, np.array([-1.0]))[0])


# ---- randomized sweep over parameters, shapes, BC kinds and flags -------------------------------------------
def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    model = [abi.LH_MODEL_COUPLED, abi.LH_MODEL_RICHARDS, abi.LH_MODEL_HEAT][seed % 3]
    nu = rng.uniform(0.25, 0.6)
    sp = lh.SoilParams(
        ν=nu, S_s=10.0 ** rng.uniform(-4, -2), ν_ss_gravel=rng.uniform(0, 0.2), ν_ss_om=rng.choice([0.0, rng.uniform(0, 0.2)]),
        ν_ss_quartz=rng.uniform(0.2, 0.9), ρc_ds=rng.uniform(0.8e6, 2.5e6), κ_solid=rng.uniform(2.0, 8.0),
        κ_sat_unfrozen=rng.uniform(0.8, 2.5), κ_sat_frozen=rng.uniform(2.0, 4.0), a=rng.uniform(0.1, 0.4), b=rng.uniform(10.0, 25.0))
    vg = lh.vanGenuchten(n=rng.choice([2.0, rng.uniform(1.2, 5.0)]), α=rng.uniform(0.5, 6.0), Ksat=10.0 ** rng.uniform(-8, -4),
                         θr=rng.choice([0.0, rng.uniform(0.0, 0.1)]))
    ice = bool(rng.integers(0, 2)) and seed % 4 == 0
    visc = lh.TemperatureDependentViscosity() if rng.integers(0, 3) == 0 else None
    imp = lh.IceImpedance() if (ice and rng.integers(0, 2)) else None
    p = w.make_params(sp, vg, viscosity=visc, impedance=imp)
    ncol, nlayer = int(rng.integers(1, 200)), int(rng.integers(1, 150))
    zmax = 0.0
    zmin = -float(rng.uniform(0.02, 0.2)) * nlayer
    S = w.saturation_profiles(2000 + seed, (0, ncol), nlayer, zmin, zmax, hi=0.97)
    ti = w.ice_profiles(2000 + seed, (0, ncol), nlayer, 0.04) if ice else np.zeros((ncol, nlayer))
    th = p.theta_r + S * ((p.nu - ti) - p.theta_r)
    fields = {0: th, 1: ti}
    if model != abi.LH_MODEL_RICHARDS:
        fields[2] = w.rho_e_int_from_T(p, th, ti, w.temperature_profiles(2000 + seed, (0, ncol), nlayer, zmin, zmax))
    th_bc = p.theta_r + rng.uniform(0.3, 0.95) * (p.nu - p.theta_r)

    def face(bottom):
        if model == abi.LH_MODEL_HEAT:
            e = [(D, rng.uniform(275, 295)), (F, rng.uniform(-5, 5))][rng.integers(0, 2)]
            return (e[0], e[1], N, 0.0)
        hk = [(D, th_bc), (F, -p.Ksat * rng.uniform(0, 0.5))] + ([(FD, 0.0)] if bottom else [])
        h = hk[rng.integers(0, len(hk))]
        if model == abi.LH_MODEL_RICHARDS:
            return (N, 0.0, h[0], h[1])
        e = [(D, rng.uniform(275, 295)), (F, rng.uniform(-5, 5))][rng.integers(0, 2)]
        return (e[0], e[1], h[0], h[1])

    aux_T = 288.0 + 6.0 * np.sin(np.linspace(0, 2, nlayer)) if (model == abi.LH_MODEL_RICHARDS and visc is not None) else None
    dz = (zmax - zmin) / nlayer
    dt = 0.02 * dz * dz / max(p.Ksat * 50.0, 1e-6)              # well inside the explicit stability limit
    wl = w.Workload(model=model, ncol=ncol, nlayer=nlayer, zmin=zmin, zmax=zmax, params=p, top=face(False), bottom=face(True),
                    dt=min(dt, 50.0), fields=fields, aux_T=aux_T, name=f"fuzz{seed}")
    flags = [0, abi.LH_FLAG_STAGE_LAUNCHES, abi.LH_FLAG_PERSISTENT, abi.LH_FLAG_GENERAL_VG][seed % 4]
    return wl, flags


@pytest.mark.parametrize("seed", range(36))
def test_randomized_parity(cuda, oracle, seed):
    """Random soil / van Genuchten / Balland-Arp parameters, column counts 1..199, 1..149 layers, every BC kind, ice and
    conductivity factors, all launch strategies: tendency 1e-12 (scaled), state 1e-10 after 4 steps."""
    wl, flags = _random_case(seed)
    g, o = _pair(cuda, oracle, wl, flags=flags)
    assert_tendency_parity(g, o, wl.model)
    assert_state_parity(g, o, wl.model, wl.dt, 4)


@pytest.mark.parametrize("kind,ncol,nlayer", [("coupled", 777, 64), ("coupled", 33, 300), ("richards", 1000, 100), ("coupled", 5, 1),
                                              ("heat", 100, 37)])
@pytest.mark.parametrize("launch", ["stage", "persistent"])
def test_fused_budgets_equal_full_pass(cuda, oracle, kind, ncol, nlayer, launch):
    """After a step the budgets come from the per-block sums the last-stage launches leave behind (fused epilogue); after
    an upload, from one pass over the state.  Both must agree (round-off of two summation orders), leave the padding
    columns out, and match the oracle."""
    make = {"coupled": w.coupled_workload, "richards": w.richards_workload, "heat": w.heat_workload}[kind]
    wl = make(ncol=ncol, nlayer=nlayer, seed=91) if kind == "heat" else make(ncol=ncol, nlayer=nlayer, seed=91, zlim=(-0.03 * nlayer, 0.0))
    flags = abi.LH_FLAG_STAGE_LAUNCHES if launch == "stage" else abi.LH_FLAG_PERSISTENT
    g, o = _pair(cuda, oracle, wl, flags=flags)
    for ctx in (g, o):
        ctx.step(0.0, wl.dt, 3)
    fused = g.budgets()                                  # right after the step
    state = {f: g.get_state(f) for f in wl.fields}
    for f, a in state.items():
        g.set_state(f, a)                                # an upload invalidates the fused sums
    full = g.budgets()
    ref = o.budgets()
    scale = np.maximum(np.abs(ref), 1e-300)
    assert np.all(np.abs(fused - full) <= 1e-14 * scale), (fused, full)
    assert np.all(np.abs(fused - ref) <= 1e-12 * scale), (fused, ref)
    g.step_with(_named(cuda, abi.LH_METHOD_SSPRK22), 0.0, wl.dt, 1)     # a generic stepper falls back to the full pass
    o.step_with(_named(oracle, abi.LH_METHOD_SSPRK22), 0.0, wl.dt, 1)
    assert np.all(np.abs(g.budgets() - o.budgets()) <= 1e-12 * np.maximum(np.abs(o.budgets()), 1e-300))


def _named(lib, method):
    t = abi.lh_soil_stepper()
    assert lib.soil_stepper_named(method, t) == abi.LH_OK
    return t
