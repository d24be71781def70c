"""How far are the closures from the EXACT values of the reference's formulas?

The formulas of SoilWaterParameterizations.jl:196-282 and SoilHeatParameterizations.jl:42-188 are evaluated with mpmath
at 50 digits on the same fp64 inputs; the oracle (libm, literal expression order) and the CUDA closures (shared
log2/exp2, Mualem identity, table-driven math) are both measured against that truth.  This bounds the oracle's own
error and shows that the re-associated device forms are not less accurate than the literal ones."""
import mpmath as mp
import numpy as np
import pytest

import workloads as w

lh, abi = w.lh, w.abi
mp.mp.dps = 50


def truth(p, th, re):
    """(K, psi, kappa, T) of one ice-free cell, exact arithmetic on the fp64 inputs."""
    f = lambda x: mp.mpf(float(x))
    nu, thr, n, m, alpha, Ksat, S_s = f(p.nu), f(p.theta_r), f(p.vg_n), f(p.vg_m), f(p.vg_alpha), f(p.Ksat), f(p.S_s)
    th, re = f(th), f(re)
    eps = mp.mpf(2) ** -52
    S = (max(th, thr + eps) - thr) / (nu - thr)
    if S <= 1:
        psi = -(((S ** (-1 / m) - 1) * alpha ** (-n)) ** (1 / n))
    else:
        psi = (th - nu) / S_s
    K = (mp.sqrt(S) * (1 - (1 - S ** (1 / m)) ** m) ** 2 if S < 1 else mp.mpf(1)) * Ksat
    tl = th if th < nu else nu
    rho_c_s = f(p.rho_c_ds) + tl * f(p.cp_l) * f(p.rho_cloud_liq)
    T = f(p.T_0) + re / rho_c_s
    S_r = tl / nu
    a_, b_ = f(p.a), f(p.b)
    om, q, g = f(p.nu_ss_om), f(p.nu_ss_quartz), f(p.nu_ss_gravel)
    K_e = S_r ** ((1 + om - a_ * q - g) / 2) * ((1 + mp.e ** (-b_ * S_r)) ** (-3) - ((1 - S_r) / 2) ** 3) ** (1 - om)
    rho_b = (1 - nu) * f(p.rho_p)
    kdp, ks, ka = f(p.kappa_dry_parameter), f(p.kappa_solid), f(p.K_therm)
    k_dry = ((kdp * ks - ka) * rho_b + ka * f(p.rho_p)) / (f(p.rho_p) - (1 - kdp) * rho_b)
    kappa = K_e * f(p.kappa_sat_unfrozen) + (1 - K_e) * k_dry
    return K, psi, kappa, T


def rel_errors(ctx, wl, ncheck=24):
    diag = {k: ctx.diagnostic(k) for k in (abi.LH_DIAG_K, abi.LH_DIAG_PSI, abi.LH_DIAG_KAPPA, abi.LH_DIAG_T)}
    worst = np.zeros(4)
    rng = np.random.default_rng(0)
    for c in rng.choice(wl.ncol, size=min(ncheck, wl.ncol), replace=False):
        for i in range(wl.nlayer):
            t = truth(wl.params, wl.fields[0][c, i], wl.fields[2][c, i])
            for k in range(4):
                worst[k] = max(worst[k], float(abs(mp.mpf(float(diag[k][c, i])) - t[k]) / abs(t[k])))
    return worst


def _workload(general):
    wl = w.coupled_workload(ncol=32, nlayer=20, seed=77, sat_hi=0.95)
    if general:
        p = wl.params
        S = wl.fields[0] / p.nu
        p.vg_n, p.vg_m, p.theta_r = 1.7, 1.0 - 1.0 / 1.7, 0.03
        wl.fields[0] = p.theta_r + S * (p.nu - p.theta_r)
    return wl


# Conditioning: 1 - S^(1/m) loses up to 1/(1 - S^(1/m)) in the literal form (S <= 0.95 here: <= ~25x for K and psi).
LIMITS = np.array([2e-14, 2e-14, 2e-15, 2e-16])        # K, psi, kappa, T


@pytest.mark.parametrize("general", [False, True], ids=["n2", "general"])
def test_oracle_against_exact_formulas(oracle, general):
    wl = _workload(general)
    ctx = lh.SoilContext(oracle, wl.config())
    wl.upload(ctx)
    worst = rel_errors(ctx, wl)
    assert np.all(worst <= LIMITS), worst


@pytest.mark.gpu
@pytest.mark.parametrize("general", [False, True], ids=["n2", "general"])
def test_cuda_against_exact_formulas(cuda, oracle, general):
    wl = _workload(general)
    g, o = lh.SoilContext(cuda, wl.config()), lh.SoilContext(oracle, wl.config())
    wl.upload(g)
    wl.upload(o)
    wg, wo = rel_errors(g, wl), rel_errors(o, wl)
    print("max relative error vs exact (K, psi, kappa, T): cuda", wg, "oracle", wo)
    assert np.all(wg <= LIMITS), wg
