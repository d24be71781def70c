"""TEST INFRASTRUCTURE ONLY — second, independent restatement of the reference path.

Pure-Python/NumPy fp64, one column at a time, written to be READ next to the Julia source: each
block quotes the reference line it follows (paths relative to /root/reference).  It exists to
cross-check ``lho_soil.c`` (two restatements written separately must agree to the last bits) and
must never be imported by the product package.  Only tests/ may import it.

Scalar transcendental calls go through ``math.pow`` / ``math.exp`` / ``math.sqrt`` (glibc), the same
libm the C oracle links, so agreement is expected to be bit-exact up to FMA-free evaluation order.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

EPS = np.finfo(np.float64).eps

RICHARDS, HEAT, COUPLED = 0, 1, 2
BC_NONE, BC_FLUX, BC_DIRICHLET, BC_FREE_DRAINAGE = 0, 1, 2, 3


def jl_max(a, b):
    if math.isnan(a) or math.isnan(b):
        return math.nan
    return a if a > b else b


def jl_pow(x, y):
    try:
        return math.pow(x, y)
    except (ValueError, OverflowError, ZeroDivisionError):
        return math.nan if x < 0 else math.inf


# --- SoilWaterParameterizations.jl -------------------------------------------------------------
def volumetric_liquid_fraction(th, nu_eff):          # :181-188
    return th if th < nu_eff else nu_eff


def effective_saturation(porosity, th, theta_r):      # :213-217
    safe = jl_max(th, theta_r + EPS)
    return (safe - theta_r) / (porosity - theta_r)


def matric_potential(p, S):                           # :196-200
    n, alpha, m = p.vg_n, p.vg_alpha, p.vg_m
    return -jl_pow((jl_pow(S, -1.0 / m) - 1.0) * jl_pow(alpha, -n), 1.0 / n)


def pressure_head(p, th, nu_eff, S_s):                # :229-242
    S_l_eff = effective_saturation(nu_eff, th, p.theta_r)
    if S_l_eff <= 1.0:
        return matric_potential(p, S_l_eff)
    return (th - nu_eff) / S_s


def hydraulic_conductivity(p, S, visc_f, imp_f):      # :269-282
    if S < 1.0:
        K = math.sqrt(S) * jl_pow(1.0 - jl_pow(1.0 - jl_pow(S, 1.0 / p.vg_m), p.vg_m), 2.0)
    else:
        K = 1.0
    return K * p.Ksat * visc_f * imp_f


def viscosity_factor(p, T):                           # :104-126
    if p.viscosity_factor == 0:
        return 1.0
    return math.exp(p.visc_gamma * (T - p.visc_T_ref))


def impedance_factor(p, f_i):                         # :76-93
    if p.impedance_factor == 0:
        return 1.0
    return jl_pow(10.0, -p.imp_Omega * f_i)


# --- SoilHeatParameterizations.jl --------------------------------------------------------------
def volumetric_heat_capacity(p, tl, ti, rho_c_ds):    # :65-79
    return rho_c_ds + tl * (p.cp_l * p.rho_cloud_liq) + ti * (p.cp_i * p.rho_cloud_ice)


def temperature_from_rho_e_int(p, re, ti, rho_c_s):   # :42-53
    return p.T_0 + (re + ti * p.rho_cloud_ice * p.LH_f0) / rho_c_s


def saturated_thermal_conductivity(tl, ti, ku, kf):   # :114-128
    tw = tl + ti
    if tw < EPS:
        return 0.0
    return jl_pow(ku, tl / tw) * jl_pow(kf, ti / tw)


def relative_saturation(tl, ti, porosity):            # :139-142
    return (tl + ti) / porosity


def kersten_number(p, ti, S_r):                       # :152-174
    if ti < EPS:
        return jl_pow(S_r, (1.0 + p.nu_ss_om - p.a * p.nu_ss_quartz - p.nu_ss_gravel) / 2.0) * jl_pow(
            jl_pow(1.0 + math.exp(-p.b * S_r), -3.0) - jl_pow((1.0 - S_r) / 2.0, 3.0), 1.0 - p.nu_ss_om
        )
    return jl_pow(S_r, 1.0 + p.nu_ss_om)


def thermal_conductivity(k_dry_, K_e, k_sat):         # :185-188
    return K_e * k_sat + (1.0 - K_e) * k_dry_


def volumetric_internal_energy_liq(p, T):             # :198-207
    return (p.cp_l * p.rho_cloud_liq) * (T - p.T_0)


def k_dry(p):                                         # :268-294
    rho_b = (1.0 - p.nu) * p.rho_p
    numerator = (p.kappa_dry_parameter * p.kappa_solid - p.K_therm) * rho_b + p.K_therm * p.rho_p
    denom = p.rho_p - (1.0 - p.kappa_dry_parameter) * rho_b
    return numerator / denom


# --- boundary_conditions.jl:470-489 ---------------------------------------------------------------
def _kappa_at(p, kd, th, ti):                         # :429-436
    nu_eff = p.nu - ti
    tl = volumetric_liquid_fraction(th, nu_eff)
    S_r = relative_saturation(tl, ti, p.nu)
    return thermal_conductivity(kd, kersten_number(p, ti, S_r),
                                saturated_thermal_conductivity(tl, ti, p.kappa_sat_unfrozen, p.kappa_sat_frozen))


def _K_psi_at(p, th, ti, T):                          # :384-393
    nu_eff = p.nu - ti
    tl = volumetric_liquid_fraction(th, nu_eff)
    f_i = ti / (tl + ti) if (tl + ti) != 0.0 else math.nan
    imp = impedance_factor(p, f_i)
    visc = viscosity_factor(p, T)
    S = effective_saturation(p.nu, th, p.theta_r)
    return hydraulic_conductivity(p, S, visc, imp), pressure_head(p, th, nu_eff, p.S_s)


def boundary_fluxes(p, model, kd, bc, val_e, val_h, is_bottom, th_c, ti_c, T_c, dzb):
    e_kind, h_kind = bc
    th = [th_c, th_c]
    T = [T_c, T_c]
    ti = [ti_c, ti_c]
    if e_kind == BC_DIRICHLET and model in (HEAT, COUPLED):      # :241-248
        T[1] = val_e
    if h_kind == BC_DIRICHLET and model in (RICHARDS, COUPLED):  # :261-268
        th[1] = val_h
    fe = fw = 0.0
    if e_kind == BC_FLUX:                                        # :295-301
        fe = val_e
    elif e_kind == BC_DIRICHLET:                                 # :416-444
        kappa = [_kappa_at(p, kd, th[k], ti[k]) for k in range(2)]
        fe = -kappa[1] * (T[1] - T[0]) / dzb
        if is_bottom:
            fe *= -1
    if h_kind == BC_FLUX:
        fw = val_h
    elif h_kind == BC_FREE_DRAINAGE:                             # :328-356
        K, _ = _K_psi_at(p, th[0], ti[0], T[0])
        fw = -K
    elif h_kind == BC_DIRICHLET:                                 # :371-401
        Kp = [_K_psi_at(p, th[k], ti[k], T[k]) for k in range(2)]
        fw = -Kp[1][0] * (Kp[1][1] - Kp[0][1] + dzb) / dzb
        if is_bottom:
            fw *= -1
    return fe, fw


# --- right_hand_side.jl ----------------------------------------------------------------------------
def column_rhs(p, model, zmin, zmax, top, bottom, bcv, th, ti, re, T_aux):
    """One column.  ``top``/``bottom`` = (energy_kind, hydrology_kind); ``bcv`` = 4 boundary values
    [top energy, top hydrology, bottom energy, bottom hydrology].  Returns (dth, dti, dre, Fw, Fe)."""
    n = len(th)
    dz = (zmax - zmin) / n
    zf = [zmin + (zmax - zmin) * j / n for j in range(n + 1)]
    zc = [(zf[i] + zf[i + 1]) / 2.0 for i in range(n)]
    kd = k_dry(p)
    K = [0.0] * n; h = [0.0] * n; kap = [0.0] * n; T = [0.0] * n; eK = [0.0] * n
    for i in range(n):
        nu_eff = p.nu - ti[i]                                            # :156, :209, :291
        tl = volumetric_liquid_fraction(th[i], nu_eff)
        T[i] = T_aux[i]
        if model in (HEAT, COUPLED):
            rho_c_s = volumetric_heat_capacity(p, tl, ti[i], p.rho_c_ds)
            T[i] = temperature_from_rho_e_int(p, re[i], ti[i], rho_c_s)
            S_r = relative_saturation(tl, ti[i], p.nu)
            kap[i] = thermal_conductivity(kd, kersten_number(p, ti[i], S_r),
                                          saturated_thermal_conductivity(tl, ti[i], p.kappa_sat_unfrozen, p.kappa_sat_frozen))
        if model in (RICHARDS, COUPLED):
            f_i = ti[i] / (tl + ti[i]) if (tl + ti[i]) != 0.0 else math.nan
            visc = viscosity_factor(p, T[i])
            imp = impedance_factor(p, f_i)
            S = effective_saturation(p.nu, th[i], p.theta_r)
            K[i] = hydraulic_conductivity(p, S, visc, imp)
            h[i] = pressure_head(p, th[i], nu_eff, p.S_s) + zc[i]
        if model == COUPLED:
            eK[i] = volumetric_internal_energy_liq(p, T[i]) * K[i]
    fe_top, fw_top = boundary_fluxes(p, model, kd, top, bcv[0], bcv[1], False, th[n - 1], ti[n - 1], T[n - 1], dz / 2.0)
    fe_bot, fw_bot = boundary_fluxes(p, model, kd, bottom, bcv[2], bcv[3], True, th[0], ti[0], T[0], dz / 2.0)
    Fw = [0.0] * (n + 1); Fe = [0.0] * (n + 1)
    Fw[0], Fw[n], Fe[0], Fe[n] = fw_bot, fw_top, fe_bot, fe_top
    for j in range(1, n):
        gh = (h[j] - h[j - 1]) / dz
        gT = (T[j] - T[j - 1]) / dz
        if model in (RICHARDS, COUPLED):
            Fw[j] = -((K[j - 1] + K[j]) / 2.0) * gh                      # :181, :358
        if model == HEAT:
            Fe[j] = -((kap[j - 1] + kap[j]) / 2.0) * gT                  # :259
        elif model == COUPLED:
            Fe[j] = -((kap[j - 1] + kap[j]) / 2.0) * gT - ((eK[j - 1] + eK[j]) / 2.0) * gh   # :361-365
    dth = [-((Fw[i + 1] - Fw[i]) / dz) if model in (RICHARDS, COUPLED) else 0.0 for i in range(n)]
    dre = [-((Fe[i + 1] - Fe[i]) / dz) if model in (HEAT, COUPLED) else 0.0 for i in range(n)]
    return np.array(dth), np.zeros(n), np.array(dre), np.array(Fw), np.array(Fe)


def ssprk33_step(rhs, u0, dt):
    """OrdinaryDiffEq v5 SSPRK33 on a tuple of arrays; rhs(u, stage) -> tuple."""
    k = rhs(u0, 1)
    u1 = tuple(a + dt * b for a, b in zip(u0, k))
    k = rhs(u1, 2)
    u2 = tuple((3 * a + b + dt * c) / 4 for a, b, c in zip(u0, u1, k))
    k = rhs(u2, 3)
    return tuple((a + 2 * b + 2 * dt * c) / 3 for a, b, c in zip(u0, u2, k))
