// lh_closures.cuh — device-side soil closures (fp64), fused into every kernel of this library.
//
// What the reference evaluates as ~16 separate broadcasts with ~12 `pow`/`exp` calls per cell
// (src/SoilModel/right_hand_side.jl:291-314; SoilWaterParameterizations.jl:196-282;
// SoilHeatParameterizations.jl:42-207) is evaluated here once per cell in registers.  The
// algebra is re-associated so that the van Genuchten retention curve and conductivity share
// ONE log(S) and ONE log(1 - S^(1/m)):
//
//     u = log(S)/m            y = S^(1/m) = e^u          w = 1 - y = -expm1(u)
//     a = log(w)
//     psi = -((S^(-1/m) - 1) alpha^(-n))^(1/n) = -(1/alpha) exp((a - u)/n)
//     K   = Ksat sqrt(S) (1 - (1 - y)^m)^2     = Ksat sqrt(S) expm1(m a)^2
//
// which is also better conditioned than the literal form (1 - y and 1 - (1-y)^m are formed by
// expm1, not by subtraction).  Results agree with the literal fp64 evaluation to a few ulp;
// the parity gate is 1e-12 (cancellation-aware norm) per tendency evaluation.
//
// The fp64 pipe (64 DFMA/clk/SM on B200), not HBM, is the binding unit for this path, so the
// elementary functions in lh_math.cuh are written for minimum DFMA count.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lh_math.cuh"

#define LH_EPS 2.220446049250313e-16

// Host-derived constants (all fp64, computed once in lh_soil_create).  Passed to kernels by
// value (__grid_constant__), so every use is a constant-bank operand, not a register.
struct LhDevParams {
    // geometry
    double dz, inv_dz, half_dz, inv_half_dz;
    // water
    double nu, theta_r, theta_r_eps;   // theta_r + eps(Float64)
    double inv_nu_thr;                 // 1 / (nu - theta_r)
    double S_s_inv;                    // 1 / S_s
    double vg_m, vg_inv_m, vg_inv_n;
    double neg_inv_alpha;              // -1 / alpha
    double Ksat;
    double visc_gamma, visc_T_ref;
    double imp_c;                      // -Omega * ln(10)
    // heat
    double rho_c_ds, rhocp_l, rhocp_i, rhoi_LH, T_0;
    double inv_nu;
    double kersten_p1;                 // (1 + nu_om - a nu_quartz - nu_gravel) / 2
    double kersten_p2;                 // 1 - nu_om
    double kersten_p3;                 // 1 + nu_om
    double neg_b;
    double k_unfrozen, k_frozen, ln_k_unfrozen, ln_k_frozen;
    double kappa_dry;
    int32_t visc_on, imp_on;
    int32_t om_zero;                   // nu_ss_om == 0: outer Kersten exponents are exactly 1
    int32_t pad_;
};

struct LhCell {
    double K;      // hydraulic conductivity
    double psi;    // pressure head
    double kappa;  // thermal conductivity
    double T;      // temperature
};

// ---------------------------------------------------------------------------------------------
// Water: K and psi of one cell.  Reference: right_hand_side.jl:156-166 / :308-313.
//   th = ϑ_l, ti = θ_i, T = temperature (only read when the viscosity factor is on).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lh_water_closures(const LhDevParams& p, double th, double ti, double T,
                                                  double& K_out, double& psi_out, double& logS_K,
                                                  double& S_K_out)
{
    const double nu_eff = p.nu - ti;
    // effective_saturation (SoilWaterParameterizations.jl:213-217); NaN-propagating max
    const double safe = (th > p.theta_r_eps || th != th) ? th : p.theta_r_eps;
    const double num = safe - p.theta_r;
    const double S_K = num * p.inv_nu_thr;                                   // porosity = nu (:163/:311)
    const bool no_ice = (ti == 0.0);
    const double S_eff = no_ice ? S_K : num / (nu_eff - p.theta_r);          // porosity = nu_eff (:235)

    // ---- pressure head (:229-242) and the shared logs
    double L_eff = lh_log(S_eff);
    double u = L_eff * p.vg_inv_m;
    double em1 = lh_expm1(u);            // y - 1
    double w = -em1;                     // 1 - S^(1/m)
    double a = lh_log(w);
    double psi;
    if (S_eff <= 1.0) {
        psi = p.neg_inv_alpha * lh_exp((a - u) * p.vg_inv_n);
    } else {
        psi = (th - nu_eff) * p.S_s_inv;
    }

    // ---- hydraulic conductivity (:269-282)
    double L_K = L_eff, a_K = a;
    if (!no_ice) {                       // S differs from S_eff only when ice is present
        L_K = lh_log(S_K);
        a_K = lh_log(-lh_expm1(L_K * p.vg_inv_m));
    }
    double Kr;
    if (S_K < 1.0) {
        const double q = lh_expm1(p.vg_m * a_K);     // (1 - y)^m - 1
        Kr = sqrt(S_K) * (q * q);
    } else {
        Kr = 1.0;
    }
    double K = Kr * p.Ksat;
    if (p.visc_on) K *= lh_exp(p.visc_gamma * (T - p.visc_T_ref));          // :117-126
    if (p.imp_on) {                                                          // :89-93, f_i :159/:308
        const double tl = (th < nu_eff) ? th : nu_eff;
        const double f_i = ti / (tl + ti);
        K *= lh_exp(p.imp_c * f_i);
    }
    K_out = K;
    psi_out = psi;
    logS_K = L_K;
    S_K_out = S_K;
}

// ---------------------------------------------------------------------------------------------
// Heat: thermal conductivity from (θ_l, θ_i).  Reference: right_hand_side.jl:296-305,
// SoilHeatParameterizations.jl:114-188.  `S_hint`/`logS_hint`: a saturation whose log is
// already known (reused when S_r == S_hint bit for bit, which holds for θr = 0, θ_i = 0).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double lh_thermal_conductivity(const LhDevParams& p, double tl, double ti,
                                                          double S_hint, double logS_hint)
{
    const double tw = tl + ti;
    const double S_r = tw * p.inv_nu;                                        // relative_saturation :139-142
    const double Lr = (S_r == S_hint) ? logS_hint : lh_log(S_r);
    double K_e;
    if (ti < LH_EPS) {                                                       // kersten_number :163-169
        const double e = lh_exp(p.neg_b * S_r);
        const double g = 1.0 + e;
        const double E3 = 1.0 / (g * g * g);
        const double c = (1.0 - S_r) * 0.5;
        double base = E3 - c * c * c;
        if (!p.om_zero) base = lh_exp(p.kersten_p2 * lh_log(base));
        K_e = lh_exp(p.kersten_p1 * Lr) * base;
    } else {                                                                 // :171
        K_e = p.om_zero ? S_r : lh_exp(p.kersten_p3 * Lr);
    }
    double k_sat;                                                            // :114-128
    if (tw < LH_EPS) k_sat = 0.0;
    else if (ti == 0.0) k_sat = p.k_unfrozen;                                // x^1 * y^0, exact
    else k_sat = lh_exp((tl * p.ln_k_unfrozen + ti * p.ln_k_frozen) / tw);
    return K_e * k_sat + (1.0 - K_e) * p.kappa_dry;                          // thermal_conductivity :185-188
}

// Temperature from ρe_int (SoilHeatParameterizations.jl:42-79).
__device__ __forceinline__ double lh_temperature(const LhDevParams& p, double tl, double ti, double re)
{
    const double rho_c_s = p.rho_c_ds + tl * p.rhocp_l + ti * p.rhocp_i;     // :65-79
    return p.T_0 + (re + ti * p.rhoi_LH) / rho_c_s;                          // :42-53
}

// All closures of one cell for model MODEL (0 Richards, 1 heat, 2 coupled).
//   Richards: T_or_re = prescribed T.   heat/coupled: T_or_re = ρe_int.
template <int MODEL>
__device__ __forceinline__ LhCell lh_cell_closures(const LhDevParams& p, double th, double ti, double T_or_re)
{
    LhCell c;
    c.K = 0.0; c.psi = 0.0; c.kappa = 0.0; c.T = T_or_re;
    const double nu_eff = p.nu - ti;
    const double tl = (th < nu_eff) ? th : nu_eff;                           // volumetric_liquid_fraction :181-188
    if (MODEL != 0) c.T = lh_temperature(p, tl, ti, T_or_re);
    double logS = 0.0, S_K = -1.0;
    if (MODEL != 1) lh_water_closures(p, th, ti, c.T, c.K, c.psi, logS, S_K);
    if (MODEL != 0) c.kappa = lh_thermal_conductivity(p, tl, ti, S_K, logS);
    return c;
}

// κ at a boundary "face state" (boundary_conditions.jl:429-436).
__device__ __forceinline__ double lh_face_kappa(const LhDevParams& p, double th, double ti)
{
    const double nu_eff = p.nu - ti;
    const double tl = (th < nu_eff) ? th : nu_eff;
    return lh_thermal_conductivity(p, tl, ti, -1.0, 0.0);
}
