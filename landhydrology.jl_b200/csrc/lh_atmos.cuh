// lh_atmos.cuh — PrescribedAtmosForcing on the device: compute_turbulent_surface_fluxes of the reference
// (src/SoilModel/boundary_conditions.jl:555-620) for the top cell of every column.
//
// PARITY UNPINNED for the two third-party pieces (SurfaceFluxes v0.1 `surface_conditions`, Thermodynamics v0.5
// `q_vap_saturation_generic`): their sources are not under the reference tree, so they follow the published
// formulations (include/lh_soil.h); everything the reference itself writes down is literal.
//
// This is a per-COLUMN computation (one Monin-Obukhov solve per column and stage, ~1 % of a 64-layer column's work), so it
// runs in its own small kernel right before the stage launch and hands the stage kernel two per-column flux arrays — the
// stage kernel then treats the top face as a spatially varying VerticalFlux (lh_soil_set_column_fluxes uses the same
// path).  Keeping the iterative solve (log / exp / atan of CUDA's math library) out of the fused stage kernel keeps that
// kernel's register budget (<= 102) and instruction cache untouched.
#pragma once

#include "lh_closures.cuh"

struct LhAtmos {
    double u_atm, theta_atm, z_atm, theta_scale, rho_a_sfc, q_atm;
    double R_v, R_d, grav, cp_d, cp_v, LH_v0, press_triple, T_triple, von_karman;
    double Pr_0, a_m, a_h;
    double cp_l, T_0, rho_l, z_0m, z_0s;      // from lh_soil_params
};

// Businger-Dyer integrated universal functions (Businger et al. 1971; Dyer 1974; Paulson 1970)
__device__ __forceinline__ double lh_most_psi_m(const LhAtmos& a, double zeta)
{
    if (zeta >= 0.0) return -a.a_m * zeta;
    const double X = sqrt(sqrt(1.0 - 15.0 * zeta));
    return 2.0 * log((1.0 + X) / 2.0) + log((1.0 + X * X) / 2.0) - 2.0 * atan(X) + 1.57079632679489661923;
}

__device__ __forceinline__ double lh_most_psi_h(const LhAtmos& a, double zeta)
{
    if (zeta >= 0.0) return -a.a_h * zeta / a.Pr_0;
    const double Y = sqrt(1.0 - 9.0 * zeta);
    return 2.0 * log((1.0 + Y) / 2.0);
}

struct LhMostEval { double us, ts, g; };

static __device__ __noinline__ LhMostEval lh_most_eval(const LhAtmos& a, double x, double du, double dth, double Lm, double Lh)
{
    LhMostEval e;
    const double k = a.von_karman, z = a.z_atm;
    e.us = k * du / (Lm - lh_most_psi_m(a, z * x) + lh_most_psi_m(a, a.z_0m * x));
    e.ts = k * dth / (a.Pr_0 * (Lh - lh_most_psi_h(a, z * x) + lh_most_psi_h(a, a.z_0s * x)));
    e.g = k * a.grav * e.ts / (e.us * e.us * a.theta_scale);
    return e;
}

// Similarity scales (u*, θ*, q*): secant iteration on F(x) = x - g(x), x = 1/L, from x0 = 0, x1 = g(0); z x in [-1000, 10].
__device__ __forceinline__ void lh_most_scales(const LhAtmos& a, double du, double dth, double dq, double& ustar, double& tstar, double& qstar)
{
    const double k = a.von_karman, z = a.z_atm;
    const double Lm = log(z / a.z_0m), Lh = log(z / a.z_0s);
    const double xmin = -1000.0 / z, xmax = 10.0 / z;
    double x = 0.0;
    LhMostEval e = lh_most_eval(a, 0.0, du, dth, Lm, Lh);
    if (dth != 0.0 && du != 0.0) {
        double x0 = 0.0, F0 = x0 - e.g;
        double x1 = e.g < xmin ? xmin : e.g > xmax ? xmax : e.g;
        e = lh_most_eval(a, x1, du, dth, Lm, Lh);
        double F1 = x1 - e.g;
        x = x1;
        for (int it = 0; it < 60 && F1 != 0.0 && F1 != F0; ++it) {
            double x2 = x1 - F1 * (x1 - x0) / (F1 - F0);
            x2 = x2 < xmin ? xmin : x2 > xmax ? xmax : x2;
            if (x2 == x1) break;
            x0 = x1; F0 = F1;
            x1 = x2;
            e = lh_most_eval(a, x1, du, dth, Lm, Lh);
            F1 = x1 - e.g;
            x = x1;
            if (fabs(F1) <= 4.0e-16 * (fabs(x1) + fabs(e.g))) break;
        }
        e = lh_most_eval(a, x, du, dth, Lm, Lh);
    }
    ustar = e.us;
    tstar = e.ts;
    qstar = k * dq / (a.Pr_0 * (Lh - lh_most_psi_h(a, z * x) + lh_most_psi_h(a, a.z_0s * x)));
}

// boundary_conditions.jl:584-617 given the matric potential psi at min(S_l_eff, 1) and the surface temperature T.
__device__ __forceinline__ void lh_atmos_fluxes(const LhAtmos& a, double psi, double T, double& heat, double& water)
{
    const double dcp = a.cp_v - a.cp_l;
    const double p_vs = a.press_triple * pow(T / a.T_triple, dcp / a.R_v) * exp((a.LH_v0 - dcp * a.T_0) / a.R_v * (1.0 / a.T_triple - 1.0 / T));
    const double q_sat = p_vs / (a.rho_a_sfc * a.R_v * T);                  // q_vap_saturation_generic(param_set, T, ρ_a_sfc, Liquid()) :584
    const double correction = exp(a.grav * psi / a.R_v / T);                // :592
    const double q_surf = q_sat * correction;                               // :593
    double ustar, tstar, qstar;
    lh_most_scales(a, a.u_atm - 0.0, a.theta_atm - T, a.q_atm - q_surf, ustar, tstar, qstar);
    const double cpm = a.cp_d + (a.cp_v - a.cp_d) * q_surf;                 // cp_m(param_set, PhasePartition(q_surf)) :606-607
    const double h_d = a.cp_d * (T - a.T_0) + a.R_d * a.T_0;                // :609
    const double E = -a.rho_a_sfc * ustar * qstar;                          // :612
    const double dry = -cpm * a.rho_a_sfc * ustar * tstar - h_d * E;        // :613
    const double vap = (a.cp_v * (T - a.T_0) + a.LH_v0) * E;                // :614-615
    water = E / a.rho_l;                                                    // :616
    heat = dry + vap;                                                       // :617
}
