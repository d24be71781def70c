// lh_derive.h — host side: the constants the closures use, derived from lh_soil_config (all fp64, computed once in
// lh_soil_create; per column / per cell in rebuild_column_params).  Shared by csrc/lh_soil_api.cu and the host build of the
// product sources that tests/support/hostemu compiles for the CPU-only test run.
#pragma once

#include <math.h>
#include <string.h>

#include <vector>

#include "lh_closures.cuh"
#include "lh_soil.h"

static inline void derive_phys(const lh_soil_config& cfg, LhPhys& d)
{
    const lh_soil_params& q = cfg.params;
    memset(&d, 0, sizeof d);
    d.dz = (cfg.zmax - cfg.zmin) / cfg.nlayer;
    d.inv_dz = 1.0 / d.dz;
    d.half_dz = d.dz / 2.0;            // boundary_cf_distance, boundary_conditions.jl:196-208 (A1)
    d.inv_half_dz = 1.0 / d.half_dz;
    d.neg_half_inv_dz = -0.5 / d.dz;
    d.nu = q.nu;
    d.theta_r = q.theta_r;
    d.theta_r_eps = q.theta_r + LH_EPS;
    d.inv_nu_thr = 1.0 / (q.nu - q.theta_r);
    d.nu_thr = q.nu - q.theta_r;
    d.S_s_inv = 1.0 / q.S_s;
    d.vg_m = q.vg_m;
    d.vg_inv_m = 1.0 / q.vg_m;
    d.vg_inv_n = 1.0 / q.vg_n;
    d.neg_inv_alpha = -1.0 / q.vg_alpha;
    d.Ksat = q.Ksat;
    d.visc_gamma_l2e = q.visc_gamma * 1.4426950408889634074;   // exp(x) = 2^(x log2 e)
    d.visc_T_ref = q.visc_T_ref;
    d.imp_c2 = -q.imp_Omega * log2(10.0);                      // 10^x = 2^(x log2 10)
    d.rho_c_ds = q.rho_c_ds;
    d.rhocp_l = q.cp_l * q.rho_cloud_liq;
    d.rhocp_i = q.cp_i * q.rho_cloud_ice;
    d.rhoi_LH = q.rho_cloud_ice * q.LH_f0;
    d.T_0 = q.T_0;
    d.inv_nu = 1.0 / q.nu;
    d.kersten_p1 = (1.0 + q.nu_ss_om - q.a * q.nu_ss_quartz - q.nu_ss_gravel) / 2.0;
    d.kersten_p2 = 1.0 - q.nu_ss_om;
    d.kersten_p3 = 1.0 + q.nu_ss_om;
    d.neg_b_l2e = -q.b * 1.4426950408889634074;
    d.k_unfrozen = q.kappa_sat_unfrozen;
    d.k_frozen = q.kappa_sat_frozen;
    d.log2_k_unfrozen = log2(q.kappa_sat_unfrozen);
    d.log2_k_frozen = log2(q.kappa_sat_frozen);
    {   // k_dry, SoilHeatParameterizations.jl:268-294 (a per-call scalar in the reference)
        const double rho_b = (1.0 - q.nu) * q.rho_p;
        const double numerator = (q.kappa_dry_parameter * q.kappa_solid - q.K_therm) * rho_b + q.K_therm * q.rho_p;
        const double denom = q.rho_p - (1.0 - q.kappa_dry_parameter) * rho_b;
        d.kappa_dry = numerator / denom;
    }
    d.k_unfrozen_minus_dry = q.kappa_sat_unfrozen - d.kappa_dry;
    d.visc_on = q.viscosity_factor != LH_FACTOR_NONE;
    d.imp_on = q.impedance_factor != LH_FACTOR_NONE;
    d.om_zero = q.nu_ss_om == 0.0;
    d.log2_Sr_sat = log2(q.nu * d.inv_nu);
    d.pow_p1_Sr_sat = (double)powl((long double)(q.nu * d.inv_nu), (long double)d.kersten_p1);
}

// Coefficients (into the parameter block) and tables of the three per-model exponents: 1/m, m and the Kersten exponent.
static inline void derive_pow(LhDevParams& d, std::vector<double>& tab)
{
    tab.assign((size_t)LHPW_COUNT * LH_POW_DOUBLES, 0.0);
    const double* log_tab = d.mc + LHC_TAB0 + LH_TAB_LOG;
    lh_pow_build(d.vg_inv_m, log_tab, &d.pw[LHPW_INVM], tab.data() + (size_t)LHPW_INVM * LH_POW_DOUBLES);
    lh_pow_build(d.vg_m, log_tab, &d.pw[LHPW_M], tab.data() + (size_t)LHPW_M * LH_POW_DOUBLES);
    lh_pow_build(d.kersten_p1, log_tab, &d.pw[LHPW_P1], tab.data() + (size_t)LHPW_P1 * LH_POW_DOUBLES);
}

static inline LhDevParams derive_params(const lh_soil_config& cfg)
{
    LhDevParams d;
    derive_phys(cfg, d);
    static const double coeffs[LHC_COUNT] = {LH_MATH_COEFFS};
    memcpy(d.mc, coeffs, sizeof coeffs);
    return d;
}

