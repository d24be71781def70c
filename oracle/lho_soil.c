/*
 * lho_soil.c — TEST INFRASTRUCTURE ONLY (see lho_soil.h).
 *
 * Plain-C fp64 restatement of the reference's soil right-hand side and the SSPRK33 stage
 * combine.  Compile with -O2 -ffp-contract=off (no FMA contraction: every expression is
 * evaluated in the order the Julia source writes it).  Columns are independent, so the only
 * parallelism is an OpenMP loop over columns.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Third-party arithmetic that is NOT in the reference tree:
 *   - ClimaCore (unpinned, ~v0.2.x) finite-difference operators InterpolateC2F / GradientC2F /
 *     DivergenceF2C+SetValue on a uniform IntervalMesh: restated as in SURVEY §3.4
 *     (arithmetic mean, centred difference / dz, flux-form divergence with the boundary face
 *     fluxes set), call sites right_hand_side.jl:170-181,249-259,337-365.
 *   - OrdinaryDiffEq v5 SSPRK33 (Shu-Osher), call site src/Simulations/simulation.jl:63-70.
 *   - CLIMAParameters v0.1 constants arrive through lh_soil_params at run time.
 */
#include "lho_soil.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Scalar closures
 * ---------------------------------------------------------------------------------------- */

/* Julia's max(a, b) propagates NaN (fmax does not). */
static inline double jl_max(double a, double b)
{
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}

/* SoilWaterParameterizations.jl:181-188 */
double lho_volumetric_liquid_fraction(double theta_l_aug, double nu_eff)
{
    double theta_l;
    if (theta_l_aug < nu_eff) theta_l = theta_l_aug;
    else theta_l = nu_eff;
    return theta_l;
}

/* SoilWaterParameterizations.jl:213-217; eps(Float64) == DBL_EPSILON */
double lho_effective_saturation(double porosity, double theta_l_aug, double theta_r)
{
    double safe = jl_max(theta_l_aug, theta_r + DBL_EPSILON);
    double S_l = (safe - theta_r) / (porosity - theta_r);
    return S_l;
}

/* SoilWaterParameterizations.jl:196-200:
 *   ψ_m = -((S^(-FT(1) / m) - FT(1)) * α^(-n))^(FT(1) / n)                                  */
double lho_matric_potential(const lh_soil_params* p, double S)
{
    double n = p->vg_n, alpha = p->vg_alpha, m = p->vg_m;
    double psi_m = -pow((pow(S, -1.0 / m) - 1.0) * pow(alpha, -n), 1.0 / n);
    return psi_m;
}

/* SoilWaterParameterizations.jl:253-258; the reference errors for ψ > 0, here NaN. */
double lho_inverse_matric_potential(const lh_soil_params* p, double psi)
{
    if (psi > 0) return NAN;
    double n = p->vg_n, m = p->vg_m, alpha = p->vg_alpha;
    double S = pow(1.0 + pow(alpha * fabs(psi), n), -m);
    return S;
}

/* SoilWaterParameterizations.jl:229-242 */
double lho_pressure_head(const lh_soil_params* p, double theta_l_aug, double nu_eff, double S_s)
{
    double S_l_eff = lho_effective_saturation(nu_eff, theta_l_aug, p->theta_r);
    double psi;
    if (S_l_eff <= 1.0) psi = lho_matric_potential(p, S_l_eff);
    else psi = (theta_l_aug - nu_eff) / S_s;
    return psi;
}

/* SoilWaterParameterizations.jl:269-282:
 *   K = sqrt(S) * (FT(1) - (FT(1) - S^(FT(1) / m))^m)^FT(2);  return K * Ksat * visc * imp   */
double lho_hydraulic_conductivity(const lh_soil_params* p, double S, double visc_f, double imp_f)
{
    double Ksat = p->Ksat, m = p->vg_m;
    double K;
    if (S < 1.0) K = sqrt(S) * pow(1.0 - pow(1.0 - pow(S, 1.0 / m), m), 2.0);
    else K = 1.0;
    return K * Ksat * visc_f * imp_f;
}

/* SoilWaterParameterizations.jl:104-126 */
double lho_viscosity_factor(const lh_soil_params* p, double T)
{
    if (p->viscosity_factor == LH_FACTOR_NONE) return 1.0;
    double factor = p->visc_gamma * (T - p->visc_T_ref);
    return exp(factor);
}

/* SoilWaterParameterizations.jl:76-93 */
double lho_impedance_factor(const lh_soil_params* p, double f_i)
{
    if (p->impedance_factor == LH_FACTOR_NONE) return 1.0;
    return pow(10.0, -p->imp_Omega * f_i);
}

/* SoilWaterParameterizations.jl:290-306 */
double lho_hydrostatic_profile(const lh_soil_params* p, double z, double z_interface, double nu,
                               double S_s)
{
    double alpha = p->vg_alpha, m = p->vg_m, n = p->vg_n, theta_r = p->theta_r;
    double theta;
    if (z > z_interface) {
        double S = pow(1.0 + pow(alpha * (z - z_interface), n), -m);
        theta = S * (nu - theta_r) + theta_r;
    } else {
        theta = -S_s * (z - z_interface) + nu;
    }
    return theta;
}

/* SoilHeatParameterizations.jl:65-79 */
double lho_volumetric_heat_capacity(const lh_soil_params* p, double theta_l, double theta_i,
                                    double rho_c_ds)
{
    double rho_i = p->rho_cloud_ice;
    double rhocp_i = p->cp_i * rho_i;
    double rho_l = p->rho_cloud_liq;
    double rhocp_l = p->cp_l * rho_l;
    double rho_c_s = rho_c_ds + theta_l * rhocp_l + theta_i * rhocp_i;
    return rho_c_s;
}

/* SoilHeatParameterizations.jl:42-53 */
double lho_temperature_from_rho_e_int(const lh_soil_params* p, double rho_e_int, double theta_i,
                                      double rho_c_s)
{
    double T = p->T_0 + (rho_e_int + theta_i * p->rho_cloud_ice * p->LH_f0) / rho_c_s;
    return T;
}

/* SoilHeatParameterizations.jl:91-102 */
double lho_volumetric_internal_energy(const lh_soil_params* p, double theta_i, double rho_c_s,
                                      double T)
{
    double rho_e_int = rho_c_s * (T - p->T_0) - theta_i * p->rho_cloud_ice * p->LH_f0;
    return rho_e_int;
}

/* SoilHeatParameterizations.jl:114-128 */
double lho_saturated_thermal_conductivity(double theta_l, double theta_i, double k_unfrozen,
                                          double k_frozen)
{
    double theta_w = theta_l + theta_i;
    double k_sat;
    if (theta_w < DBL_EPSILON) k_sat = 0.0;
    else k_sat = pow(k_unfrozen, theta_l / theta_w) * pow(k_frozen, theta_i / theta_w);
    return k_sat;
}

/* SoilHeatParameterizations.jl:139-142 */
double lho_relative_saturation(double theta_l, double theta_i, double porosity)
{
    return (theta_l + theta_i) / porosity;
}

/* SoilHeatParameterizations.jl:152-174 */
double lho_kersten_number(const lh_soil_params* p, double theta_i, double S_r)
{
    double a = p->a, b = p->b;
    double nu_om = p->nu_ss_om, nu_q = p->nu_ss_quartz, nu_g = p->nu_ss_gravel;
    double K_e;
    if (theta_i < DBL_EPSILON) {
        K_e = pow(S_r, (1.0 + nu_om - a * nu_q - nu_g) / 2.0) *
              pow(pow(1.0 + exp(-b * S_r), -3.0) - pow((1.0 - S_r) / 2.0, 3.0), 1.0 - nu_om);
    } else {
        K_e = pow(S_r, 1.0 + nu_om);
    }
    return K_e;
}

/* SoilHeatParameterizations.jl:185-188 */
double lho_thermal_conductivity(double kappa_dry, double K_e, double kappa_sat)
{
    return K_e * kappa_sat + (1.0 - K_e) * kappa_dry;
}

/* SoilHeatParameterizations.jl:198-207 */
double lho_volumetric_internal_energy_liq(const lh_soil_params* p, double T)
{
    double rhocp_l = p->cp_l * p->rho_cloud_liq;
    return rhocp_l * (T - p->T_0);
}

/* SoilHeatParameterizations.jl:223-233 */
double lho_k_solid(double nu_ss_om, double nu_ss_quartz, double k_quartz, double k_minerals,
                   double k_om)
{
    return pow(k_om, nu_ss_om) * pow(k_quartz, nu_ss_quartz) *
           pow(k_minerals, 1.0 - nu_ss_om - nu_ss_quartz);
}

/* SoilHeatParameterizations.jl:245-247 */
double lho_ksat_frozen(double k_solid, double porosity, double k_ice)
{
    return pow(k_solid, 1.0 - porosity) * pow(k_ice, porosity);
}

/* SoilHeatParameterizations.jl:258-260 */
double lho_ksat_unfrozen(double k_solid, double porosity, double k_liq)
{
    return pow(k_solid, 1.0 - porosity) * pow(k_liq, porosity);
}

/* SoilHeatParameterizations.jl:268-294 (ρb_ss and k_dry) */
double lho_k_dry(const lh_soil_params* p)
{
    double kdp = p->kappa_dry_parameter;
    double porosity = p->nu, rho_p = p->rho_p, k_solid = p->kappa_solid;
    double k_air = p->K_therm;
    double rho_b = (1.0 - porosity) * rho_p;
    double numerator = (kdp * k_solid - k_air) * rho_b + k_air * rho_p;
    double denom = rho_p - (1.0 - kdp) * rho_b;
    return numerator / denom;
}

/* ------------------------------------------------------------------------------------------
 * PrescribedAtmosForcing: compute_turbulent_surface_fluxes, boundary_conditions.jl:555-620.
 * PARITY UNPINNED for the two third-party pieces (include/lh_soil.h): they are restated from the
 * published formulations, everything the reference itself writes down is literal.
 * ---------------------------------------------------------------------------------------- */
/* Saturation vapour pressure over liquid, Clausius-Clapeyron with constant Δcp = cp_v - cp_l
 * (Romps 2008 eq. 10; Thermodynamics.jl `saturation_vapor_pressure(param_set, T, LH_0, Δcp)`):
 *   p_vs = p_tr (T/T_tr)^(Δcp/R_v) exp((LH_v0 - Δcp T_0)/R_v (1/T_tr - 1/T));  q_vs = p_vs / (ρ R_v T). */
double lho_q_vap_saturation_liquid(const lh_soil_params* p, const lh_soil_atmos* a, double T, double rho)
{
    const double dcp = a->cp_v - p->cp_l;
    const double p_vs = a->press_triple * pow(T / a->T_triple, dcp / a->R_v) *
                        exp((a->LH_v0 - dcp * p->T_0) / a->R_v * (1.0 / a->T_triple - 1.0 / T));
    return p_vs / (rho * a->R_v * T);
}

/* Businger-Dyer integrated universal functions (Businger et al. 1971; Dyer 1974; Paulson 1970 for the
 * integrated unstable forms).                                                                 */
static double most_psi_m(const lh_soil_atmos* a, double zeta)
{
    if (zeta >= 0.0) return -a->a_m * zeta;
    const double X = sqrt(sqrt(1.0 - 15.0 * zeta));
    return 2.0 * log((1.0 + X) / 2.0) + log((1.0 + X * X) / 2.0) - 2.0 * atan(X) + 1.57079632679489661923;
}

static double most_psi_h(const lh_soil_atmos* a, double zeta)
{
    if (zeta >= 0.0) return -a->a_h * zeta / a->Pr_0;
    const double Y = sqrt(1.0 - 9.0 * zeta);
    return 2.0 * log((1.0 + Y) / 2.0);
}

/* Similarity scales for differences (du, dth, dq) between z_atm and the surface, roughness lengths z0m / z0s.
 * Unknown x = 1/L (so that the neutral limit is x = 0, no division by θ*):
 *   u*(x) = κ du / (ln(z/z0m) - ψ_m(z x) + ψ_m(z0m x))
 *   θ*(x) = κ dθ / (Pr_0 (ln(z/z0s) - ψ_h(z x) + ψ_h(z0s x))),  q* likewise
 *   x = κ g θ* / (u*^2 θ_scale)            (Obukhov length from the θ flux)
 * solved by a secant iteration on F(x) = x - g(x) from x0 = 0, x1 = g(0), with z x kept in [-1000, 10].    */
static void most_scales(const lh_soil_atmos* a, double du, double dth, double dq, double z0m, double z0s,
                        double* ustar, double* tstar, double* qstar)
{
    const double k = a->von_karman, z = a->z_atm;
    const double Lm = log(z / z0m), Lh = log(z / z0s);
    const double xmin = -1000.0 / z, xmax = 10.0 / z;
#define MOST_EVAL(x, us, ts, gx)                                                                  \
    do {                                                                                          \
        us = k * du / (Lm - most_psi_m(a, z * (x)) + most_psi_m(a, z0m * (x)));                   \
        ts = k * dth / (a->Pr_0 * (Lh - most_psi_h(a, z * (x)) + most_psi_h(a, z0s * (x))));      \
        gx = k * a->grav * ts / (us * us * a->theta_scale);                                       \
    } while (0)
    double x = 0.0, us, ts, g0;
    MOST_EVAL(0.0, us, ts, g0);
    if (dth != 0.0 && du != 0.0) {
        double x0 = 0.0, F0 = x0 - g0;
        double x1 = g0 < xmin ? xmin : g0 > xmax ? xmax : g0, g1;
        MOST_EVAL(x1, us, ts, g1);
        double F1 = x1 - g1;
        x = x1;
        for (int it = 0; it < 60 && F1 != 0.0 && F1 != F0; ++it) {
            double x2 = x1 - F1 * (x1 - x0) / (F1 - F0);
            x2 = x2 < xmin ? xmin : x2 > xmax ? xmax : x2;
            if (x2 == x1) break;
            x0 = x1; F0 = F1;
            x1 = x2;
            MOST_EVAL(x1, us, ts, g1);
            F1 = x1 - g1;
            x = x1;
            if (fabs(F1) <= 4.0e-16 * (fabs(x1) + fabs(g1))) break;
        }
        MOST_EVAL(x, us, ts, g1);
    }
#undef MOST_EVAL
    *ustar = us;
    *tstar = ts;
    *qstar = k * dq / (a->Pr_0 * (Lh - most_psi_h(a, z * x) + most_psi_h(a, z0s * x)));
}

/* compute_turbulent_surface_fluxes(energy, hydrology, model, ϑ_l, θ_i, T)  boundary_conditions.jl:555-620 */
void lho_turbulent_surface_fluxes(const lh_soil_params* p, const lh_soil_atmos* a, double theta_l_aug, double theta_i,
                                  double T, double* heat_flux, double* water_flux)
{
    const double q_sat = lho_q_vap_saturation_liquid(p, a, T, a->rho_a_sfc);                 /* :584 */
    const double nu_eff = p->nu - theta_i;                                                     /* :588 */
    const double theta_l = lho_volumetric_liquid_fraction(theta_l_aug, nu_eff);                /* :589 */
    double S_l_eff = lho_effective_saturation(nu_eff, theta_l, p->theta_r);                    /* :590 */
    if (!(S_l_eff < 1.0)) S_l_eff = 1.0;                                                       /* min(., 1) */
    const double psi = lho_matric_potential(p, S_l_eff);                                       /* :591 */
    const double correction = exp(a->grav * psi / a->R_v / T);                                 /* :592 */
    const double q_surf = q_sat * correction;                                                  /* :593 */
    double ustar, tstar, qstar;
    most_scales(a, a->u_atm - 0.0, a->theta_atm - T, a->q_atm - q_surf, p->z_0m, p->z_0s, &ustar, &tstar, &qstar);
    const double cpm = a->cp_d + (a->cp_v - a->cp_d) * q_surf;                                 /* cp_m(PhasePartition(q_surf)) :606-607 */
    const double T_ref = p->T_0;
    const double h_d = a->cp_d * (T - T_ref) + a->R_d * T_ref;                                 /* :609 */
    const double E = -a->rho_a_sfc * ustar * qstar;                                            /* :612 */
    const double dry = -cpm * a->rho_a_sfc * ustar * tstar - h_d * E;                          /* :613 */
    const double vap = (a->cp_v * (T - T_ref) + a->LH_v0) * E;                                 /* :614-615 */
    *water_flux = E / p->rho_cloud_liq;                                                        /* :616 */
    *heat_flux = dry + vap;                                                                    /* :617 */
}

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */

struct lho_soil_ctx {
    lh_soil_config cfg;
    int64_t ncol;
    int32_t nlayer;
    double dz;
    double* zc;           /* nlayer                                              */
    double* f[LH_NUM_FIELDS]; /* [col * nlayer + layer]                          */
    double* u1[3];        /* stage buffer (all three prognostic slots, like the
                             reference's generic axpy over the whole FieldVector) */
    double* tend[3];
    double* Fw;           /* face fluxes of the last rhs call, [col*(nlayer+1)+j] */
    double* Fe;
    double* colp[12];     /* per-column nu, theta_r, vg_n, vg_alpha, Ksat | rho_c_ds, kappa_sat_unfrozen, kappa_sat_frozen, kappa_solid,
                             nu_ss_om, nu_ss_quartz, nu_ss_gravel (NULL: uniform); lho_soil_set_column_params / _heat_params */
    double bcv[4];
    double* cellp[5];     /* lho_soil_set_cell_params: per-cell nu, theta_r, vg_n, vg_alpha, Ksat, [col * nlayer + layer] (NULL: per column) */
    double* flux_cols[4]; /* lho_soil_set_column_fluxes: per-column VerticalFlux values, LH_BCV_* order (NULL: scalar) */
    lh_soil_atmos atmos;  /* lho_soil_set_atmos_forcing                                 */
    int atmos_on;
    double* aux_tab[LH_NUM_FIELDS]; /* lho_soil_set_aux_table: [nrows][nlayer] per prescribed field */
    int64_t aux_rows[LH_NUM_FIELDS];
    int64_t aux_row;
    double async_bud[8][2];  /* lho_soil_budgets_async results by ticket % 8 */
    int64_t async_ticket[8];
    int64_t next_ticket;
    double last_ms;
    int64_t last_launches;
    char err[256];
};

static _Thread_local char g_create_err[256];

static int32_t fail(lho_soil_ctx* ctx, int32_t code, const char* msg)
{
    if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", msg);
    else snprintf(g_create_err, sizeof g_create_err, "%s", msg);
    return code;
}

int32_t lho_soil_abi_version(void) { return LH_SOIL_ABI_VERSION; }

const char* lho_soil_last_error(const lho_soil_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int32_t lho_soil_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void lho_soil_set_num_threads(int32_t n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static int has_water(int model) { return model == LH_MODEL_RICHARDS || model == LH_MODEL_COUPLED; }
static int has_heat(int model) { return model == LH_MODEL_HEAT || model == LH_MODEL_COUPLED; }

/* The reference's method table for vertical_flux (boundary_conditions.jl:295-444):
 *   dynamic energy   : VerticalFlux, Dirichlet            (NoBC -> `nothing` into SetValue)
 *   dynamic hydrology: VerticalFlux, Dirichlet, FreeDrainage
 *   prescribed       : NoBC, VerticalFlux (value unused)
 * anything else is a MethodError.                                                           */
static int32_t validate_face(const lh_soil_face_bc* bc, int model, const char* face)
{
    char msg[200];
    int ek = bc->energy_kind, hk = bc->hydrology_kind;
    if (ek < 0 || ek > 3 || hk < 0 || hk > 3) {
        snprintf(msg, sizeof msg, "%s: unknown BC kind", face);
        return fail(NULL, LH_ERR_INVALID_ARG, msg);
    }
    if (has_heat(model)) {
        if (!(ek == LH_BC_FLUX || ek == LH_BC_DIRICHLET)) {
            snprintf(msg, sizeof msg, "%s: energy BC kind %d has no vertical_flux method for SoilEnergyModel", face, ek);
            return fail(NULL, LH_ERR_UNSUPPORTED_BC, msg);
        }
    } else if (!(ek == LH_BC_NONE || ek == LH_BC_FLUX)) {
        snprintf(msg, sizeof msg, "%s: energy BC kind %d has no vertical_flux method for PrescribedTemperatureModel", face, ek);
        return fail(NULL, LH_ERR_UNSUPPORTED_BC, msg);
    }
    if (has_water(model)) {
        if (!(hk == LH_BC_FLUX || hk == LH_BC_DIRICHLET || hk == LH_BC_FREE_DRAINAGE)) {
            snprintf(msg, sizeof msg, "%s: hydrology BC kind %d has no vertical_flux method for SoilHydrologyModel", face, hk);
            return fail(NULL, LH_ERR_UNSUPPORTED_BC, msg);
        }
    } else if (!(hk == LH_BC_NONE || hk == LH_BC_FLUX)) {
        snprintf(msg, sizeof msg, "%s: hydrology BC kind %d has no vertical_flux method for PrescribedHydrologyModel", face, hk);
        return fail(NULL, LH_ERR_UNSUPPORTED_BC, msg);
    }
    return LH_OK;
}

int32_t lho_soil_create(const lh_soil_config* cfg, lho_soil_ctx** out)
{
    if (!cfg || !out) return fail(NULL, LH_ERR_INVALID_ARG, "cfg/out is NULL");
    *out = NULL;
    if (cfg->struct_size != (int32_t)sizeof(lh_soil_config))
        return fail(NULL, LH_ERR_INVALID_ARG, "lh_soil_config.struct_size mismatch");
    if (cfg->ncol < 1 || cfg->nlayer < 1) return fail(NULL, LH_ERR_INVALID_ARG, "ncol and nlayer must be >= 1");
    if (cfg->model < 0 || cfg->model > 2) return fail(NULL, LH_ERR_INVALID_ARG, "unknown model kind");
    /* Domains/domain.jl:30  @assert zlim[1] < zlim[2] */
    if (!(cfg->zmin < cfg->zmax)) return fail(NULL, LH_ERR_DOMAIN, "zlim[1] < zlim[2] violated");
    int32_t st;
    if ((st = validate_face(&cfg->top, cfg->model, "top")) != LH_OK) return st;
    if ((st = validate_face(&cfg->bottom, cfg->model, "bottom")) != LH_OK) return st;

    lho_soil_ctx* c = (lho_soil_ctx*)calloc(1, sizeof *c);
    if (!c) return fail(NULL, LH_ERR_INVALID_ARG, "out of memory");
    c->cfg = *cfg;
    c->ncol = cfg->ncol;
    c->nlayer = cfg->nlayer;
    int n = cfg->nlayer;
    /* Domains/domain.jl:58-69: uniform IntervalMesh; faces zmin + j (zmax - zmin) / n, centres
     * are face midpoints (pinned: zc == -1.95:0.1:-0.05 for (-2, 0), n = 20; coupled.jl:198). */
    c->dz = (cfg->zmax - cfg->zmin) / n;
    size_t cells = (size_t)c->ncol * (size_t)n;
    c->zc = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) {
        double zf0 = cfg->zmin + (cfg->zmax - cfg->zmin) * (double)i / (double)n;
        double zf1 = cfg->zmin + (cfg->zmax - cfg->zmin) * (double)(i + 1) / (double)n;
        c->zc[i] = (zf0 + zf1) / 2.0;
    }
    for (int k = 0; k < LH_NUM_FIELDS; ++k) c->f[k] = (double*)calloc(cells, sizeof(double));
    for (int k = 0; k < 3; ++k) {
        c->u1[k] = (double*)calloc(cells, sizeof(double));
        c->tend[k] = (double*)calloc(cells, sizeof(double));
    }
    c->Fw = (double*)calloc((size_t)c->ncol * (n + 1), sizeof(double));
    c->Fe = (double*)calloc((size_t)c->ncol * (n + 1), sizeof(double));
    /* PrescribedTemperatureModel default T ≡ 288 (models.jl:51-54) */
    for (size_t i = 0; i < cells; ++i) c->f[LH_FIELD_T][i] = 288.0;
    c->bcv[LH_BCV_TOP_ENERGY] = cfg->top.energy_value;
    c->bcv[LH_BCV_TOP_HYDROLOGY] = cfg->top.hydrology_value;
    c->bcv[LH_BCV_BOTTOM_ENERGY] = cfg->bottom.energy_value;
    c->bcv[LH_BCV_BOTTOM_HYDROLOGY] = cfg->bottom.hydrology_value;
    *out = c;
    return LH_OK;
}

int32_t lho_soil_destroy(lho_soil_ctx* c)
{
    if (!c) return LH_OK;
    free(c->zc);
    for (int k = 0; k < LH_NUM_FIELDS; ++k) free(c->f[k]);
    for (int k = 0; k < 3; ++k) { free(c->u1[k]); free(c->tend[k]); }
    free(c->Fw); free(c->Fe);
    for (int k = 0; k < 12; ++k) free(c->colp[k]);
    for (int k = 0; k < LH_NUM_FIELDS; ++k) free(c->aux_tab[k]);
    for (int k = 0; k < 4; ++k) free(c->flux_cols[k]);
    for (int k = 0; k < 5; ++k) free(c->cellp[k]);
    free(c);
    return LH_OK;
}

int32_t lho_soil_get_zc(const lho_soil_ctx* c, double* zc_out)
{
    if (!c || !zc_out) return LH_ERR_INVALID_ARG;
    memcpy(zc_out, c->zc, sizeof(double) * c->nlayer);
    return LH_OK;
}

static int32_t copy_in(lho_soil_ctx* c, double* dst, const double* host, int64_t cs, int64_t ls)
{
    if (!host) return fail(c, LH_ERR_INVALID_ARG, "host pointer is NULL");
    int n = c->nlayer;
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
    for (int64_t col = 0; col < c->ncol; ++col)
        for (int i = 0; i < n; ++i) dst[col * n + i] = host[col * cs + i * ls];
    return LH_OK;
}

static int32_t copy_out(lho_soil_ctx* c, const double* src, double* host, int64_t cs, int64_t ls)
{
    if (!host) return fail(c, LH_ERR_INVALID_ARG, "host pointer is NULL");
    int n = c->nlayer;
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
    for (int64_t col = 0; col < c->ncol; ++col)
        for (int i = 0; i < n; ++i) host[col * cs + i * ls] = src[col * n + i];
    return LH_OK;
}

int32_t lho_soil_set_state(lho_soil_ctx* c, int32_t field, const double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (field < 0 || field >= LH_NUM_FIELDS) return fail(c, LH_ERR_INVALID_ARG, "bad field id");
    return copy_in(c, c->f[field], host, cs, ls);
}

int32_t lho_soil_set_aux(lho_soil_ctx* c, int32_t field, const double* host, int64_t cs, int64_t ls)
{
    return lho_soil_set_state(c, field, host, cs, ls);
}

int32_t lho_soil_get_state(lho_soil_ctx* c, int32_t field, double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (field < 0 || field >= LH_NUM_FIELDS) return fail(c, LH_ERR_INVALID_ARG, "bad field id");
    return copy_out(c, c->f[field], host, cs, ls);
}

int32_t lho_soil_get_tendency(lho_soil_ctx* c, int32_t field, double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (field < 0 || field >= 3) return fail(c, LH_ERR_INVALID_ARG, "bad field id");
    return copy_out(c, c->tend[field], host, cs, ls);
}

int32_t lho_soil_set_bc_values(lho_soil_ctx* c, const double values[4])
{
    if (!c || !values) return LH_ERR_INVALID_ARG;
    memcpy(c->bcv, values, sizeof c->bcv);
    return LH_OK;
}

/* ------------------------------------------------------------------------------------------
 * The right-hand side of one column
 * ---------------------------------------------------------------------------------------- */

typedef struct { double K, psi, kappa, T; } cell_closures;

/* Pointwise block of right_hand_side.jl:156-167 (Richards), :209-224 (heat), :291-314
 * (coupled), evaluated for one cell.  T_in is the prescribed T for the Richards model.      */
static cell_closures closures_at(const lh_soil_params* p, int model, double kappa_dry,
                                 double th, double ti, double re, double T_in)
{
    cell_closures c = {0.0, 0.0, 0.0, T_in};
    double nu = p->nu;
    double nu_eff = nu - ti;
    double tl = lho_volumetric_liquid_fraction(th, nu_eff);
    if (has_heat(model)) {
        double rho_c_s = lho_volumetric_heat_capacity(p, tl, ti, p->rho_c_ds);
        c.T = lho_temperature_from_rho_e_int(p, re, ti, rho_c_s);
        double S_r = lho_relative_saturation(tl, ti, nu);
        double kersten = lho_kersten_number(p, ti, S_r);
        double k_sat = lho_saturated_thermal_conductivity(tl, ti, p->kappa_sat_unfrozen,
                                                          p->kappa_sat_frozen);
        c.kappa = lho_thermal_conductivity(kappa_dry, kersten, k_sat);
    }
    if (has_water(model)) {
        double f_i = ti / (tl + ti);
        double visc = lho_viscosity_factor(p, c.T);
        double imp = lho_impedance_factor(p, f_i);
        double S = lho_effective_saturation(nu, th, p->theta_r);
        c.K = lho_hydraulic_conductivity(p, S, visc, imp);
        c.psi = lho_pressure_head(p, th, nu_eff, p->S_s);
    }
    return c;
}

/* boundary_fluxes(X, bc::SoilComponentBC, face, model, cs, t)  boundary_conditions.jl:470-489.
 * (th_c, ti_c, T_c) are interior_values (:174-190) at the cell next to the face; dzb is
 * boundary_cf_distance (:196-208) = dz/2 on the uniform mesh (assumption A1).
 * is_bottom flips the sign of the Dirichlet fluxes (:396-398, :439-441).                    */
static void boundary_fluxes(const lh_soil_params* p, int model, double kappa_dry,
                            const lh_soil_face_bc* bc, double val_e, double val_h, int is_bottom,
                            double th_c, double ti_c, double T_c, double dzb,
                            double* f_energy, double* f_water)
{
    /* initialize_boundary_values :218-228: pairs [centre, face], face := centre */
    double th[2] = {th_c, th_c}, T[2] = {T_c, T_c}, ti[2] = {ti_c, ti_c};
    /* set_boundary_values! :241-288, energy first then hydrology (:482-483); the Dirichlet
     * methods exist only for the dynamic components.                                        */
    if (bc->energy_kind == LH_BC_DIRICHLET && has_heat(model)) T[1] = val_e;
    if (bc->hydrology_kind == LH_BC_DIRICHLET && has_water(model)) th[1] = val_h;

    double nu = p->nu;
    *f_energy = 0.0;
    *f_water = 0.0;

    /* energy: vertical_flux(bc.energy, energy, X_cf, model, dz, face) */
    if (bc->energy_kind == LH_BC_FLUX) {
        *f_energy = val_e;                                        /* :295-301 */
    } else if (bc->energy_kind == LH_BC_DIRICHLET) {              /* :416-444 */
        double kappa[2];
        for (int k = 0; k < 2; ++k) {
            double nu_eff = nu - ti[k];
            double tl = lho_volumetric_liquid_fraction(th[k], nu_eff);
            double S_r = lho_relative_saturation(tl, ti[k], nu);
            double kersten = lho_kersten_number(p, ti[k], S_r);
            double k_sat = lho_saturated_thermal_conductivity(tl, ti[k], p->kappa_sat_unfrozen,
                                                              p->kappa_sat_frozen);
            kappa[k] = lho_thermal_conductivity(kappa_dry, kersten, k_sat);
        }
        double flux = -kappa[1] * (T[1] - T[0]) / dzb;
        if (is_bottom) flux *= -1;
        *f_energy = flux;
    }

    /* hydrology: vertical_flux(bc.hydrology, hydrology, X_cf, model, dz, face) */
    if (bc->hydrology_kind == LH_BC_FLUX) {
        *f_water = val_h;                                         /* :295-301 */
    } else if (bc->hydrology_kind == LH_BC_FREE_DRAINAGE) {       /* :328-356, centre values */
        double nu_eff = nu - ti[0];
        double tl = lho_volumetric_liquid_fraction(th[0], nu_eff);
        double f_i = ti[0] / (tl + ti[0]);
        double imp = lho_impedance_factor(p, f_i);
        double visc = lho_viscosity_factor(p, T[0]);
        double S = lho_effective_saturation(nu, th[0], p->theta_r);
        double K = lho_hydraulic_conductivity(p, S, visc, imp);
        *f_water = -K;
    } else if (bc->hydrology_kind == LH_BC_DIRICHLET) {           /* :371-401 */
        double K[2], psi[2];
        for (int k = 0; k < 2; ++k) {
            double nu_eff = nu - ti[k];
            double tl = lho_volumetric_liquid_fraction(th[k], nu_eff);
            double f_i = ti[k] / (tl + ti[k]);
            double imp = lho_impedance_factor(p, f_i);
            double visc = lho_viscosity_factor(p, T[k]);
            double S = lho_effective_saturation(nu, th[k], p->theta_r);
            K[k] = lho_hydraulic_conductivity(p, S, visc, imp);
            psi[k] = lho_pressure_head(p, th[k], nu_eff, p->S_s);
        }
        double flux = -K[1] * (psi[1] - psi[0] + dzb) / dzb;
        if (is_bottom) flux *= -1;
        *f_water = flux;
    }
}

/* The parameter set of one column: the model's SoilParams / vanGenuchten with the per-column overrides of
 * lho_soil_set_column_params (m = 1 - 1/n as the reference constructor computes it).                          */
static lh_soil_params params_of_column(const lho_soil_ctx* c, int64_t col)
{
    lh_soil_params p = c->cfg.params;
    if (c->colp[0]) p.nu = c->colp[0][col];
    if (c->colp[1]) p.theta_r = c->colp[1][col];
    if (c->colp[2]) { p.vg_n = c->colp[2][col]; p.vg_m = 1.0 - 1.0 / p.vg_n; }
    if (c->colp[3]) p.vg_alpha = c->colp[3][col];
    if (c->colp[4]) p.Ksat = c->colp[4][col];
    if (c->colp[5]) p.rho_c_ds = c->colp[5][col];
    if (c->colp[6]) p.kappa_sat_unfrozen = c->colp[6][col];
    if (c->colp[7]) p.kappa_sat_frozen = c->colp[7][col];
    if (c->colp[8]) p.kappa_solid = c->colp[8][col];
    if (c->colp[9]) p.nu_ss_om = c->colp[9][col];
    if (c->colp[10]) p.nu_ss_quartz = c->colp[10][col];
    if (c->colp[11]) p.nu_ss_gravel = c->colp[11][col];
    return p;
}

/* The parameter set of one CELL: the column's, with the per-cell overrides of lho_soil_set_cell_params. */
static lh_soil_params params_of_cell(const lho_soil_ctx* c, const lh_soil_params* pcol, int64_t col, int i)
{
    lh_soil_params p = *pcol;
    const size_t o = (size_t)col * c->nlayer + i;
    if (c->cellp[0]) p.nu = c->cellp[0][o];
    if (c->cellp[1]) p.theta_r = c->cellp[1][o];
    if (c->cellp[2]) { p.vg_n = c->cellp[2][o]; p.vg_m = 1.0 - 1.0 / p.vg_n; }
    if (c->cellp[3]) p.vg_alpha = c->cellp[3][o];
    if (c->cellp[4]) p.Ksat = c->cellp[4][o];
    return p;
}

static int has_cell_params(const lho_soil_ctx* c)
{
    return c->cellp[0] || c->cellp[1] || c->cellp[2] || c->cellp[3] || c->cellp[4];
}

/* One column.  u_th/u_ti/u_re/u_T: nlayer values each (layer 0 = bottom).  work: 5*nlayer.
 * Fw/Fe: nlayer+1 face fluxes (outputs).                                                    */
static void column_rhs(const lho_soil_ctx* c, int64_t col, const lh_soil_params* p, const double bcv[4], double kappa_dry,
                       const double* u_th, const double* u_ti, const double* u_re,
                       const double* u_T, double* d_th, double* d_ti, double* d_re,
                       double* Fw, double* Fe, double* work)
{
    const int model = c->cfg.model;
    const int n = c->nlayer;
    const double dz = c->dz;
    double* K = work;
    double* h = work + n;
    double* kappa = work + 2 * n;
    double* T = work + 3 * n;
    double* eK = work + 4 * n;

    const int layered = has_cell_params(c);
    lh_soil_params p_bot = *p, p_top = *p;          /* the parameters the boundary fluxes see: those of the cell next to the face */
    double kd_bot = kappa_dry, kd_top = kappa_dry;
    for (int i = 0; i < n; ++i) {
        lh_soil_params pi = *p;
        double kd = kappa_dry;
        if (layered) {
            pi = params_of_cell(c, p, col, i);
            kd = lho_k_dry(&pi);
            if (i == 0) { p_bot = pi; kd_bot = kd; }
            if (i == n - 1) { p_top = pi; kd_top = kd; }
        }
        cell_closures cc = closures_at(&pi, model, kd, u_th[i], u_ti[i], u_re[i], u_T[i]);
        K[i] = cc.K;
        h[i] = cc.psi + c->zc[i];                                   /* :167, :314 */
        kappa[i] = cc.kappa;
        T[i] = cc.T;
        /* ρe_int_l * K, :306 and :364 */
        eK[i] = (model == LH_MODEL_COUPLED) ? lho_volumetric_internal_energy_liq(&pi, cc.T) * cc.K : 0.0;
    }

    double fe_top, fw_top, fe_bot, fw_bot;
    boundary_fluxes(&p_top, model, kd_top, &c->cfg.top, bcv[LH_BCV_TOP_ENERGY],
                    bcv[LH_BCV_TOP_HYDROLOGY], 0, u_th[n - 1], u_ti[n - 1], T[n - 1], dz / 2.0,
                    &fe_top, &fw_top);
    boundary_fluxes(&p_bot, model, kd_bot, &c->cfg.bottom, bcv[LH_BCV_BOTTOM_ENERGY],
                    bcv[LH_BCV_BOTTOM_HYDROLOGY], 1, u_th[0], u_ti[0], T[0], dz / 2.0,
                    &fe_bot, &fw_bot);
    /* boundary_fluxes(X, bc::PrescribedAtmosForcing, :top, ...) :516-536: from interior_values at the top cell */
    if (c->atmos_on) lho_turbulent_surface_fluxes(&p_top, &c->atmos, u_th[n - 1], u_ti[n - 1], T[n - 1], &fe_top, &fw_top);

    /* Face fluxes.  Interior face j sits between cells j-1 and j (0-based), j = 1..n-1:
     *   water  :181/:358   -interpc2f(K) * gradc2f(h)
     *   energy :259        -interpc2f(κ) * gradc2f(T)
     *          :361-365    -interpc2f(κ) * gradc2f(T) - interpc2f(ρe_int_l K) * gradc2f(h)
     * Boundary faces 0 and n carry the SetValue fluxes.                                     */
    Fw[0] = fw_bot; Fw[n] = fw_top;
    Fe[0] = fe_bot; Fe[n] = fe_top;
    for (int j = 1; j < n; ++j) {
        double grad_h = (h[j] - h[j - 1]) / dz;
        double grad_T = (T[j] - T[j - 1]) / dz;
        Fw[j] = has_water(model) ? -((K[j - 1] + K[j]) / 2.0) * grad_h : 0.0;
        if (model == LH_MODEL_HEAT) Fe[j] = -((kappa[j - 1] + kappa[j]) / 2.0) * grad_T;
        else if (model == LH_MODEL_COUPLED)
            Fe[j] = -((kappa[j - 1] + kappa[j]) / 2.0) * grad_T - ((eK[j - 1] + eK[j]) / 2.0) * grad_h;
        else Fe[j] = 0.0;
    }
    for (int i = 0; i < n; ++i) {
        d_th[i] = has_water(model) ? -((Fw[i + 1] - Fw[i]) / dz) : 0.0;
        d_ti[i] = 0.0;                                              /* :182, :359 */
        d_re[i] = has_heat(model) ? -((Fe[i + 1] - Fe[i]) / dz) : 0.0;
    }
}

/* rhs over all columns: input state arrays (th, ti, re, T) -> c->tend[] */
static void rhs_all(lho_soil_ctx* c, const double* th, const double* ti, const double* re,
                    const double* T)
{
    const int n = c->nlayer;
#pragma omp parallel if (c->ncol >= 64)
    {
        double* work = (double*)malloc(sizeof(double) * 5 * n);
#pragma omp for schedule(static)
        for (int64_t col = 0; col < c->ncol; ++col) {
            size_t o = (size_t)col * n;
            const lh_soil_params pc = params_of_column(c, col);
            const double kappa_dry = lho_k_dry(&pc);                /* :214, :295 */
            double bcv[4];
            for (int k = 0; k < 4; ++k) bcv[k] = c->flux_cols[k] ? c->flux_cols[k][col] : c->bcv[k];
            column_rhs(c, col, &pc, bcv, kappa_dry, th + o, ti + o, re + o, T + o, c->tend[0] + o,
                       c->tend[1] + o, c->tend[2] + o, c->Fw + (size_t)col * (n + 1),
                       c->Fe + (size_t)col * (n + 1), work);
        }
        free(work);
    }
}

int32_t lho_soil_rhs(lho_soil_ctx* c, double t)
{
    (void)t;
    if (!c) return LH_ERR_INVALID_ARG;
    rhs_all(c, c->f[0], c->f[1], c->f[2], c->f[3]);
    return LH_OK;
}

int32_t lho_soil_face_fluxes(lho_soil_ctx* c, int64_t col, double* Fw_out, double* Fe_out)
{
    if (!c || col < 0 || col >= c->ncol) return LH_ERR_INVALID_ARG;
    size_t o = (size_t)col * (c->nlayer + 1);
    if (Fw_out) memcpy(Fw_out, c->Fw + o, sizeof(double) * (c->nlayer + 1));
    if (Fe_out) memcpy(Fe_out, c->Fe + o, sizeof(double) * (c->nlayer + 1));
    return LH_OK;
}

/* ------------------------------------------------------------------------------------------
 * SSPRK33 (OrdinaryDiffEq v5, Shu-Osher form).  The combine runs over every
 * prognostic field of the model's FieldVector, θ_i included, like the reference's broadcast.
 * ---------------------------------------------------------------------------------------- */

static int prognostic(int model, int field)
{
    if (model == LH_MODEL_RICHARDS) return field == 0 || field == 1;   /* (ϑ_l, θ_i)         */
    if (model == LH_MODEL_HEAT) return field == 2;                     /* (ρe_int)           */
    return 1;                                                          /* (ϑ_l, θ_i, ρe_int) */
}

/* update_aux! before a stage (right_hand_side.jl:54-81) from the rows the host evaluated ahead (lh_soil_set_aux_table). */
static int32_t apply_aux_tables(lho_soil_ctx* c)
{
    int any = 0;
    for (int f = 0; f < LH_NUM_FIELDS; ++f) {
        if (!c->aux_tab[f]) continue;
        if (c->aux_row >= c->aux_rows[f]) return fail(c, LH_ERR_STATE, "prescribed-profile table exhausted");
        const double* row = c->aux_tab[f] + c->aux_row * c->nlayer;
        for (int64_t col = 0; col < c->ncol; ++col)
            for (int i = 0; i < c->nlayer; ++i) c->f[f][col * c->nlayer + i] = row[i];
        any = 1;
    }
    if (any) ++c->aux_row;
    return LH_OK;
}

int32_t lho_soil_set_aux_table(lho_soil_ctx* c, int32_t field, const double* table, int64_t nrows)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (field < 0 || field >= LH_NUM_FIELDS) return fail(c, LH_ERR_INVALID_ARG, "bad field id");
    const int model = c->cfg.model;
    const int prescribed = (model == LH_MODEL_RICHARDS && field == LH_FIELD_T) ||
                           (model == LH_MODEL_HEAT && (field == LH_FIELD_THETA_L || field == LH_FIELD_THETA_I));
    if (!prescribed) return fail(c, LH_ERR_INVALID_ARG, "field is not a prescribed profile of this model");
    free(c->aux_tab[field]);
    c->aux_tab[field] = NULL;
    c->aux_rows[field] = 0;
    c->aux_row = 0;
    if (!table || nrows <= 0) return LH_OK;
    const size_t cnt = (size_t)nrows * c->nlayer;
    c->aux_tab[field] = (double*)malloc(cnt * sizeof(double));
    memcpy(c->aux_tab[field], table, cnt * sizeof(double));
    c->aux_rows[field] = nrows;
    return LH_OK;
}

int32_t lho_soil_stage_ssprk33(lho_soil_ctx* c, int32_t stage, double dt)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (stage < 1 || stage > 3) return fail(c, LH_ERR_INVALID_ARG, "stage must be 1, 2 or 3");
    { int32_t st_ = apply_aux_tables(c); if (st_ != LH_OK) return st_; }
    const int model = c->cfg.model;
    const size_t cells = (size_t)c->ncol * c->nlayer;
    /* stage input: u0 for stage 1, the stage buffer otherwise (non-prognostic slots always
     * come from f[]).                                                                       */
    const double* in[3];
    for (int k = 0; k < 3; ++k) in[k] = (stage == 1 || !prognostic(model, k)) ? c->f[k] : c->u1[k];
    rhs_all(c, in[0], in[1], in[2], c->f[3]);
    for (int k = 0; k < 3; ++k) {
        if (!prognostic(model, k)) continue;
        double* u0 = c->f[k];
        double* u = c->u1[k];
        const double* kk = c->tend[k];
        if (stage == 1) {
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
            for (size_t i = 0; i < cells; ++i) u[i] = u0[i] + dt * kk[i];
        } else if (stage == 2) {
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
            for (size_t i = 0; i < cells; ++i) u[i] = (3 * u0[i] + u[i] + dt * kk[i]) / 4;
        } else {
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
            for (size_t i = 0; i < cells; ++i) u0[i] = (u0[i] + 2 * u[i] + 2 * dt * kk[i]) / 3;
        }
    }
    return LH_OK;
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int32_t lho_soil_step_ssprk33(lho_soil_ctx* c, double t, double dt, int64_t nsteps,
                              const double* bc_table)
{
    (void)t;
    if (!c) return LH_ERR_INVALID_ARG;
    if (nsteps < 0) return fail(c, LH_ERR_INVALID_ARG, "nsteps < 0");
    double t0 = now_ms();
    for (int64_t s = 0; s < nsteps; ++s) {
        for (int stage = 1; stage <= 3; ++stage) {
            if (bc_table) memcpy(c->bcv, bc_table + (s * 3 + (stage - 1)) * 4, sizeof c->bcv);
            int32_t st = lho_soil_stage_ssprk33(c, stage, dt);
            if (st != LH_OK) return st;
        }
    }
    c->last_ms = now_ms() - t0;
    c->last_launches = 0;
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) {
        const size_t cells = (size_t)c->ncol * c->nlayer;
        for (int k = 0; k < 3; ++k)
            for (size_t i = 0; i < cells; ++i)
                if (!isfinite(c->f[k][i])) return fail(c, LH_ERR_NONFINITE, "non-finite state");
    }
    return LH_OK;
}

/* ------------------------------------------------------------------------------------------
 * Other explicit low-storage steppers (include/lh_soil.h, "other explicit steppers").  The
 * combine runs over every prognostic field, like the reference's broadcast over the FieldVector.
 *   Shu-Osher, two registers:  u_i = a[i] u^n + b[i] u_{i-1} + g[i] dt f(u_{i-1})
 *   Williamson 2N:             r = a[i] r + dt f(u);  u = u + b[i] r
 * Built-in tables: forward Euler; SSPRK22 and SSPRK33 (Shu & Osher 1988); SSPRK43 (Kraaijevanger
 * 1991 / Spiteri & Ruuth 2002, the 4-stage third-order SSP method with C = 2);
 * CarpenterKennedy2N54 (Carpenter & Kennedy 1994, NASA TM 109112, the rational coefficients of
 * its table for the (5,4) 2N-storage scheme).
 * ---------------------------------------------------------------------------------------- */
int32_t lho_soil_stepper_named(int32_t method, lh_soil_stepper* out)
{
    if (!out) return LH_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    switch (method) {
    case LH_METHOD_EULER:
        out->kind = LH_STEPPER_SHU_OSHER; out->nstages = 1;
        out->a[0] = 0.0; out->b[0] = 1.0; out->g[0] = 1.0; out->c[0] = 0.0;
        return LH_OK;
    case LH_METHOD_SSPRK22:
        out->kind = LH_STEPPER_SHU_OSHER; out->nstages = 2;
        out->a[0] = 0.0; out->b[0] = 1.0; out->g[0] = 1.0; out->c[0] = 0.0;
        out->a[1] = 0.5; out->b[1] = 0.5; out->g[1] = 0.5; out->c[1] = 1.0;
        return LH_OK;
    case LH_METHOD_SSPRK33:
        out->kind = LH_STEPPER_SHU_OSHER; out->nstages = 3;
        out->a[0] = 0.0; out->b[0] = 1.0; out->g[0] = 1.0; out->c[0] = 0.0;
        out->a[1] = 0.75; out->b[1] = 0.25; out->g[1] = 0.25; out->c[1] = 1.0;
        out->a[2] = 1.0 / 3.0; out->b[2] = 2.0 / 3.0; out->g[2] = 2.0 / 3.0; out->c[2] = 0.5;
        return LH_OK;
    case LH_METHOD_SSPRK43:
        out->kind = LH_STEPPER_SHU_OSHER; out->nstages = 4;
        out->a[0] = 0.0; out->b[0] = 1.0; out->g[0] = 0.5; out->c[0] = 0.0;
        out->a[1] = 0.0; out->b[1] = 1.0; out->g[1] = 0.5; out->c[1] = 0.5;
        out->a[2] = 2.0 / 3.0; out->b[2] = 1.0 / 3.0; out->g[2] = 1.0 / 6.0; out->c[2] = 1.0;
        out->a[3] = 0.0; out->b[3] = 1.0; out->g[3] = 0.5; out->c[3] = 0.5;
        return LH_OK;
    case LH_METHOD_CK2N54:
        out->kind = LH_STEPPER_2N; out->nstages = 5;
        out->a[0] = 0.0;
        out->a[1] = -567301805773.0 / 1357537059087.0;
        out->a[2] = -2404267990393.0 / 2016746695238.0;
        out->a[3] = -3550918686646.0 / 2091501179385.0;
        out->a[4] = -1275806237668.0 / 842570457699.0;
        out->b[0] = 1432997174477.0 / 9575080441755.0;
        out->b[1] = 5161836677717.0 / 13612068292357.0;
        out->b[2] = 1720146321549.0 / 2090206949498.0;
        out->b[3] = 3134564353537.0 / 4481467310338.0;
        out->b[4] = 2277821191437.0 / 14882151754819.0;
        out->c[0] = 0.0;
        out->c[1] = 1432997174477.0 / 9575080441755.0;
        out->c[2] = 2526269341429.0 / 6820363962896.0;
        out->c[3] = 2006345519317.0 / 3224310063776.0;
        out->c[4] = 2802321613138.0 / 2924317926251.0;
        return LH_OK;
    default:
        return LH_ERR_INVALID_ARG;
    }
}

int32_t lho_soil_step(lho_soil_ctx* c, const lh_soil_stepper* sp, double t, double dt, int64_t nsteps,
                      const double* bc_table)
{
    (void)t;
    if (!c || !sp) return LH_ERR_INVALID_ARG;
    if (nsteps < 0) return fail(c, LH_ERR_INVALID_ARG, "nsteps < 0");
    if (sp->nstages < 1 || sp->nstages > LH_MAX_STAGES) return fail(c, LH_ERR_INVALID_ARG, "bad stage count");
    if (sp->kind != LH_STEPPER_SHU_OSHER && sp->kind != LH_STEPPER_2N) return fail(c, LH_ERR_INVALID_ARG, "bad stepper kind");
    if (sp->kind == LH_STEPPER_2N && sp->a[0] != 0.0) return fail(c, LH_ERR_INVALID_ARG, "2N scheme needs a[0] == 0");
    const int model = c->cfg.model;
    const int ns = sp->nstages;
    const size_t cells = (size_t)c->ncol * c->nlayer;
    double t0 = now_ms();
    for (int64_t s = 0; s < nsteps; ++s) {
        for (int i = 0; i < ns; ++i) {
            if (bc_table) memcpy(c->bcv, bc_table + (s * ns + i) * 4, sizeof c->bcv);
            { int32_t st_ = apply_aux_tables(c); if (st_ != LH_OK) return st_; }
            if (sp->kind == LH_STEPPER_SHU_OSHER) {
                /* stage input: u^n for the first stage, the stage register otherwise; the last stage
                 * writes u^{n+1} over u^n                                                           */
                const double* in[3];
                for (int k = 0; k < 3; ++k) in[k] = (i == 0 || !prognostic(model, k)) ? c->f[k] : c->u1[k];
                rhs_all(c, in[0], in[1], in[2], c->f[3]);
                const double a = sp->a[i], b = sp->b[i], g = sp->g[i];
                for (int k = 0; k < 3; ++k) {
                    if (!prognostic(model, k)) continue;
                    double* u0 = c->f[k];
                    const double* v = in[k];
                    double* out = (i == ns - 1) ? c->f[k] : c->u1[k];
                    const double* kk = c->tend[k];
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
                    for (size_t j = 0; j < cells; ++j) out[j] = a * u0[j] + b * v[j] + (g * dt) * kk[j];
                }
            } else {
                rhs_all(c, c->f[0], c->f[1], c->f[2], c->f[3]);
                const double A = sp->a[i], B = sp->b[i];
                for (int k = 0; k < 3; ++k) {
                    if (!prognostic(model, k)) continue;
                    double* u = c->f[k];
                    double* r = c->u1[k];
                    const double* kk = c->tend[k];
#pragma omp parallel for schedule(static) if (c->ncol >= 64)
                    for (size_t j = 0; j < cells; ++j) {
                        const double rn = (i == 0) ? dt * kk[j] : A * r[j] + dt * kk[j];
                        r[j] = rn;
                        u[j] = u[j] + B * rn;
                    }
                }
            }
        }
    }
    c->last_ms = now_ms() - t0;
    c->last_launches = 0;
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) {
        for (int k = 0; k < 3; ++k)
            for (size_t i = 0; i < cells; ++i)
                if (!isfinite(c->f[k][i])) return fail(c, LH_ERR_NONFINITE, "non-finite state");
    }
    return LH_OK;
}

/* Budgets (new in this build): W = Σ ϑ_l dz, E = Σ ρe_int dz.  Neumaier-compensated
 * so the oracle's sum is accurate to ~1 ulp independent of order.                           */
static void comp_add(double* s, double* comp, double x)
{
    double t = *s + x;
    if (fabs(*s) >= fabs(x)) *comp += (*s - t) + x;
    else *comp += (x - t) + *s;
    *s = t;
}

int32_t lho_soil_budgets(lho_soil_ctx* c, double out[2])
{
    if (!c || !out) return LH_ERR_INVALID_ARG;
    const size_t cells = (size_t)c->ncol * c->nlayer;
    double s0 = 0, c0 = 0, s1 = 0, c1 = 0;
    for (size_t i = 0; i < cells; ++i) {
        comp_add(&s0, &c0, c->f[0][i]);
        comp_add(&s1, &c1, c->f[2][i]);
    }
    out[0] = (s0 + c0) * c->dz;
    out[1] = (s1 + c1) * c->dz;
    return LH_OK;
}

int32_t lho_soil_budgets_allreduce(lho_soil_ctx* c, double out[2]) { return lho_soil_budgets(c, out); }

/* lh_soil_run: the plain loop the device version overlaps. */
int32_t lho_soil_run(lho_soil_ctx* c, double t0, double dt, int64_t nsteps, const lh_soil_run_opts* o)
{
    if (!c || !o) return LH_ERR_INVALID_ARG;
    if (o->struct_size != (int32_t)sizeof(lh_soil_run_opts)) return fail(c, LH_ERR_INVALID_ARG, "lh_soil_run_opts.struct_size mismatch");
    if (nsteps < 0 || o->budget_every < 0 || o->save_every < 0) return fail(c, LH_ERR_INVALID_ARG, "negative step count or cadence");
    if (o->budget_every > 0 && !o->budgets_out) return fail(c, LH_ERR_INVALID_ARG, "budget_every > 0 needs budgets_out");
    const int saving = o->save_every > 0 || o->save_first;
    if (saving && (!o->save_out || o->nsave_fields < 1 || o->nsave_fields > LH_NUM_FIELDS)) return fail(c, LH_ERR_INVALID_ARG, "snapshots need save_out and fields");
    int64_t nsnap = 0, nbud = 0;
    int32_t st;
    if (o->save_first) {
        for (int k = 0; k < o->nsave_fields; ++k)
            if ((st = lho_soil_get_state(c, o->save_fields[k], o->save_out + nsnap * o->snapshot_stride + k * o->field_stride, o->col_stride, o->layer_stride)) != LH_OK) return st;
        ++nsnap;
    }
    for (int64_t s = 0; s < nsteps; ++s) {
        if ((st = lho_soil_step_ssprk33(c, t0 + s * dt, dt, 1, o->bc_table ? o->bc_table + s * 12 : NULL)) != LH_OK) return st;
        if (o->budget_every > 0 && (s + 1) % o->budget_every == 0) { lho_soil_budgets(c, o->budgets_out + 2 * nbud); ++nbud; }
        if (o->save_every > 0 && (s + 1) % o->save_every == 0) {
            for (int k = 0; k < o->nsave_fields; ++k)
                if ((st = lho_soil_get_state(c, o->save_fields[k], o->save_out + nsnap * o->snapshot_stride + k * o->field_stride, o->col_stride, o->layer_stride)) != LH_OK) return st;
            ++nsnap;
        }
    }
    return LH_OK;
}

/* Checkpoints of the oracle: its own layout, same entry points. */
int64_t lho_soil_checkpoint_bytes(const lho_soil_ctx* c)
{
    if (!c) return LH_ERR_INVALID_ARG;
    return (int64_t)(sizeof(int64_t) * 4 + sizeof(double) * 4 + (size_t)LH_NUM_FIELDS * c->ncol * c->nlayer * sizeof(double));
}

int32_t lho_soil_checkpoint_save(lho_soil_ctx* c, void* buf, int64_t cap)
{
    if (!c || !buf) return LH_ERR_INVALID_ARG;
    if (cap < lho_soil_checkpoint_bytes(c)) return fail(c, LH_ERR_INVALID_ARG, "checkpoint buffer too small");
    char* p = (char*)buf;
    int64_t h[4] = {0x4c484f43, c->ncol, c->nlayer, c->aux_row};
    memcpy(p, h, sizeof h); p += sizeof h;
    memcpy(p, c->bcv, sizeof c->bcv); p += sizeof c->bcv;
    const size_t fb = (size_t)c->ncol * c->nlayer * sizeof(double);
    for (int k = 0; k < LH_NUM_FIELDS; ++k) { memcpy(p, c->f[k], fb); p += fb; }
    return LH_OK;
}

int32_t lho_soil_checkpoint_load(lho_soil_ctx* c, const void* buf, int64_t bytes)
{
    if (!c || !buf) return LH_ERR_INVALID_ARG;
    if (bytes < lho_soil_checkpoint_bytes(c)) return fail(c, LH_ERR_INVALID_ARG, "checkpoint truncated");
    const char* p = (const char*)buf;
    int64_t h[4];
    memcpy(h, p, sizeof h); p += sizeof h;
    if (h[0] != 0x4c484f43 || h[1] != c->ncol || h[2] != c->nlayer) return fail(c, LH_ERR_INVALID_ARG, "checkpoint is of another problem");
    c->aux_row = h[3];
    memcpy(c->bcv, p, sizeof c->bcv); p += sizeof c->bcv;
    const size_t fb = (size_t)c->ncol * c->nlayer * sizeof(double);
    for (int k = 0; k < LH_NUM_FIELDS; ++k) { memcpy(c->f[k], p, fb); p += fb; }
    return LH_OK;
}

int32_t lho_soil_alloc_host(int64_t bytes, void** out)
{
    if (!out || bytes < 0) return LH_ERR_INVALID_ARG;
    *out = malloc((size_t)(bytes > 0 ? bytes : 1));
    return *out ? LH_OK : LH_ERR_INVALID_ARG;
}

int32_t lho_soil_free_host(void* p) { free(p); return LH_OK; }

/* The CPU has no stream: the "asynchronous" form computes at once and hands the result out on wait. */
int32_t lho_soil_budgets_async(lho_soil_ctx* c, int64_t* ticket_out)
{
    if (!c || !ticket_out) return LH_ERR_INVALID_ARG;
    if (c->next_ticket < 1) c->next_ticket = 1;
    const int64_t t = c->next_ticket;
    const int slot = (int)(t % 8);
    if (c->async_ticket[slot] != 0) return fail(c, LH_ERR_STATE, "lho_soil_budgets_async: 8 results outstanding");
    int32_t st = lho_soil_budgets(c, c->async_bud[slot]);
    if (st != LH_OK) return st;
    c->async_ticket[slot] = t;
    c->next_ticket = t + 1;
    *ticket_out = t;
    return LH_OK;
}

int32_t lho_soil_budgets_wait(lho_soil_ctx* c, int64_t ticket, double out[2])
{
    if (!c || !out) return LH_ERR_INVALID_ARG;
    const int slot = (int)(ticket % 8);
    if (ticket <= 0 || c->async_ticket[slot] != ticket) return fail(c, LH_ERR_STATE, "lho_soil_budgets_wait: unknown ticket");
    out[0] = c->async_bud[slot][0];
    out[1] = c->async_bud[slot][1];
    c->async_ticket[slot] = 0;
    return LH_OK;
}

int32_t lho_soil_diagnostic(lho_soil_ctx* c, int32_t which, double* host, int64_t cs, int64_t ls)
{
    if (!c || !host) return LH_ERR_INVALID_ARG;
    if (which < 0 || which >= LH_NUM_DIAGS) return fail(c, LH_ERR_INVALID_ARG, "bad diagnostic id");
    const int n = c->nlayer;
    const int model = c->cfg.model;
    /* K/ψ are defined whenever ϑ_l, θ_i exist (all models); κ/T from ρe_int need an energy
     * model, otherwise T is the prescribed aux and κ is evaluated with it.                  */
    for (int64_t col = 0; col < c->ncol; ++col) {
        const lh_soil_params pc = params_of_column(c, col);
        const double kappa_dry = lho_k_dry(&pc);
        for (int i = 0; i < n; ++i) {
            size_t o = (size_t)col * n + i;
            int m = has_heat(model) ? LH_MODEL_COUPLED : LH_MODEL_RICHARDS;
            const lh_soil_params pi = has_cell_params(c) ? params_of_cell(c, &pc, col, i) : pc;
            const double kd = has_cell_params(c) ? lho_k_dry(&pi) : kappa_dry;
            cell_closures cc = closures_at(&pi, m, kd, c->f[0][o], c->f[1][o],
                                           c->f[2][o], c->f[3][o]);
            if (which == LH_DIAG_KAPPA && !has_heat(model)) {
                cell_closures ch = closures_at(&pi, LH_MODEL_HEAT, kd, c->f[0][o],
                                               c->f[1][o], 0.0, c->f[3][o]);
                cc.kappa = ch.kappa;
            }
            double v = which == LH_DIAG_K ? cc.K : which == LH_DIAG_PSI ? cc.psi
                     : which == LH_DIAG_KAPPA ? cc.kappa : cc.T;
            host[col * cs + i * ls] = v;
        }
    }
    return LH_OK;
}

/* Per-column hydraulic parameters (include/lh_soil.h): each array ncol doubles, NULL keeps the model's value. */
int32_t lho_soil_set_column_params(lho_soil_ctx* c, const double* nu, const double* theta_r, const double* vg_n,
                                   const double* vg_alpha, const double* Ksat)
{
    if (!c) return LH_ERR_INVALID_ARG;
    const double* src[5] = {nu, theta_r, vg_n, vg_alpha, Ksat};
    for (int k = 0; k < 5; ++k) {
        free(c->colp[k]);
        c->colp[k] = NULL;
        if (src[k]) {
            c->colp[k] = (double*)malloc(sizeof(double) * (size_t)c->ncol);
            if (!c->colp[k]) return fail(c, LH_ERR_INVALID_ARG, "out of memory");
            memcpy(c->colp[k], src[k], sizeof(double) * (size_t)c->ncol);
        }
    }
    return LH_OK;
}

int32_t lho_soil_set_cell_params(lho_soil_ctx* c, const double* nu, const double* theta_r, const double* vg_n,
                                 const double* vg_alpha, const double* Ksat, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    const double* src[5] = {nu, theta_r, vg_n, vg_alpha, Ksat};
    const int n = c->nlayer;
    for (int k = 0; k < 5; ++k) {
        free(c->cellp[k]);
        c->cellp[k] = NULL;
        if (!src[k]) continue;
        c->cellp[k] = (double*)malloc(sizeof(double) * (size_t)c->ncol * n);
        for (int64_t col = 0; col < c->ncol; ++col)
            for (int i = 0; i < n; ++i) c->cellp[k][(size_t)col * n + i] = src[k][col * cs + (int64_t)i * ls];
    }
    return LH_OK;
}

int32_t lho_soil_set_column_heat_params(lho_soil_ctx* c, const double* rho_c_ds, const double* kappa_sat_unfrozen,
                                        const double* kappa_sat_frozen, const double* kappa_solid, const double* nu_ss_om,
                                        const double* nu_ss_quartz, const double* nu_ss_gravel)
{
    if (!c) return LH_ERR_INVALID_ARG;
    const double* src[7] = {rho_c_ds, kappa_sat_unfrozen, kappa_sat_frozen, kappa_solid, nu_ss_om, nu_ss_quartz, nu_ss_gravel};
    if (c->cfg.model == LH_MODEL_RICHARDS)
        for (int k = 0; k < 7; ++k) if (src[k]) return fail(c, LH_ERR_INVALID_ARG, "the Richards model has no energy equation");
    for (int k = 0; k < 7; ++k) {
        free(c->colp[5 + k]);
        c->colp[5 + k] = NULL;
        if (src[k]) {
            c->colp[5 + k] = (double*)malloc(sizeof(double) * (size_t)c->ncol);
            if (!c->colp[5 + k]) return fail(c, LH_ERR_INVALID_ARG, "out of memory");
            memcpy(c->colp[5 + k], src[k], sizeof(double) * (size_t)c->ncol);
        }
    }
    return LH_OK;
}

int32_t lho_soil_set_column_fluxes(lho_soil_ctx* c, const double* const values[4])
{
    if (!c) return LH_ERR_INVALID_ARG;
    const int kinds[4] = {c->cfg.top.energy_kind, c->cfg.top.hydrology_kind, c->cfg.bottom.energy_kind, c->cfg.bottom.hydrology_kind};
    for (int k = 0; k < 4; ++k)
        if (values && values[k] && kinds[k] != LH_BC_FLUX) return fail(c, LH_ERR_INVALID_ARG, "per-column fluxes need a face of kind LH_BC_FLUX");
    for (int k = 0; k < 4; ++k) {
        free(c->flux_cols[k]);
        c->flux_cols[k] = NULL;
        if (values && values[k]) {
            c->flux_cols[k] = (double*)malloc(sizeof(double) * (size_t)c->ncol);
            memcpy(c->flux_cols[k], values[k], sizeof(double) * (size_t)c->ncol);
        }
    }
    return LH_OK;
}

int32_t lho_soil_set_atmos_forcing(lho_soil_ctx* c, const lh_soil_atmos* a)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (!a) { c->atmos_on = 0; return LH_OK; }
    if (a->struct_size != (int32_t)sizeof(lh_soil_atmos)) return fail(c, LH_ERR_INVALID_ARG, "lh_soil_atmos.struct_size mismatch");
    if (c->cfg.model != LH_MODEL_COUPLED)     /* boundary_conditions.jl:103-112: both components must be prognostic */
        return fail(c, LH_ERR_UNSUPPORTED_BC, "PrescribedAtmosForcing needs SoilEnergyModel + SoilHydrologyModel");
    if (!(a->z_atm > 0.0) || !(a->rho_a_sfc > 0.0) || !(a->theta_scale > 0.0) || !(c->cfg.params.z_0m > 0.0) || !(c->cfg.params.z_0s > 0.0))
        return fail(c, LH_ERR_INVALID_ARG, "PrescribedAtmosForcing needs z_atm, rho_a_sfc, theta_scale, z_0m, z_0s > 0");
    c->atmos = *a;
    c->atmos_on = 1;
    return LH_OK;
}

int32_t lho_soil_atmos_fluxes(lho_soil_ctx* c, const double* th, const double* ti, const double* T, int64_t n,
                              double* heat, double* water)
{
    if (!c || !th || !ti || !T || !heat || !water || n < 0) return LH_ERR_INVALID_ARG;
    if (!c->atmos_on) return fail(c, LH_ERR_STATE, "lh_soil_set_atmos_forcing has not been called");
    for (int64_t i = 0; i < n; ++i) lho_turbulent_surface_fluxes(&c->cfg.params, &c->atmos, th[i], ti[i], T[i], &heat[i], &water[i]);
    return LH_OK;
}

int32_t lho_soil_sync(lho_soil_ctx* c) { (void)c; return LH_OK; }

int32_t lho_soil_kernel_info(lho_soil_ctx* c, char* buf, int64_t cap)
{
    if (!c || !buf || cap < 1) return LH_ERR_INVALID_ARG;
    snprintf(buf, (size_t)cap, "oracle: CPU restatement of the reference path (no kernels), %d OpenMP thread(s)", lho_soil_num_threads());
    return LH_OK;
}

/* libm versions of the functions the device library hand-writes */
int32_t lho_soil_eval_math(lho_soil_ctx* c, int32_t fn, const double* x, double* y, int64_t n)
{
    (void)c;
    if (!x || !y || n < 0) return LH_ERR_INVALID_ARG;
    for (int64_t i = 0; i < n; ++i) {
        switch (fn) {
        case LH_MATH_LOG2: y[i] = log2(x[i]); break;
        case LH_MATH_EXP2: y[i] = exp2(x[i]); break;
        case LH_MATH_EXP2M1: y[i] = exp2(x[i]) - 1.0; break;
        case LH_MATH_SQRT: y[i] = sqrt(x[i]); break;
        case LH_MATH_RSQRT: y[i] = 1.0 / sqrt(x[i]); break;
        case LH_MATH_RCP: case LH_MATH_RCP_SEED: y[i] = 1.0 / x[i]; break;
        case LH_MATH_RSQRT_SEED: y[i] = 1.0 / sqrt(x[i]); break;
        case LH_MATH_DIV: y[i] = x[i] / x[n + i]; break;
        default: return LH_ERR_INVALID_ARG;
        }
    }
    return LH_OK;
}

int32_t lho_soil_last_step_timing(lho_soil_ctx* c, double* ms_out, int64_t* launches_out)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (ms_out) *ms_out = c->last_ms;
    if (launches_out) *launches_out = c->last_launches;
    return LH_OK;
}

int32_t lho_soil_device_ptr(lho_soil_ctx* c, int32_t field, void** dptr, int64_t* ncol_padded)
{
    (void)field; (void)dptr; (void)ncol_padded;
    return fail(c, LH_ERR_NO_DEVICE, "the oracle has no device memory");
}

int32_t lho_soil_comm_unique_id(uint8_t id_out[128]) { memset(id_out, 0, 128); return LH_OK; }

int32_t lho_soil_comm_init(lho_soil_ctx* c, int32_t nranks, int32_t rank, const uint8_t id[128])
{
    (void)c; (void)nranks; (void)rank; (void)id;
    return LH_OK;
}
