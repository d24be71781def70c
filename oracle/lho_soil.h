/*
 * lho_soil.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (the parity oracle) of the soil RHS + SSPRK33 path of
 * CliMA/LandHydrology.jl.  It exports the SAME C ABI as include/lh_soil.h under the
 * prefix `lho_` (so one ctypes harness drives both libraries) plus the reference's scalar
 * parameterisation functions for closure-level tests.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library — as the checker, never as the product.  The product
 * (landhydrology.jl_b200) never imports, links or executes anything under oracle/.
 *
 * Pinning: the reference cannot run here (no Julia).  The oracle is pinned against the
 * reference's own known-answer tests (tests/test_oracle_pins.py cites each one); the ulp-level
 * evaluation order inside the un-vendored ClimaCore operators and OrdinaryDiffEq's
 * `@muladd` stage combine is NOT pinned by any reference test ("parity unpinned" at the ulp
 * level; pinned at the reference's own test tolerances).
 */
#ifndef LHO_SOIL_H
#define LHO_SOIL_H

#define lh_soil_ctx               lho_soil_ctx
#define lh_soil_abi_version       lho_soil_abi_version
#define lh_soil_create            lho_soil_create
#define lh_soil_destroy           lho_soil_destroy
#define lh_soil_last_error        lho_soil_last_error
#define lh_soil_get_zc            lho_soil_get_zc
#define lh_soil_set_state         lho_soil_set_state
#define lh_soil_get_state         lho_soil_get_state
#define lh_soil_set_aux           lho_soil_set_aux
#define lh_soil_set_bc_values     lho_soil_set_bc_values
#define lh_soil_rhs               lho_soil_rhs
#define lh_soil_get_tendency      lho_soil_get_tendency
#define lh_soil_stage_ssprk33     lho_soil_stage_ssprk33
#define lh_soil_step_ssprk33      lho_soil_step_ssprk33
#define lh_soil_set_column_params lho_soil_set_column_params
#define lh_soil_set_column_heat_params lho_soil_set_column_heat_params
#define lh_soil_set_column_fluxes  lho_soil_set_column_fluxes
#define lh_soil_set_cell_params    lho_soil_set_cell_params
#define lh_soil_set_atmos_forcing  lho_soil_set_atmos_forcing
#define lh_soil_atmos_fluxes       lho_soil_atmos_fluxes
#define lh_soil_stepper_named     lho_soil_stepper_named
#define lh_soil_step              lho_soil_step
#define lh_soil_budgets           lho_soil_budgets
#define lh_soil_diagnostic        lho_soil_diagnostic
#define lh_soil_sync              lho_soil_sync
#define lh_soil_eval_math         lho_soil_eval_math
#define lh_soil_last_step_timing  lho_soil_last_step_timing
#define lh_soil_device_ptr        lho_soil_device_ptr
#define lh_soil_comm_unique_id    lho_soil_comm_unique_id
#define lh_soil_comm_init         lho_soil_comm_init
#define lh_soil_budgets_allreduce lho_soil_budgets_allreduce
#define lh_soil_kernel_info       lho_soil_kernel_info
#define lh_soil_budgets_async     lho_soil_budgets_async
#define lh_soil_budgets_wait      lho_soil_budgets_wait
#define lh_soil_set_aux_table     lho_soil_set_aux_table
#define lh_soil_run               lho_soil_run
#define lh_soil_checkpoint_bytes  lho_soil_checkpoint_bytes
#define lh_soil_checkpoint_save   lho_soil_checkpoint_save
#define lh_soil_checkpoint_load   lho_soil_checkpoint_load
#define lh_soil_alloc_host        lho_soil_alloc_host
#define lh_soil_free_host         lho_soil_free_host

#include "../include/lh_soil.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Scalar closures, literal restatements (expression order kept) of
 * src/SoilModel/SoilWaterParameterizations.jl and SoilHeatParameterizations.jl.            */
double lho_volumetric_liquid_fraction(double theta_l_aug, double nu_eff);
double lho_effective_saturation(double porosity, double theta_l_aug, double theta_r);
double lho_matric_potential(const lh_soil_params* p, double S);
double lho_inverse_matric_potential(const lh_soil_params* p, double psi); /* NaN for psi > 0 */
double lho_pressure_head(const lh_soil_params* p, double theta_l_aug, double nu_eff, double S_s);
double lho_hydraulic_conductivity(const lh_soil_params* p, double S, double visc_f, double imp_f);
double lho_viscosity_factor(const lh_soil_params* p, double T);
double lho_impedance_factor(const lh_soil_params* p, double f_i);
double lho_hydrostatic_profile(const lh_soil_params* p, double z, double z_interface,
                               double nu, double S_s);
double lho_volumetric_heat_capacity(const lh_soil_params* p, double theta_l, double theta_i,
                                    double rho_c_ds);
double lho_temperature_from_rho_e_int(const lh_soil_params* p, double rho_e_int, double theta_i,
                                      double rho_c_s);
double lho_volumetric_internal_energy(const lh_soil_params* p, double theta_i, double rho_c_s,
                                      double T);
double lho_saturated_thermal_conductivity(double theta_l, double theta_i, double k_unfrozen,
                                          double k_frozen);
double lho_relative_saturation(double theta_l, double theta_i, double porosity);
double lho_kersten_number(const lh_soil_params* p, double theta_i, double S_r);
double lho_thermal_conductivity(double kappa_dry, double K_e, double kappa_sat);
double lho_volumetric_internal_energy_liq(const lh_soil_params* p, double T);
double lho_k_solid(double nu_ss_om, double nu_ss_quartz, double k_quartz, double k_minerals,
                   double k_om);
double lho_ksat_frozen(double k_solid, double porosity, double k_ice);
double lho_ksat_unfrozen(double k_solid, double porosity, double k_liq);
double lho_k_dry(const lh_soil_params* p);
/* PrescribedAtmosForcing pieces (parity unpinned, include/lh_soil.h) */
double lho_q_vap_saturation_liquid(const lh_soil_params* p, const lh_soil_atmos* a, double T, double rho);
void lho_turbulent_surface_fluxes(const lh_soil_params* p, const lh_soil_atmos* a, double theta_l_aug, double theta_i,
                                  double T, double* heat_flux, double* water_flux);

/* Face fluxes of the last rhs call for ONE column (nlayer+1 values each, bottom face first):
 * the parity tests scale their tolerance by max|flux|.                         */
int32_t lho_soil_face_fluxes(lho_soil_ctx* ctx, int64_t col, double* Fw_out, double* Fe_out);
/* Number of OpenMP threads the oracle will use (cpu_baseline.cores).                        */
int32_t lho_soil_num_threads(void);
void    lho_soil_set_num_threads(int32_t n);

#ifdef __cplusplus
}
#endif
#endif
