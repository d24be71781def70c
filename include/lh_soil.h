/*
 * lh_soil.h — C ABI of the B200-native soil right-hand-side + SSPRK33 stage path.
 *
 * This is the drop-in boundary for ONE hot path of CliMA/LandHydrology.jl: the closure
 * `rhs!(dY, Y, Ya, t)` returned by `make_rhs(model::SoilModel)`
 * (reference src/SoilModel/right_hand_side.jl:33-44) and the `step!` / `run!` loop that
 * `Simulation` drives (reference src/Simulations/simulation.jl:34-87, SSPRK33 stage
 * combine of OrdinaryDiffEq v5).  The reference has no FFI of its own (it is 100 % Julia),
 * so every entry point below names the Julia interface it replaces.  A Julia host binds
 * them with `ccall` (julia/LandHydrologyB200.jl); the tested twin is the
 * Python/ctypes host in landhydrology.jl_b200/.
 *
 * Conventions
 *   - plain C: pointers, sizes, doubles; no C++/torch types cross this boundary.
 *   - every function returns an int32 status: LH_OK (0) or a negative LH_ERR_* code; nothing
 *     throws or aborts.  `lh_soil_last_error(ctx)` gives the message (ctx == NULL: the message
 *     of the last failed `lh_soil_create` on this thread).
 *   - a ctx owns its device memory and one CUDA stream; calls on one ctx must be serialised
 *     by the caller (the reference's rhs! is not re-entrant either); different ctxs are
 *     independent.  Host pointers are borrowed for the duration of the call only.
 *   - arithmetic is fp64 throughout.  Vertical index 0 is the BOTTOM cell, nlayer-1 the top
 *     (reference boundary_conditions.jl:182-185).
 *   - host arrays are addressed as host[col * col_stride + layer * layer_stride] (strides in
 *     elements), so the reference's per-column `parent(Y.soil)` matrix (layer fastest, one
 *     field after the other) and a column-fastest SoA block are both expressible.
 */
#ifndef LH_SOIL_H
#define LH_SOIL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LH_SOIL_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------- */
#define LH_OK                   0
#define LH_ERR_INVALID_ARG     -1  /* NULL pointer, bad field id, bad stride, bad stage ...            */
#define LH_ERR_DOMAIN          -2  /* zlim[1] < zlim[2] violated (reference Domains/domain.jl:30)      */
#define LH_ERR_UNSUPPORTED_BC  -3  /* BC/component pair with no `vertical_flux` method in the reference
                                      (a Julia MethodError, boundary_conditions.jl:295-444)            */
#define LH_ERR_CUDA            -4  /* CUDA runtime failure (message has the CUDA error string)         */
#define LH_ERR_NO_DEVICE       -5  /* no CUDA device / bad ordinal: this library has NO CPU fallback   */
#define LH_ERR_NCCL            -6  /* NCCL missing or failed                                           */
#define LH_ERR_NONFINITE       -7  /* NaN/Inf in the state (the reference raises DomainError from `^`,
                                      SoilWaterParameterizations.jl:209-211)                           */
#define LH_ERR_STATE           -8  /* call sequence error (e.g. comm used before comm_init)            */

/* ---- enums -------------------------------------------------------------------------- */
/* Which (energy, hydrology) component pair `make_rhs` dispatches on (right_hand_side.jl).  */
#define LH_MODEL_RICHARDS 0   /* PrescribedTemperatureModel + SoilHydrologyModel  (:118-186) */
#define LH_MODEL_HEAT     1   /* SoilEnergyModel + PrescribedHydrologyModel       (:192-263) */
#define LH_MODEL_COUPLED  2   /* SoilEnergyModel + SoilHydrologyModel             (:269-369) */

/* Boundary-condition kinds (boundary_conditions.jl:27,43-46,61-64,77).                      */
#define LH_BC_NONE          0   /* NoBC          */
#define LH_BC_FLUX          1   /* VerticalFlux  */
#define LH_BC_DIRICHLET     2   /* Dirichlet     */
#define LH_BC_FREE_DRAINAGE 3   /* FreeDrainage  */
/* PrescribedAtmosForcing is not a per-component kind: it replaces the whole top SoilComponentBC (lh_soil_set_atmos_forcing). */

/* Conductivity factors (SoilWaterParameterizations.jl:38-65).                               */
#define LH_FACTOR_NONE        0 /* NoEffect                       */
#define LH_FACTOR_VISCOSITY   1 /* TemperatureDependentViscosity  */
#define LH_FACTOR_IMPEDANCE   1 /* IceImpedance                   */

/* Cell-centred fields.  theta_l is the reference's augmented liquid fraction ϑ_l.           */
#define LH_FIELD_THETA_L   0    /* ϑ_l     prognostic (Richards, coupled) / prescribed (heat) */
#define LH_FIELD_THETA_I   1    /* θ_i     tendency identically 0 (right_hand_side.jl:182,359) */
#define LH_FIELD_RHO_E_INT 2    /* ρe_int  prognostic (heat, coupled)                          */
#define LH_FIELD_T         3    /* T       prescribed aux (Richards only; Ya.soil.T)           */
#define LH_NUM_FIELDS      4

/* Index into the 4-vector of boundary values: [top energy, top hydrology, bottom energy,
 * bottom hydrology].  For LH_BC_FLUX the value is the flux (positive along +z at BOTH faces,
 * boundary_conditions.jl:44); for LH_BC_DIRICHLET it is `state_value(t)` (T or ϑ_l).        */
#define LH_BCV_TOP_ENERGY       0
#define LH_BCV_TOP_HYDROLOGY    1
#define LH_BCV_BOTTOM_ENERGY    2
#define LH_BCV_BOTTOM_HYDROLOGY 3

/* ---- parameter block ------------------------------------------------------------------
 * All constants are runtime doubles supplied by the host; nothing is hard-coded in kernels.
 *   SoilParams{FT}         reference src/SoilModel/parameters.jl:11-43
 *   vanGenuchten{FT}       reference SoilWaterParameterizations.jl:151-170
 *   conductivity factors   reference SoilWaterParameterizations.jl:46-65
 *   CLIMAParameters        values read at SoilHeatParameterizations.jl:12-13
 */
typedef struct lh_soil_params {
    /* SoilParams */
    double nu;                 /* ν   porosity                                  */
    double S_s;                /* specific storage                              */
    double nu_ss_gravel;
    double nu_ss_om;
    double nu_ss_quartz;
    double rho_c_ds;           /* ρc_ds volumetric heat capacity of dry soil    */
    double kappa_solid;
    double rho_p;              /* ρp particle density                           */
    double kappa_sat_unfrozen;
    double kappa_sat_frozen;
    double a;                  /* Balland & Arp                                 */
    double b;
    double kappa_dry_parameter;
    double z_0m;               /* carried for API completeness; unused on this path */
    double z_0s;
    /* vanGenuchten */
    double vg_n;
    double vg_alpha;
    double vg_m;               /* 1 - 1/n, as stored by the reference constructor (anything else: LH_ERR_INVALID_ARG) */
    double theta_r;
    double Ksat;
    /* conductivity factors */
    int32_t viscosity_factor;  /* LH_FACTOR_NONE | LH_FACTOR_VISCOSITY */
    int32_t impedance_factor;  /* LH_FACTOR_NONE | LH_FACTOR_IMPEDANCE */
    double visc_gamma;         /* γ     (default 2.64e-2) */
    double visc_T_ref;         /* T_ref (default 288)     */
    double imp_Omega;          /* Ω     (default 7)       */
    /* earth parameter set */
    double rho_cloud_liq;
    double rho_cloud_ice;
    double cp_l;
    double cp_i;
    double T_0;
    double LH_f0;
    double K_therm;
} lh_soil_params;

/* One face of SoilColumnBC (boundary_conditions.jl:95-101,144-161). */
typedef struct lh_soil_face_bc {
    int32_t energy_kind;       /* LH_BC_*                                       */
    int32_t hydrology_kind;    /* LH_BC_*                                       */
    double  energy_value;      /* flux, or Dirichlet T at create time           */
    double  hydrology_value;   /* flux, or Dirichlet ϑ_l at create time         */
} lh_soil_face_bc;

typedef struct lh_soil_config {
    int32_t struct_size;       /* sizeof(lh_soil_config): ABI guard                         */
    int32_t device;            /* CUDA ordinal                                              */
    int64_t ncol;              /* laterally independent columns (Column: 1; HybridBox: nx*ny) */
    int32_t nlayer;            /* Column.nelements                                          */
    int32_t model;             /* LH_MODEL_*                                                */
    double  zmin, zmax;        /* Column.zlim                                               */
    lh_soil_params  params;
    lh_soil_face_bc top;
    lh_soil_face_bc bottom;
    int32_t flags;             /* LH_FLAG_*                                                 */
    int32_t reserved;
} lh_soil_config;

#define LH_FLAG_CHECK_FINITE 1 /* rhs/step return LH_ERR_NONFINITE when NaN/Inf appears     */
#define LH_FLAG_GENERAL_VG   2 /* never use the van Genuchten n == 2 (m == 1/2) square-root
                                  specialisation: always evaluate the general-n log/exp form   */
/* lh_soil_step_ssprk33 launch strategy.  Default: one launch per stage (chained with programmatic dependent
 * launch), except for small, launch-bound grids (<= ~1.5 waves of resident blocks) where ONE persistent launch
 * runs all 3 nsteps stages, every block keeping its columns (L2-resident stage registers, no launch gaps).
 * Results are bit-identical either way.                                                          */
#define LH_FLAG_STAGE_LAUNCHES 4 /* always one launch per stage                                  */
#define LH_FLAG_PERSISTENT     8 /* always the persistent launch                                 */
/* Consecutive stage launches are chained block to block (block j of a stage starts as soon as block j of the previous
 * stage has published its results, instead of after the whole previous grid): the launch tail is paid once per call,
 * not once per stage.  Results are bit-identical.  This flag restores the whole-grid dependency (for measurements). */
#define LH_FLAG_NO_CHAIN      16

typedef struct lh_soil_ctx lh_soil_ctx;

/* ---- lifetime ------------------------------------------------------------------------ */
int32_t lh_soil_abi_version(void);

/* Replaces: Column(...), SoilParams{FT}(...), vanGenuchten{FT}(...), SoilColumnBC(...),
 * SoilModel(FT; ...) (models.jl:115-135) + make_function_space (domain.jl:58-69).
 * Validates like the reference's method table: an unsupported BC/component pair is
 * LH_ERR_UNSUPPORTED_BC, zmin >= zmax is LH_ERR_DOMAIN.                                     */
int32_t lh_soil_create(const lh_soil_config* cfg, lh_soil_ctx** out);
int32_t lh_soil_destroy(lh_soil_ctx* ctx);
const char* lh_soil_last_error(const lh_soil_ctx* ctx);

/* Cell-centre coordinates Ya.zc (right_hand_side.jl:7-8): writes nlayer doubles.            */
int32_t lh_soil_get_zc(const lh_soil_ctx* ctx, double* zc_out);

/* ---- state / aux transfer (replaces initialize_states, initial_conditions.jl:101-107) - */
/* set_state borrows `host` for the call only: it returns once the last byte has left the host buffer.  The layout transform of
 * the last blocks may still be queued on the ctx stream then; every later call on this ctx is ordered behind it.  (Only a
 * caller that reads the raw pointer of lh_soil_device_ptr on a stream of its own must call lh_soil_sync first.)            */
int32_t lh_soil_set_state(lh_soil_ctx* ctx, int32_t field, const double* host,
                          int64_t col_stride, int64_t layer_stride);
int32_t lh_soil_get_state(lh_soil_ctx* ctx, int32_t field, double* host,
                          int64_t col_stride, int64_t layer_stride);
/* Prescribed profiles (make_update_aux, right_hand_side.jl:54-81), evaluated by the host.
 * col_stride == 0 broadcasts one nlayer-long profile to every column.  Same storage as
 * set_state; kept as its own entry point because the reference keeps Y and Ya apart.        */
int32_t lh_soil_set_aux(lh_soil_ctx* ctx, int32_t field, const double* host,
                        int64_t col_stride, int64_t layer_stride);

/* Heterogeneous soils: per-column hydraulic parameters (new; the reference's SoilParams / vanGenuchten are one
 * set of scalars per model, parameters.jl:11-43, SoilWaterParameterizations.jl:151-170 — a land model over many
 * columns needs them per column).  Each array holds ncol doubles; NULL keeps the model's scalar for that parameter;
 * all NULL returns to the homogeneous kernels.  m = 1 - 1/n per column, as the reference constructor computes it;
 * κ_dry follows ν per column (k_dry, SoilHeatParameterizations.jl:280-294).  The kernels then read these values per
 * lane (one load per column and launch, +96 B per column, nothing per cell).                                    */
int32_t lh_soil_set_column_params(lh_soil_ctx* ctx, const double* nu, const double* theta_r,
                                  const double* vg_n, const double* vg_alpha, const double* Ksat);

/* Per-column HEAT parameters (new, same reason): ρc_ds, κ_sat_unfrozen, κ_sat_frozen, κ_solid and the solid fractions
 * ν_ss_om, ν_ss_quartz, ν_ss_gravel of SoilParams (parameters.jl:11-43).  κ_dry follows κ_solid and ν per column
 * (k_dry, SoilHeatParameterizations.jl:280-294), the Kersten exponents follow the fractions (:152-174).  Each array ncol
 * doubles, NULL keeps the model's scalar; combines with lh_soil_set_column_params (either call keeps what the other set).
 * Models without an energy equation return LH_ERR_INVALID_ARG.                                                    */
int32_t lh_soil_set_column_heat_params(lh_soil_ctx* ctx, const double* rho_c_ds, const double* kappa_sat_unfrozen,
                                       const double* kappa_sat_frozen, const double* kappa_solid, const double* nu_ss_om,
                                       const double* nu_ss_quartz, const double* nu_ss_gravel);

/* Layered soils: per-CELL hydraulic parameters (new): ν, θr, van Genuchten n and α, K_sat as fields over (column, layer),
 * addressed host[col * col_stride + layer * layer_stride] like a state field; NULL keeps the per-column value (or the
 * model's scalar) for that parameter, all NULL removes the fields.  The kernels then re-read nine derived values per cell and
 * stage (+72 B per cell and stage on top of the state traffic: the path becomes HBM-heavier).  κ_dry follows
 * ν cell by cell (k_dry, SoilHeatParameterizations.jl:280-294).  Heat parameters stay per column.                      */
int32_t lh_soil_set_cell_params(lh_soil_ctx* ctx, const double* nu, const double* theta_r, const double* vg_n,
                                const double* vg_alpha, const double* Ksat, int64_t col_stride, int64_t layer_stride);

/* Spatially varying prescribed fluxes (new): per-column values for the faces whose kind is LH_BC_FLUX (VerticalFlux,
 * boundary_conditions.jl:43-46,295-301), indexed like the boundary-value 4-vector: values[LH_BCV_*] is an array of ncol
 * doubles or NULL (the scalar of lh_soil_set_bc_values / the bc table applies).  E.g. a precipitation or ground heat flux
 * field over the columns of a HybridBox.  All NULL removes the arrays.                                                */
int32_t lh_soil_set_column_fluxes(lh_soil_ctx* ctx, const double* const values[4]);

/* ---- PrescribedAtmosForcing (boundary_conditions.jl:103-131, 516-620) -------------------------------------------------
 * The reference's only physically driven surface condition (experiments/SoilModel/surface_fluxes.jl): turbulent fluxes of
 * energy and water volume between the top soil cell and a prescribed atmospheric state, from Monin-Obukhov similarity.
 * PARITY UNPINNED: the arithmetic of `surface_conditions` (SurfaceFluxes v0.1) and `q_vap_saturation_generic` / `cp_m`
 * (Thermodynamics v0.5) is not under the reference tree.  What the reference itself fixes is restated exactly
 * (compute_turbulent_surface_fluxes :555-620): the pore-air humidity q_surf = q_sat(T, ρ_a) exp(g ψ / (R_v T)) with ψ the
 * matric potential at min(S_l_eff, 1); E = -ρ_a u* q*; the dry / vapour static-energy fluxes; Ẽ = E / ρ_l.  The similarity
 * solution (u*, θ*, q*) follows the published formulation: Businger-Dyer universal functions (Businger et al. 1971, Dyer
 * 1974) with Pr_0, a_m, a_h below, the Obukhov length from the θ flux alone, solved per column to 1e-15 by a secant
 * iteration on 1/L.  Saturation vapour pressure: Clausius-Clapeyron with constant Δcp = cp_v - cp_l
 * (Romps 2008; the closed form Thermodynamics.jl documents).  Its structural tests are the reference's own
 * (test_prescribed_atmos_bc.jl:75-79,155,161-194).
 * Valid for LH_MODEL_COUPLED only (both components dynamic, :103-112), top face only (:525-527): otherwise
 * LH_ERR_UNSUPPORTED_BC.  NULL restores the configured top boundary conditions.  May be called between steps with new
 * atmospheric values.  z_0m, z_0s come from lh_soil_params.                                                          */
typedef struct lh_soil_atmos {
    int32_t struct_size;       /* sizeof(lh_soil_atmos)                                       */
    int32_t reserved;
    double u_atm;              /* wind speed at z_atm                                         */
    double theta_atm;          /* potential temperature at z_atm                              */
    double z_atm;
    double theta_scale;
    double rho_a_sfc;          /* moist air density at the surface                            */
    double q_atm;              /* specific humidity at z_atm                                  */
    /* CLIMAParameters.Planet / SubgridScale values the reference reads at boundary_conditions.jl:575-617 */
    double R_v, R_d, grav, cp_d, cp_v, LH_v0, press_triple, T_triple, von_karman;
    /* Businger universal functions */
    double Pr_0, a_m, a_h;
} lh_soil_atmos;
int32_t lh_soil_set_atmos_forcing(lh_soil_ctx* ctx, const lh_soil_atmos* atmos);
/* The fluxes compute_turbulent_surface_fluxes returns (:555-620) for ONE surface state, evaluated on the device with the
 * ctx's parameters: out[0] = heat flux (W/m^2, positive upward), out[1] = Ẽ (m/s).  n states.                         */
int32_t lh_soil_atmos_fluxes(lh_soil_ctx* ctx, const double* theta_l, const double* theta_i, const double* T, int64_t n,
                             double* heat_flux_out, double* water_flux_out);

/* Boundary values for the NEXT rhs/stage call: the host evaluates Dirichlet
 * `state_value(t)` closures (boundary_conditions.jl:247,267) and passes 4 doubles indexed
 * by LH_BCV_*.                                                                              */
int32_t lh_soil_set_bc_values(lh_soil_ctx* ctx, const double values[4]);

/* ---- the hot path -------------------------------------------------------------------- */
/* Replaces rhs!(dY, Y, Ya, t) (right_hand_side.jl:37-42): tendency of the current state into
 * the ctx's tendency buffers; fetch with lh_soil_get_tendency.  This is the 1e-12 parity
 * entry point.                                                                              */
int32_t lh_soil_rhs(lh_soil_ctx* ctx, double t);
int32_t lh_soil_get_tendency(lh_soil_ctx* ctx, int32_t field, double* host,
                             int64_t col_stride, int64_t layer_stride);

/* One fused RHS + SSPRK33 stage (stage = 1, 2, 3), using the current bc values / aux:
 *   1: u1 = u0 + dt f(u0)      2: u2 = (3 u0 + u1 + dt f(u1)) / 4
 *   3: u  = (u0 + 2 u2 + 2 dt f(u2)) / 3
 * Replaces one third of OrdinaryDiffEq's SSPRK33 perform_step! driven by step!
 * (simulation.jl:79-80).  Stage times are t, t+dt, t+dt/2.                                  */
int32_t lh_soil_stage_ssprk33(lh_soil_ctx* ctx, int32_t stage, double dt);

/* nsteps full SSPRK33 steps.  bc_table is NULL (boundary values constant = the current
 * ones) or nsteps*3*4 doubles: for each step and stage the LH_BCV_* 4-vector evaluated by
 * the host at that stage's time.  Replaces step!/run! (simulation.jl:79-87).                */
int32_t lh_soil_step_ssprk33(lh_soil_ctx* ctx, double t, double dt, int64_t nsteps,
                             const double* bc_table);

/* Time-dependent prescribed profiles (make_update_aux, right_hand_side.jl:54-81: T(z,t) of PrescribedTemperatureModel,
 * ϑ_l(z,t) / θ_i(z,t) of PrescribedHydrologyModel) without a host round trip per stage: the host evaluates the profile
 * closures AHEAD for the stage times of the coming steps and uploads them as `nrows` rows of nlayer doubles; every stage
 * launch of lh_soil_stage_ssprk33 / lh_soil_step_ssprk33 / lh_soil_step / lh_soil_run then consumes the next row
 * (row r = r-th stage launch after this call; SSPRK33: step s, stage i -> row 3 s + i, times t, t + dt, t + dt/2),
 * broadcasting it to every column on the device before the launch.  table == NULL removes the table (the field keeps
 * its last values).  Running out of rows is LH_ERR_STATE.                                                            */
int32_t lh_soil_set_aux_table(lh_soil_ctx* ctx, int32_t field, const double* table, int64_t nrows);

/* ---- run!(simulation) with saveat and per-step diagnostics in ONE call ------------------------------------------
 * Replaces the solve! loop behind run! (simulation.jl:86-87) together with the `saveat` / callback keywords every
 * reference test passes (simulation.jl:64-70; richards_equation.jl:66-78, coupled.jl:94-98): nsteps SSPRK33 steps,
 * the budgets every `budget_every` steps and a snapshot of the listed fields every `save_every` steps (plus the initial
 * state with save_first, DiffEq's save_start).  Snapshots are copied device-to-device on the compute stream, which goes
 * on stepping at once; layout transform and PCIe transfer run on the ctx's copy stream, two snapshots deep, so a
 * snapshot costs the run nothing unless PCIe is the bottleneck.  Budgets arrive through a pinned ring, 16 bytes per
 * budget point.  save_out should be pinned host memory (lh_soil_alloc_host) for the copies to be asynchronous.
 * Snapshot s, field save_fields[k], column c, layer l lands at
 *     save_out[s * snapshot_stride + k * field_stride + c * col_stride + l * layer_stride]
 * with (col_stride, layer_stride) = (nlayer, 1) (the reference's parent(Y.soil) layout) or (1, >= ncol).
 * The call returns when every snapshot and budget has arrived.                                                       */
typedef struct lh_soil_run_opts {
    int32_t struct_size;          /* sizeof(lh_soil_run_opts)                                        */
    int32_t save_first;           /* != 0: snapshot 0 is the state before the first step             */
    const double* bc_table;       /* NULL or nsteps*3*4 boundary values (as lh_soil_step_ssprk33)    */
    int64_t budget_every;         /* 0: no budgets; k: after steps k, 2k, ...                        */
    double* budgets_out;          /* [nsteps / budget_every][2]                                      */
    int64_t save_every;           /* 0: no periodic snapshots; k: after steps k, 2k, ...             */
    int32_t nsave_fields;
    int32_t save_fields[LH_NUM_FIELDS];
    int32_t reserved;
    double* save_out;
    int64_t snapshot_stride, field_stride, col_stride, layer_stride;   /* in elements                */
} lh_soil_run_opts;
int32_t lh_soil_run(lh_soil_ctx* ctx, double t0, double dt, int64_t nsteps, const lh_soil_run_opts* opts);

/* Checkpoint / restart: everything a ctx needs to continue a run bit for bit (fields in device layout, current boundary
 * values, position in the prescribed-profile tables).  Parameters, BC kinds and per-column parameters are part of the
 * configuration, not of the checkpoint: load into a ctx created (and configured) the same way.                       */
int64_t lh_soil_checkpoint_bytes(const lh_soil_ctx* ctx);
int32_t lh_soil_checkpoint_save(lh_soil_ctx* ctx, void* buf, int64_t capacity);
int32_t lh_soil_checkpoint_load(lh_soil_ctx* ctx, const void* buf, int64_t bytes);

/* Pinned (page-locked) host memory for state buffers: transfers from / to it are asynchronous and run at PCIe speed. */
int32_t lh_soil_alloc_host(int64_t bytes, void** out);
int32_t lh_soil_free_host(void* p);

/* ---- other explicit steppers on the same fused kernel ----------------------------------
 * The reference's test driver also imports SSPRK73 and CarpenterKennedy2N54 next to SSPRK33
 * (test/runtests.jl:5-10); only SSPRK33 is ever used.  The fused RHS+stage kernel generalises by
 * coefficients to the two low-storage families those methods belong to, each stage still ONE launch:
 *   LH_STEPPER_SHU_OSHER  two-register Shu-Osher form
 *        u_i = a[i] u^n + b[i] u_{i-1} + g[i] dt f(u_{i-1}, t + c[i] dt),  u_0 = u^n,  u^{n+1} = u_s
 *        (Euler, SSPRK22, SSPRK33, SSPRK43, any method whose stages only need u^n and u_{i-1})
 *   LH_STEPPER_2N         Williamson 2N storage
 *        r = a[i] r + dt f(u, t + c[i] dt);   u = u + b[i] r          (a[0] must be 0)
 *        (CarpenterKennedy2N54 and the other LowStorageRK2N methods of OrdinaryDiffEq)
 * SSPRK73's coefficients are numerically optimised constants that live only in OrdinaryDiffEq (not
 * vendored under the reference): a Julia host passes them through this table API. */
#define LH_MAX_STAGES        16
#define LH_STEPPER_SHU_OSHER 0
#define LH_STEPPER_2N        1
typedef struct lh_soil_stepper {
    int32_t kind;                 /* LH_STEPPER_*                                   */
    int32_t nstages;              /* 1 .. LH_MAX_STAGES                             */
    double  a[LH_MAX_STAGES];
    double  b[LH_MAX_STAGES];
    double  g[LH_MAX_STAGES];     /* Shu-Osher only                                 */
    double  c[LH_MAX_STAGES];     /* stage time offsets: row s of bc_table is for t + c[s] dt */
} lh_soil_stepper;

/* Built-in tables (OrdinaryDiffEq names). */
#define LH_METHOD_EULER    0
#define LH_METHOD_SSPRK22  1
#define LH_METHOD_SSPRK33  2      /* same scheme as lh_soil_step_ssprk33, through the generic stage kernel */
#define LH_METHOD_SSPRK43  3
#define LH_METHOD_CK2N54   4      /* CarpenterKennedy2N54 */
int32_t lh_soil_stepper_named(int32_t method, lh_soil_stepper* out);

/* nsteps steps of `stepper`.  bc_table is NULL or nsteps * nstages * 4 doubles (LH_BCV_* per stage).
 * Replaces step!/run! for a Simulation built with another OrdinaryDiffEq method (simulation.jl:34-87). */
int32_t lh_soil_step(lh_soil_ctx* ctx, const lh_soil_stepper* stepper, double t, double dt,
                     int64_t nsteps, const double* bc_table);

/* ---- diagnostics ---------------------------------------------------------------------
 * Water and energy budgets of THIS ctx's columns: out[0] = sum ϑ_l Δz, out[1] = sum ρe_int Δz
 * (deterministic fixed-tree reduction).  New in this build.                     */
int32_t lh_soil_budgets(lh_soil_ctx* ctx, double out[2]);
/* Right after lh_soil_step_ssprk33 / lh_soil_stage_ssprk33(3) this costs one small reduction: the last stage has
 * already summed, per thread block, the values it wrote.  After an upload or a generic stepper it is one pass over the
 * state.  Both are fixed-shape trees (bitwise reproducible for a given shard).                                     */
/* Non-blocking form: enqueues the reduction and a 16-byte copy into a pinned slot of the ctx behind whatever is already
 * on the ctx stream and returns a ticket at once; lh_soil_budgets_wait blocks until THAT result has arrived (not until the
 * stream is idle).  A host that reads the budgets after every step (the reference's callbacks do, simulation.jl:64-70)
 * can enqueue step n+1 before it collects the budgets of step n, so the stream never drains.  At most 8 tickets may be
 * outstanding (LH_ERR_STATE beyond that); results are identical to lh_soil_budgets.                                  */
int32_t lh_soil_budgets_async(lh_soil_ctx* ctx, int64_t* ticket_out);
int32_t lh_soil_budgets_wait(lh_soil_ctx* ctx, int64_t ticket, double out[2]);
/* Column-integrated boundary fluxes of the last rhs/stage call are not stored; conservation
 * tests use budgets before/after a step.                                                    */

/* Derived cell-centre fields of the CURRENT state, evaluated on the device with the same
 * closures the RHS kernel uses (host-side in the reference: users call the
 * parameterisation functions on `parent(Y.soil.*)`, e.g. test/SoilModel/coupled.jl:97-100). */
#define LH_DIAG_K      0   /* hydraulic_conductivity   (SoilWaterParameterizations.jl:269-282) */
#define LH_DIAG_PSI    1   /* pressure_head            (:229-242)                              */
#define LH_DIAG_KAPPA  2   /* thermal_conductivity     (SoilHeatParameterizations.jl:185-188)  */
#define LH_DIAG_T      3   /* temperature_from_ρe_int  (:42-53) or the prescribed T            */
#define LH_NUM_DIAGS   4
int32_t lh_soil_diagnostic(lh_soil_ctx* ctx, int32_t which, double* host,
                           int64_t col_stride, int64_t layer_stride);

int32_t lh_soil_sync(lh_soil_ctx* ctx);

/* Evaluates one of the device elementary functions of csrc/lh_math.cuh on n host values
 * (y[i] = f(x[i]) computed ON THE GPU): the accuracy of the hand-written log2/exp2/exp2m1/sqrt/
 * rsqrt/rcp/div that replace the reference's `^`, `exp`, `sqrt`, `/` is thereby testable against
 * a high-precision host reference.  For LH_MATH_DIV, x holds n numerators followed by n
 * denominators.                                                                              */
#define LH_MATH_LOG2   0   /* log2(x)  */
#define LH_MATH_EXP2   1   /* 2^x      */
#define LH_MATH_EXP2M1 2   /* 2^x - 1  */
#define LH_MATH_SQRT  3
#define LH_MATH_RSQRT 4
#define LH_MATH_RCP   5
#define LH_MATH_DIV   6
#define LH_MATH_RCP_SEED   7   /* the raw MUFU.RCP64H seed (rcp.approx.ftz.f64)     */
#define LH_MATH_RSQRT_SEED 8   /* the raw MUFU.RSQ64H seed (rsqrt.approx.ftz.f64)   */
int32_t lh_soil_eval_math(lh_soil_ctx* ctx, int32_t fn, const double* x, double* y, int64_t n);

/* Device time (ms) spent in the kernels of the last lh_soil_step_ssprk33 call, measured with
 * CUDA events on the ctx stream, and the number of kernels it launched.                     */
int32_t lh_soil_last_step_timing(lh_soil_ctx* ctx, double* ms_out, int64_t* launches_out);

/* One line describing what lh_soil_step_ssprk33 launches for the current state of the ctx: kernel template and variant
 * flags (ICE / GEN / VG2 / HET), block and grid shape, launch strategy, and the HBM bytes per cell-step that variant
 * moves.  For benchmark records (bench.py prints it next to the roofline figures).                                  */
int32_t lh_soil_kernel_info(lh_soil_ctx* ctx, char* buf, int64_t cap);

/* Raw device pointer of a field's column-fastest SoA block [layer][ncol_padded] and the
 * padded column stride, for zero-copy interop (e.g. wrapping in a torch tensor).  Work this ctx has enqueued (uploads, steps)
 * is ordered on the ctx's own stream: call lh_soil_sync before touching the memory from another stream.                  */
int32_t lh_soil_device_ptr(lh_soil_ctx* ctx, int32_t field, void** dptr, int64_t* ncol_padded);

/* ---- multi-GPU: column shards, one ctx per GPU/process ------------------------------- */
/* 128-byte NCCL unique id, produced on rank 0 and distributed by the host's own plumbing.   */
int32_t lh_soil_comm_unique_id(uint8_t id_out[128]);
int32_t lh_soil_comm_init(lh_soil_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t id[128]);
/* Global budgets over all ranks' shards: local fixed-tree reduction, then ncclAllReduce(sum)
 * of 2 doubles on the ctx stream.  The ONLY collective on this path (columns never exchange
 * halos: right_hand_side.jl uses vertical operators only).                                  */
int32_t lh_soil_budgets_allreduce(lh_soil_ctx* ctx, double out[2]);

#ifdef __cplusplus
}
#endif
#endif /* LH_SOIL_H */
