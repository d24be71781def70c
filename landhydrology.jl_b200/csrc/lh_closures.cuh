// lh_closures.cuh — device-side soil closures (fp64), fused into every kernel of this library.
//
// What the reference evaluates as ~16 separate broadcasts with ~12 `pow`/`exp` calls per cell
// (src/SoilModel/right_hand_side.jl:291-314; SoilWaterParameterizations.jl:196-282;
// SoilHeatParameterizations.jl:42-207) is evaluated here once per cell in registers.  The
// algebra is re-associated so that the van Genuchten retention curve and conductivity share
// ONE log(S) and ONE log(1 - S^(1/m)):
//
//     u = log2(S)/m           y = S^(1/m) = 2^u          w = 1 - y = -exp2m1(u)
//     a = log2(w)
//     psi = -((S^(-1/m) - 1) alpha^(-n))^(1/n) = -(1/alpha) exp2((a - u)/n)
//         = -(1/alpha) w S / (y W),  W = (1 - y)^m,  when m = 1 - 1/n (Mualem, the reference's constructor)
//     K   = Ksat sqrt(S) (1 - (1 - y)^m)^2     = Ksat sqrt(S) exp2m1(m a)^2
//
// which is also better conditioned than the literal form (1 - y and 1 - (1-y)^m are formed by
// exp2m1, not by subtraction).  Results agree with the literal fp64 evaluation to a few ulp;
// the parity gate is 1e-12 (cancellation-aware norm) per tendency evaluation.
//
// The fp64 pipe (64 DFMA/clk/SM on B200), not HBM, is the binding unit for this path, so the
// elementary functions in lh_math.cuh are written for minimum DFMA count.  `tab` points at the exp2 / log2
// tables of lh_math.cuh, staged in shared memory by every kernel (lh_stage_tables).
#pragma once

#include <stdint.h>

#include "lh_math.cuh"

// The same source compiles on the host (LH_MATH_HOST, as lh_math.cuh): tests/support/hostemu builds the whole library for the
// CPU-only test run (every kernel variant against the oracle).  Test infrastructure only; on the device LH_DEVFN is exactly
// `__device__` + `__forceinline__` and `lh_ldg` is `__ldg`.
#ifdef LH_MATH_HOST
#define LH_DEVFN static inline
#define lh_ldg(p) (*(p))
#else
#include <cuda_runtime.h>
#define LH_DEVFN __device__ __forceinline__
#define lh_ldg(p) __ldg(p)
#endif

#define LH_EPS 2.220446049250313e-16

// Host-derived constants (all fp64, computed once in lh_soil_create).  Passed to kernels by
// value (__grid_constant__), so every use is a constant-bank operand, not a register.
struct LhPhys {
    // geometry
    double dz, inv_dz, half_dz, inv_half_dz;
    double neg_half_inv_dz;            // -1 / (2 dz): interior face fluxes
    // water
    double nu, theta_r, theta_r_eps;   // theta_r + eps(Float64)
    double inv_nu_thr;                 // 1 / (nu - theta_r)
    double nu_thr;                     // nu - theta_r
    double S_s_inv;                    // 1 / S_s
    double vg_m, vg_inv_m, vg_inv_n;
    double neg_inv_alpha;              // -1 / alpha
    double Ksat;
    double visc_gamma_l2e, visc_T_ref; // gamma * log2(e)
    double imp_c2;                     // -Omega * log2(10)
    // heat
    double rho_c_ds, rhocp_l, rhocp_i, rhoi_LH, T_0;
    double inv_nu;
    double kersten_p1;                 // (1 + nu_om - a nu_quartz - nu_gravel) / 2
    double kersten_p2;                 // 1 - nu_om
    double kersten_p3;                 // 1 + nu_om
    double neg_b_l2e;                  // -b * log2(e)
    double k_unfrozen, k_frozen, log2_k_unfrozen, log2_k_frozen;
    double kappa_dry;
    double k_unfrozen_minus_dry;       // kappa_sat_unfrozen - kappa_dry
    double log2_Sr_sat;                // log2(nu * (1/nu)): relative saturation of a saturated, ice-free cell
    double pow_p1_Sr_sat;              // (nu * (1/nu))^kersten_p1
    LhPowCoef pw[3];                   // fixed-exponent powers (lh_math.cuh): LHPW_INVM x^(1/m), LHPW_M x^m, LHPW_P1 x^kersten_p1
    int32_t visc_on, imp_on;
    int32_t om_zero;                   // nu_ss_om == 0: outer Kersten exponents are exactly 1
    int32_t pad_;
};

enum { LHPW_INVM, LHPW_M, LHPW_P1, LHPW_COUNT };
// Shared-memory table block of every kernel: the exp2 / log2 tables, then one table per fixed exponent.
#define LH_TAB_ALL (LH_TAB_DOUBLES + LHPW_COUNT * LH_POW_DOUBLES)

// The uniform parameter block of a launch: physics + the elementary-function coefficients and tables.
// POW: the exponents 1/m, m and the Kersten exponent are per-model constants, so x^c is one table-driven evaluation
// (lh_pow_fixed) instead of exp2(c log2 x).
struct LhDevParams : LhPhys {
    static constexpr bool POW = true;
    double mc[LHC_COUNT];              // lh_math.cuh
};

// The same physics seen by ONE lane when the hydraulic parameters differ from column to column
// (lh_soil_set_column_params): a per-thread copy whose column-dependent members were overwritten; the members
// that stay uniform still come from the parameter block (constant propagation through the copy).
struct LhLaneParams : LhPhys {
    static constexpr bool POW = false; // per-column van Genuchten n: the exponents differ from lane to lane
    const double* mc;
};

// Per-column derived parameters, one [ncol_pad] array each (lh_soil_api.cu derive_column_params).
enum { LHCP_NU, LHCP_THETA_R, LHCP_THETA_R_EPS, LHCP_INV_NU_THR, LHCP_NU_THR, LHCP_VG_M, LHCP_VG_INV_M, LHCP_VG_INV_N,
       LHCP_NEG_INV_ALPHA, LHCP_KSAT, LHCP_INV_NU, LHCP_KAPPA_DRY,
       // heat parameters (lh_soil_set_column_heat_params; read by the HETH variants only)
       LHCP_RHO_C_DS, LHCP_KERSTEN_P1, LHCP_KERSTEN_P2, LHCP_KERSTEN_P3, LHCP_K_UNFROZEN, LHCP_LOG2_K_UNFROZEN, LHCP_LOG2_K_FROZEN,
       LHCP_COUNT };

// Per-cell derived parameter fields, [LHCELL_COUNT][nlayer][ncol_pad] (lh_soil_api.cu rebuild_cell_params); the cheap
// derivatives (theta_r + eps, nu - theta_r, 1 - m, kappa_unfrozen - kappa_dry) are formed in the kernel.
enum { LHCELL_NU, LHCELL_THETA_R, LHCELL_INV_NU_THR, LHCELL_VG_M, LHCELL_VG_INV_M, LHCELL_NEG_INV_ALPHA, LHCELL_KSAT,
       LHCELL_INV_NU, LHCELL_KAPPA_DRY, LHCELL_COUNT };

// Overwrites the cell-dependent members of a lane's parameter view from the per-cell fields: cp points at this lane's
// element of field 0 for the cell, fs is the field stride (nlayer * ncol_pad).
LH_DEVFN void lh_load_cell_params(LhLaneParams& p, const double* __restrict__ cp, int64_t fs)
{
    p.nu = lh_ldg(cp + LHCELL_NU * fs);
    p.theta_r = lh_ldg(cp + LHCELL_THETA_R * fs);
    p.inv_nu_thr = lh_ldg(cp + LHCELL_INV_NU_THR * fs);
    p.vg_m = lh_ldg(cp + LHCELL_VG_M * fs);
    p.vg_inv_m = lh_ldg(cp + LHCELL_VG_INV_M * fs);
    p.neg_inv_alpha = lh_ldg(cp + LHCELL_NEG_INV_ALPHA * fs);
    p.Ksat = lh_ldg(cp + LHCELL_KSAT * fs);
    p.inv_nu = lh_ldg(cp + LHCELL_INV_NU * fs);
    p.kappa_dry = lh_ldg(cp + LHCELL_KAPPA_DRY * fs);
    p.theta_r_eps = p.theta_r + LH_EPS;
    p.nu_thr = p.nu - p.theta_r;
    p.vg_inv_n = 1.0 - p.vg_m;                       // 1/n = 1 - m (Mualem), within 1 ulp of the quotient
    p.k_unfrozen_minus_dry = p.k_unfrozen - p.kappa_dry;
}

// Copies the exp2 / log2 tables from the parameter block, and the fixed-exponent power tables from `pow_tab`
// (LHPW_COUNT * LH_POW_DOUBLES doubles in global memory, written once by lh_soil_create), to shared memory
// (LH_TAB_ALL doubles, 16-byte aligned destination); the caller must __syncthreads().
LH_DEVFN void lh_stage_tables(const LhDevParams& p, const double* __restrict__ pow_tab, double* tab_smem,
                                                int linear_tid, int nthreads)
{
    for (int k = linear_tid; k < LH_TAB_DOUBLES; k += nthreads) tab_smem[k] = p.mc[LHC_TAB0 + k];
    if (pow_tab) for (int k = linear_tid; k < LHPW_COUNT * LH_POW_DOUBLES; k += nthreads) tab_smem[LH_TAB_DOUBLES + k] = pow_tab[k];
}

struct LhCell {
    double K;      // hydraulic conductivity
    double psi;    // pressure head
    double kappa;  // thermal conductivity
    double T;      // temperature
    double dT;     // T - T_0 as computed (the quotient of temperature_from_ρe_int), for ρe_int_l = ρc_l (T - T_0)
};

// Kernel variants (template flags).  Both are properties of the uploaded problem, decided on the
// host: θ_i never changes (dθ_i ≡ 0, right_hand_side.jl:182,359), so "no ice anywhere" is known
// once θ_i has been uploaded.
//   ICE : θ_i may be non-zero somewhere.  !ICE: θ_i ≡ 0, the kernels do not even read the field.
//   GEN : general parameters (viscosity/impedance factor on, ν_ss_om != 0 or θr != 0).  !GEN is the
//         branch-free fast path: NoEffect factors (the reference's defaults, models.jl:31-32),
//         ν_ss_om = 0 (outer Kersten exponents exactly 1) and θr = 0 (S_r == S bit for bit, so the
//         Kersten number reuses log S).
//   VG2 : van Genuchten n == 2 (m == 1/2) exactly, as in the reference's coupled tests
//         (test/SoilModel/coupled.jl:6): S^(1/m) = S^2 and x^m = sqrt(x), so the retention curve and the
//         conductivity need no log/exp at all:
//             w = 1 - S^2 = (1 - S)(1 + S)       psi = -(1/alpha) sqrt(w) / S
//             1 - sqrt(w) = S^2 / (1 + sqrt(w))  K = Ksat sqrt(S) (S^2 / (1 + sqrt(w)))^2
//         (exact identities; the second also removes the cancellation of the literal form).
#define LH_FLAG_ICE 1
#define LH_FLAG_GEN 2
#define LH_FLAG_VG2 4
//   HET : per-column hydraulic parameters (lh_soil_set_column_params): nu, theta_r, van Genuchten n and alpha, Ksat
//         and what derives from them are per-lane values.  Implies GEN and excludes VG2.
#define LH_FLAG_HET 8
//   HETH: per-column HEAT parameters too (lh_soil_set_column_heat_params): rho_c_ds, kappa_sat_unfrozen / frozen, and the
//         Kersten exponents that follow nu_ss_om / nu_ss_quartz / nu_ss_gravel.  Implies HET; nu_ss_om may differ from column
//         to column, so the outer Kersten exponents are always evaluated (no om_zero shortcut).
#define LH_FLAG_HETH 16
//   CELLP: per-CELL hydraulic parameters (lh_soil_set_cell_params: layered soils): the lane's nu, theta_r, van Genuchten and
//         Ksat values (and what derives from them) are re-read from LHCELL_COUNT fields before every cell (+72 B per cell and
//         stage).  Implies HET (and HETH for the models with an energy equation).
#define LH_FLAG_CELLP 32

// ---------------------------------------------------------------------------------------------
// Water: K and psi of one cell.  Reference: right_hand_side.jl:156-166 / :308-313.
//   th = ϑ_l, ti = θ_i, T = temperature (only read when the viscosity factor is on).
// The two data-dependent conditionals of the reference (S_l_eff <= 1, S < 1) are selects on the
// results: the unsaturated expressions are evaluated unconditionally (they are the common case and
// yield NaN/garbage only where the select discards them).
// ---------------------------------------------------------------------------------------------
template <bool ICE, bool GEN, bool VG2, bool NEED_LOG, bool THR0 = false, class P = LhDevParams>
LH_DEVFN void lh_water_closures(const P& p, const double* __restrict__ tab,
                                                  double th, double ti, double T,
                                                  double& K_out, double& psi_out, double& logS_K, LhPowArg& argS_K)
{
    const double* __restrict__ mc = p.mc;
    const double nu_eff = ICE ? p.nu - ti : p.nu;
    // effective_saturation (SoilWaterParameterizations.jl:213-217); NaN stays NaN
    const double safe = !(th <= p.theta_r_eps) ? th : p.theta_r_eps;
    const double num = THR0 ? safe : safe - p.theta_r;                       // THR0: theta_r == 0 is known
    const double S_K = num * p.inv_nu_thr;                                   // porosity = nu (:163/:311)
    // The ICE variants take the two-saturation path for EVERY cell (straight-line code): a per-lane test for ti != 0 made the
    // layer loop divergent (10 reconvergence points, spills) and saved nothing in a warp that holds both kinds of cells.
    constexpr bool icy = ICE;
    double S_eff = S_K;                                                      // porosity = nu_eff (:235)
    const double den_K = THR0 ? p.nu : p.nu_thr;
    double den_eff = den_K;
    if (icy) { den_eff = THR0 ? nu_eff : nu_eff - p.theta_r; S_eff = lh_div(num, den_eff); }
    // The reference branches on the ROUNDED quotients (S_l_eff <= 1, S < 1).  A correctly rounded num/den is < 1
    // exactly when num < den, and at num == den both branches of the pressure head give 0, so the branches are
    // decided on num and den themselves: the product num (1/den) used for S here can be 1 - 2^-53 where the
    // quotient is exactly 1 (it is for ~15 % of (nu, theta_r) pairs), and a cell that sits exactly at saturation
    // would otherwise get psi = -(1/alpha) sqrt(2^-52) and K (1 - 3e-8) instead of 0 and K_sat.
    const bool unsat_eff = num < den_eff, unsat_K = num < den_K;
    const double psi_sat = (th - nu_eff) * p.S_s_inv;
    double psi_unsat, Kr_unsat, L_K = 0.0;

    if (VG2) {
        // ---- n = 2: square roots only.  One rsqrt(S) gives both sqrt(S) = S r and 1/S = r r.
        const double rS = lh_rsqrt(S_K);
        const double sw = lh_sqrt_fast((1.0 - S_eff) * (1.0 + S_eff));       // (1 - S^(1/m))^m
        double inv_S_eff = rS * rS, swK = sw;
        if (icy) {
            inv_S_eff = lh_rcp(S_eff);
            swK = lh_sqrt_fast((1.0 - S_K) * (1.0 + S_K));
        }
        psi_unsat = p.neg_inv_alpha * (sw * inv_S_eff);                      // :196-200
        const double t = (S_K * S_K) * lh_rcp(1.0 + swK);                    // 1 - (1 - S^(1/m))^m
        Kr_unsat = (S_K * rS) * (t * t);                                     // :277
        if (NEED_LOG) {                                      // Kersten exponent only; a NaN state already poisons psi and K
            if constexpr (P::POW) argS_K = lh_pow_arg(tab, S_K);
            else L_K = lh_log2<false>(mc, tab, S_K);
        }
    } else if constexpr (P::POW) {
        // ---- general n, exponents fixed per model: y = S^(1/m), W = (1 - y)^m are fixed-exponent powers.  Nothing is
        // checked: S_eff > 0 always; 1 - y <= 0 only for S_eff >= 1, where both selects below discard the unsaturated
        // expressions (lh_pow_* return finite garbage there); a NaN state reaches psi and K through their factors of S.
        const double* __restrict__ pt_invm = tab + LH_TAB_DOUBLES + LHPW_INVM * LH_POW_DOUBLES;
        const double* __restrict__ pt_m = tab + LH_TAB_DOUBLES + LHPW_M * LH_POW_DOUBLES;
        const LhPowArg aS = lh_pow_arg(tab, S_eff);
        const LhExpParts ey = lh_pow_parts(p.pw[LHPW_INVM], pt_invm, aS);    // y = S^(1/m) = s (1 + p)
        const double w = lh_fma(-ey.s, ey.p, 1.0 - ey.s);                    // 1 - y (exactly -p when S is within 2^-9 of 1)
        const LhPowArg aw = lh_pow_arg(tab, w);
        const double y = lh_fma(ey.s, ey.p, ey.s);
        argS_K = aS;
        if (!icy) {
            const LhExpParts eW = lh_pow_parts(p.pw[LHPW_M], pt_m, aw);      // W = (1 - y)^m
            const double q = lh_fma(eW.s, eW.p, eW.s - 1.0);                 // W - 1 (exactly p for dry cells: y < 2^-9)
            Kr_unsat = (S_K * lh_rsqrt(S_K)) * (q * q);                      // :277; sqrt(S) = S rsqrt(S): S >= eps > 0 here
            // Mualem's m = 1 - 1/n (what the reference constructor stores; lh_soil_create rejects anything else):
            // 1/n = 1 - m and y^m = S, so ((1 - y)/y)^(1/n) = (w/y) (y/w)^m = w S / (y W): a reciprocal, no third power.
            const double W = lh_fma(eW.s, eW.p, eW.s);
            psi_unsat = p.neg_inv_alpha * ((w * S_eff) * lh_rcp(y * W));     // :196-200
        } else {                           // S (porosity nu) differs from S_eff (nu - theta_i) only when ice is present
            const double W = lh_pow_fixed(p.pw[LHPW_M], pt_m, aw);
            psi_unsat = p.neg_inv_alpha * ((w * S_eff) * lh_rcp(y * W));
            argS_K = lh_pow_arg(tab, S_K);
            const LhExpParts eyK = lh_pow_parts(p.pw[LHPW_INVM], pt_invm, argS_K);
            const LhPowArg awK = lh_pow_arg(tab, lh_fma(-eyK.s, eyK.p, 1.0 - eyK.s));
            const LhExpParts eW = lh_pow_parts(p.pw[LHPW_M], pt_m, awK);
            const double q = lh_fma(eW.s, eW.p, eW.s - 1.0);
            Kr_unsat = (S_K * lh_rsqrt(S_K)) * (q * q);
        }
    } else {
        // ---- general n: pressure head (:229-242) and the shared logs.  The logs run unchecked (no NaN flag for a
        // negative argument): S_eff > 0 always; 1 - y < 0 only for S_eff > 1, where both selects below discard the
        // unsaturated expressions; and a NaN state still reaches psi and K through their explicit factors of S.
        const double L_eff = lh_log2<false>(mc, tab, S_eff);
        const double u = L_eff * p.vg_inv_m;
        const LhExpParts ey = lh_exp2_parts(mc, tab, u);                     // y = S^(1/m) = s (1 + p)
        const double w = lh_fma(-ey.s, ey.p, 1.0 - ey.s);                    // 1 - y; its zero is +0
        const double a = lh_log2<false>(mc, tab, w);
        // ---- hydraulic conductivity (:269-282)
        double a_K = a;
        L_K = L_eff;
        if (icy) {                       // S differs from S_eff only when ice is present
            L_K = lh_log2<false>(mc, tab, S_K);
            a_K = lh_log2<false>(mc, tab, lh_one_minus_exp2(mc, tab, L_K * p.vg_inv_m));
        }
        const LhExpParts eW = lh_exp2_parts(mc, tab, p.vg_m * a_K);          // W = (1 - y)^m
        const double q = lh_fma(eW.s, eW.p, eW.s - 1.0);                     // W - 1
        Kr_unsat = (S_K * lh_rsqrt(S_K)) * (q * q);                          // sqrt(S) = S rsqrt(S): S >= eps > 0 here
        if (!icy) {
            // Mualem's m = 1 - 1/n (what the reference constructor stores, SoilWaterParameterizations.jl:162-169;
            // lh_soil_create rejects anything else):
            // 1/n = 1 - m and y^m = S, so ((1 - y)/y)^(1/n) = (w/y) (y/w)^m = w S / (y W) with the W of the
            // conductivity: a reciprocal instead of a third exp2.
            const double y = lh_fma(ey.s, ey.p, ey.s);
            const double W = lh_fma(eW.s, eW.p, eW.s);
            psi_unsat = p.neg_inv_alpha * ((w * S_eff) * lh_rcp(y * W));
        } else {
            psi_unsat = p.neg_inv_alpha * lh_exp2(mc, tab, (a - u) * p.vg_inv_n);
        }
    }
    const double psi = unsat_eff ? psi_unsat : psi_sat;
    const double Kr = unsat_K ? Kr_unsat : 1.0;
    double K = Kr * p.Ksat;
    if (GEN) {
        if (p.visc_on) K *= lh_exp2(mc, tab, p.visc_gamma_l2e * (T - p.visc_T_ref));   // :117-126
        if (p.imp_on) {                                                      // :89-93, f_i :159/:308
            const double tl = (th < nu_eff) ? th : nu_eff;
            const double f_i = lh_div(ti, tl + ti);
            K *= lh_exp2(mc, tab, p.imp_c2 * f_i);                              // 10^(-Omega f_i)
        }
    }
    K_out = K;
    psi_out = psi;
    logS_K = L_K;
}

// ---------------------------------------------------------------------------------------------
// Heat: thermal conductivity from (θ_l, θ_i).  Reference: right_hand_side.jl:296-305,
// SoilHeatParameterizations.jl:114-188.
//   REUSE: the caller has log2 S of the SAME cell (coupled model, fast path): with θr = 0 and no ice,
//   S_r = θ_l/ν equals S = ϑ_l/ν bit for bit while ϑ_l < ν, and is the constant ν(1/ν) once the cell
//   is saturated, so no second log is needed.  (For ϑ_l <= eps the two differ, but there the
//   Kersten base E3 - c^3 is ~0 and K_e vanishes either way.)
// ---------------------------------------------------------------------------------------------
template <bool ICE, bool GEN, bool REUSE, class P = LhDevParams>
LH_DEVFN double lh_thermal_conductivity(const P& p, const double* __restrict__ tab,
                                                          double tl, double ti, bool unsat, double logS, const LhPowArg& argS)
{
    const double* __restrict__ mc = p.mc;
    const double tw = ICE ? tl + ti : tl;
    const double S_r = tw * p.inv_nu;                                        // relative_saturation :139-142
    constexpr bool REUSED = REUSE && !ICE && !GEN;
    double Lr = 0.0;
    if constexpr (!P::POW) {
        if (REUSED) Lr = unsat ? logS : p.log2_Sr_sat;
        else Lr = lh_log2(mc, tab, S_r);         // checked: a negative water content must not yield a finite kappa
    }
    double K_e;
    if (!ICE || ti < LH_EPS) {                                               // kersten_number :163-169
        const double e = lh_exp2(mc, tab, p.neg_b_l2e * S_r);                // exp(-b S_r)
        const double g = 1.0 + e;
        const double E3 = lh_rcp(g * g * g);
        const double c = (1.0 - S_r) * 0.5;
        double base = E3 - c * c * c;
        if (GEN && !p.om_zero) base = lh_exp2(mc, tab, p.kersten_p2 * lh_log2(mc, tab, base));
        if constexpr (P::POW) {
            // S_r^p1 with the per-model exponent table; REUSED: the water closure already reduced S == S_r
            const double* __restrict__ pt = tab + LH_TAB_DOUBLES + LHPW_P1 * LH_POW_DOUBLES;
            double Sp;
            if (REUSED) {
                Sp = lh_pow_fixed(p.pw[LHPW_P1], pt, argS);
                Sp = unsat ? Sp : p.pow_p1_Sr_sat;
            } else {
                Sp = lh_pow_fixed(p.pw[LHPW_P1], pt, lh_pow_arg(tab, S_r));
                // a negative water content (or a NaN) must not yield a finite kappa
                Sp = lh_mk(lh_pow_bad(S_r) ? 0x7ff80000 : lh_hi(Sp), lh_lo(Sp));
            }
            K_e = Sp * base;
        } else {
            K_e = lh_exp2(mc, tab, p.kersten_p1 * Lr) * base;
        }
    } else {                                                                 // :171
        if constexpr (P::POW) {
            K_e = S_r;
            if (GEN && !p.om_zero) K_e = lh_exp2(mc, tab, p.kersten_p3 * lh_log2(mc, tab, S_r));
            else K_e = lh_mk(lh_pow_bad(S_r) ? 0x7ff80000 : lh_hi(K_e), lh_lo(K_e));
        } else {
            K_e = (!GEN || p.om_zero) ? S_r : lh_exp2(mc, tab, p.kersten_p3 * Lr);
        }
    }
    if (!ICE) {
        // kappa_sat = kappa_unfrozen^1 * kappa_frozen^0 exactly (:114-128), so
        // K_e kappa_sat + (1 - K_e) kappa_dry = kappa_dry + K_e (kappa_unfrozen - kappa_dry): one FMA.
        // The theta_w < eps case (kappa_sat = 0) needs no select: there S_r -> 0, the Kersten base
        // E3 - c^3 -> 1/8 - 1/8 and K_e vanishes, so kappa = kappa_dry either way.
        return lh_fma(K_e, p.k_unfrozen_minus_dry, p.kappa_dry);   // (the per-column view overrides both members)
    }
    double k_sat = p.k_unfrozen;                                             // :114-128; x^1 * y^0 is exact
    k_sat = lh_exp2(mc, tab, lh_div_fast(tl * p.log2_k_unfrozen + ti * p.log2_k_frozen, tw));   // ti == 0: 2^(log2 k_u), 1 ulp from k_u
    k_sat = (tw < LH_EPS) ? 0.0 : k_sat;
    return K_e * k_sat + (1.0 - K_e) * p.kappa_dry;                          // thermal_conductivity :185-188
}

// Temperature from ρe_int (SoilHeatParameterizations.jl:42-79): returns T - T_0 (the quotient); T = T_0 + it.
template <bool ICE, class P = LhDevParams>
LH_DEVFN double lh_temperature_minus_T0(const P& p, double tl, double ti, double re)
{
    if (ICE) {
        const double rho_c_s = p.rho_c_ds + tl * p.rhocp_l + ti * p.rhocp_i; // :65-79
        return lh_div_fast(re + ti * p.rhoi_LH, rho_c_s);                    // :42-53 (quotient << T_0: 2 ulp of it is below ulp(T))
    }
    const double rho_c_s = p.rho_c_ds + tl * p.rhocp_l;
    return lh_div_fast(re, rho_c_s);
}

// All closures of one cell for model MODEL (0 Richards, 1 heat, 2 coupled).
//   Richards: T_or_re = prescribed T.   heat/coupled: T_or_re = ρe_int.
template <int MODEL, int FLAGS, class P = LhDevParams>
LH_DEVFN LhCell lh_cell_closures(const P& p, const double* __restrict__ tab,
                                                   double th, double ti, double T_or_re)
{
    constexpr bool ICE = (FLAGS & LH_FLAG_ICE) != 0, GEN = (FLAGS & LH_FLAG_GEN) != 0, VG2 = (FLAGS & LH_FLAG_VG2) != 0;
    constexpr bool REUSE = (MODEL == 2) && !ICE && !GEN;   // the Kersten number takes log S from the water closure
    LhCell c;
    c.K = 0.0; c.psi = 0.0; c.kappa = 0.0; c.T = T_or_re; c.dT = 0.0;
    const double nu_eff = ICE ? p.nu - ti : p.nu;
    const bool unsat = th < nu_eff;
    const double tl = unsat ? th : nu_eff;                                   // volumetric_liquid_fraction :181-188
    if (MODEL != 0) {
        // ρe_int_l = ρc_l (T - T_0) (SoilHeatParameterizations.jl:198-207) takes the quotient itself: the
        // reference's T - T_0 re-rounds it to ulp(T) ~ 5.7e-14, which this form does not
        c.dT = lh_temperature_minus_T0<ICE>(p, tl, ti, T_or_re);
        c.T = p.T_0 + c.dT;
    }
    double logS = 0.0;
    LhPowArg argS;
    argS.t = 0.0; argS.j = 0; argS.be = 0;
    // the coupled !GEN variants are only launched when theta_r == 0 (update_kernel_flags)
    if (MODEL != 1) lh_water_closures<ICE, GEN, VG2, REUSE, (MODEL == 2 && !GEN)>(p, tab, th, ti, c.T, c.K, c.psi, logS, argS);
    if (MODEL == 1) c.kappa = lh_thermal_conductivity<ICE, GEN, false>(p, tab, tl, ti, unsat, 0.0, argS);
    if (MODEL == 2) c.kappa = lh_thermal_conductivity<ICE, GEN, true>(p, tab, tl, ti, unsat, logS, argS);
    return c;
}

// κ at a boundary "face state" (boundary_conditions.jl:429-436).
template <int FLAGS, class P = LhDevParams>
LH_DEVFN double lh_face_kappa(const P& p, const double* __restrict__ tab, double th, double ti)
{
    constexpr bool ICE = (FLAGS & LH_FLAG_ICE) != 0, GEN = (FLAGS & LH_FLAG_GEN) != 0;
    const double nu_eff = ICE ? p.nu - ti : p.nu;
    const bool unsat = th < nu_eff;
    const double tl = unsat ? th : nu_eff;
    LhPowArg none;
    none.t = 0.0; none.j = 0; none.be = 0;
    return lh_thermal_conductivity<ICE, GEN, false>(p, tab, tl, ti, unsat, 0.0, none);
}
