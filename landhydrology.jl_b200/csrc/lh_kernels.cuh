// lh_kernels.cuh — kernel argument block and launch entry points (implemented in lh_kernels.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "lh_closures.cuh"

// Register budget of the stage kernels.  Once the layer loop is provably warp-uniform (lh_stage_kernel.cuh) the
// coupled n = 2 variant needs ~96 registers and every model runs best with 20 resident warps per SM (<= 102
// registers; profiles/r01_i_*: coupled 80 %, general-n 61 %, Richards 51 % of the HBM roofline, against 79 / 58 / 49 %
// at 16 warps/SM).  The bound is expressed as __launch_bounds__(640, 1): the same register cap as (128, 5) and
// byte-identical code, but blocks may hold up to 20 warps, which tall columns use for more chunks per column.
// The per-column-parameter (HET) variants keep ~24 more registers live: (512, 1), 128 registers, 16 warps/SM.
#ifndef LH_PDL
#define LH_PDL 1      // programmatic dependent launch between consecutive stage kernels
#endif
#define LH_WARPS_PER_SM 20
#define LH_WARPS_PER_SM_HET 16
template <int FLAGS> struct LhBounds { static constexpr int max_threads = ((FLAGS & LH_FLAG_HET) ? LH_WARPS_PER_SM_HET : LH_WARPS_PER_SM) * 32; };   // CELLP implies HET

// What one stage reads and writes (device pointers to column-fastest SoA blocks) and its scalars.
struct LhStageIO {
    const double* in_th;   // stage input ϑ_l            (state U, or stage buffer V)
    const double* in_ti;   // θ_i                         (always U: its tendency is 0)
    const double* in_re;   // stage input ρe_int          (U or V)
    const double* aux_T;   // prescribed T                (Richards; read only if viscosity is on)
    const double* u0_th;   // U, for the stage-2/3 combine
    const double* u0_re;
    double* out_th;        // V (stage 1, 2) / U (stage 3) / tendency buffer (stage 0)
    double* out_re;
    double* out2_th;       // 2N stages (STAGE 5): the residual register r, updated in place (== u0_th)
    double* out2_re;
    double bcv[4];         // LH_BCV_* boundary values for THIS stage
    const double* flux_cols[4];   // per-column VerticalFlux values (LH_BCV_* order) or NULL: the scalar in bcv applies.
                                  // Also how PrescribedAtmosForcing arrives: lh_atmos_flux_kernel fills two arrays per stage.
    double dt;
    double sa, sb, sg;     // stage coefficients of the generic steppers (STAGE 4: a, b, g; STAGE 5: a, b)
    int32_t first2n;       // STAGE 5, first stage: r is not read (a == 0 and r may hold anything)
    int32_t budget;        // STAGE 6: this is the last stage of a step (accumulate the budgets of what it writes)
};

struct LhKernelArgs {
    LhDevParams p;
    LhStageIO io;          // one-stage launches: this stage; persistent SSPRK33 launches: stage 1 (in = U, out = V)
    const double* zc;      // nlayer centre coordinates
    const double* pow_tab; // LHPW_COUNT fixed-exponent power tables (lh_math.cuh), written once by lh_soil_create
    const double* colp;    // HET variants: [LHCP_COUNT][ncol_pad] per-column derived parameters
    const double* cellp;   // CELLP variants: [LHCELL_COUNT][nlayer][ncol_pad] per-cell derived parameters
    double* budget_partials;   // [nblocks][2]: per-block sums of the ϑ_l and ρe_int values the last stage writes (or NULL)
    int64_t ncol;          // valid columns (<= ncol_pad): the padding is left out of the budgets
    int64_t ncol_pad;
    int32_t nlayer;
    int32_t Lc;            // layers per thread (vertical chunk)
    int32_t W;             // chunks per column = blockDim.y
    int32_t top_e_kind, top_h_kind, bot_e_kind, bot_h_kind;
    // Stage-to-stage chaining (one-stage launches).  Block j of a stage launch touches only its own column groups, and
    // so does block j of the next launch (same launch shape), so the only true dependency between consecutive stage
    // launches is block j -> block j.  Every block publishes chain_flags[j] = chain_set when its stores are visible
    // (fence + release store); with chain_wait != 0 the next launch's block j waits for chain_flags[j] == chain_wait
    // INSTEAD of griddepcontrol.wait (whole previous grid completed and flushed): the next stage starts filling the SM
    // slots the previous one frees during its last wave, and the per-launch tail (~half a block time per SM, 8 % of a
    // 131 072-column shard's launch) is paid once per call instead of once per stage.  chain_wait == 0: full dependency.
    int32_t* chain_flags;  // [nblocks] or NULL
    int32_t chain_wait, chain_set;
    // persistent SSPRK33 launches only
    int64_t nsteps;
    const double* bc_dev;  // NULL (io.bcv for every stage) or nsteps * 3 * 4 boundary values in device memory
};

struct LhLaunchShape {
    int32_t Lc, W, G;      // chunk length, chunks per column, column groups (of 32) per block
    int64_t nblocks;
    size_t smem_bytes;
    int32_t warp_budget;   // resident warps per SM the variant's register cap allows (20, HET: 16)
    double waves;          // nblocks / (SMs x resident blocks per SM)
};

LhLaunchShape lh_choose_shape(int model, int64_t ncol_pad, int32_t nlayer, int sm_count, bool het);

// stage 0 = tendency only; 1..3 = fused RHS + SSPRK33 stage; 4 = generic Shu-Osher stage; 5 = 2N stage.  flags: LH_FLAG_ICE | LH_FLAG_GEN
// (lh_closures.cuh) select the compiled kernel variant.
cudaError_t lh_launch_stage(int model, int stage, int flags, const LhKernelArgs& args,
                            const LhLaunchShape& shape, cudaStream_t stream);
// args.nsteps whole SSPRK33 steps in ONE launch: every block keeps its column group for all 3 nsteps stages.
cudaError_t lh_launch_ssprk33_persistent(int model, int flags, const LhKernelArgs& args,
                                         const LhLaunchShape& shape, cudaStream_t stream);
// *flag (int32, device) := 1 if any element of x[0..n) is non-zero (NaN counts), else unchanged.
cudaError_t lh_launch_any_nonzero(const double* x, int64_t n, int* flag, cudaStream_t stream);

// Pointwise diagnostics (LH_DIAG_*): out[layer*ncol_pad+col].
cudaError_t lh_launch_diagnostic(int model, int which, const LhDevParams& p, const double* pow_tab, const double* th,
                                 const double* ti, const double* re, const double* T, double* out,
                                 int64_t ncells_pad, const double* colp, int64_t ncol_pad, int heat, const double* cellp, cudaStream_t stream);

// PrescribedAtmosForcing (lh_atmos.cuh): per column, from the top cell (layer nlayer-1) of the stage input, the turbulent
// heat flux and water volume flux into flux_e[col], flux_w[col].
struct LhAtmos;
cudaError_t lh_launch_atmos_fluxes(const LhDevParams& p, const double* pow_tab, const LhAtmos& atm, const double* th_top,
                                   const double* ti_top, const double* re_top, double* flux_e, double* flux_w, int64_t ncol_pad,
                                   const double* colp, int heat_cols, const double* cellp_top, int64_t cell_fs, cudaStream_t stream);
// The same for n given surface states (theta_l, theta_i, T): lh_soil_atmos_fluxes.
cudaError_t lh_launch_atmos_eval(const LhDevParams& p, const double* pow_tab, const LhAtmos& atm, const double* th, const double* ti,
                                 const double* T, double* heat, double* water, int64_t n, cudaStream_t stream);

// Deterministic budgets: out2[0] = sum ϑ_l dz, out2[1] = sum ρe_int dz over columns < ncol.
cudaError_t lh_launch_budgets(const double* th, const double* re, int64_t ncol, int64_t ncol_pad,
                              int32_t nlayer, double dz, double* partials, int32_t npartials,
                              double* out2, cudaStream_t stream);
// The same from the per-block sums a last-stage launch left behind (fixed-shape tree over the blocks).
cudaError_t lh_launch_budgets_from_partials(const double* partials, int64_t npartials, double dz, double* out2, cudaStream_t stream);

// Layout transforms between a dense host-layout staging block [col][layer] (layer fastest) and
// the device SoA [layer][ncol_pad].
cudaError_t lh_launch_to_soa(const double* staged, double* soa, int64_t col0, int64_t ncols,
                             int32_t nlayer, int64_t ncol_pad, cudaStream_t stream);
cudaError_t lh_launch_from_soa(const double* soa, double* staged, int64_t col0, int64_t ncols,
                               int32_t nlayer, int64_t ncol_pad, cudaStream_t stream);
// Broadcast an nlayer profile to all columns; replicate the last valid column into the padding.
cudaError_t lh_launch_fill_profile(const double* profile, double* soa, int32_t nlayer, int64_t ncol_pad,
                                   cudaStream_t stream);
cudaError_t lh_launch_fill_padding(double* soa, int64_t ncol, int64_t ncol_pad, int32_t nlayer,
                                   cudaStream_t stream);
// y[i] = f(x[i]) for the elementary function `fn` (LH_MATH_*), all device pointers.
cudaError_t lh_launch_eval_math(const LhDevParams& p, int fn, const double* x, double* y, int64_t n, cudaStream_t stream);
// Counts non-finite values of a field into *count (uint64).
cudaError_t lh_launch_count_nonfinite(const double* soa, int64_t n, unsigned long long* count,
                                      cudaStream_t stream);
