// lh_kernels.cu — the fused soil RHS (+ SSPRK33 stage) kernels for sm_100a.
//
// One launch = one pass over the state:  pointwise closures -> cell-to-face interpolation and
// gradient -> Darcy / conductive / advective face fluxes -> boundary fluxes -> face-to-cell
// divergence -> Runge-Kutta stage combine -> store.  Nothing but the state itself touches HBM
// (reference: ~16 array temporaries per RHS, right_hand_side.jl:291-314, and separate axpy
// passes per stage in OrdinaryDiffEq).
//
// Mapping.  Columns are laterally independent and stored column-fastest, so lane = column:
// a warp owns 32 adjacent columns and every global access is two full 128-byte lines.  The
// vertical direction is cut into W chunks of Lc layers; chunk w of a column group is warp w of
// the block.  Each thread marches its chunk bottom -> top with a register sliding window
// (closures of cell i-1 and i, the flux of the face below), so no cell is read twice and no
// closure is evaluated twice.  The only coupling between chunks is the face shared by the top
// cell of chunk w and the bottom cell of chunk w+1: both threads publish those cells' closures
// to shared memory, meet at ONE __syncthreads, and each finishes its own first and last cell
// (the shared face flux is computed by both, bit-identically, so the flux-form divergence
// telescopes exactly and budgets are conserved to round-off).
//
// Stage buffers are updated in place: a thread only ever reads and writes cells of its own chunk
// in global memory, and writes a cell after its last read of it.
#include "lh_kernels.cuh"

#include "lh_atmos.cuh"
#include "lh_ptx.cuh"
#include "lh_soil.h"

#include <stdlib.h>

#ifndef LH_MIN_CHUNK
#define LH_MIN_CHUNK 16
#endif

#define LH_DECL_STAGE(M) \
    cudaError_t lh_launch_stage_m##M##_g0(int, int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t); \
    cudaError_t lh_launch_stage_m##M##_g1(int, int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t); \
    cudaError_t lh_launch_stage_m##M##_g2(int, int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t);
LH_DECL_STAGE(0)
LH_DECL_STAGE(1)
LH_DECL_STAGE(2)
cudaError_t lh_launch_persistent_m0(int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t);
cudaError_t lh_launch_persistent_m1(int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t);
cudaError_t lh_launch_persistent_m2(int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t);

LhLaunchShape lh_choose_shape(int model, int64_t ncol_pad, int32_t nlayer, int sm_count, bool het)
{
    LhLaunchShape s;
    const int64_t groups = ncol_pad / 32;
    const int budget = het ? LH_WARPS_PER_SM_HET : LH_WARPS_PER_SM;
    // Chunks per column (= warps per column group): divisors of the warp budget, so that whole blocks fill it.
    // Long chunks are cheaper (the first two cells of a chunk and its two faces are handled outside the layer loop:
    // one thread per whole 64-layer column runs at 94 % of the HBM roofline, 16-layer chunks at 90 %,
    // profiles/r01_z_chunk_length.log), so the column is cut only as far as needed for >= 2.5 waves of resident warps
    // (column shards of a multi-GPU run included: 131 072 columns run 3 % faster in 32-layer than in 16-layer chunks),
    // and never below LH_MIN_CHUNK layers per thread — unless there are too few columns to even put 8 warps on
    // every SM, when the chunks shrink down to 2 layers.
    static const int cand20[] = {1, 2, 4, 5, 10, 20}, cand16[] = {1, 2, 4, 8, 16};
    const int* cand = het ? cand16 : cand20;
    const int ncand = het ? 5 : 6;
    // tuning knob, read ONCE per process (not on every call of the product path)
    static const int min_chunk = [] { const char* e = getenv("LH_MIN_CHUNK"); const int v = e ? atoi(e) : 0; return v > 0 ? v : LH_MIN_CHUNK; }();
    const int64_t enough_warps = (int64_t)5 * sm_count * budget / 2;
    int k = 0;
    while (k + 1 < ncand && groups * cand[k] < enough_warps && nlayer >= min_chunk * cand[k + 1]) ++k;
    const int64_t want_warps = (int64_t)sm_count * 8;
    while (k + 1 < ncand && groups * cand[k] < want_warps && nlayer >= 2 * cand[k + 1]) ++k;
    int W = cand[k];
    const int Lc = (nlayer + W - 1) / W;
    W = (nlayer + Lc - 1) / Lc;          // no empty chunks
    int G = 1;
    while (W * G * 2 <= 4 && (int64_t)G * 2 <= groups) G *= 2;     // 4-warp blocks at least (when there are that many groups)
    s.Lc = Lc; s.W = W; s.G = G;
    s.nblocks = (groups + G - 1) / G;
    const int nq = model == LH_MODEL_COUPLED ? 5 : 2;
    s.smem_bytes = (LH_TAB_ALL + 2 * LH_WARPS_PER_SM + (size_t)G * W * ((2 * nq + 6) * 32 + 4 * 5 * 32)) * sizeof(double);   // + cp.async input ring
    s.warp_budget = budget;
    const int by_smem = (int)((227 * 1024) / (s.smem_bytes + 1024));
    const int by_regs = budget / (W * G);
    int resident = by_smem < by_regs ? by_smem : by_regs;
    if (resident < 1) resident = 1;
    s.waves = (double)s.nblocks / ((double)sm_count * resident);
    return s;
}

cudaError_t lh_launch_stage(int model, int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape,
                            cudaStream_t stream)
{
    if (model < 0 || model > 2) return cudaErrorInvalidValue;
    if (shape.W * shape.G > shape.warp_budget || shape.smem_bytes > 226 * 1024) return cudaErrorInvalidConfiguration;
    if (stage < 0 || stage > 5) return cudaErrorInvalidValue;
    typedef cudaError_t (*launch_fn)(int, int, const LhKernelArgs&, const LhLaunchShape&, cudaStream_t);
    static const launch_fn table[3][3] = {{lh_launch_stage_m0_g0, lh_launch_stage_m0_g1, lh_launch_stage_m0_g2},
                                          {lh_launch_stage_m1_g0, lh_launch_stage_m1_g1, lh_launch_stage_m1_g2},
                                          {lh_launch_stage_m2_g0, lh_launch_stage_m2_g1, lh_launch_stage_m2_g2}};
    return table[model][stage / 2](stage, flags, args, shape, stream);
}

cudaError_t lh_launch_ssprk33_persistent(int model, int flags, const LhKernelArgs& args, const LhLaunchShape& shape,
                                         cudaStream_t stream)
{
    if (model < 0 || model > 2) return cudaErrorInvalidValue;
    if (shape.W * shape.G > shape.warp_budget || shape.smem_bytes > 226 * 1024) return cudaErrorInvalidConfiguration;
    switch (model) {
    case 0: return lh_launch_persistent_m0(flags, args, shape, stream);
    case 1: return lh_launch_persistent_m1(flags, args, shape, stream);
    default: return lh_launch_persistent_m2(flags, args, shape, stream);
    }
}

// -------------------------------------------------------------------------------------------------
// Diagnostics
// -------------------------------------------------------------------------------------------------
namespace {
// The per-lane parameter view of column `col` (see lh_stage_body).
__device__ __forceinline__ LhLaneParams lh_lane_params(const LhDevParams& p, const double* colp, int64_t col, int64_t st, bool heat)
{
    LhLaneParams pl;
    static_cast<LhPhys&>(pl) = static_cast<const LhPhys&>(p);
    pl.mc = p.mc;
    if (colp) {
        const double* cp = colp + col;
        pl.nu = cp[LHCP_NU * st];
        pl.theta_r = cp[LHCP_THETA_R * st];
        pl.theta_r_eps = cp[LHCP_THETA_R_EPS * st];
        pl.inv_nu_thr = cp[LHCP_INV_NU_THR * st];
        pl.nu_thr = cp[LHCP_NU_THR * st];
        pl.vg_m = cp[LHCP_VG_M * st];
        pl.vg_inv_m = cp[LHCP_VG_INV_M * st];
        pl.vg_inv_n = cp[LHCP_VG_INV_N * st];
        pl.neg_inv_alpha = cp[LHCP_NEG_INV_ALPHA * st];
        pl.Ksat = cp[LHCP_KSAT * st];
        pl.inv_nu = cp[LHCP_INV_NU * st];
        pl.kappa_dry = cp[LHCP_KAPPA_DRY * st];
        if (heat) {
            pl.rho_c_ds = cp[LHCP_RHO_C_DS * st];
            pl.kersten_p1 = cp[LHCP_KERSTEN_P1 * st];
            pl.kersten_p2 = cp[LHCP_KERSTEN_P2 * st];
            pl.kersten_p3 = cp[LHCP_KERSTEN_P3 * st];
            pl.k_unfrozen = cp[LHCP_K_UNFROZEN * st];
            pl.log2_k_unfrozen = cp[LHCP_LOG2_K_UNFROZEN * st];
            pl.log2_k_frozen = cp[LHCP_LOG2_K_FROZEN * st];
            pl.om_zero = 0;
        }
        pl.k_unfrozen_minus_dry = pl.k_unfrozen - pl.kappa_dry;
    }
    return pl;
}

template <int MODEL, class P>
__device__ __forceinline__ double lh_diag_value(const P& p, const double* tab, int which, double th, double ti, double x)
{
    // K/ψ exist for every model; κ/T come from ρe_int when there is an energy model, otherwise from the prescribed T.
    if (MODEL == 0) {
        const LhCell c = lh_cell_closures<0, 3>(p, tab, th, ti, x);
        if (which == LH_DIAG_KAPPA) {
            const double nu_eff = p.nu - ti;
            const double tl = th < nu_eff ? th : nu_eff;
            LhPowArg none;
            none.t = 0.0; none.j = 0; none.be = 0;
            return lh_thermal_conductivity<true, true, false>(p, tab, tl, ti, th < nu_eff, 0.0, none);
        }
        return which == LH_DIAG_K ? c.K : which == LH_DIAG_PSI ? c.psi : c.T;
    }
    const LhCell c = lh_cell_closures<2, 3>(p, tab, th, ti, x);
    return which == LH_DIAG_K ? c.K : which == LH_DIAG_PSI ? c.psi : which == LH_DIAG_KAPPA ? c.kappa : c.T;
}

// HET: per-column parameters (the per-lane view, general log2/exp2 closures); otherwise the uniform parameter block and
// the fixed-exponent power tables, exactly what the stage kernels of a homogeneous soil evaluate.
template <int MODEL, bool HET>
__global__ void lh_diag_kernel(const __grid_constant__ LhDevParams pu, const double* __restrict__ pow_tab, int which,
                               const double* __restrict__ th, const double* __restrict__ ti, const double* __restrict__ re,
                               const double* __restrict__ T, double* __restrict__ out, int64_t n,
                               const double* __restrict__ colp, int64_t ncol_pad, int heat, const double* __restrict__ cellp)
{
    __shared__ __align__(16) double tab[LH_TAB_ALL];
    lh_stage_tables(pu, pow_tab, tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = MODEL == 0 ? T[i] : re[i];
    if constexpr (HET) {
        LhLaneParams p = lh_lane_params(pu, colp, i % ncol_pad, ncol_pad, heat != 0);
        if (cellp) lh_load_cell_params(p, cellp + i, n);          // n = nlayer * ncol_pad = the field stride
        out[i] = lh_diag_value<MODEL>(p, tab, which, th[i], ti[i], x);
    } else {
        out[i] = lh_diag_value<MODEL>(pu, tab, which, th[i], ti[i], x);
    }
}
}  // namespace

cudaError_t lh_launch_diagnostic(int model, int which, const LhDevParams& p, const double* pow_tab, const double* th,
                                 const double* ti, const double* re, const double* T, double* out,
                                 int64_t n, const double* colp, int64_t ncol_pad, int heat, const double* cellp, cudaStream_t stream)
{
    const int block = 256;
    const unsigned grid = (unsigned)((n + block - 1) / block);
    if (model == LH_MODEL_RICHARDS) {
        if (colp) LH_LAUNCH((lh_diag_kernel<0, true>), grid, block, 0, stream, p, pow_tab, which, th, ti, re, T, out, n, colp, ncol_pad, heat, cellp);
        else LH_LAUNCH((lh_diag_kernel<0, false>), grid, block, 0, stream, p, pow_tab, which, th, ti, re, T, out, n, colp, ncol_pad, heat, cellp);
    } else {
        if (colp) LH_LAUNCH((lh_diag_kernel<2, true>), grid, block, 0, stream, p, pow_tab, which, th, ti, re, T, out, n, colp, ncol_pad, heat, cellp);
        else LH_LAUNCH((lh_diag_kernel<2, false>), grid, block, 0, stream, p, pow_tab, which, th, ti, re, T, out, n, colp, ncol_pad, heat, cellp);
    }
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// PrescribedAtmosForcing: turbulent surface fluxes per column (lh_atmos.cuh)
// -------------------------------------------------------------------------------------------------
namespace {
// Matric potential at min(S_l_eff, 1) (boundary_conditions.jl:588-591) from the general water closures.
template <class P>
__device__ __forceinline__ double lh_surface_matric_potential(const P& p, const double* tab, double th, double ti)
{
    double K, psi, l_;
    LhPowArg a_;
    lh_water_closures<true, true, false, false>(p, tab, th, ti, 288.0, K, psi, l_, a_);
    const double nu_eff = p.nu - ti;
    return th < nu_eff ? psi : 0.0;            // saturated or oversaturated: S_l_eff = 1, matric_potential(1) = 0
}

template <bool HET>
__global__ void lh_atmos_flux_kernel(const __grid_constant__ LhDevParams pu, const double* __restrict__ pow_tab, const LhAtmos atm,
                                     const double* __restrict__ th_top, const double* __restrict__ ti_top,
                                     const double* __restrict__ re_top, double* __restrict__ flux_e, double* __restrict__ flux_w,
                                     int64_t ncol_pad, const double* __restrict__ colp, int heat_cols,
                                     const double* __restrict__ cellp_top, int64_t cell_fs)
{
    __shared__ __align__(16) double tab[LH_TAB_ALL];
    lh_stage_tables(pu, pow_tab, tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol_pad) return;
    const double th = th_top[col], ti = ti_top[col], re = re_top[col];
    double psi, dT;
    if constexpr (HET) {
        LhLaneParams p = lh_lane_params(pu, colp, col, ncol_pad, heat_cols != 0);
        if (cellp_top) lh_load_cell_params(p, cellp_top + col, cell_fs);
        const double nu_eff = p.nu - ti;
        dT = lh_temperature_minus_T0<true>(p, th < nu_eff ? th : nu_eff, ti, re);
        psi = lh_surface_matric_potential(p, tab, th, ti);
    } else {
        const double nu_eff = pu.nu - ti;
        dT = lh_temperature_minus_T0<true>(pu, th < nu_eff ? th : nu_eff, ti, re);
        psi = lh_surface_matric_potential(pu, tab, th, ti);
    }
    double heat, water;
    lh_atmos_fluxes(atm, psi, pu.T_0 + dT, heat, water);
    flux_e[col] = heat;
    flux_w[col] = water;
}

__global__ void lh_atmos_eval_kernel(const __grid_constant__ LhDevParams pu, const double* __restrict__ pow_tab, const LhAtmos atm,
                                     const double* __restrict__ th, const double* __restrict__ ti, const double* __restrict__ T,
                                     double* __restrict__ heat, double* __restrict__ water, int64_t n)
{
    __shared__ __align__(16) double tab[LH_TAB_ALL];
    lh_stage_tables(pu, pow_tab, tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double psi = lh_surface_matric_potential(pu, tab, th[i], ti[i]);
    double h, w;
    lh_atmos_fluxes(atm, psi, T[i], h, w);
    heat[i] = h;
    water[i] = w;
}
}  // namespace

cudaError_t lh_launch_atmos_fluxes(const LhDevParams& p, const double* pow_tab, const LhAtmos& atm, const double* th_top,
                                   const double* ti_top, const double* re_top, double* flux_e, double* flux_w, int64_t ncol_pad,
                                   const double* colp, int heat_cols, const double* cellp_top, int64_t cell_fs, cudaStream_t stream)
{
    const int block = 128;
    const unsigned grid = (unsigned)((ncol_pad + block - 1) / block);
    if (colp) LH_LAUNCH((lh_atmos_flux_kernel<true>), grid, block, 0, stream, p, pow_tab, atm, th_top, ti_top, re_top, flux_e, flux_w, ncol_pad, colp, heat_cols, cellp_top, cell_fs);
    else LH_LAUNCH((lh_atmos_flux_kernel<false>), grid, block, 0, stream, p, pow_tab, atm, th_top, ti_top, re_top, flux_e, flux_w, ncol_pad, colp, heat_cols, cellp_top, cell_fs);
    return cudaGetLastError();
}

cudaError_t lh_launch_atmos_eval(const LhDevParams& p, const double* pow_tab, const LhAtmos& atm, const double* th, const double* ti,
                                 const double* T, double* heat, double* water, int64_t n, cudaStream_t stream)
{
    LH_LAUNCH((lh_atmos_eval_kernel), (unsigned)((n + 127) / 128), 128, 0, stream, p, pow_tab, atm, th, ti, T, heat, water, n);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// Budgets: fixed-shape two-level tree (no atomics on doubles) => bitwise reproducible for a
// given shard; different shardings agree to ~1e-13 relative.
// -------------------------------------------------------------------------------------------------
namespace {
constexpr int BUDGET_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;
}

__global__ void lh_budget_partial_kernel(const double* __restrict__ th, const double* __restrict__ re,
                                         int64_t ncol, int64_t ncol_pad, int32_t nlayer,
                                         double* __restrict__ partials)
{
    __shared__ double sh[32];
    // each block owns a fixed contiguous range of columns; each thread sums whole columns
    // bottom -> top, columns strided by blockDim within the block's range.
    const int64_t per_block = (ncol + gridDim.x - 1) / gridDim.x;
    const int64_t c0 = (int64_t)blockIdx.x * per_block;
    const int64_t c1 = min(ncol, c0 + per_block);
    double sw = 0.0, se = 0.0;
    for (int64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        double cw = 0.0, ce = 0.0;
        for (int i = 0; i < nlayer; ++i) {
            cw += th[(int64_t)i * ncol_pad + c];
            ce += re[(int64_t)i * ncol_pad + c];
        }
        sw += cw; se += ce;
    }
    const double bw = block_sum(sw, sh);
    const double be = block_sum(se, sh);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = bw; partials[2 * blockIdx.x + 1] = be; }
}

__global__ void lh_budget_final_kernel(const double* __restrict__ partials, int64_t npartials, double dz,
                                       double* __restrict__ out2)
{
    __shared__ double sh[32];
    double sw = 0.0, se = 0.0;
    for (int64_t i = threadIdx.x; i < npartials; i += blockDim.x) { sw += partials[2 * i]; se += partials[2 * i + 1]; }
    const double bw = block_sum(sw, sh);
    const double be = block_sum(se, sh);
    if (threadIdx.x == 0) { out2[0] = bw * dz; out2[1] = be * dz; }
}
}  // namespace

cudaError_t lh_launch_budgets(const double* th, const double* re, int64_t ncol, int64_t ncol_pad,
                              int32_t nlayer, double dz, double* partials, int32_t npartials,
                              double* out2, cudaStream_t stream)
{
    LH_LAUNCH((lh_budget_partial_kernel), npartials, BUDGET_THREADS, 0, stream, th, re, ncol, ncol_pad, nlayer, partials);
    LH_LAUNCH((lh_budget_final_kernel), 1, BUDGET_THREADS, 0, stream, partials, npartials, dz, out2);
    return cudaGetLastError();
}

cudaError_t lh_launch_budgets_from_partials(const double* partials, int64_t npartials, double dz, double* out2, cudaStream_t stream)
{
    LH_LAUNCH((lh_budget_final_kernel), 1, BUDGET_THREADS, 0, stream, partials, npartials, dz, out2);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// Layout transforms (shared-memory tiled transposes, both sides coalesced)
// -------------------------------------------------------------------------------------------------
namespace {
// staged: [ncols][nlayer] (layer fastest), a block of columns starting at absolute column col0.
__global__ void lh_to_soa_kernel(const double* __restrict__ staged, double* __restrict__ soa, int64_t col0,
                                 int64_t ncols, int32_t nlayer, int64_t ncol_pad)
{
    __shared__ double tile[32][33];
    const int64_t cb = (int64_t)blockIdx.x * 32;   // column tile (relative)
    const int lb = blockIdx.y * 32;                // layer tile
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {   // r: column within tile, x: layer
        const int64_t c = cb + r;
        const int l = lb + threadIdx.x;
        if (c < ncols && l < nlayer) tile[r][threadIdx.x] = staged[c * nlayer + l];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {   // r: layer within tile, x: column
        const int64_t c = cb + threadIdx.x;
        const int l = lb + r;
        if (c < ncols && l < nlayer) soa[(int64_t)l * ncol_pad + col0 + c] = tile[threadIdx.x][r];
    }
}

__global__ void lh_from_soa_kernel(const double* __restrict__ soa, double* __restrict__ staged, int64_t col0,
                                   int64_t ncols, int32_t nlayer, int64_t ncol_pad)
{
    __shared__ double tile[32][33];
    const int64_t cb = (int64_t)blockIdx.x * 32;
    const int lb = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {   // r: layer, x: column
        const int64_t c = cb + threadIdx.x;
        const int l = lb + r;
        if (c < ncols && l < nlayer) tile[r][threadIdx.x] = soa[(int64_t)l * ncol_pad + col0 + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {   // r: column, x: layer
        const int64_t c = cb + r;
        const int l = lb + threadIdx.x;
        if (c < ncols && l < nlayer) staged[c * nlayer + l] = tile[threadIdx.x][r];
    }
}

__global__ void lh_fill_profile_kernel(const double* __restrict__ profile, double* __restrict__ soa,
                                       int32_t nlayer, int64_t ncol_pad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nlayer * ncol_pad) return;
    soa[i] = profile[i / ncol_pad];
}

__global__ void lh_fill_padding_kernel(double* __restrict__ soa, int64_t ncol, int64_t ncol_pad, int32_t nlayer)
{
    const int64_t npad = ncol_pad - ncol;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npad * nlayer) return;
    const int64_t l = i / npad, c = ncol + i % npad;
    soa[l * ncol_pad + c] = soa[l * ncol_pad + ncol - 1];
}

__global__ void lh_any_nonzero_kernel(const double* __restrict__ x, int64_t n, int* flag)
{
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        any |= !(x[i] == 0.0);
    if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1;
}

__global__ void lh_count_nonfinite_kernel(const double* __restrict__ x, int64_t n, unsigned long long* count)
{
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!isfinite(x[i])) ++local;
    if (local) atomicAdd(count, local);
}
}  // namespace

cudaError_t lh_launch_to_soa(const double* staged, double* soa, int64_t col0, int64_t ncols,
                             int32_t nlayer, int64_t ncol_pad, cudaStream_t stream)
{
    dim3 block(32, 8), grid((unsigned)((ncols + 31) / 32), (unsigned)((nlayer + 31) / 32));
    LH_LAUNCH((lh_to_soa_kernel), grid, block, 0, stream, staged, soa, col0, ncols, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_from_soa(const double* soa, double* staged, int64_t col0, int64_t ncols,
                               int32_t nlayer, int64_t ncol_pad, cudaStream_t stream)
{
    dim3 block(32, 8), grid((unsigned)((ncols + 31) / 32), (unsigned)((nlayer + 31) / 32));
    LH_LAUNCH((lh_from_soa_kernel), grid, block, 0, stream, soa, staged, col0, ncols, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_fill_profile(const double* profile, double* soa, int32_t nlayer, int64_t ncol_pad,
                                   cudaStream_t stream)
{
    const int64_t n = (int64_t)nlayer * ncol_pad;
    LH_LAUNCH((lh_fill_profile_kernel), (unsigned)((n + 255) / 256), 256, 0, stream, profile, soa, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_fill_padding(double* soa, int64_t ncol, int64_t ncol_pad, int32_t nlayer,
                                   cudaStream_t stream)
{
    const int64_t n = (ncol_pad - ncol) * nlayer;
    if (n <= 0) return cudaSuccess;
    LH_LAUNCH((lh_fill_padding_kernel), (unsigned)((n + 255) / 256), 256, 0, stream, soa, ncol, ncol_pad, nlayer);
    return cudaGetLastError();
}

namespace {
__global__ void lh_eval_math_kernel(const __grid_constant__ LhDevParams p, int fn, const double* __restrict__ x,
                                    double* __restrict__ y, int64_t n)
{
    __shared__ __align__(16) double tab[LH_TAB_DOUBLES];
    lh_stage_tables(p, nullptr, tab, threadIdx.x, blockDim.x);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double r;
    switch (fn) {
    case LH_MATH_LOG2: r = lh_log2(p.mc, tab, v); break;
    case LH_MATH_EXP2: r = lh_exp2(p.mc, tab, v); break;
    case LH_MATH_EXP2M1: r = lh_exp2m1(p.mc, tab, v); break;
    case LH_MATH_SQRT: r = lh_sqrt(v); break;
    case LH_MATH_RSQRT: r = lh_rsqrt(v); break;
    case LH_MATH_RCP: r = lh_rcp(v); break;
    case LH_MATH_RCP_SEED: r = lh_rcp_seed(v); break;
    case LH_MATH_RSQRT_SEED: r = lh_rsqrt_seed(v); break;
    default: r = lh_div(v, x[n + i]); break;
    }
    y[i] = r;
}
}  // namespace

cudaError_t lh_launch_eval_math(const LhDevParams& p, int fn, const double* x, double* y, int64_t n, cudaStream_t stream)
{
    LH_LAUNCH((lh_eval_math_kernel), (unsigned)((n + 255) / 256), 256, 0, stream, p, fn, x, y, n);
    return cudaGetLastError();
}

cudaError_t lh_launch_any_nonzero(const double* x, int64_t n, int* flag, cudaStream_t stream)
{
    LH_LAUNCH((lh_any_nonzero_kernel), 592, 256, 0, stream, x, n, flag);
    return cudaGetLastError();
}

cudaError_t lh_launch_count_nonfinite(const double* soa, int64_t n, unsigned long long* count,
                                      cudaStream_t stream)
{
    LH_LAUNCH((lh_count_nonfinite_kernel), 592, 256, 0, stream, soa, n, count);
    return cudaGetLastError();
}
