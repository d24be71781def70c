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

#include "lh_soil.h"

namespace {

// Quantities exchanged across a chunk face / kept in the sliding window.
template <int MODEL> struct Q;
template <> struct Q<0> { double K, h; };                       // Richards
template <> struct Q<1> { double kappa, T; };                   // heat
template <> struct Q<2> { double K, h, kappa, T, eK; };         // coupled

template <int MODEL> struct NQ { static constexpr int value = sizeof(Q<MODEL>) / sizeof(double); };

struct Flux { double w, e; };

template <int MODEL>
__device__ __forceinline__ void q_store(double* sm, const Q<MODEL>& q)
{
    // sm points at this lane's slot; quantities are strided by 32 lanes (conflict-free)
    if constexpr (MODEL == 0) { sm[0] = q.K; sm[32] = q.h; }
    else if constexpr (MODEL == 1) { sm[0] = q.kappa; sm[32] = q.T; }
    else { sm[0] = q.K; sm[32] = q.h; sm[64] = q.kappa; sm[96] = q.T; sm[128] = q.eK; }
}

template <int MODEL>
__device__ __forceinline__ Q<MODEL> q_load(const double* sm)
{
    Q<MODEL> q;
    if constexpr (MODEL == 0) { q.K = sm[0]; q.h = sm[32]; }
    else if constexpr (MODEL == 1) { q.kappa = sm[0]; q.T = sm[32]; }
    else { q.K = sm[0]; q.h = sm[32]; q.kappa = sm[64]; q.T = sm[96]; q.eK = sm[128]; }
    return q;
}

// Interior face between cell `lo` (below) and `hi` (above).
//   water  right_hand_side.jl:181/:358   -interpc2f(K) * gradc2f(h)
//   energy :259 / :361-365               -interpc2f(κ) * gradc2f(T) - interpc2f(ρe_int_l K) * gradc2f(h)
template <int MODEL>
__device__ __forceinline__ Flux face_flux(const LhDevParams& p, const Q<MODEL>& lo, const Q<MODEL>& hi)
{
    Flux f;
    f.w = 0.0; f.e = 0.0;
    if constexpr (MODEL == 0) {
        const double gh = (hi.h - lo.h) * p.inv_dz;
        f.w = -(0.5 * (lo.K + hi.K)) * gh;
    } else if constexpr (MODEL == 1) {
        const double gT = (hi.T - lo.T) * p.inv_dz;
        f.e = -(0.5 * (lo.kappa + hi.kappa)) * gT;
    } else {
        const double gh = (hi.h - lo.h) * p.inv_dz;
        const double gT = (hi.T - lo.T) * p.inv_dz;
        f.w = -(0.5 * (lo.K + hi.K)) * gh;
        f.e = -(0.5 * (lo.kappa + hi.kappa)) * gT - (0.5 * (lo.eK + hi.eK)) * gh;
    }
    return f;
}

// boundary_fluxes(X, bc::SoilComponentBC, face, ...) boundary_conditions.jl:470-489 for one face.
// (th, ti) raw centre values, `c` the centre closures (c.T is the centre temperature).
template <int MODEL>
__device__ __forceinline__ Flux boundary_flux(const LhDevParams& p, int e_kind, int h_kind, double val_e,
                                              double val_h, bool is_bottom, double th, double ti,
                                              const LhCell& c)
{
    Flux f;
    f.w = 0.0; f.e = 0.0;
    // X_cf face values (:218-228, :241-288): Dirichlet overrides, energy first then hydrology
    const double th_f = (MODEL != 1 && h_kind == LH_BC_DIRICHLET) ? val_h : th;
    const double T_f = (MODEL != 0 && e_kind == LH_BC_DIRICHLET) ? val_e : c.T;
    if constexpr (MODEL != 1) {
        if (h_kind == LH_BC_FLUX) {
            f.w = val_h;                                                     // :295-301
        } else if (h_kind == LH_BC_FREE_DRAINAGE) {
            f.w = -c.K;                                                      // :328-356 (K of the centre cell)
        } else if (h_kind == LH_BC_DIRICHLET) {                              // :371-401
            double K_f, psi_f, l_, s_;
            lh_water_closures(p, th_f, ti, T_f, K_f, psi_f, l_, s_);
            double flux = (-K_f * (psi_f - c.psi + p.half_dz)) * p.inv_half_dz;
            f.w = is_bottom ? -flux : flux;
        }
    }
    if constexpr (MODEL != 0) {
        if (e_kind == LH_BC_FLUX) {
            f.e = val_e;
        } else if (e_kind == LH_BC_DIRICHLET) {                              // :416-444
            const double kappa_f = lh_face_kappa(p, th_f, ti);
            double flux = (-kappa_f * (T_f - c.T)) * p.inv_half_dz;
            f.e = is_bottom ? -flux : flux;
        }
    }
    return f;
}

template <int STAGE>
__device__ __forceinline__ double stage_base(double v, double u0)
{
    if constexpr (STAGE == 2) return fma(3.0, u0, v);       // 3 u0 + u1
    else if constexpr (STAGE == 3) return fma(2.0, v, u0);  // u0 + 2 u2
    else return v;
}

template <int STAGE>
__device__ __forceinline__ double stage_out(double base, double k, double dt)
{
    if constexpr (STAGE == 0) return k;
    else if constexpr (STAGE == 1) return fma(dt, k, base);
    else if constexpr (STAGE == 2) return 0.25 * fma(dt, k, base);
    else return (1.0 / 3.0) * fma(2.0 * dt, k, base);
}

struct Base { double th, re; };

template <int MODEL, int STAGE>
__global__ void __launch_bounds__(512, 1)
lh_soil_stage_kernel(const __grid_constant__ LhKernelArgs A)
{
    extern __shared__ double smem[];
    constexpr int NQv = NQ<MODEL>::value;
    const LhDevParams& p = A.p;
    const int lane = threadIdx.x, w = threadIdx.y, g = threadIdx.z;
    const int W = blockDim.y;
    const int64_t col = ((int64_t)blockIdx.x * blockDim.z + g) * 32 + lane;
    const bool valid = col < A.ncol_pad;      // whole column groups are valid or not (ncol_pad % 32 == 0)
    const int n = A.nlayer;
    const int a = w * A.Lc;
    const int b = min(n, a + A.Lc);
    const int64_t stride = A.ncol_pad;

    // shared slots: [g][w][bot|top][NQ][32]
    double* sm_bot = smem + ((size_t)(g * W + w) * 2 + 0) * NQv * 32 + lane;
    double* sm_top = smem + ((size_t)(g * W + w) * 2 + 1) * NQv * 32 + lane;

    Q<MODEL> prev;            // closures of the cell below the current one
    Base base_prev;           // stage-combine base of that cell
    Base base_first;
    Flux F_below, F_first_up, F_bc_bot, F_bc_top;
    F_below.w = F_below.e = 0.0;
    F_first_up = F_below; F_bc_bot = F_below; F_bc_top = F_below;
    base_prev.th = base_prev.re = 0.0;
    base_first = base_prev;

    if (valid && a < n) {
        const double* pth = A.in_th + col;
        const double* pti = A.in_ti + col;
        const double* pre = A.in_re + col;
        const double* pT = A.aux_T + col;
        const double* p0th = A.u0_th + col;
        const double* p0re = A.u0_re + col;
        double* oth = A.out_th + col;
        double* ore = A.out_re + col;
        const bool need_T = (MODEL == 0) && p.visc_on;

        // software prefetch of the next layer's raw values
        double n_th = pth[(int64_t)a * stride];
        double n_ti = pti[(int64_t)a * stride];
        double n_x = (MODEL != 0) ? pre[(int64_t)a * stride] : (need_T ? pT[(int64_t)a * stride] : 288.0);
        double n_u0th = 0.0, n_u0re = 0.0;
        if constexpr (STAGE >= 2) {
            if constexpr (MODEL != 1) n_u0th = p0th[(int64_t)a * stride];
            if constexpr (MODEL != 0) n_u0re = p0re[(int64_t)a * stride];
        }

        for (int i = a; i < b; ++i) {
            const double th = n_th, ti = n_ti, x = n_x, u0th = n_u0th, u0re = n_u0re;
            if (i + 1 < b) {
                const int64_t o = (int64_t)(i + 1) * stride;
                n_th = pth[o];
                n_ti = pti[o];
                if constexpr (MODEL != 0) n_x = pre[o];
                else if (need_T) n_x = pT[o];
                if constexpr (STAGE >= 2) {
                    if constexpr (MODEL != 1) n_u0th = p0th[o];
                    if constexpr (MODEL != 0) n_u0re = p0re[o];
                }
            }
            // ---- pointwise closures of cell i
            const LhCell c = lh_cell_closures<MODEL>(p, th, ti, x);
            Q<MODEL> cur;
            if constexpr (MODEL == 0) { cur.K = c.K; cur.h = c.psi + A.zc[i]; }
            else if constexpr (MODEL == 1) { cur.kappa = c.kappa; cur.T = c.T; }
            else {
                cur.K = c.K; cur.h = c.psi + A.zc[i]; cur.kappa = c.kappa; cur.T = c.T;
                cur.eK = (p.rhocp_l * (c.T - p.T_0)) * c.K;                 // ρe_int_l * K (:306, :364)
            }
            Base base_cur;
            base_cur.th = (MODEL != 1) ? stage_base<STAGE>(th, u0th) : 0.0;
            base_cur.re = (MODEL != 0) ? stage_base<STAGE>(x, u0re) : 0.0;

            if (i == 0)
                F_bc_bot = boundary_flux<MODEL>(p, A.bot_e_kind, A.bot_h_kind, A.bcv[LH_BCV_BOTTOM_ENERGY],
                                                A.bcv[LH_BCV_BOTTOM_HYDROLOGY], true, th, ti, c);
            if (i == n - 1)
                F_bc_top = boundary_flux<MODEL>(p, A.top_e_kind, A.top_h_kind, A.bcv[LH_BCV_TOP_ENERGY],
                                                A.bcv[LH_BCV_TOP_HYDROLOGY], false, th, ti, c);

            if (i == a) {
                q_store<MODEL>(sm_bot, cur);
                base_first = base_cur;
            } else {
                const Flux F = face_flux<MODEL>(p, prev, cur);
                if (i - 1 == a) {
                    F_first_up = F;        // first cell of the chunk waits for the face below it
                } else {
                    const int64_t o = (int64_t)(i - 1) * stride;
                    if constexpr (MODEL != 1) oth[o] = stage_out<STAGE>(base_prev.th, -(F.w - F_below.w) * p.inv_dz, A.dt);
                    if constexpr (MODEL != 0) ore[o] = stage_out<STAGE>(base_prev.re, -(F.e - F_below.e) * p.inv_dz, A.dt);
                }
                F_below = F;
            }
            prev = cur;
            base_prev = base_cur;
        }
        q_store<MODEL>(sm_top, prev);
    }
    __syncthreads();
    if (valid && a < n) {
        double* oth = A.out_th + col;
        double* ore = A.out_re + col;
        const Q<MODEL> first = q_load<MODEL>(sm_bot);
        Flux F_lo, F_hi;
        if (a == 0) F_lo = F_bc_bot;
        else F_lo = face_flux<MODEL>(p, q_load<MODEL>(sm_bot - 1 * NQv * 32), first);   // top slot of chunk w-1
        if (b == n) F_hi = F_bc_top;
        else F_hi = face_flux<MODEL>(p, prev, q_load<MODEL>(sm_top + 1 * NQv * 32));    // bot slot of chunk w+1
        const int64_t oa = (int64_t)a * stride, ob = (int64_t)(b - 1) * stride;
        if (b - a == 1) {
            if constexpr (MODEL != 1) oth[oa] = stage_out<STAGE>(base_first.th, -(F_hi.w - F_lo.w) * p.inv_dz, A.dt);
            if constexpr (MODEL != 0) ore[oa] = stage_out<STAGE>(base_first.re, -(F_hi.e - F_lo.e) * p.inv_dz, A.dt);
        } else {
            if constexpr (MODEL != 1) {
                oth[oa] = stage_out<STAGE>(base_first.th, -(F_first_up.w - F_lo.w) * p.inv_dz, A.dt);
                oth[ob] = stage_out<STAGE>(base_prev.th, -(F_hi.w - F_below.w) * p.inv_dz, A.dt);
            }
            if constexpr (MODEL != 0) {
                ore[oa] = stage_out<STAGE>(base_first.re, -(F_first_up.e - F_lo.e) * p.inv_dz, A.dt);
                ore[ob] = stage_out<STAGE>(base_prev.re, -(F_hi.e - F_below.e) * p.inv_dz, A.dt);
            }
        }
    }
}

template <int MODEL>
cudaError_t launch_model(int stage, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    dim3 block(32, s.W, s.G);
    dim3 grid((unsigned)s.nblocks);
    switch (stage) {
    case 0: lh_soil_stage_kernel<MODEL, 0><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 1: lh_soil_stage_kernel<MODEL, 1><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 2: lh_soil_stage_kernel<MODEL, 2><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 3: lh_soil_stage_kernel<MODEL, 3><<<grid, block, s.smem_bytes, stream>>>(args); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

LhLaunchShape lh_choose_shape(int model, int64_t ncol_pad, int32_t nlayer, int sm_count)
{
    LhLaunchShape s;
    const int64_t groups = ncol_pad / 32;
    // Chunk length: 16 layers per thread amortises the chunk-face exchange; shorter chunks when
    // there are too few columns to fill the machine (>= ~8 warps per SM wanted).
    int Lc = 16;
    const int64_t want_warps = (int64_t)sm_count * 8;
    while (Lc > 2 && groups * ((nlayer + Lc - 1) / Lc) < want_warps) Lc >>= 1;
    int W = (nlayer + Lc - 1) / Lc;
    if (W > 16) { W = 16; }           // <= 512 threads per block (128 registers per thread)
    Lc = (nlayer + W - 1) / W;
    W = (nlayer + Lc - 1) / Lc;          // no empty chunks
    int G = 1;
    while (W * G * 2 <= 8 && (int64_t)G * 2 <= groups) G *= 2;   // at least ~8 warps per block when possible
    s.Lc = Lc; s.W = W; s.G = G;
    s.nblocks = (groups + G - 1) / G;
    const int nq = model == LH_MODEL_COUPLED ? 5 : 2;
    s.smem_bytes = (size_t)G * W * 2 * nq * 32 * sizeof(double);
    return s;
}

cudaError_t lh_launch_stage(int model, int stage, const LhKernelArgs& args, const LhLaunchShape& shape,
                            cudaStream_t stream)
{
    if (model < 0 || model > 2) return cudaErrorInvalidValue;
    if (shape.W * shape.G * 32 > 512 || shape.smem_bytes > 48 * 1024) return cudaErrorInvalidConfiguration;
    switch (model) {
    case 0: return launch_model<0>(stage, args, shape, stream);
    case 1: return launch_model<1>(stage, args, shape, stream);
    default: return launch_model<2>(stage, args, shape, stream);
    }
}

// -------------------------------------------------------------------------------------------------
// Diagnostics
// -------------------------------------------------------------------------------------------------
namespace {
template <int MODEL>
__global__ void lh_diag_kernel(const __grid_constant__ LhDevParams p, int which, const double* __restrict__ th,
                               const double* __restrict__ ti, const double* __restrict__ re,
                               const double* __restrict__ T, double* __restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // K/ψ exist for every model; κ/T come from ρe_int when there is an energy model, otherwise
    // from the prescribed T.
    double v;
    if (MODEL == 0) {
        const LhCell c = lh_cell_closures<0>(p, th[i], ti[i], T[i]);
        if (which == LH_DIAG_KAPPA) {
            const double nu_eff = p.nu - ti[i];
            const double tl = th[i] < nu_eff ? th[i] : nu_eff;
            v = lh_thermal_conductivity(p, tl, ti[i], -1.0, 0.0);
        } else v = which == LH_DIAG_K ? c.K : which == LH_DIAG_PSI ? c.psi : c.T;
    } else {
        const LhCell c = lh_cell_closures<2>(p, th[i], ti[i], re[i]);
        v = which == LH_DIAG_K ? c.K : which == LH_DIAG_PSI ? c.psi : which == LH_DIAG_KAPPA ? c.kappa : c.T;
    }
    out[i] = v;
}
}  // namespace

cudaError_t lh_launch_diagnostic(int model, int which, const LhDevParams& p, const double* th,
                                 const double* ti, const double* re, const double* T, double* out,
                                 int64_t n, cudaStream_t stream)
{
    const int block = 256;
    const unsigned grid = (unsigned)((n + block - 1) / block);
    if (model == LH_MODEL_RICHARDS) lh_diag_kernel<0><<<grid, block, 0, stream>>>(p, which, th, ti, re, T, out, n);
    else lh_diag_kernel<2><<<grid, block, 0, stream>>>(p, which, th, ti, re, T, out, n);
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

__global__ void lh_budget_final_kernel(const double* __restrict__ partials, int32_t npartials, double dz,
                                       double* __restrict__ out2)
{
    __shared__ double sh[32];
    double sw = 0.0, se = 0.0;
    for (int i = threadIdx.x; i < npartials; i += blockDim.x) { sw += partials[2 * i]; se += partials[2 * i + 1]; }
    const double bw = block_sum(sw, sh);
    const double be = block_sum(se, sh);
    if (threadIdx.x == 0) { out2[0] = bw * dz; out2[1] = be * dz; }
}
}  // namespace

cudaError_t lh_launch_budgets(const double* th, const double* re, int64_t ncol, int64_t ncol_pad,
                              int32_t nlayer, double dz, double* partials, int32_t npartials,
                              double* out2, cudaStream_t stream)
{
    lh_budget_partial_kernel<<<npartials, BUDGET_THREADS, 0, stream>>>(th, re, ncol, ncol_pad, nlayer, partials);
    lh_budget_final_kernel<<<1, BUDGET_THREADS, 0, stream>>>(partials, npartials, dz, out2);
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
    lh_to_soa_kernel<<<grid, block, 0, stream>>>(staged, soa, col0, ncols, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_from_soa(const double* soa, double* staged, int64_t col0, int64_t ncols,
                               int32_t nlayer, int64_t ncol_pad, cudaStream_t stream)
{
    dim3 block(32, 8), grid((unsigned)((ncols + 31) / 32), (unsigned)((nlayer + 31) / 32));
    lh_from_soa_kernel<<<grid, block, 0, stream>>>(soa, staged, col0, ncols, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_fill_profile(const double* profile, double* soa, int32_t nlayer, int64_t ncol_pad,
                                   cudaStream_t stream)
{
    const int64_t n = (int64_t)nlayer * ncol_pad;
    lh_fill_profile_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(profile, soa, nlayer, ncol_pad);
    return cudaGetLastError();
}

cudaError_t lh_launch_fill_padding(double* soa, int64_t ncol, int64_t ncol_pad, int32_t nlayer,
                                   cudaStream_t stream)
{
    const int64_t n = (ncol_pad - ncol) * nlayer;
    if (n <= 0) return cudaSuccess;
    lh_fill_padding_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(soa, ncol, ncol_pad, nlayer);
    return cudaGetLastError();
}

cudaError_t lh_launch_count_nonfinite(const double* soa, int64_t n, unsigned long long* count,
                                      cudaStream_t stream)
{
    lh_count_nonfinite_kernel<<<592, 256, 0, stream>>>(soa, n, count);
    return cudaGetLastError();
}
