// lh_stage_kernel.cuh — the fused soil RHS (+ SSPRK33 stage) kernel template and its launcher.
// Included by one translation unit per model (lh_kernels_m0/m1/m2.cu) so the 3 x 4 x 8 kernel variants
// compile in parallel.  See lh_kernels.cu for the design notes.
#pragma once

#include "lh_kernels.cuh"

#include "lh_soil.h"

namespace {

// Quantities exchanged across a chunk face / kept in the sliding window.
template <int MODEL> struct Q;
// The hydraulic head h = psi + z_c (right_hand_side.jl:166/:313) only ever appears as a difference between
// vertically adjacent centres, and those are exactly dz apart on the uniform mesh: the window carries psi
// and the face flux adds dz, so z_c is never loaded (and the boundary flux reads psi directly).
template <> struct Q<0> { double K, psi; };                     // Richards
template <> struct Q<1> { double kappa, T; };                   // heat
template <> struct Q<2> { double K, psi, kappa, T, eK; };       // coupled

template <int MODEL> struct NQ { static constexpr int value = sizeof(Q<MODEL>) / sizeof(double); };

struct Flux { double w, e; };

template <int MODEL>
__device__ __forceinline__ void q_store(double* sm, const Q<MODEL>& q)
{
    // sm points at this lane's slot; quantities are strided by 32 lanes (conflict-free)
    if constexpr (MODEL == 0) { sm[0] = q.K; sm[32] = q.psi; }
    else if constexpr (MODEL == 1) { sm[0] = q.kappa; sm[32] = q.T; }
    else { sm[0] = q.K; sm[32] = q.psi; sm[64] = q.kappa; sm[96] = q.T; sm[128] = q.eK; }
}

template <int MODEL>
__device__ __forceinline__ Q<MODEL> q_load(const double* sm)
{
    Q<MODEL> q;
    if constexpr (MODEL == 0) { q.K = sm[0]; q.psi = sm[32]; }
    else if constexpr (MODEL == 1) { q.kappa = sm[0]; q.T = sm[32]; }
    else { q.K = sm[0]; q.psi = sm[32]; q.kappa = sm[64]; q.T = sm[96]; q.eK = sm[128]; }
    return q;
}

// Interior face between cell `lo` (below) and `hi` (above).
//   water  right_hand_side.jl:181/:358   -interpc2f(K) * gradc2f(h)
//   energy :259 / :361-365               -interpc2f(κ) * gradc2f(T) - interpc2f(ρe_int_l K) * gradc2f(h)
template <int MODEL>
__device__ __forceinline__ Flux face_flux(const LhDevParams& p, const Q<MODEL>& lo, const Q<MODEL>& hi)
{
    // -interp(a) * grad(b) = -(a_lo + a_hi)/2 * (b_hi - b_lo)/dz = c (a_lo + a_hi)(b_hi - b_lo),  c = -1/(2 dz)
    Flux f;
    f.w = 0.0; f.e = 0.0;
    const double c = p.neg_half_inv_dz;
    if constexpr (MODEL == 0) {
        f.w = (c * (lo.K + hi.K)) * ((hi.psi - lo.psi) + p.dz);
    } else if constexpr (MODEL == 1) {
        f.e = (c * (lo.kappa + hi.kappa)) * (hi.T - lo.T);
    } else {
        const double dh = (hi.psi - lo.psi) + p.dz;
        f.w = (c * (lo.K + hi.K)) * dh;
        f.e = c * fma(lo.kappa + hi.kappa, hi.T - lo.T, (lo.eK + hi.eK) * dh);
    }
    return f;
}

// boundary_fluxes(X, bc::SoilComponentBC, face, ...) boundary_conditions.jl:470-489 for one face.
// (th, ti) raw centre values, `c` the centre closures (c.T is the centre temperature).
template <int MODEL, int FLAGS>
__device__ __forceinline__ Flux boundary_flux(const LhDevParams& p, const double* __restrict__ tab, int e_kind,
                                              int h_kind, double val_e, double val_h, bool is_bottom,
                                              double th, double ti, const LhCell& c)
{
    constexpr bool ICE = (FLAGS & LH_FLAG_ICE) != 0, GEN = (FLAGS & LH_FLAG_GEN) != 0;
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
            double K_f, psi_f, l_;
            lh_water_closures<ICE, GEN, (FLAGS & LH_FLAG_VG2) != 0, false>(p, tab, th_f, ti, T_f, K_f, psi_f, l_);
            double flux = (-K_f * (psi_f - c.psi + p.half_dz)) * p.inv_half_dz;
            f.w = is_bottom ? -flux : flux;
        }
    }
    if constexpr (MODEL != 0) {
        if (e_kind == LH_BC_FLUX) {
            f.e = val_e;
        } else if (e_kind == LH_BC_DIRICHLET) {                              // :416-444
            const double kappa_f = lh_face_kappa<FLAGS>(p, tab, th_f, ti);
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

// Stage combine with the flux-form divergence folded in:  k = -(F_hi - F_lo)/dz and
//   stage 0: k     1: base + dt k     2: (base + dt k)/4     3: (base + 2 dt k)/3
// are evaluated as  s (base + cdt (F_hi - F_lo))  with cdt = -dt/dz (stage 3: -2 dt/dz; stage 0: -1/dz).
template <int STAGE>
__device__ __forceinline__ double stage_out(double base, double dF, double cdt)
{
    if constexpr (STAGE == 0) return cdt * dF;
    else if constexpr (STAGE == 1) return fma(cdt, dF, base);
    else if constexpr (STAGE == 2) return 0.25 * fma(cdt, dF, base);
    else return (1.0 / 3.0) * fma(cdt, dF, base);
}

struct Base { double th, re; };
struct Raw { double th, ti, x, u0th, u0re; };
template <int MODEL> struct Cell { Q<MODEL> q; Base base; };

// Shared-memory slot of one (column group, chunk): [bot: NQ][top: NQ][pending: 4], each x32 lanes.
// Input ring (cp.async): RING_DEPTH cells x RING_FIELDS fields x 32 lanes per warp.
constexpr int RING_DEPTH = 4, RING_FIELDS = 5, RING_DOUBLES = RING_DEPTH * RING_FIELDS * 32;
constexpr uint32_t RING_CELL_BYTES = RING_FIELDS * 256;

template <int MODEL> struct Slot { static constexpr int NQv = NQ<MODEL>::value; static constexpr int doubles = (2 * NQv + 4) * 32; };

__device__ __forceinline__ void lh_cp8(uint32_t dst, const double* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
template <int OFF>
__device__ __forceinline__ double lh_lds(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}

template <int MODEL, int STAGE, int FLAGS>
__global__ void __launch_bounds__(LhBounds<MODEL>::max_threads, LhBounds<MODEL>::min_blocks)
lh_soil_stage_kernel(const __grid_constant__ LhKernelArgs A)
{
    extern __shared__ __align__(16) double smem[];
    constexpr int NQv = NQ<MODEL>::value;
    constexpr bool ICE = (FLAGS & LH_FLAG_ICE) != 0;
    const LhDevParams& p = A.p;
    // A warp is one (column group g, chunk w) pair, so w, g and everything derived from them (loop
    // bounds, `active`) are warp-uniform.  ptxas cannot see that from threadIdx.y/z and then treats the
    // layer loop as divergent: no uniform-register operands inside it, every constant copied to vector
    // registers (128 registers, 260 instructions per cell).  Reading them through a lane-0 shuffle
    // proves uniformity: 100 registers, 230 instructions per cell, the coefficients stay in UR.
    const int lane = threadIdx.x;
    const int w = __shfl_sync(0xffffffffu, (int)threadIdx.y, 0), g = __shfl_sync(0xffffffffu, (int)threadIdx.z, 0);
    const int W = blockDim.y;
    const int64_t col0 = ((int64_t)blockIdx.x * blockDim.z + g) * 32;
    const int64_t col = col0 + lane;
    const bool valid = col0 < A.ncol_pad;     // whole column groups are valid or not (ncol_pad % 32 == 0)
    const int n = A.nlayer;
    const int a = w * A.Lc;
    const int b = min(n, a + A.Lc);
    const int64_t stride = A.ncol_pad;
    const bool active = valid && a < n;
    const bool need_T = (MODEL == 0) && (FLAGS & LH_FLAG_GEN) && p.visc_on;

    // shared memory: the exp2 / log2 tables, then per (g, w): one Slot (chunk-face exchange) and one input ring
    const double* tab = smem;
    lh_stage_tables(p, smem, (threadIdx.z * blockDim.y + threadIdx.y) * 32 + threadIdx.x, blockDim.x * blockDim.y * blockDim.z);
    double* warp_base = smem + LH_TAB_DOUBLES + (size_t)(g * W + w) * (Slot<MODEL>::doubles + RING_DOUBLES);
    double* slot = warp_base + lane;
    double* sm_bot = slot;                               // Q of the chunk's first cell
    double* sm_top = slot + NQv * 32;                    // Q of the chunk's last cell
    double* sm_pend = slot + 2 * NQv * 32;               // base.th, base.re, F_first_up.w, F_first_up.e
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(warp_base + Slot<MODEL>::doubles + lane);   // [RING_DEPTH][RING_FIELDS][32]
    __syncthreads();

    // Stage 2 reads and writes V, stage 3 reads U (as u0) and writes it (lh_soil_api.cu fill_args): the
    // store addresses are formed from the same base registers as the loads.
    const double* pth = A.in_th + col;
    const double* pti = A.in_ti + col;
    const double* pre = A.in_re + col;
    const double* pT = A.aux_T + col;
    const double* p0th = A.u0_th + col;
    const double* p0re = A.u0_re + col;
    double* oth = STAGE == 2 ? const_cast<double*>(pth) : STAGE == 3 ? const_cast<double*>(p0th) : A.out_th + col;
    double* ore = STAGE == 2 ? const_cast<double*>(pre) : STAGE == 3 ? const_cast<double*>(p0re) : A.out_re + col;

    // Input pipeline.  The raw values of cell i travel global -> shared with cp.async (LDGSTS): no
    // registers are held while the copy is in flight (a register software pipeline was tried: ptxas
    // sinks such loads down to the next possibly-aliasing store and spills them; prefetch.global.L1
    // was tried too and only reaches L2 on this part — 3 % L1 hit rate, profiles/r01_d_*).  Each lane
    // copies and later reads back only ITS OWN 8 bytes, so cp.async.wait_group is the only
    // synchronisation needed.  Ring of RING_DEPTH cells; one commit group per cell, always committed
    // (empty past the end of the chunk) so that wait_group's constant stays valid.  Cell a + k sits in
    // ring cell k mod 4; the layer loop handles two cells per trip, so its ring cells are a PAIR
    // (0,1) or (2,3) and every ring address is `pair base + constant`; the pair base toggles with one
    // subtraction per trip.
    int64_t o_ld = (int64_t)a * stride;       // element offset of the next cell to request
    auto issue = [&](uint32_t d, int i) {     // request cell i (whose offset o_ld is) into ring address d
        if (i < b) {
            lh_cp8(d, pth + o_ld);
            if (ICE) lh_cp8(d + 256, pti + o_ld);
            if (MODEL != 0) lh_cp8(d + 512, pre + o_ld);
            else if (need_T) lh_cp8(d + 512, pT + o_ld);
            if constexpr (STAGE >= 2) {
                if constexpr (MODEL != 1) lh_cp8(d + 768, p0th + o_ld);
                if constexpr (MODEL != 0) lh_cp8(d + 1024, p0re + o_ld);
            }
            o_ld += stride;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // wait for the oldest outstanding cell, read it from ring address d, then request cell i_next into d_next
    auto load_raw = [&](uint32_t d, uint32_t d_next, int i_next) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RING_DEPTH - 2) : "memory");
        Raw r;
        r.th = lh_lds<0>(d);
        r.ti = ICE ? lh_lds<256>(d) : 0.0;
        r.x = (MODEL != 0 || need_T) ? lh_lds<512>(d) : 288.0;
        r.u0th = 0.0; r.u0re = 0.0;
        if constexpr (STAGE >= 2) {
            if constexpr (MODEL != 1) r.u0th = lh_lds<768>(d);
            if constexpr (MODEL != 0) r.u0re = lh_lds<1024>(d);
        }
        issue(d_next, i_next);
        return r;
    };
    auto eval = [&](const Raw& r) {
        const LhCell c = lh_cell_closures<MODEL, FLAGS>(p, tab, r.th, r.ti, r.x);
        Cell<MODEL> o;
        if constexpr (MODEL == 0) { o.q.K = c.K; o.q.psi = c.psi; }
        else if constexpr (MODEL == 1) { o.q.kappa = c.kappa; o.q.T = c.T; }
        else {
            o.q.K = c.K; o.q.psi = c.psi; o.q.kappa = c.kappa; o.q.T = c.T;
            o.q.eK = (p.rhocp_l * c.dT) * c.K;                               // ρe_int_l * K (:306, :364)
        }
        o.base.th = (MODEL != 1) ? stage_base<STAGE>(r.th, r.u0th) : 0.0;
        o.base.re = (MODEL != 0) ? stage_base<STAGE>(r.x, r.u0re) : 0.0;
        return o;
    };
    const double cdt = STAGE == 0 ? -p.inv_dz : (STAGE == 3 ? -2.0 * (A.dt * p.inv_dz) : -(A.dt * p.inv_dz));
    int64_t o_st = (int64_t)(a + 1) * stride;   // element offset of the next cell to store (cell a is stored last)
    auto write_next = [&](const Base& base, const Flux& lo, const Flux& hi) {
        if constexpr (MODEL != 1) oth[o_st] = stage_out<STAGE>(base.th, hi.w - lo.w, cdt);
        if constexpr (MODEL != 0) ore[o_st] = stage_out<STAGE>(base.re, hi.e - lo.e, cdt);
        o_st += stride;
    };
    auto write_at = [&](int i, const Base& base, const Flux& lo, const Flux& hi) {
        const int64_t o = (int64_t)i * stride;
        if constexpr (MODEL != 1) oth[o] = stage_out<STAGE>(base.th, hi.w - lo.w, cdt);
        if constexpr (MODEL != 0) ore[o] = stage_out<STAGE>(base.re, hi.e - lo.e, cdt);
    };

    Q<MODEL> prev;            // closures of the last evaluated cell
    Base base_prev;
    Flux F_below;
    F_below.w = F_below.e = 0.0;
    base_prev.th = base_prev.re = 0.0;

#if LH_LDG_PIPE
    if (active) {
        // Register pipeline: the inputs of cell i + 1 are requested (plain LDG) before cell i is evaluated and
        // consumed one cell later; two cells per trip alternate between two fixed register sets.
        auto ldg_raw = [&](int i) {
            Raw r;
            r.th = 0.0; r.ti = 0.0; r.x = 288.0; r.u0th = 0.0; r.u0re = 0.0;
            if (i < b) {
                r.th = pth[o_ld];
                if (ICE) r.ti = pti[o_ld];
                if (MODEL != 0) r.x = pre[o_ld];
                else if (need_T) r.x = pT[o_ld];
                if constexpr (STAGE >= 2) {
                    if constexpr (MODEL != 1) r.u0th = p0th[o_ld];
                    if constexpr (MODEL != 0) r.u0re = p0re[o_ld];
                }
                o_ld += stride;
            }
            return r;
        };
        int i = a;
        Raw rA = ldg_raw(i);
        Raw rB = ldg_raw(i + 1);
        {   // first cell of the chunk: no face below it yet -> park what its update needs in shared memory
            const Cell<MODEL> c = eval(rA);
            rA = ldg_raw(i + 2);
            q_store<MODEL>(sm_bot, c.q);
            sm_pend[0] = c.base.th; sm_pend[32] = c.base.re;
            prev = c.q; base_prev = c.base;
            ++i;
        }
        if (i < b) {   // second cell: the face above the first cell
            const Cell<MODEL> c = eval(rB);
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            sm_pend[64] = F.w; sm_pend[96] = F.e;
            F_below = F; prev = c.q; base_prev = c.base;
            ++i;
        }
        for (; i + 1 < b; i += 2) {   // two cells per trip: no sliding-window register moves
            rB = ldg_raw(i + 1);
            const Cell<MODEL> c0 = eval(rA);
            const Flux F0 = face_flux<MODEL>(p, prev, c0.q);
            write_next(base_prev, F_below, F0);
            rA = ldg_raw(i + 2);
            const Cell<MODEL> c1 = eval(rB);
            const Flux F1 = face_flux<MODEL>(p, c0.q, c1.q);
            write_next(c0.base, F0, F1);
            F_below = F1; prev = c1.q; base_prev = c1.base;
        }
        if (i < b) {   // odd tail
            const Cell<MODEL> c = eval(rA);
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            write_next(base_prev, F_below, F);
            F_below = F; prev = c.q; base_prev = c.base;
        }
#else
    if (active) {
        constexpr uint32_t CB = RING_CELL_BYTES;
        for (int k = 0; k < RING_DEPTH - 1; ++k) issue(ring_s + k * CB, a + k);
        int i = a;
        {   // first cell of the chunk: no face below it yet -> park what its update needs in shared memory
            const Cell<MODEL> c = eval(load_raw(ring_s, ring_s + 3 * CB, i + 3));
            q_store<MODEL>(sm_bot, c.q);
            sm_pend[0] = c.base.th; sm_pend[32] = c.base.re;
            prev = c.q; base_prev = c.base;
            ++i;
        }
        if (i < b) {   // second cell: the face above the first cell
            const Cell<MODEL> c = eval(load_raw(ring_s + CB, ring_s, i + 3));
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            sm_pend[64] = F.w; sm_pend[96] = F.e;
            F_below = F; prev = c.q; base_prev = c.base;
            ++i;
        }
        uint32_t dA = ring_s + 2 * CB;                    // ring pair of cells (i, i + 1)
        const uint32_t dsum = 2 * ring_s + 2 * CB;        // pair (0,1) base + pair (2,3) base
        for (; i + 1 < b; i += 2) {   // two cells per trip: no sliding-window register moves
            const uint32_t dB = dsum - dA;                // the other pair: cells (i + 2, i + 3)
#if LH_ILP2
            // both cells' inputs first, so that the two (independent) closure chains can be interleaved
            const Raw r0 = load_raw(dA, dB + CB, i + 3);
            const Raw r1 = load_raw(dA + CB, dA, i + 4);
            const Cell<MODEL> c0 = eval(r0);
            const Cell<MODEL> c1 = eval(r1);
            const Flux F0 = face_flux<MODEL>(p, prev, c0.q);
            const Flux F1 = face_flux<MODEL>(p, c0.q, c1.q);
            write_next(base_prev, F_below, F0);
            write_next(c0.base, F0, F1);
#else
            const Cell<MODEL> c0 = eval(load_raw(dA, dB + CB, i + 3));
            const Flux F0 = face_flux<MODEL>(p, prev, c0.q);
            write_next(base_prev, F_below, F0);
            const Cell<MODEL> c1 = eval(load_raw(dA + CB, dA, i + 4));
            const Flux F1 = face_flux<MODEL>(p, c0.q, c1.q);
            write_next(c0.base, F0, F1);
#endif
            F_below = F1; prev = c1.q; base_prev = c1.base;
            dA = dB;
        }
        if (i < b) {   // odd tail
            const Cell<MODEL> c = eval(load_raw(dA, dA, b));       // nothing left to request: empty group
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            write_next(base_prev, F_below, F);
            F_below = F; prev = c.q; base_prev = c.base;
        }
#endif
        q_store<MODEL>(sm_top, prev);
    }
    __syncthreads();
    if (active) {
        const Q<MODEL> first = q_load<MODEL>(sm_bot);
        Flux F_lo, F_hi;
        if (a == 0) {
            // bottom boundary flux from the first cell (its raw values are still unwritten in global memory)
            LhCell c;
            c.K = 0.0; c.psi = 0.0; c.kappa = 0.0; c.T = 288.0; c.dT = 0.0;
            if constexpr (MODEL != 1) { c.K = first.K; c.psi = first.psi; }
            if constexpr (MODEL != 0) c.T = first.T;
            else if (need_T) c.T = pT[0];
            F_lo = boundary_flux<MODEL, FLAGS>(p, tab, A.bot_e_kind, A.bot_h_kind, A.bcv[LH_BCV_BOTTOM_ENERGY],
                                               A.bcv[LH_BCV_BOTTOM_HYDROLOGY], true, pth[0], ICE ? pti[0] : 0.0, c);
        } else {
            F_lo = face_flux<MODEL>(p, q_load<MODEL>(sm_top - (Slot<MODEL>::doubles + RING_DOUBLES)), first);   // top of chunk w-1
        }
        if (b == n) {
            const int64_t o = (int64_t)(n - 1) * stride;
            LhCell c;
            c.K = 0.0; c.psi = 0.0; c.kappa = 0.0; c.T = 288.0; c.dT = 0.0;
            if constexpr (MODEL != 1) { c.K = prev.K; c.psi = prev.psi; }
            if constexpr (MODEL != 0) c.T = prev.T;
            else if (need_T) c.T = pT[o];
            F_hi = boundary_flux<MODEL, FLAGS>(p, tab, A.top_e_kind, A.top_h_kind, A.bcv[LH_BCV_TOP_ENERGY],
                                               A.bcv[LH_BCV_TOP_HYDROLOGY], false, pth[o], ICE ? pti[o] : 0.0, c);
        } else {
            F_hi = face_flux<MODEL>(p, prev, q_load<MODEL>(sm_bot + (Slot<MODEL>::doubles + RING_DOUBLES)));    // bot of chunk w+1
        }
        Base base_first;
        base_first.th = sm_pend[0]; base_first.re = sm_pend[32];
        if (b - a == 1) {
            write_at(a, base_first, F_lo, F_hi);
        } else {
            Flux F_first_up;
            F_first_up.w = sm_pend[64]; F_first_up.e = sm_pend[96];
            write_at(a, base_first, F_lo, F_first_up);
            write_at(b - 1, base_prev, F_below, F_hi);
        }
    }
}

template <int MODEL, int FLAGS>
cudaError_t launch_variant(int stage, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    dim3 block(32, s.W, s.G);
    dim3 grid((unsigned)s.nblocks);
    if (s.smem_bytes > 48 * 1024) {
        cudaError_t e;
        const int bytes = (int)s.smem_bytes;
        if ((e = cudaFuncSetAttribute(lh_soil_stage_kernel<MODEL, 0, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(lh_soil_stage_kernel<MODEL, 1, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(lh_soil_stage_kernel<MODEL, 2, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(lh_soil_stage_kernel<MODEL, 3, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
    }
    switch (stage) {
    case 0: lh_soil_stage_kernel<MODEL, 0, FLAGS><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 1: lh_soil_stage_kernel<MODEL, 1, FLAGS><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 2: lh_soil_stage_kernel<MODEL, 2, FLAGS><<<grid, block, s.smem_bytes, stream>>>(args); break;
    case 3: lh_soil_stage_kernel<MODEL, 3, FLAGS><<<grid, block, s.smem_bytes, stream>>>(args); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_model(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    if (MODEL == 1) flags &= ~LH_FLAG_VG2;   // the heat-only model has no water closures
    switch (flags & 7) {
    case 0: return launch_variant<MODEL, 0>(stage, args, s, stream);
    case 1: return launch_variant<MODEL, 1>(stage, args, s, stream);
    case 2: return launch_variant<MODEL, 2>(stage, args, s, stream);
    case 3: return launch_variant<MODEL, 3>(stage, args, s, stream);
    case 4: return launch_variant<MODEL, (MODEL == 1 ? 0 : 4)>(stage, args, s, stream);
    case 5: return launch_variant<MODEL, (MODEL == 1 ? 1 : 5)>(stage, args, s, stream);
    case 6: return launch_variant<MODEL, (MODEL == 1 ? 2 : 6)>(stage, args, s, stream);
    default: return launch_variant<MODEL, (MODEL == 1 ? 3 : 7)>(stage, args, s, stream);
    }
}

}  // namespace

