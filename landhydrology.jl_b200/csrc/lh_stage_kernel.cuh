// lh_stage_kernel.cuh — the fused soil RHS (+ Runge-Kutta stage) kernel template, the persistent SSPRK33 kernel and
// their launchers.  Included by one translation unit per (model, stage group) (lh_kernels_m<M>_<a|b|c|p>.cu) so that the
// variants (3 models x 7 stage kinds x up to 14 flag combinations) compile in parallel.  See lh_kernels.cu for the design notes.
#pragma once

#include "lh_kernels.cuh"
#include "lh_ptx.cuh"

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
template <int MODEL, class P>
__device__ __forceinline__ Flux face_flux(const P& p, const Q<MODEL>& lo, const Q<MODEL>& hi)
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
template <int MODEL, int FLAGS, class P>
__device__ __forceinline__ Flux boundary_flux(const P& p, const double* __restrict__ tab, int e_kind,
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
            LhPowArg a_;
            lh_water_closures<ICE, GEN, (FLAGS & LH_FLAG_VG2) != 0, false>(p, tab, th_f, ti, T_f, K_f, psi_f, l_, a_);
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

// STAGE: 0 tendency; 1, 2, 3 the SSPRK33 stages; 4 a generic two-register Shu-Osher stage
// (a u^n + b u_{i-1} + g dt f); 5 a Williamson 2N stage (r = a r + dt f, u = u + b r), include/lh_soil.h;
// 6 an SSPRK33 stage chosen at RUN time, sg (sa u0 + sb v + cdt dF) with (sa, sb, sg) = (0, 1, 1), (3, 1, 1/4),
// (1, 2, 1/3): bit-identical to stages 1, 2, 3 (the products with 0, 1, 2 are exact) — one loop body for the
// persistent kernel instead of three.
template <int STAGE>
__device__ __forceinline__ double stage_base(const LhStageIO& io, double v, double u0)
{
    if constexpr (STAGE == 2) return fma(3.0, u0, v);       // 3 u0 + u1
    else if constexpr (STAGE == 3) return fma(2.0, v, u0);  // u0 + 2 u2
    else if constexpr (STAGE == 4) return fma(io.sa, u0, io.sb * v);
    else if constexpr (STAGE == 5) return io.first2n ? 0.0 : io.sa * u0;   // a r (r is not read as a number in the first stage)
    else if constexpr (STAGE == 6) return fma(io.sa, u0, io.sb * v);
    else return v;
}

// x / 3, correctly rounded (Markstein: one residual correction of x * fl(1/3)).  A bare multiplication by fl(1/3)
// is biased by -5.6e-17 relative, EVERY step, on the whole state: over the 138 240 steps of the reference's coupled
// equilibrium test (coupled.jl:36-39) the water budget drifts by 8e-12 — the reference divides by 3.
__device__ __forceinline__ double lh_third(double x)
{
    const double q = x * (1.0 / 3.0);
    return fma(fma(-3.0, q, x), 1.0 / 3.0, q);
}

// Stage combine with the flux-form divergence folded in:  k = -(F_hi - F_lo)/dz and
//   stage 0: k     1: base + dt k     2: (base + dt k)/4     3: (base + 2 dt k)/3     4: base + g dt k     5: base + dt k
// are evaluated as  s (base + cdt (F_hi - F_lo))  with cdt = -dt/dz (stage 3: -2 dt/dz; stage 0: -1/dz; stage 4: -g dt/dz).
template <int STAGE>
__device__ __forceinline__ double stage_out(double base, double dF, double cdt, double sg)
{
    if constexpr (STAGE == 6) {
        const double x = fma(cdt, dF, base);
        return sg == 1.0 / 3.0 ? lh_third(x) : sg * x;     // uniform select; 1 and 1/4 scale exactly
    }
    if constexpr (STAGE == 0) return cdt * dF;
    else if constexpr (STAGE == 1 || STAGE == 4 || STAGE == 5) return fma(cdt, dF, base);
    else if constexpr (STAGE == 2) return 0.25 * fma(cdt, dF, base);
    else return lh_third(fma(cdt, dF, base));
}

struct Base { double th, re, th2, re2; };   // th2/re2: the state itself, 2N stages only
struct Raw { double th, ti, x, u0th, u0re; };
template <int MODEL> struct Cell { Q<MODEL> q; Base base; };

// Shared-memory slot of one (column group, chunk): [bot: NQ][top: NQ][pending: 6], each x32 lanes.
// Input ring (cp.async): RING_DEPTH cells x RING_FIELDS fields x 32 lanes per warp.
constexpr int RING_DEPTH = 4, RING_FIELDS = 5, RING_DOUBLES = RING_DEPTH * RING_FIELDS * 32;
constexpr uint32_t RING_CELL_BYTES = RING_FIELDS * 256;
constexpr int LH_RED_DOUBLES = 2 * LH_WARPS_PER_SM;      // per-warp budget partial sums (after the tables)

template <int MODEL> struct Slot { static constexpr int NQv = NQ<MODEL>::value; static constexpr int doubles = (2 * NQv + 6) * 32; };

// One stage over this block's column groups.  `smem` = the block's dynamic shared memory with the exp2 / log2
// tables already staged at its start.  Contains ONE __syncthreads (the chunk-face exchange): every thread of
// the block must call it.
template <int MODEL, int STAGE, int FLAGS, class P>
__device__ __forceinline__ void lh_stage_body_impl(const LhKernelArgs& A, const LhStageIO& io, double* smem, P& p)
{
    constexpr int NQv = NQ<MODEL>::value;
    constexpr bool ICE = (FLAGS & LH_FLAG_ICE) != 0;
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
    double* warp_base = smem + LH_TAB_ALL + LH_RED_DOUBLES + (size_t)(g * W + w) * (Slot<MODEL>::doubles + RING_DOUBLES);
    double* slot = warp_base + lane;
    double* sm_bot = slot;                               // Q of the chunk's first cell
    double* sm_top = slot + NQv * 32;                    // Q of the chunk's last cell
    double* sm_pend = slot + 2 * NQv * 32;               // base.th, base.re, F_first_up.w, F_first_up.e, base.th2, base.re2

    const double* pth = io.in_th + col;
    const double* pti = io.in_ti + col;
    const double* pT = io.aux_T + col;
    double* oth = io.out_th + col;
    double* ore = io.out_re + col;

    // Input pipeline.  The raw values of cell i travel global -> shared with cp.async (LDGSTS): no
    // registers are held while the copy is in flight (a register software pipeline was tried: ptxas
    // sinks such loads down to the next possibly-aliasing store and spills them; prefetch.global.L1
    // was tried too and only reaches L2 on this part — 3 % L1 hit rate in the ncu capture of that build).
    // One field of one layer is a 256-byte row (32 columns); every lane copies 16 bytes, so ONE
    // instruction moves two rows (lanes 0-15 the first, lanes 16-31 the second) and the copy can take
    // the L1-bypassing .cg path: with 8-byte .ca copies the ~60 KB in flight per SM had to be resident
    // in L1, which shrinks to 22 KB once five blocks' shared memory is carved out.  A lane later reads
    // back values copied by OTHER lanes, hence the __syncwarp after cp.async.wait_group.
    // Ring of RING_DEPTH cells; one commit group per cell, always committed (empty past the end of the
    // chunk) so that wait_group's constant stays valid.  Cell a + k sits in ring cell k mod 4; the layer
    // loop handles two cells per trip, so its ring cells are a PAIR (0,1) or (2,3) and every ring address
    // is `pair base + constant`; the pair base toggles with one subtraction per trip.
    constexpr bool HAS_X = MODEL != 0 || (FLAGS & LH_FLAG_GEN) != 0;       // ρe_int, or the prescribed T
    constexpr int NROWS = 1 + (ICE ? 1 : 0) + (HAS_X ? 1 : 0) + ((STAGE >= 2 && MODEL != 1) ? 1 : 0) + ((STAGE >= 2 && MODEL != 0) ? 1 : 0);
    constexpr int NG = (NROWS + 1) / 2;
    const double* rowptr[5];
    int rowpos[5];                            // position of the row inside a ring cell: th 0, ti 1, x 2, u0th 3, u0re 4
    {
        int k = 0;
        rowptr[k] = io.in_th; rowpos[k++] = 0;
        if (ICE) { rowptr[k] = io.in_ti; rowpos[k++] = 1; }
        if (HAS_X) { rowptr[k] = MODEL != 0 ? io.in_re : io.aux_T; rowpos[k++] = 2; }
        if (STAGE >= 2 && MODEL != 1) { rowptr[k] = io.u0_th; rowpos[k++] = 3; }
        if (STAGE >= 2 && MODEL != 0) { rowptr[k] = io.u0_re; rowpos[k++] = 4; }
    }
    const bool upper = lane >= 16;
    const int sub = lane & 15;
    // Running pointers, advanced by one layer (stride_b bytes) per request: a 64-bit add each, instead of rebuilding
    // base + 8 * (column + layer * stride) for every copy (6 integer instructions per copy in the first build).
    const int64_t stride_b = stride * (int64_t)sizeof(double);
    const char* sp[NG];                       // this lane's source (column pair) of copy instruction g, next layer to request
    uint32_t doff[NG];                        // and its destination offset inside a ring cell
    bool pred[NG];
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) {
        const int lo = 2 * gi, hi = 2 * gi + 1 < NROWS ? 2 * gi + 1 : 2 * gi;
        sp[gi] = reinterpret_cast<const char*>((upper ? rowptr[hi] : rowptr[lo]) + col0 + 2 * sub + (int64_t)a * stride);
        doff[gi] = (uint32_t)((upper ? rowpos[hi] : rowpos[lo]) * 256 + sub * 16);
        pred[gi] = !upper || 2 * gi + 1 < NROWS;
    }
    const uint32_t ring_w = (uint32_t)__cvta_generic_to_shared(warp_base + Slot<MODEL>::doubles);   // warp's ring, lane 0
    auto issue = [&](uint32_t d, int i) {     // request cell i (the next layer of sp[]) into the ring cell at d (lane-0 address)
        if (i < b) {
#pragma unroll
            for (int gi = 0; gi < NG; ++gi) { lh_cp16(d + doff[gi], sp[gi], pred[gi]); sp[gi] += stride_b; }
        }
        lh_cp_commit();
    };
    // wait for the oldest outstanding cell, read it from the ring cell at d, then request cell i_next into d_next
    const uint32_t lane8 = (uint32_t)lane * 8;
    auto load_raw = [&](uint32_t d, uint32_t d_next, int i_next) {
        lh_cp_wait<RING_DEPTH - 2>();
        __syncwarp();
        const uint32_t dl = d + lane8;
        Raw r;
        r.th = lh_lds<0>(dl);
        r.ti = ICE ? lh_lds<256>(dl) : 0.0;
        r.x = HAS_X ? lh_lds<512>(dl) : 288.0;
        r.u0th = 0.0; r.u0re = 0.0;
        if constexpr (STAGE >= 2) {
            if constexpr (MODEL != 1) r.u0th = lh_lds<768>(dl);
            if constexpr (MODEL != 0) r.u0re = lh_lds<1024>(dl);
        }
        issue(d_next, i_next);
        return r;
    };
    // CELLP: the lane's hydraulic parameters change from cell to cell: re-read them right before the closures
    constexpr bool CELLP = (FLAGS & LH_FLAG_CELLP) != 0;
    const int64_t cell_fs = (int64_t)A.nlayer * stride;
    const double* cellp_next = CELLP ? A.cellp + col + (int64_t)a * stride : nullptr;      // cell the next eval() is for
    auto eval = [&](const Raw& r) {
        if constexpr (CELLP) { lh_load_cell_params(p, cellp_next, cell_fs); cellp_next += stride; }
        const LhCell c = lh_cell_closures<MODEL, FLAGS>(p, tab, r.th, r.ti, r.x);
        Cell<MODEL> o;
        if constexpr (MODEL == 0) { o.q.K = c.K; o.q.psi = c.psi; }
        else if constexpr (MODEL == 1) { o.q.kappa = c.kappa; o.q.T = c.T; }
        else {
            o.q.K = c.K; o.q.psi = c.psi; o.q.kappa = c.kappa; o.q.T = c.T;
            o.q.eK = (p.rhocp_l * c.dT) * c.K;                               // ρe_int_l * K (:306, :364)
        }
        o.base.th = (MODEL != 1) ? stage_base<STAGE>(io, r.th, r.u0th) : 0.0;
        o.base.re = (MODEL != 0) ? stage_base<STAGE>(io, r.x, r.u0re) : 0.0;
        o.base.th2 = (STAGE == 5 && MODEL != 1) ? r.th : 0.0;
        o.base.re2 = (STAGE == 5 && MODEL != 0) ? r.x : 0.0;
        return o;
    };
    const double cdt = STAGE == 6 ? io.dt /* the caller passes cdt itself */ : STAGE == 0 ? -p.inv_dz : (STAGE == 3 ? -2.0 * (io.dt * p.inv_dz) : STAGE == 4 ? -((io.sg * io.dt) * p.inv_dz) : -(io.dt * p.inv_dz));
    // The last SSPRK33 stage writes the new state: its values are summed on the way out (water and energy budgets of this
    // thread's cells), so that lh_soil_budgets after a step needs no pass over the state.
    constexpr bool BUDGET = STAGE == 3 || STAGE == 6;
    const bool budget = BUDGET && (STAGE == 3 || io.budget) && A.budget_partials != nullptr;     // block-uniform
    double bud_w = 0.0, bud_e = 0.0;
    // Output addresses are byte offsets from the column's layer-0 element; the layer loop advances ONE running offset
    // per trip (write_next) instead of rebuilding 8 * (column + layer * stride) per store.
    char* const oth_c = reinterpret_cast<char*>(oth);
    char* const ore_c = reinterpret_cast<char*>(ore);
    char* const o2th_c = reinterpret_cast<char*>(io.out2_th + col);     // 2N stages: the residual register r
    char* const o2re_c = reinterpret_cast<char*>(io.out2_re + col);
    auto store_cell = [&](int64_t ob, const Base& base, const Flux& lo, const Flux& hi) {
        if constexpr (MODEL != 1) {
            const double v = stage_out<STAGE>(base.th, hi.w - lo.w, cdt, io.sg);
            if constexpr (STAGE == 5) { *reinterpret_cast<double*>(o2th_c + ob) = v; *reinterpret_cast<double*>(oth_c + ob) = fma(io.sb, v, base.th2); }
            else *reinterpret_cast<double*>(oth_c + ob) = v;
            if constexpr (BUDGET) bud_w += v;
        }
        if constexpr (MODEL != 0) {
            const double v = stage_out<STAGE>(base.re, hi.e - lo.e, cdt, io.sg);
            if constexpr (STAGE == 5) { *reinterpret_cast<double*>(o2re_c + ob) = v; *reinterpret_cast<double*>(ore_c + ob) = fma(io.sb, v, base.re2); }
            else *reinterpret_cast<double*>(ore_c + ob) = v;
            if constexpr (BUDGET) bud_e += v;
        }
    };
    int64_t o_st = (int64_t)(a + 1) * stride_b;   // byte offset of the next cell to store (cell a is stored last)
    auto write_next = [&](const Base& base, const Flux& lo, const Flux& hi) {
        store_cell(o_st, base, lo, hi);
        o_st += stride_b;
    };
    auto write_at = [&](int i, const Base& base, const Flux& lo, const Flux& hi) {
        store_cell((int64_t)i * stride_b, base, lo, hi);
    };

    Q<MODEL> prev;            // closures of the last evaluated cell
    Base base_prev;
    Flux F_below;
    F_below.w = F_below.e = 0.0;
    base_prev.th = base_prev.re = base_prev.th2 = base_prev.re2 = 0.0;

    if (active) {
        constexpr uint32_t CB = RING_CELL_BYTES;
        for (int k = 0; k < RING_DEPTH - 1; ++k) issue(ring_w + k * CB, a + k);
        int i = a;
        {   // first cell of the chunk: no face below it yet -> park what its update needs in shared memory
            const Cell<MODEL> c = eval(load_raw(ring_w, ring_w + 3 * CB, i + 3));
            q_store<MODEL>(sm_bot, c.q);
            sm_pend[0] = c.base.th; sm_pend[32] = c.base.re;
            if constexpr (STAGE == 5) { sm_pend[128] = c.base.th2; sm_pend[160] = c.base.re2; }
            prev = c.q; base_prev = c.base;
            ++i;
        }
        if (i < b) {   // second cell: the face above the first cell
            const Cell<MODEL> c = eval(load_raw(ring_w + CB, ring_w, i + 3));
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            sm_pend[64] = F.w; sm_pend[96] = F.e;
            F_below = F; prev = c.q; base_prev = c.base;
            ++i;
        }
        uint32_t dA = ring_w + 2 * CB;                    // ring pair of cells (i, i + 1)
        const uint32_t dsum = 2 * ring_w + 2 * CB;        // pair (0,1) base + pair (2,3) base
        for (; i + 1 < b; i += 2) {   // two cells per trip: no sliding-window register moves
            const uint32_t dB = dsum - dA;                // the other pair: cells (i + 2, i + 3)
            const Cell<MODEL> c0 = eval(load_raw(dA, dB + CB, i + 3));
            const Flux F0 = face_flux<MODEL>(p, prev, c0.q);
            write_next(base_prev, F_below, F0);
            const Cell<MODEL> c1 = eval(load_raw(dA + CB, dA, i + 4));
            const Flux F1 = face_flux<MODEL>(p, c0.q, c1.q);
            write_next(c0.base, F0, F1);
            F_below = F1; prev = c1.q; base_prev = c1.base;
            dA = dB;
        }
        if (i < b) {   // odd tail
            const Cell<MODEL> c = eval(load_raw(dA, dA, b));       // nothing left to request: empty group
            const Flux F = face_flux<MODEL>(p, prev, c.q);
            write_next(base_prev, F_below, F);
            F_below = F; prev = c.q; base_prev = c.base;
        }
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
            else if (need_T) c.T = __ldcg(pT);
            if constexpr (CELLP) lh_load_cell_params(p, A.cellp + col, cell_fs);              // the bottom cell's parameters
            // per-column prescribed fluxes (lh_soil_set_column_fluxes) replace the scalar boundary value
            const double ve = io.flux_cols[LH_BCV_BOTTOM_ENERGY] ? __ldcg(io.flux_cols[LH_BCV_BOTTOM_ENERGY] + col) : io.bcv[LH_BCV_BOTTOM_ENERGY];
            const double vh = io.flux_cols[LH_BCV_BOTTOM_HYDROLOGY] ? __ldcg(io.flux_cols[LH_BCV_BOTTOM_HYDROLOGY] + col) : io.bcv[LH_BCV_BOTTOM_HYDROLOGY];
            F_lo = boundary_flux<MODEL, FLAGS>(p, tab, A.bot_e_kind, A.bot_h_kind, ve, vh, true, __ldcg(pth), ICE ? __ldcg(pti) : 0.0, c);
        } else {
            F_lo = face_flux<MODEL>(p, q_load<MODEL>(sm_top - (Slot<MODEL>::doubles + RING_DOUBLES)), first);   // top of chunk w-1
        }
        if (b == n) {
            const int64_t o = (int64_t)(n - 1) * stride;
            LhCell c;
            c.K = 0.0; c.psi = 0.0; c.kappa = 0.0; c.T = 288.0; c.dT = 0.0;
            if constexpr (MODEL != 1) { c.K = prev.K; c.psi = prev.psi; }
            if constexpr (MODEL != 0) c.T = prev.T;
            else if (need_T) c.T = __ldcg(pT + o);
            if constexpr (CELLP) lh_load_cell_params(p, A.cellp + col + o, cell_fs);          // the top cell's parameters
            // per-column fluxes: prescribed fields, or the atmospheric fluxes lh_atmos_flux_kernel left for this stage
            const double ve = io.flux_cols[LH_BCV_TOP_ENERGY] ? __ldcg(io.flux_cols[LH_BCV_TOP_ENERGY] + col) : io.bcv[LH_BCV_TOP_ENERGY];
            const double vh = io.flux_cols[LH_BCV_TOP_HYDROLOGY] ? __ldcg(io.flux_cols[LH_BCV_TOP_HYDROLOGY] + col) : io.bcv[LH_BCV_TOP_HYDROLOGY];
            F_hi = boundary_flux<MODEL, FLAGS>(p, tab, A.top_e_kind, A.top_h_kind, ve, vh, false, __ldcg(pth + o), ICE ? __ldcg(pti + o) : 0.0, c);
        } else {
            F_hi = face_flux<MODEL>(p, prev, q_load<MODEL>(sm_bot + (Slot<MODEL>::doubles + RING_DOUBLES)));    // bot of chunk w+1
        }
        Base base_first;
        base_first.th = sm_pend[0]; base_first.re = sm_pend[32];
        base_first.th2 = base_first.re2 = 0.0;
        if constexpr (STAGE == 5) { base_first.th2 = sm_pend[128]; base_first.re2 = sm_pend[160]; }
        if (b - a == 1) {
            write_at(a, base_first, F_lo, F_hi);
        } else {
            Flux F_first_up;
            F_first_up.w = sm_pend[64]; F_first_up.e = sm_pend[96];
            write_at(a, base_first, F_lo, F_first_up);
            write_at(b - 1, base_prev, F_below, F_hi);
        }
    }
    if constexpr (BUDGET) {
        if (budget) {
            // fixed-shape tree: lanes (padding columns masked), then the block's warps in order -> one pair per block
            if (!active || col >= A.ncol) { bud_w = 0.0; bud_e = 0.0; }
            for (int o = 16; o > 0; o >>= 1) {
                bud_w += __shfl_down_sync(0xffffffffu, bud_w, o);
                bud_e += __shfl_down_sync(0xffffffffu, bud_e, o);
            }
            double* red = smem + LH_TAB_ALL;
            const int wib = g * W + w;
            if (lane == 0) { red[2 * wib] = bud_w; red[2 * wib + 1] = bud_e; }
            __syncthreads();
            if (wib == 0 && lane == 0) {
                double sw = 0.0, se = 0.0;
                const int nw = (int)(blockDim.y * blockDim.z);
                for (int k = 0; k < nw; ++k) { sw += red[2 * k]; se += red[2 * k + 1]; }
                A.budget_partials[2 * (int64_t)blockIdx.x] = sw;
                A.budget_partials[2 * (int64_t)blockIdx.x + 1] = se;
            }
        }
    }
}

// Homogeneous soils: the parameters are the launch's uniform block.  HET (lh_soil_set_column_params): every lane
// overrides the column-dependent members with its own column's values, loaded once per stage.
template <int MODEL, int STAGE, int FLAGS>
__device__ __forceinline__ void lh_stage_body(const LhKernelArgs& A, const LhStageIO& io, double* smem)
{
    if constexpr ((FLAGS & LH_FLAG_HET) == 0) {
        lh_stage_body_impl<MODEL, STAGE, FLAGS, const LhDevParams>(A, io, smem, A.p);
    } else {
        const int g = __shfl_sync(0xffffffffu, (int)threadIdx.z, 0);
        int64_t col = ((int64_t)blockIdx.x * blockDim.z + g) * 32 + threadIdx.x;
        if (col >= A.ncol_pad) col = A.ncol_pad - 1;           // an idle column group of the last block
        const double* cp = A.colp + col;
        const int64_t st = A.ncol_pad;
        LhLaneParams pl;
        static_cast<LhPhys&>(pl) = static_cast<const LhPhys&>(A.p);
        pl.mc = A.p.mc;
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
        if constexpr ((FLAGS & LH_FLAG_HETH) != 0) {
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
        lh_stage_body_impl<MODEL, STAGE, FLAGS, LhLaneParams>(A, io, smem, pl);
    }
}

template <int MODEL, int STAGE, int FLAGS>
__global__ void __launch_bounds__(LhBounds<FLAGS>::max_threads, 1)
lh_soil_stage_kernel(const __grid_constant__ LhKernelArgs A)
{
    LH_DYN_SMEM(double, smem);
#if LH_PDL
    // Programmatic dependent launch: the NEXT stage's blocks may be scheduled as soon as this grid's blocks have
    // all started, into the SM slots its last wave frees; they stage their tables (parameter block only) and then
    // wait here until the previous stage has completed and flushed.  Hides the launch gap and the block prologue
    // behind the previous stage's tail — what matters when a launch is only ~80 us (column shards at 8 GPUs).
    lh_pdl_launch_dependents();
#endif
    const int tid = (threadIdx.z * blockDim.y + threadIdx.y) * 32 + threadIdx.x;
    lh_stage_tables(A.p, A.pow_tab, smem, tid, blockDim.x * blockDim.y * blockDim.z);
    if (A.chain_flags != nullptr && A.chain_wait != 0) {
        // this block's predecessor: block blockIdx.x of the previous stage launch (same columns); see LhKernelArgs
        if (tid == 0) {
            const int32_t* f = A.chain_flags + blockIdx.x;
            for (;;) {
                if (lh_ld_acquire(f) == A.chain_wait) break;
                __nanosleep(64);
            }
        }
    } else {
#if LH_PDL
        lh_pdl_wait();
#endif
    }
    __syncthreads();
    lh_stage_body<MODEL, STAGE, FLAGS>(A, A.io, smem);
    if (A.chain_flags != nullptr) {
        __threadfence();                  // every thread's stores are visible device-wide before the flag is
        __syncthreads();
        if ((threadIdx.x | threadIdx.y | threadIdx.z) == 0)      // (not `tid`: keeping it live across the layer loop costs 4 moves per cell)
            lh_st_release(A.chain_flags + blockIdx.x, A.chain_set);
    }
}

// A.nsteps whole SSPRK33 steps in one launch.  A block keeps its column groups for all 3 nsteps stages: columns
// are laterally independent, so there is no grid-wide dependency between stages, only the block's own — the
// stage register V and the state U of the resident blocks (~60 MB for 740 blocks of 64 layers) stay in the
// 126 MB L2 between stages, there are no launch gaps and only one tail per call instead of one per stage.
// Used for grids of few waves (column shards of a multi-GPU run, small domains), where launch gaps and wave
// quantisation cost more than 5 %; results are bit-identical to the per-stage launches.
template <int MODEL, int FLAGS>
__global__ void __launch_bounds__(LhBounds<FLAGS>::max_threads, 1)
lh_soil_ssprk33_persistent_kernel(const __grid_constant__ LhKernelArgs A)
{
    LH_DYN_SMEM(double, smem);
    lh_stage_tables(A.p, A.pow_tab, smem, (threadIdx.z * blockDim.y + threadIdx.y) * 32 + threadIdx.x, blockDim.x * blockDim.y * blockDim.z);
    __syncthreads();
    // stage 1 as passed: in = U, out = V
    const double* Uth = A.io.in_th;
    const double* Ure = A.io.in_re;
    double* Vth = A.io.out_th;
    double* Vre = A.io.out_re;
    const double cdt1 = -(A.io.dt * A.p.inv_dz), cdt3 = -2.0 * (A.io.dt * A.p.inv_dz);    // as stages 1/2 and 3 form them
    int st = 0;                                   // 0, 1, 2: which SSPRK33 stage
    for (int64_t k = 0; k < 3 * A.nsteps; ++k) {
        LhStageIO io = A.io;
        if (MODEL != 1) io.in_th = st == 0 ? Uth : Vth;     // the heat-only model reads the prescribed ϑ_l from U in every stage
        if (MODEL != 0) io.in_re = st == 0 ? Ure : Vre;
        io.u0_th = Uth; io.u0_re = Ure;
        io.out_th = st == 2 ? const_cast<double*>(Uth) : Vth;
        io.out_re = st == 2 ? const_cast<double*>(Ure) : Vre;
        io.sa = st == 0 ? 0.0 : st == 1 ? 3.0 : 1.0;
        io.sb = st == 2 ? 2.0 : 1.0;
        io.sg = st == 0 ? 1.0 : st == 1 ? 0.25 : 1.0 / 3.0;
        io.dt = st == 2 ? cdt3 : cdt1;
        io.budget = st == 2;
        if (A.bc_dev) {
            const double* b = A.bc_dev + k * 4;
            io.bcv[0] = b[0]; io.bcv[1] = b[1]; io.bcv[2] = b[2]; io.bcv[3] = b[3];
        }
        lh_stage_body<MODEL, 6, FLAGS>(A, io, smem);
        __syncthreads();                          // the exchange slots and rings are reused by the next stage
        st = st == 2 ? 0 : st + 1;
    }
}

// The stage kinds are compiled in GROUPS of two (0: tendency + SSPRK33 stage 1, 1: SSPRK33 stages 2 and 3, 2: the generic
// Shu-Osher and 2N stages), one translation unit per (model, group), so that the ~200 kernel variants build in parallel.
template <int MODEL, int FLAGS, int GROUP>
cudaError_t launch_variant(int stage, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    dim3 block(32, s.W, s.G);
    dim3 grid((unsigned)s.nblocks);
    // Once per variant and device: allow > 48 KB of dynamic shared memory, and ask for the largest shared
    // memory carve-out.  Five resident blocks need 5 x (38.5 + 1) KB = 197.5 KB, just above the 196 KB
    // configuration; left to its default the driver picks a smaller carve-out and only four blocks fit.
    static int configured_smem[64];                      // per device ordinal: largest size configured so far
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && configured_smem[dev] < (int)s.smem_bytes + 1) {
        cudaError_t e;
        const int bytes = (int)s.smem_bytes > 48 * 1024 ? (int)s.smem_bytes : 48 * 1024;
        auto configure = [&](auto kernel) -> cudaError_t {
            if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
            return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        };
        if ((e = configure(lh_soil_stage_kernel<MODEL, 2 * GROUP, FLAGS>))) return e;
        if ((e = configure(lh_soil_stage_kernel<MODEL, 2 * GROUP + 1, FLAGS>))) return e;
        configured_smem[dev] = (int)s.smem_bytes + 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = s.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = LH_PDL;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (stage == 2 * GROUP) return cudaLaunchKernelEx(&cfg, lh_soil_stage_kernel<MODEL, 2 * GROUP, FLAGS>, args);
    if (stage == 2 * GROUP + 1) return cudaLaunchKernelEx(&cfg, lh_soil_stage_kernel<MODEL, 2 * GROUP + 1, FLAGS>, args);
    return cudaErrorInvalidValue;
}

template <int MODEL, int GROUP>
cudaError_t launch_model(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    if (MODEL == 1) flags &= ~LH_FLAG_VG2;   // the heat-only model has no water closures
    if (flags & LH_FLAG_HET) {               // per-column parameters: always the general closures
        if (flags & LH_FLAG_CELLP) {         // per-cell hydraulic parameters (with the per-column heat parameters where there is heat)
            constexpr int F = LH_FLAG_CELLP | LH_FLAG_HET | LH_FLAG_GEN | (MODEL == 0 ? 0 : LH_FLAG_HETH);
            return (flags & LH_FLAG_ICE) ? launch_variant<MODEL, F | LH_FLAG_ICE, GROUP>(stage, args, s, stream)
                                         : launch_variant<MODEL, F, GROUP>(stage, args, s, stream);
        }
        if (MODEL != 0 && (flags & LH_FLAG_HETH)) {      // per-column heat parameters (the Richards model has no heat closures)
            switch (flags & LH_FLAG_ICE) {
            case 0: return launch_variant<MODEL, (MODEL == 0 ? 0 : LH_FLAG_HETH) | LH_FLAG_HET | LH_FLAG_GEN, GROUP>(stage, args, s, stream);
            default: return launch_variant<MODEL, (MODEL == 0 ? 0 : LH_FLAG_HETH) | LH_FLAG_HET | LH_FLAG_GEN | LH_FLAG_ICE, GROUP>(stage, args, s, stream);
            }
        }
        switch (flags & LH_FLAG_ICE) {
        case 0: return launch_variant<MODEL, LH_FLAG_HET | LH_FLAG_GEN, GROUP>(stage, args, s, stream);
        default: return launch_variant<MODEL, LH_FLAG_HET | LH_FLAG_GEN | LH_FLAG_ICE, GROUP>(stage, args, s, stream);
        }
    }
    switch (flags & 7) {
    case 0: return launch_variant<MODEL, 0, GROUP>(stage, args, s, stream);
    case 1: return launch_variant<MODEL, 1, GROUP>(stage, args, s, stream);
    case 2: return launch_variant<MODEL, 2, GROUP>(stage, args, s, stream);
    case 3: return launch_variant<MODEL, 3, GROUP>(stage, args, s, stream);
    case 4: return launch_variant<MODEL, (MODEL == 1 ? 0 : 4), GROUP>(stage, args, s, stream);
    case 5: return launch_variant<MODEL, (MODEL == 1 ? 1 : 5), GROUP>(stage, args, s, stream);
    case 6: return launch_variant<MODEL, (MODEL == 1 ? 2 : 6), GROUP>(stage, args, s, stream);
    default: return launch_variant<MODEL, (MODEL == 1 ? 3 : 7), GROUP>(stage, args, s, stream);
    }
}

template <int MODEL, int FLAGS>
cudaError_t launch_persistent_variant(const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    dim3 block(32, s.W, s.G);
    dim3 grid((unsigned)s.nblocks);
    static int configured_smem[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && configured_smem[dev] < (int)s.smem_bytes + 1) {
        cudaError_t e;
        const int bytes = (int)s.smem_bytes > 48 * 1024 ? (int)s.smem_bytes : 48 * 1024;
        if ((e = cudaFuncSetAttribute(lh_soil_ssprk33_persistent_kernel<MODEL, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))) return e;
        if ((e = cudaFuncSetAttribute(lh_soil_ssprk33_persistent_kernel<MODEL, FLAGS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared))) return e;
        configured_smem[dev] = (int)s.smem_bytes + 1;
    }
    LH_LAUNCH((lh_soil_ssprk33_persistent_kernel<MODEL, FLAGS>), grid, block, s.smem_bytes, stream, args);
    return cudaGetLastError();
}

template <int MODEL>
cudaError_t launch_persistent_model(int flags, const LhKernelArgs& args, const LhLaunchShape& s, cudaStream_t stream)
{
    if (MODEL == 1) flags &= ~LH_FLAG_VG2;
    if (flags & LH_FLAG_HET) {
        if (flags & LH_FLAG_CELLP) {
            constexpr int F = LH_FLAG_CELLP | LH_FLAG_HET | LH_FLAG_GEN | (MODEL == 0 ? 0 : LH_FLAG_HETH);
            return (flags & LH_FLAG_ICE) ? launch_persistent_variant<MODEL, F | LH_FLAG_ICE>(args, s, stream)
                                         : launch_persistent_variant<MODEL, F>(args, s, stream);
        }
        if (MODEL != 0 && (flags & LH_FLAG_HETH)) {
            switch (flags & LH_FLAG_ICE) {
            case 0: return launch_persistent_variant<MODEL, (MODEL == 0 ? 0 : LH_FLAG_HETH) | LH_FLAG_HET | LH_FLAG_GEN>(args, s, stream);
            default: return launch_persistent_variant<MODEL, (MODEL == 0 ? 0 : LH_FLAG_HETH) | LH_FLAG_HET | LH_FLAG_GEN | LH_FLAG_ICE>(args, s, stream);
            }
        }
        switch (flags & LH_FLAG_ICE) {
        case 0: return launch_persistent_variant<MODEL, LH_FLAG_HET | LH_FLAG_GEN>(args, s, stream);
        default: return launch_persistent_variant<MODEL, LH_FLAG_HET | LH_FLAG_GEN | LH_FLAG_ICE>(args, s, stream);
        }
    }
    switch (flags & 7) {
    case 0: return launch_persistent_variant<MODEL, 0>(args, s, stream);
    case 1: return launch_persistent_variant<MODEL, 1>(args, s, stream);
    case 2: return launch_persistent_variant<MODEL, 2>(args, s, stream);
    case 3: return launch_persistent_variant<MODEL, 3>(args, s, stream);
    case 4: return launch_persistent_variant<MODEL, (MODEL == 1 ? 0 : 4)>(args, s, stream);
    case 5: return launch_persistent_variant<MODEL, (MODEL == 1 ? 1 : 5)>(args, s, stream);
    case 6: return launch_persistent_variant<MODEL, (MODEL == 1 ? 2 : 6)>(args, s, stream);
    default: return launch_persistent_variant<MODEL, (MODEL == 1 ? 3 : 7)>(args, s, stream);
    }
}

}  // namespace
