// lh_math.cuh — branch-free fp64 elementary functions for the soil closures (sm_100a).
//
// Why not libdevice's log/exp/expm1/pow: the ncu profile of the first kernel (profiles/r01_a_*)
// showed only 32 % of issued instructions on the fp64 pipe — 157 UMOV + 117 IMAD per cell were
// nvcc materialising 64-bit polynomial constants as pairs of 32-bit immediates, and ~77 were the
// branches of libdevice's special-case slow paths.  Here
//   * every coefficient is read from a table the caller passes in; the kernels keep that table in
//     their __grid_constant__ parameter block (constant bank 0), which sm_100a loads with LDCU into
//     UNIFORM registers and hoists out of the layer loop (a __constant__ array in bank 3 was tried:
//     ptxas emits one LDC.64 into a vector register per use, which costs an issue slot and spills);
//   * special cases are handled by selects on the result, never by branches;
//   * reciprocal / rsqrt seeds come from the MUFU pipe (rcp.approx.ftz.f64 / rsqrt.approx.ftz.f64)
//     and are refined by Newton steps on the fp64 pipe.
// Accuracy (tests/test_device_math.py, host emulation of the same code against mpmath): <= 2 ulp on
// the closures' ranges.  Coefficients: tools/gen_math_coeffs.py.
//
// The same source compiles on the host (LH_MATH_HOST) so the algorithms can be verified here
// without a GPU; that build is test infrastructure only.
#pragma once

#include <stdint.h>
#include <string.h>

#include "lh_math_coeffs.inc"

#ifdef LH_MATH_HOST
#include <math.h>
#define LH_DEV static inline
static const double lh_c_host[LHC_COUNT] = {LH_MATH_COEFFS};
LH_DEV double lh_fma(double a, double b, double c) { return fma(a, b, c); }
LH_DEV int32_t lh_hi(double x) { int64_t b; memcpy(&b, &x, 8); return (int32_t)(b >> 32); }
LH_DEV int32_t lh_lo(double x) { int64_t b; memcpy(&b, &x, 8); return (int32_t)(b & 0xffffffff); }
LH_DEV double lh_mk(int32_t hi, int32_t lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x; }
// MUFU.RCP64H / RSQ64H emulation: the high word only (20 mantissa bits, truncated), low word zero —
// the measured accuracy of the hardware seeds (2^-19.95 / 2^-20.06, tools/seed_accuracy.py).
LH_DEV double lh_rcp_seed(double x) { double r = 1.0 / x; return lh_mk(lh_hi(r), 0); }
LH_DEV double lh_rsqrt_seed(double x) { double r = 1.0 / sqrt(x); return lh_mk(lh_hi(r), 0); }
#define LH_INF (INFINITY)
#define LH_NAN (NAN)
#else
#include <cuda_runtime.h>
#define LH_DEV __device__ __forceinline__
LH_DEV double lh_fma(double a, double b, double c) { return fma(a, b, c); }
LH_DEV int32_t lh_hi(double x) { return __double2hiint(x); }
LH_DEV int32_t lh_lo(double x) { return __double2loint(x); }
LH_DEV double lh_mk(int32_t hi, int32_t lo) { return __hiloint2double(hi, lo); }
LH_DEV double lh_rcp_seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
LH_DEV double lh_rsqrt_seed(double x) { double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
#define LH_INF (__longlong_as_double(0x7ff0000000000000LL))
#define LH_NAN (__longlong_as_double(0x7ff8000000000000LL))
#endif

// The MUFU seeds are accurate to 2^-20 (measured on B200: tools/seed_accuracy.py gives 2^-19.95 for
// rcp.approx.ftz.f64 and 2^-20.06 for rsqrt.approx.ftz.f64), so ONE third-order step reaches fp64:
// the truncation error is e^3 ~ 2^-60, well below 2^-53.  That is 3 (rcp) / 5 (rsqrt) dependent fp64
// operations instead of the 4 / 6+ of two Newton steps.

// 1/x for normal, finite, non-zero x:  r0 (1 + e + e^2),  e = 1 - x r0.   ~1 ulp.
LH_DEV double lh_rcp(double x)
{
    const double r0 = lh_rcp_seed(x);
    const double e = lh_fma(-x, r0, 1.0);
    return lh_fma(r0, lh_fma(e, e, e), r0);
}

// a/b for normal finite b != 0, |a/b| in range: reciprocal + one residual correction (<= 1 ulp).
LH_DEV double lh_div(double a, double b)
{
    const double r = lh_rcp(b);
    const double q = a * r;
    return lh_fma(lh_fma(-q, b, a), r, q);
}

// a / b without the residual correction (~2 ulp): for quotients that are added to something larger.
LH_DEV double lh_div_fast(double a, double b) { return a * lh_rcp(b); }

// 1/sqrt(x) from a seed r0:  r0 (1 + e/2 + 3 e^2/8),  e = 1 - x r0^2.
LH_DEV double lh_rsqrt_refine(double x, double r0)
{
    const double e = lh_fma(-(x * r0), r0, 1.0);
    const double p = lh_fma(0.375, e, 0.5) * e;
    return lh_fma(r0, p, r0);
}

// 1/sqrt(x) for normal x > 0 (~1.5 ulp): sqrt(x) = x r and 1/x = r r then cost one multiply each.
LH_DEV double lh_rsqrt(double x) { return lh_rsqrt_refine(x, lh_rsqrt_seed(x)); }

// sqrt(x), x >= 0 finite.  The seed's high word is clamped below 0x7fe00000 with an INTEGER min (an fp64
// compare would occupy the fp64 pipe): rsqrt(0) = +inf would turn x r into NaN; clamped to ~9e307
// every step below stays finite and gives exactly 0 for x == 0.  Negative x yields NaN.
LH_DEV double lh_sqrt(double x)
{
    const double s0 = lh_rsqrt_seed(x);
    const int32_t hi = lh_hi(s0);
    const double r0 = lh_mk(((uint32_t)hi > 0x7fe00000u && hi > 0) ? 0x7fe00000 : hi, lh_lo(s0));
    const double r = lh_rsqrt_refine(x, r0);
    const double g = x * r;
    return lh_fma(lh_fma(-g, g, x), 0.5 * r, g);       // one correction of g = x r: <= 1 ulp
}

// sqrt(x) = x rsqrt(x) without the final correction (~1.5 ulp, 3 fp64 operations fewer): for the square roots whose
// result is multiplied into a product that is itself a few ulp from the literal form.  Same zero / negative handling.
LH_DEV double lh_sqrt_fast(double x)
{
    const double s0 = lh_rsqrt_seed(x);
    const int32_t hi = lh_hi(s0);
    const double r0 = lh_mk(((uint32_t)hi > 0x7fe00000u && hi > 0) ? 0x7fe00000 : hi, lh_lo(s0));
    return x * lh_rsqrt_refine(x, r0);
}

// ---------------------------------------------------------------------------------------------------
// Every transcendental on the soil path is a power x^c, so base 2 serves everywhere and saves the ln 2
// scalings of a natural log/exp pair.  Both functions are table driven; `tab` points at the LH_TAB_DOUBLES
// constants (lh_math_coeffs.inc from LHC_TAB0 on) that every kernel stages in SHARED memory:
//   tab[0 .. 63]            2^(j/64)
//   tab[64 + 2j], [65 + 2j] r_j, -log2 r_j for the 128 mantissa intervals of lh_log2 (one 16-byte load)
// The stage kernel is issue bound and an fp64 instruction holds the issue port for two cycles, so a table
// load (one LDS, bank conflicts cost LSU cycles, not issue slots) is worth more than one polynomial term.
//
// exp2 core: x = k/64 + r, |r| <= 1/128 (the reduction r = x - k/64 is EXACT);
//   2^x = 2^(k>>6) * T[k&63] * (1 + p),  p = r g(r) (degree-4 g).
// ---------------------------------------------------------------------------------------------------
#define LH_TAB_EXP 0
#define LH_TAB_LOG LH_EXP_TAB

struct LhExpParts { double s, p; };

LH_DEV LhExpParts lh_exp2_parts(const double* __restrict__ lh_c, const double* __restrict__ tab, double x)
{
    // Lower clamp in the INTEGER domain (an fp64 compare+select costs a DSETP on the fp64 pipe plus
    // two FSELs): for negative x the unsigned high word grows with |x|, so hi > 0xC08FF000 <=> x < -1022
    // (or x is a negative-signed NaN, which then reads as -1022).  Positive NaN propagates through the
    // arithmetic; x >= 1024 is outside the contract (the closures never produce it from a finite state).
    int32_t xhi = lh_hi(x);
    xhi = ((uint32_t)xhi > 0xC08FF000u) ? (int32_t)0xC08FF000 : xhi;
    const double xc = lh_mk(xhi, lh_lo(x));
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: low word of t is k = rint(64 x)
    const double t = lh_fma(xc, (double)LH_EXP_TAB, MAGIC);
    const int32_t k = lh_lo(t);
    const double r = lh_fma(t - MAGIC, -1.0 / LH_EXP_TAB, xc);   // exact
    const double T = tab[LH_TAB_EXP + (k & (LH_EXP_TAB - 1))];
    double g = lh_c[LHC_EXP2_G4];
    g = lh_fma(g, r, lh_c[LHC_EXP2_G3]);
    g = lh_fma(g, r, lh_c[LHC_EXP2_G2]);
    g = lh_fma(g, r, lh_c[LHC_EXP2_G1]);
    g = lh_fma(g, r, lh_c[LHC_EXP2_G0]);
    LhExpParts o;
    o.p = r * g;
    o.s = lh_mk(lh_hi(T) + ((k >> 6) << 20), lh_lo(T));     // T * 2^(k>>6): exponent-field add
    return o;
}

// 2^x for x < 1024.  x < -1022 (incl. -inf) returns 2^-1022 (not exactly 0: nothing downstream
// distinguishes them), NaN -> NaN.
LH_DEV double lh_exp2(const double* __restrict__ lh_c, const double* __restrict__ tab, double x)
{
    const LhExpParts e = lh_exp2_parts(lh_c, tab, x);
    return lh_fma(e.s, e.p, e.s);
}

// 2^x - 1.  |x| <= 1/128: k == 0, s == 1 and the result is p itself (a few ulp where 1 - 2^x cancels; the
// literal fp64 form 1 - x^c the reference uses is off by 1.1e-16 / |1 - 2^x| there).
// Otherwise s - 1 carries the rounding of T[j]: relative error <= 2^-53 / |2^x - 1| < 2e-14.
LH_DEV double lh_exp2m1(const double* __restrict__ lh_c, const double* __restrict__ tab, double x)
{
    const LhExpParts e = lh_exp2_parts(lh_c, tab, x);
    return lh_fma(e.s, e.p, e.s - 1.0);
}

// 1 - 2^x, same accuracy.  Its zero is +0 (fma(-1, 0, +0)), never -0, which lh_log2 would flag as negative.
LH_DEV double lh_one_minus_exp2(const double* __restrict__ lh_c, const double* __restrict__ tab, double x)
{
    const LhExpParts e = lh_exp2_parts(lh_c, tab, x);
    return lh_fma(-e.s, e.p, 1.0 - e.s);
}

// ---------------------------------------------------------------------------------------------------
// log2(x), table driven (Tang): x = 2^e m, m in [0.709, 1.418); j = top 7 bits of m's position in
// that range; t = m r_j - 1 by ONE FMA (|t| <= 2^-8, relative error 2^-53 whatever r_j's rounding was);
//   log2 x = (e + L_j) + t Q(t),  L_j = -log2 r_j,  Q of degree 5.
// The interval around m = 1 has r = 1, L = 0 exactly, so the RELATIVE accuracy is kept as x -> 1 (the
// closures need log2 S for S -> 1-).  8 fp64 operations, no division.
// CHECK: x < 0, -0.0, NaN (and +inf) -> NaN.  x == +0 and subnormals read as 2^-1023 m: log2(0) ~ -1023 instead of
// -inf, which is what the closures need (exp2 of it underflows; nothing tests for -inf).
// ---------------------------------------------------------------------------------------------------
template <bool CHECK = true>
LH_DEV double lh_log2(const double* __restrict__ lh_c, const double* __restrict__ tab, double x)
{
    // exponent such that the mantissa lands in [0.709, 1.418), and the interval index of that mantissa.
    // No sign masking: for a negative x everything below stays finite garbage and the flag overrides it.
    const int32_t off = lh_hi(x) - LH_LOG_HI0;
    const int32_t e = off >> 20;
    const int32_t j = (off >> LH_LOG_SHIFT) & (LH_LOG_TAB - 1);
    const double m = lh_mk(LH_LOG_HI0 + (off & 0xfffff), lh_lo(x));
#ifdef LH_MATH_HOST
    const double r = tab[LH_TAB_LOG + 2 * j], L = tab[LH_TAB_LOG + 2 * j + 1];
#else
    const double2 rl = *reinterpret_cast<const double2*>(tab + LH_TAB_LOG + 2 * j);
    const double r = rl.x, L = rl.y;
#endif
    const double t = lh_fma(m, r, -1.0);
    double Q = lh_c[LHC_LOG_Q5];
    Q = lh_fma(Q, t, lh_c[LHC_LOG_Q4]);
    Q = lh_fma(Q, t, lh_c[LHC_LOG_Q3]);
    Q = lh_fma(Q, t, lh_c[LHC_LOG_Q2]);
    Q = lh_fma(Q, t, lh_c[LHC_LOG_Q1]);
    Q = lh_fma(Q, t, lh_c[LHC_LOG_Q0]);
    const double y = lh_fma(t, Q, (double)e + L);
    if (!CHECK) return y;
    // sign bit set (x < 0, and -0.0: callers form 1 - 2^u with lh_one_minus_exp2, whose zero is +0), NaN,
    // +inf: force NaN with ONE unsigned compare on the high word (an fp64 compare occupies the fp64 pipe).
    const bool bad = (uint32_t)lh_hi(x) > 0x7fefffffu;
    return lh_mk(bad ? 0x7ff80000 : lh_hi(y), lh_lo(y));
}

// ---------------------------------------------------------------------------------------------------
// x^c for an exponent c that is FIXED per parameter set (the van Genuchten 1/m and m, the Kersten exponent):
// the log2 / exp2 pair collapses into one table-driven evaluation.  With the decomposition of lh_log2,
//   x = 2^e m,  m = (1 + t) / r_j,  t = m r_j - 1  (|t| <= 2^-8, ONE FMA)
//   x^c = 2^(e c) * r_j^(-c) * (1 + t)^c = B_e A_j (1 + t g(t)),   g of degree LH_POW_DEG
// A_j = r_j^-c (128 entries), B_e = 2^(e c) (e = -64 .. 1) and the coefficients of g are built on the host for
// each exponent when the context is created (lh_pow_build, long double) and staged in shared memory next to
// the exp2 / log2 tables.  8 fp64 operations (7 for a second power of the same x) instead of the 20 of
// exp2(c log2 x), and half the integer glue.  The interval around m = 1 has r = A = 1 exactly and B_0 = 1, so for
// x in [1 - 2^-8, 1 + 2^-8] the parts are s = 1, p = (1 + t)^c - 1 with t = x - 1 exact: x^c - 1 and 1 - x^c keep
// their RELATIVE accuracy there, which the conductivity 1 - (1 - S^(1/m))^m needs for dry cells.
// x must be positive and normal; 0 <= x < 2^-64 returns 0 (B table entry 0); x >= 2.83, negative x and NaN give
// FINITE garbage (the closures discard those lanes: oversaturated cells take the saturated branches, and a NaN state
// reaches the results through their explicit factors of S; lh_pow_bad flags them where that is not so).
// ---------------------------------------------------------------------------------------------------
#define LH_POW_DEG 4
#define LH_POW_EMIN 64
#define LH_POW_NB (LH_POW_EMIN + 2)
#define LH_POW_DOUBLES (LH_LOG_TAB + LH_POW_NB)     // one exponent's table: A_j, then B_e

struct LhPowCoef { double c[LH_POW_DEG + 1]; };     // (1 + t)^c = 1 + t (c[0] + c[1] t + ... + c[4] t^4)

struct LhPowArg { double t; int32_t j, be; };       // reduced argument: shared by every power of the same x

LH_DEV LhPowArg lh_pow_arg(const double* __restrict__ tab, double x)
{
    const int32_t off = lh_hi(x) - LH_LOG_HI0;
    int32_t e = off >> 20;
    e = e < -LH_POW_EMIN ? -LH_POW_EMIN : e;
    e = e > 1 ? 1 : e;
    LhPowArg a;
    a.be = LH_LOG_TAB + LH_POW_EMIN + e;
    a.j = (off >> LH_LOG_SHIFT) & (LH_LOG_TAB - 1);
    const double m = lh_mk(LH_LOG_HI0 + (off & 0xfffff), lh_lo(x));
    a.t = lh_fma(m, tab[LH_TAB_LOG + 2 * a.j], -1.0);
    return a;
}

// x < 0 (incl. -0.0), NaN, +inf: one unsigned compare on the high word.
LH_DEV bool lh_pow_bad(double x) { return (uint32_t)lh_hi(x) > 0x7fefffffu; }

// x^c = s (1 + p); ptab = this exponent's table.
LH_DEV LhExpParts lh_pow_parts(const LhPowCoef& k, const double* __restrict__ ptab, const LhPowArg& a)
{
    double g = k.c[4];
    g = lh_fma(g, a.t, k.c[3]);
    g = lh_fma(g, a.t, k.c[2]);
    g = lh_fma(g, a.t, k.c[1]);
    g = lh_fma(g, a.t, k.c[0]);
    LhExpParts o;
    o.p = a.t * g;
    o.s = ptab[a.j] * ptab[a.be];
    return o;
}

LH_DEV double lh_pow_fixed(const LhPowCoef& k, const double* __restrict__ ptab, const LhPowArg& a)
{
    const LhExpParts e = lh_pow_parts(k, ptab, a);
    return lh_fma(e.s, e.p, e.s);
}

// Host side: coefficients and table of one exponent (long double; entries are correctly rounded in all but ~1e-3 of
// the cases, 0.5 ulp + 2^-64 otherwise).  g interpolates ((1 + t)^c - 1)/t at the 5 Chebyshev nodes of the t range:
// error <= |binom(c, 6)| a^5 / 16 relative to 1 (a = 2^-8): < 2e-18 for |c| <= 4.
#include <math.h>
static inline void lh_pow_build(double c, const double* log_tab /* (r_j, L_j) pairs */, LhPowCoef* coef, double* ptab)
{
    const long double cl = (long double)c;
    for (int j = 0; j < LH_LOG_TAB; ++j) ptab[j] = (double)powl((long double)log_tab[2 * j], -cl);
    ptab[LH_LOG_TAB] = 0.0;                                        // e <= -64: x^c -> 0
    for (int e = -LH_POW_EMIN + 1; e <= 1; ++e) ptab[LH_LOG_TAB + LH_POW_EMIN + e] = (double)exp2l((long double)e * cl);
    // t range of the decomposition: [lo r - 1, hi r - 1] over the intervals; symmetric bound a
    const long double a = 0.0039215L * 1.02L;
    const int n = LH_POW_DEG + 1;
    long double x[LH_POW_DEG + 1], M[LH_POW_DEG + 1][LH_POW_DEG + 2];
    for (int i = 0; i < n; ++i) {
        x[i] = cosl((2 * i + 1) * 3.14159265358979323846264338327950288L / (2 * n));   // in [-1, 1]; t = a x
        const long double t = a * x[i];
        const long double gt = expm1l(cl * log1pl(t)) / t;
        long double pw = 1.0L;
        for (int k = 0; k < n; ++k) { M[i][k] = pw; pw *= x[i]; }
        M[i][n] = gt;
    }
    for (int col = 0; col < n; ++col) {                             // Gauss-Jordan with partial pivoting
        int piv = col;
        for (int r = col + 1; r < n; ++r) if (fabsl(M[r][col]) > fabsl(M[piv][col])) piv = r;
        for (int k = 0; k <= n; ++k) { const long double tmp = M[col][k]; M[col][k] = M[piv][k]; M[piv][k] = tmp; }
        for (int r = 0; r < n; ++r) {
            if (r == col) continue;
            const long double f = M[r][col] / M[col][col];
            for (int k = col; k <= n; ++k) M[r][k] -= f * M[col][k];
        }
    }
    long double sc = 1.0L;
    for (int k = 0; k < n; ++k) { coef->c[k] = (double)(M[k][n] / M[k][k] / sc); sc *= a; }
}
