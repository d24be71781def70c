// Host build of csrc/lh_math.cuh (LH_MATH_HOST): the SAME source the kernels compile, with the
// MUFU seeds emulated at 20-bit precision, so the algorithms' accuracy can be checked on a box
// without a GPU (tests/test_device_math.py).  Test infrastructure only.
#define LH_MATH_HOST 1
#include "lh_math.cuh"

extern "C" {
void lhm_log2(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_log2(lh_c_host, lh_c_host + LHC_TAB0, x[i]); }
void lhm_exp2(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_exp2(lh_c_host, lh_c_host + LHC_TAB0, x[i]); }
void lhm_exp2m1(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_exp2m1(lh_c_host, lh_c_host + LHC_TAB0, x[i]); }
void lhm_one_minus_exp2(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_one_minus_exp2(lh_c_host, lh_c_host + LHC_TAB0, x[i]); }
void lhm_sqrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_sqrt(x[i]); }
void lhm_rsqrt(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_rsqrt(x[i]); }
void lhm_rcp(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_rcp(x[i]); }
void lhm_div(const double* a, const double* b, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = lh_div(a[i], b[i]); }
// x^c with the fixed-exponent tables built for c[0]; mode 0: x^c, 1: x^c - 1 (from the parts), 2: 1 - x^c
void lhm_pow(const double* x, const double* c, double* y, long n)
{
    LhPowCoef k;
    double ptab[LH_POW_DOUBLES];
    lh_pow_build(c[0], lh_c_host + LHC_TAB0 + LH_TAB_LOG, &k, ptab);
    const int mode = (int)c[1];
    for (long i = 0; i < n; ++i) {
        const LhPowArg a = lh_pow_arg(lh_c_host + LHC_TAB0, x[i]);
        const LhExpParts e = lh_pow_parts(k, ptab, a);
        y[i] = mode == 0 ? lh_fma(e.s, e.p, e.s) : mode == 1 ? lh_fma(e.s, e.p, e.s - 1.0) : lh_fma(-e.s, e.p, 1.0 - e.s);
    }
}
}
