// Host build of the PRODUCT's closures (csrc/lh_closures.cuh with LH_MATH_HOST: the same source the kernels compile, the MUFU
// seeds emulated at 20-bit precision) so that every kernel variant's K, psi, kappa, T can be compared with the oracle on a box
// without a GPU (tests/test_device_closures_host.py).  Test infrastructure only — nothing in the product links this file.
#define LH_MATH_HOST 1
#include "lh_derive.h"

#include <vector>

namespace {
template <int MODEL, int FLAGS, class P>
void eval_all(const P& p, const double* tab, long n, const double* th, const double* ti, const double* x, double* K, double* psi,
              double* kappa, double* T)
{
    for (long i = 0; i < n; ++i) {
        const LhCell c = lh_cell_closures<MODEL, FLAGS>(p, tab, th[i], ti[i], x[i]);
        K[i] = c.K; psi[i] = c.psi; kappa[i] = c.kappa; T[i] = c.T;
    }
}

template <int MODEL, class P>
int dispatch_flags(int flags, const P& p, const double* tab, long n, const double* th, const double* ti, const double* x, double* K,
                   double* psi, double* kappa, double* T)
{
#define CASE(F) case F: eval_all<MODEL, F>(p, tab, n, th, ti, x, K, psi, kappa, T); return 0;
    switch (flags) {
        CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7)
    default: return -1;
    }
#undef CASE
}

template <int MODEL>
int dispatch_het(int flags, const LhLaneParams& p, const double* tab, long n, const double* th, const double* ti, const double* x,
                 double* K, double* psi, double* kappa, double* T)
{
    // the HET variants: general closures with per-lane parameters (LH_FLAG_HET | LH_FLAG_GEN [| LH_FLAG_ICE])
    if (flags & LH_FLAG_ICE) eval_all<MODEL, LH_FLAG_HET | LH_FLAG_GEN | LH_FLAG_ICE>(p, tab, n, th, ti, x, K, psi, kappa, T);
    else eval_all<MODEL, LH_FLAG_HET | LH_FLAG_GEN>(p, tab, n, th, ti, x, K, psi, kappa, T);
    return 0;
}
}  // namespace

extern "C" {
// K, psi, kappa, T of n cells with the closures of kernel variant (model, flags); flags as lh_closures.cuh (ICE 1, GEN 2, VG2 4,
// HET 8).  x = rho_e_int (heat / coupled) or the prescribed T (Richards).  Returns 0, or -1 for a combination no kernel is
// compiled for / the parameters do not allow (the same rules as update_kernel_flags in csrc/lh_soil_api.cu).
int lhm_cell_closures(const lh_soil_params* q, int model, int flags, long n, const double* th, const double* ti, const double* x,
                      double* K, double* psi, double* kappa, double* T)
{
    lh_soil_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.params = *q;
    cfg.nlayer = 10; cfg.zmin = -1.0; cfg.zmax = 0.0; cfg.ncol = 1; cfg.model = model;
    LhDevParams d = derive_params(cfg);
    std::vector<double> pow_tab;
    derive_pow(d, pow_tab);
    std::vector<double> tab(LH_TAB_ALL);
    lh_stage_tables(d, pow_tab.data(), tab.data(), 0, 1);
    const bool gen_needed = d.visc_on || d.imp_on || !d.om_zero || (model == LH_MODEL_COUPLED && q->theta_r != 0.0);
    if (!(flags & LH_FLAG_GEN) && gen_needed) return -1;                                  // the !GEN fast path does not apply
    if ((flags & LH_FLAG_VG2) && !(q->vg_n == 2.0 && q->vg_m == 0.5)) return -1;
    if (flags & LH_FLAG_HET) {
        LhLaneParams pl;
        static_cast<LhPhys&>(pl) = static_cast<const LhPhys&>(d);
        pl.mc = d.mc;
        if (flags & LH_FLAG_HETH) pl.om_zero = 0;
        switch (model) {
        case 0: return dispatch_het<0>(flags, pl, tab.data(), n, th, ti, x, K, psi, kappa, T);
        case 1: return dispatch_het<1>(flags, pl, tab.data(), n, th, ti, x, K, psi, kappa, T);
        default: return dispatch_het<2>(flags, pl, tab.data(), n, th, ti, x, K, psi, kappa, T);
        }
    }
    switch (model) {
    case 0: return dispatch_flags<0>(flags, d, tab.data(), n, th, ti, x, K, psi, kappa, T);
    case 1: return dispatch_flags<1>(flags & ~LH_FLAG_VG2, d, tab.data(), n, th, ti, x, K, psi, kappa, T);
    default: return dispatch_flags<2>(flags, d, tab.data(), n, th, ti, x, K, psi, kappa, T);
    }
}
}
