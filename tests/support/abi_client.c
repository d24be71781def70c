/*
 * abi_client.c — a NON-Python client of the C ABI (test infrastructure).
 *
 * Includes nothing of this repository but include/lh_soil.h, loads a library that exports that ABI with dlopen, and drives
 * the path the way a reference host would (reference call sites: SoilModel(...) models.jl:115-135, initialize_states
 * initial_conditions.jl:101-107, make_rhs / rhs! right_hand_side.jl:33-44, Simulation / step! / run! simulation.jl:34-87):
 *     create -> set_state (the reference's per-column n x nfields layout, batched columns) -> rhs -> get_tendency
 *            -> step_ssprk33 -> run (snapshots + budgets) -> get_state -> budgets -> checkpoint -> destroy
 * and writes every result to a binary file.  tests/test_abi_client.py runs it against the oracle (CPU) and against the
 * CUDA library (GPU) and compares the two files.
 *
 *     abi_client <library.so> <symbol prefix: lh_ | lho_> <out.bin> [ncol] [nlayer]
 */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lh_soil.h"

#define FN(ret, name, ...) typedef ret (*name##_t)(__VA_ARGS__); static name##_t p_##name
FN(int32_t, lh_soil_abi_version, void);
FN(int32_t, lh_soil_create, const lh_soil_config*, lh_soil_ctx**);
FN(int32_t, lh_soil_destroy, lh_soil_ctx*);
FN(const char*, lh_soil_last_error, const lh_soil_ctx*);
FN(int32_t, lh_soil_get_zc, const lh_soil_ctx*, double*);
FN(int32_t, lh_soil_set_state, lh_soil_ctx*, int32_t, const double*, int64_t, int64_t);
FN(int32_t, lh_soil_get_state, lh_soil_ctx*, int32_t, double*, int64_t, int64_t);
FN(int32_t, lh_soil_rhs, lh_soil_ctx*, double);
FN(int32_t, lh_soil_get_tendency, lh_soil_ctx*, int32_t, double*, int64_t, int64_t);
FN(int32_t, lh_soil_step_ssprk33, lh_soil_ctx*, double, double, int64_t, const double*);
FN(int32_t, lh_soil_run, lh_soil_ctx*, double, double, int64_t, const lh_soil_run_opts*);
FN(int32_t, lh_soil_budgets, lh_soil_ctx*, double*);
FN(int64_t, lh_soil_checkpoint_bytes, const lh_soil_ctx*);
FN(int32_t, lh_soil_checkpoint_save, lh_soil_ctx*, void*, int64_t);
FN(int32_t, lh_soil_checkpoint_load, lh_soil_ctx*, const void*, int64_t);

static void* resolve(void* h, const char* prefix, const char* name)
{
    char sym[128];
    snprintf(sym, sizeof sym, "%s%s", prefix, name + 3);   /* name is "lh_..." */
    void* p = dlsym(h, sym);
    if (!p) { fprintf(stderr, "missing symbol %s\n", sym); exit(3); }
    return p;
}
#define LOAD(name) p_##name = (name##_t)resolve(h, prefix, #name)

static lh_soil_ctx* g_ctx;
#define CHECK(expr)                                                                                     \
    do {                                                                                                \
        int32_t st_ = (expr);                                                                           \
        if (st_ != LH_OK) {                                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #expr, st_, p_lh_soil_last_error(g_ctx));                 \
            exit(4);                                                                                    \
        }                                                                                               \
    } while (0)

/* test/SoilModel/coupled.jl:3-32 */
static void coupled_params(lh_soil_params* q)
{
    memset(q, 0, sizeof *q);
    const double nu = 0.5, nu_om = 0.0, nu_q = 0.92;
    /* k_solid, ksat_unfrozen, ksat_frozen (SoilHeatParameterizations.jl:223-260) */
    const double k_solid = pow(0.25, nu_om) * pow(7.7, nu_q) * pow(2.5, 1.0 - nu_om - nu_q);
    q->nu = nu; q->S_s = 1e-3; q->nu_ss_gravel = 0.0; q->nu_ss_om = nu_om; q->nu_ss_quartz = nu_q;
    q->rho_c_ds = (1.0 - nu) * 1.926e6; q->kappa_solid = k_solid; q->rho_p = 2700.0;
    q->kappa_sat_unfrozen = pow(k_solid, 1.0 - nu) * pow(0.57, nu);
    q->kappa_sat_frozen = pow(k_solid, 1.0 - nu) * pow(2.29, nu);
    q->a = 0.24; q->b = 18.1; q->kappa_dry_parameter = 0.053; q->z_0m = 0.001; q->z_0s = 0.001;
    q->vg_n = 2.0; q->vg_alpha = 2.6; q->vg_m = 1.0 - 1.0 / 2.0; q->theta_r = 0.0; q->Ksat = 0.0443 / 3600 / 100;
    q->viscosity_factor = LH_FACTOR_NONE; q->impedance_factor = LH_FACTOR_NONE;
    q->visc_gamma = 2.64e-2; q->visc_T_ref = 288.0; q->imp_Omega = 7.0;
    q->rho_cloud_liq = 1000.0; q->rho_cloud_ice = 916.7; q->cp_l = 4181.0; q->cp_i = 2100.0; q->T_0 = 273.16;
    q->LH_f0 = 333600.0; q->K_therm = 0.024;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s lib.so prefix out.bin [ncol] [nlayer]\n", argv[0]); return 2; }
    const char* prefix = argv[2];
    const int64_t ncol = argc > 4 ? atoll(argv[4]) : 48;
    const int32_t n = argc > 5 ? atoi(argv[5]) : 20;
    void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
    LOAD(lh_soil_abi_version); LOAD(lh_soil_create); LOAD(lh_soil_destroy); LOAD(lh_soil_last_error); LOAD(lh_soil_get_zc);
    LOAD(lh_soil_set_state); LOAD(lh_soil_get_state); LOAD(lh_soil_rhs); LOAD(lh_soil_get_tendency);
    LOAD(lh_soil_step_ssprk33); LOAD(lh_soil_run); LOAD(lh_soil_budgets); LOAD(lh_soil_checkpoint_bytes);
    LOAD(lh_soil_checkpoint_save); LOAD(lh_soil_checkpoint_load);
    if (p_lh_soil_abi_version() != LH_SOIL_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 3; }

    lh_soil_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.device = 0; cfg.ncol = ncol; cfg.nlayer = n; cfg.model = LH_MODEL_COUPLED; cfg.zmin = -2.0; cfg.zmax = 0.0;
    coupled_params(&cfg.params);
    cfg.top.energy_kind = LH_BC_DIRICHLET; cfg.top.energy_value = 288.0;
    cfg.top.hydrology_kind = LH_BC_DIRICHLET; cfg.top.hydrology_value = 0.4;
    cfg.bottom.energy_kind = LH_BC_FLUX; cfg.bottom.energy_value = 0.0;
    cfg.bottom.hydrology_kind = LH_BC_FREE_DRAINAGE;
    if (p_lh_soil_create(&cfg, &g_ctx) != LH_OK) { fprintf(stderr, "create: %s\n", p_lh_soil_last_error(NULL)); return 4; }

    /* the reference's state: per column an n x nfields matrix, layer fastest, fields (ϑ_l, θ_i, ρe_int) one after the other
     * (initial_conditions.jl:101-107); columns one after the other (HybridBox = batch of columns) */
    const int64_t cs = 3 * (int64_t)n;
    double* Y = (double*)malloc((size_t)(ncol * cs) * sizeof(double));
    double* zc = (double*)malloc((size_t)n * sizeof(double));
    CHECK(p_lh_soil_get_zc(g_ctx, zc));
    const lh_soil_params* q = &cfg.params;
    for (int64_t c = 0; c < ncol; ++c) {
        for (int i = 0; i < n; ++i) {
            const double S = 0.45 + 0.25 * sin(3.0 * zc[i] + 0.37 * (double)c) + 0.1 * cos(1.7 * (double)c);
            const double th = q->nu * S, ti = 0.0;
            const double T = 285.0 + 6.0 * (0.5 + 0.5 * sin(1.3 * zc[i] - 0.11 * (double)c));
            const double rho_c_s = q->rho_c_ds + th * (q->cp_l * q->rho_cloud_liq) + ti * (q->cp_i * q->rho_cloud_ice);
            Y[c * cs + i] = th;
            Y[c * cs + n + i] = ti;
            Y[c * cs + 2 * n + i] = rho_c_s * (T - q->T_0) - ti * q->rho_cloud_ice * q->LH_f0;
        }
    }
    for (int f = 0; f < 3; ++f) CHECK(p_lh_soil_set_state(g_ctx, f, Y + f * n, cs, 1));

    FILE* out = fopen(argv[3], "wb");
    if (!out) { perror("fopen"); return 5; }
    const size_t cells = (size_t)(ncol * n);
    double* buf = (double*)malloc(cells * 2 * sizeof(double));

    /* rhs!(dY, Y, Ya, t) */
    CHECK(p_lh_soil_rhs(g_ctx, 0.0));
    CHECK(p_lh_soil_get_tendency(g_ctx, LH_FIELD_THETA_L, buf, n, 1));
    CHECK(p_lh_soil_get_tendency(g_ctx, LH_FIELD_RHO_E_INT, buf + cells, n, 1));
    fwrite(buf, sizeof(double), cells * 2, out);

    /* step!(sim) x 5 with Dirichlet values evaluated by the host at the stage times */
    const double dt = 20.0;
    double table[5 * 3 * 4];
    for (int s = 0; s < 5; ++s)
        for (int st = 0; st < 3; ++st) {
            const double t = (s + (st == 0 ? 0.0 : st == 1 ? 1.0 : 0.5)) * dt;
            double* r = table + (s * 3 + st) * 4;
            r[LH_BCV_TOP_ENERGY] = 288.0 + 1e-3 * t; r[LH_BCV_TOP_HYDROLOGY] = 0.4; r[LH_BCV_BOTTOM_ENERGY] = 0.0; r[LH_BCV_BOTTOM_HYDROLOGY] = 0.0;
        }
    CHECK(p_lh_soil_step_ssprk33(g_ctx, 0.0, dt, 5, table));
    CHECK(p_lh_soil_get_state(g_ctx, LH_FIELD_THETA_L, buf, n, 1));
    CHECK(p_lh_soil_get_state(g_ctx, LH_FIELD_RHO_E_INT, buf + cells, n, 1));
    fwrite(buf, sizeof(double), cells * 2, out);

    /* checkpoint, run!(sim) with saveat every 2 of 6 steps + budgets every step, then restore and repeat: same bits */
    const int64_t ckb = p_lh_soil_checkpoint_bytes(g_ctx);
    void* ck = malloc((size_t)ckb);
    CHECK(p_lh_soil_checkpoint_save(g_ctx, ck, ckb));
    double budgets[2][6 * 2];
    double* snaps[2];
    for (int rep = 0; rep < 2; ++rep) {
        snaps[rep] = (double*)malloc(4 * 2 * cells * sizeof(double));
        lh_soil_run_opts o;
        memset(&o, 0, sizeof o);
        o.struct_size = (int32_t)sizeof o;
        o.save_first = 1; o.budget_every = 1; o.budgets_out = budgets[rep]; o.save_every = 2;
        o.nsave_fields = 2; o.save_fields[0] = LH_FIELD_THETA_L; o.save_fields[1] = LH_FIELD_RHO_E_INT;
        o.save_out = snaps[rep]; o.snapshot_stride = 2 * (int64_t)cells; o.field_stride = (int64_t)cells; o.col_stride = n; o.layer_stride = 1;
        CHECK(p_lh_soil_run(g_ctx, 5 * dt, dt, 6, &o));
        if (rep == 0) CHECK(p_lh_soil_checkpoint_load(g_ctx, ck, ckb));
    }
    if (memcmp(budgets[0], budgets[1], sizeof budgets[0]) != 0 || memcmp(snaps[0], snaps[1], 4 * 2 * cells * sizeof(double)) != 0) {
        fprintf(stderr, "restart from the checkpoint did not reproduce the run bit for bit\n");
        return 6;
    }
    fwrite(snaps[0], sizeof(double), 4 * 2 * cells, out);
    fwrite(budgets[0], sizeof(double), 12, out);
    double b2[2];
    CHECK(p_lh_soil_budgets(g_ctx, b2));
    fwrite(b2, sizeof(double), 2, out);
    fclose(out);
    CHECK(p_lh_soil_destroy(g_ctx));
    printf("abi_client ok: %lld columns x %d layers\n", (long long)ncol, n);
    return 0;
}
