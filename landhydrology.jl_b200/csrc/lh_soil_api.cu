// lh_soil_api.cu — the C ABI of include/lh_soil.h over the sm_100a kernels of lh_kernels.cu.
//
// There is NO CPU path in this library: without a CUDA device lh_soil_create returns
// LH_ERR_NO_DEVICE.  NCCL is resolved lazily with dlopen("libnccl.so.2") so a single-GPU host
// needs no NCCL at all; it is used for exactly one thing, the 2-double budget all-reduce.
#include "lh_soil.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "lh_atmos.cuh"
#include "lh_derive.h"
#include "lh_kernels.cuh"

// ------------------------------------------------------------------------------------------------
// NCCL, resolved at run time
// ------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm* ncclComm_t_;
struct NcclUniqueId_ { char internal[128]; };
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId_*) = nullptr;
    int (*CommInitRank)(ncclComm_t_*, int, NcclUniqueId_, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
    int (*CommDestroy)(ncclComm_t_) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble
constexpr int NCCL_SUM = 0;       // ncclSum

NcclApi& nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // RTLD_NOLOAD first: reuse the libnccl already mapped by the host process (e.g. torch's)
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) {
            api.GetUniqueId = (int (*)(NcclUniqueId_*))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (int (*)(ncclComm_t_*, int, NcclUniqueId_, int))dlsym(api.handle, "ncclCommInitRank");
            api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t_, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
            api.CommDestroy = (int (*)(ncclComm_t_))dlsym(api.handle, "ncclCommDestroy");
            api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
            api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
        }
    }
    return api;
}

thread_local char g_create_err[256] = "";
}  // namespace

struct lh_soil_ctx;

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
struct lh_soil_ctx {
    lh_soil_config cfg;
    int device = 0;
    int sm_count = 148;
    int64_t ncol = 0, ncol_pad = 0;
    int32_t nlayer = 0;
    int model = 0;
    LhDevParams dp;
    LhLaunchShape shape;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t flag_stream = nullptr;          // small read-backs that must neither wait in front of kernels nor in front of uploads
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_copy[2] = {nullptr, nullptr}, ev_xpose[2] = {nullptr, nullptr};
    double* U[LH_NUM_FIELDS] = {nullptr, nullptr, nullptr, nullptr};   // ϑ_l, θ_i, ρe_int, T
    double* V[3] = {nullptr, nullptr, nullptr};                        // stage buffer (ϑ_l, -, ρe_int)
    double* tend[3] = {nullptr, nullptr, nullptr};                     // lazily allocated (ϑ_l, -, ρe_int)
    double* zc_dev = nullptr;
    std::vector<double> zc;
    double* stage_dev[2] = {nullptr, nullptr};   // layout staging blocks [chunk_cols][nlayer]
    double* stage_host = nullptr;                // pinned, 2 * chunk
    int64_t chunk_cols = 0;
    double* partials = nullptr;
    int32_t npartials = 0;
    double* budget_dev = nullptr;                // 2 doubles (+2 for the all-reduce result)
    // lh_soil_budgets_async: a ring of pinned result slots, one event each
    static constexpr int BUDGET_SLOTS = 8;
    double* budget_ring_dev = nullptr;           // [BUDGET_SLOTS][2]
    double* budget_ring_host = nullptr;          // pinned, [BUDGET_SLOTS][2]
    cudaEvent_t budget_ev[BUDGET_SLOTS] = {};
    int64_t budget_ticket[BUDGET_SLOTS] = {};    // ticket whose result the slot holds (0: free)
    int64_t budget_next_ticket = 1;
    double* colp_dev = nullptr;                  // [LHCP_COUNT][ncol_pad] per-column derived parameters (heterogeneous soils)
    // what the host supplied: nu, theta_r, vg_n, vg_alpha, Ksat | rho_c_ds, kappa_sat_unfrozen, kappa_sat_frozen, kappa_solid,
    // nu_ss_om, nu_ss_quartz, nu_ss_gravel (empty: the model's scalar)
    std::vector<double> col_user[12];
    bool col_heat = false;                       // some heat parameter is per column -> LH_FLAG_HETH
    // lh_soil_set_cell_params: nu, theta_r, vg_n, vg_alpha, Ksat per CELL, dense [col * nlayer + layer] (empty: per column / scalar)
    std::vector<double> cell_user[5];
    double* cellp_dev = nullptr;                 // [LHCELL_COUNT][nlayer][ncol_pad] derived per-cell parameters -> LH_FLAG_CELLP
    double* pow_tab_dev = nullptr;               // LHPW_COUNT fixed-exponent power tables (lh_math.cuh), built at create
    std::vector<double> pow_tab;                 // their host copy
    double* diag_dev = nullptr;                  // scratch field of lh_soil_diagnostic (lazily allocated)
    bool theta_i_ptr_out = false;                // lh_soil_device_ptr handed out θ_i: the ICE kernels stay selected
    double* fused_partials = nullptr;            // [shape.nblocks][2] budget sums left by the last-stage launches
    int64_t fused_nblocks = 0;
    bool budget_fresh = false;                   // fused_partials describe the current state U
    bool external_writes = false;                // lh_soil_device_ptr handed out U: never trust the fused sums
    double* flux_cols_dev[4] = {nullptr, nullptr, nullptr, nullptr};   // lh_soil_set_column_fluxes: [ncol_pad] each, LH_BCV_* order
    bool atmos_on = false;                       // lh_soil_set_atmos_forcing: the top face takes the turbulent surface fluxes
    LhAtmos atmos;
    double* atm_flux_dev[2] = {nullptr, nullptr};   // per-column heat / water fluxes of the current stage (lh_atmos_flux_kernel)
    int32_t* chain_dev = nullptr;                // per-block completion flags of the stage launches (LhKernelArgs::chain_flags)
    int64_t chain_cap = 0;                       // blocks the flag array holds
    int32_t chain_seq = 0;                       // value the last stage launch published; only ever grows, so that a flag
                                                 // left over from an earlier launch can never equal a value waited for
    bool chain_break = true;                     // the next stage launch takes the whole-grid dependency (first launch, new shape)
    // lh_soil_set_aux_table: time-dependent prescribed profiles, one row of nlayer values per stage launch
    double* aux_tab_dev[LH_NUM_FIELDS] = {nullptr, nullptr, nullptr, nullptr};
    int64_t aux_tab_rows[LH_NUM_FIELDS] = {0, 0, 0, 0};
    int64_t aux_row = 0;                         // next row to consume
    // lh_soil_run: budget history and double-buffered snapshots
    double* hist_dev = nullptr;                  // [hist_cap][2]
    double* hist_host = nullptr;                 // pinned
    int64_t hist_cap = 0;
    double* snap_dev[2][LH_NUM_FIELDS] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    cudaEvent_t ev_snap_ready[2] = {nullptr, nullptr}, ev_snap_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_hist = nullptr;               // lh_soil_run: a step's budgets are in the history buffer (the copy stream waits for it)
    double* bc_dev = nullptr;                    // boundary-value table of a persistent launch
    int64_t bc_dev_steps = 0;
    unsigned long long* nonfinite_dev = nullptr;
    double bcv[4] = {0, 0, 0, 0};
    ncclComm_t_ comm = nullptr;
    int nranks = 1, rank = 0;
    bool has_ice = false;        // some θ_i != 0 (θ_i is constant in time: dθ_i ≡ 0), re-evaluated on every θ_i upload
    // The answer to "any ice?" after a θ_i upload is fetched lazily: the upload only enqueues the scan and a 4-byte read into
    // pinned memory (its own stream, behind an event); the first call that needs the kernel variant waits for it (resolve_ice).
    int* ice_flag_dev = nullptr;
    int* ice_flag_host = nullptr;                // pinned
    cudaEvent_t ev_ice_scan = nullptr, ev_ice = nullptr;
    bool ice_pending = false;
    int kernel_flags = 0;        // LH_FLAG_ICE | LH_FLAG_GEN | LH_FLAG_VG2 -> compiled kernel variant
    bool force_general_vg = false;   // LH_FLAG_GENERAL_VG: never take the n == 2 shortcut (benchmark the general path)
    bool timing_valid = false;
    int64_t last_launches = 0;
    char err[512] = "";
};

namespace {

int32_t fail(lh_soil_ctx* ctx, int32_t code, const char* fmt, ...)
{
    char* dst = ctx ? ctx->err : g_create_err;
    size_t cap = ctx ? sizeof(ctx->err) : sizeof(g_create_err);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, cap, fmt, ap);
    va_end(ap);
    return code;
}

#define LH_CUDA(ctx, expr)                                                                          \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(ctx, LH_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int32_t resolve_ice(lh_soil_ctx* c);
#define LH_RESOLVE(ctx)                                                                             \
    do {                                                                                            \
        int32_t r_ = resolve_ice(ctx);                                                              \
        if (r_ != LH_OK) return r_;                                                                 \
    } while (0)

bool has_water(int m) { return m == LH_MODEL_RICHARDS || m == LH_MODEL_COUPLED; }
bool has_heat(int m) { return m == LH_MODEL_HEAT || m == LH_MODEL_COUPLED; }

// The reference's vertical_flux method table (boundary_conditions.jl:295-444).
int32_t validate_face(const lh_soil_face_bc& bc, int model, const char* face)
{
    const int ek = bc.energy_kind, hk = bc.hydrology_kind;
    if (ek < 0 || ek > 3 || hk < 0 || hk > 3) return fail(nullptr, LH_ERR_INVALID_ARG, "%s: unknown BC kind", face);
    if (has_heat(model)) {
        if (!(ek == LH_BC_FLUX || ek == LH_BC_DIRICHLET))
            return fail(nullptr, LH_ERR_UNSUPPORTED_BC, "%s: energy BC kind %d has no vertical_flux method for SoilEnergyModel", face, ek);
    } else if (!(ek == LH_BC_NONE || ek == LH_BC_FLUX)) {
        return fail(nullptr, LH_ERR_UNSUPPORTED_BC, "%s: energy BC kind %d has no vertical_flux method for PrescribedTemperatureModel", face, ek);
    }
    if (has_water(model)) {
        if (!(hk == LH_BC_FLUX || hk == LH_BC_DIRICHLET || hk == LH_BC_FREE_DRAINAGE))
            return fail(nullptr, LH_ERR_UNSUPPORTED_BC, "%s: hydrology BC kind %d has no vertical_flux method for SoilHydrologyModel", face, hk);
    } else if (!(hk == LH_BC_NONE || hk == LH_BC_FLUX)) {
        return fail(nullptr, LH_ERR_UNSUPPORTED_BC, "%s: hydrology BC kind %d has no vertical_flux method for PrescribedHydrologyModel", face, hk);
    }
    return LH_OK;
}

void free_all(lh_soil_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    for (auto& p : c->U) if (p) cudaFree(p);
    for (auto& p : c->V) if (p) cudaFree(p);
    for (auto& p : c->tend) if (p) cudaFree(p);
    for (auto& p : c->stage_dev) if (p) cudaFree(p);
    if (c->stage_host) cudaFreeHost(c->stage_host);
    if (c->zc_dev) cudaFree(c->zc_dev);
    if (c->partials) cudaFree(c->partials);
    if (c->budget_dev) cudaFree(c->budget_dev);
    if (c->budget_ring_dev) cudaFree(c->budget_ring_dev);
    if (c->budget_ring_host) cudaFreeHost(c->budget_ring_host);
    for (auto& e : c->budget_ev) if (e) cudaEventDestroy(e);
    if (c->bc_dev) cudaFree(c->bc_dev);
    if (c->chain_dev) cudaFree(c->chain_dev);
    for (auto& p : c->flux_cols_dev) if (p) cudaFree(p);
    for (auto& p : c->atm_flux_dev) if (p) cudaFree(p);
    for (auto& p : c->aux_tab_dev) if (p) cudaFree(p);
    if (c->hist_dev) cudaFree(c->hist_dev);
    if (c->hist_host) cudaFreeHost(c->hist_host);
    for (auto& b : c->snap_dev) for (auto& p : b) if (p) cudaFree(p);
    for (auto& e : c->ev_snap_ready) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_snap_done) if (e) cudaEventDestroy(e);
    if (c->ev_hist) cudaEventDestroy(c->ev_hist);
    if (c->ice_flag_dev) cudaFree(c->ice_flag_dev);
    if (c->ice_flag_host) cudaFreeHost(c->ice_flag_host);
    if (c->ev_ice_scan) cudaEventDestroy(c->ev_ice_scan);
    if (c->ev_ice) cudaEventDestroy(c->ev_ice);
    if (c->fused_partials) cudaFree(c->fused_partials);
    if (c->colp_dev) cudaFree(c->colp_dev);
    if (c->cellp_dev) cudaFree(c->cellp_dev);
    if (c->pow_tab_dev) cudaFree(c->pow_tab_dev);
    if (c->diag_dev) cudaFree(c->diag_dev);
    if (c->nonfinite_dev) cudaFree(c->nonfinite_dev);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    for (auto& e : c->ev_copy) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_xpose) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->flag_stream) cudaStreamDestroy(c->flag_stream);
    delete c;
}

size_t field_bytes(const lh_soil_ctx* c) { return (size_t)c->ncol_pad * c->nlayer * sizeof(double); }

bool field_ok(int f) { return f >= 0 && f < LH_NUM_FIELDS; }

int32_t ensure_staging(lh_soil_ctx* c)
{
    if (c->stage_dev[0] && c->stage_dev[1] && c->stage_host) return LH_OK;     // (all three: an earlier attempt may have run out of memory half way)
    // ~32 MiB blocks: large enough for PCIe efficiency, small enough to pipeline copy and transpose.  (LH_STAGE_BLOCK_BYTES,
    // read here, once per ctx: a tuning knob, and how the tests push many small blocks through the two-buffer pipeline.)
    const char* env_block = getenv("LH_STAGE_BLOCK_BYTES");
    const long long env_bytes = env_block ? atoll(env_block) : 0;
    const int64_t block_bytes = env_bytes > 0 ? (int64_t)env_bytes : (int64_t)(32ll << 20);
    if (c->chunk_cols == 0) {                    // (a retry after a failed attempt keeps the block size of the blocks it already has)
        int64_t cols = std::max<int64_t>(32, block_bytes / ((int64_t)c->nlayer * 8));
        c->chunk_cols = std::min<int64_t>((cols + 31) / 32 * 32, c->ncol_pad);
    }
    const size_t bytes = (size_t)c->chunk_cols * c->nlayer * sizeof(double);
    for (int k = 0; k < 2; ++k) if (!c->stage_dev[k]) LH_CUDA(c, cudaMalloc(&c->stage_dev[k], bytes));
    if (!c->stage_host) LH_CUDA(c, cudaMallocHost(&c->stage_host, 2 * bytes));
    return LH_OK;
}

// host (col_stride, layer_stride) -> device SoA.  Pipelined in column blocks over two staging
// buffers: [gather into pinned block if the host layout is not dense] -> H2D -> transpose kernel.
int32_t upload_field(lh_soil_ctx* c, double* soa, const double* host, int64_t cs, int64_t ls)
{
    if (!host) return fail(c, LH_ERR_INVALID_ARG, "host pointer is NULL");
    LH_CUDA(c, cudaSetDevice(c->device));
    const int n = c->nlayer;
    if (cs == 0) {   // one nlayer profile broadcast to all columns
        std::vector<double> prof(n);
        for (int i = 0; i < n; ++i) prof[i] = host[(int64_t)i * ls];
        LH_CUDA(c, cudaMemcpyAsync((c->zc_dev + c->nlayer), prof.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));   // prof is a stack vector
        LH_CUDA(c, lh_launch_fill_profile((c->zc_dev + c->nlayer), soa, n, c->ncol_pad, c->stream));
        return LH_OK;
    }
    if (cs == 1 && ls == c->ncol_pad && c->ncol == c->ncol_pad) {   // already device layout
        LH_CUDA(c, cudaMemcpyAsync(soa, host, field_bytes(c), cudaMemcpyHostToDevice, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        return LH_OK;
    }
    if (cs == 1 && n == 1) {   // a single layer: the row of columns is contiguous whatever the layer stride says
        LH_CUDA(c, cudaMemcpyAsync(soa, host, c->ncol * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LH_CUDA(c, lh_launch_fill_padding(soa, c->ncol, c->ncol_pad, n, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        return LH_OK;
    }
    if (cs == 1) {   // column-fastest host block with its own layer stride
        LH_CUDA(c, cudaMemcpy2DAsync(soa, c->ncol_pad * sizeof(double), host, ls * sizeof(double),
                                     c->ncol * sizeof(double), n, cudaMemcpyHostToDevice, c->stream));
        LH_CUDA(c, lh_launch_fill_padding(soa, c->ncol, c->ncol_pad, n, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        return LH_OK;
    }
    int32_t st = ensure_staging(c);
    if (st != LH_OK) return st;
    const bool dense = (ls == 1 && cs == n);
    const size_t blk = (size_t)c->chunk_cols * n;
    int k = 0;
    for (int64_t c0 = 0; c0 < c->ncol; c0 += c->chunk_cols, k ^= 1) {
        const int64_t m = std::min<int64_t>(c->chunk_cols, c->ncol - c0);
        // buffer k is free once the transpose that last read it has finished
        LH_CUDA(c, cudaEventSynchronize(c->ev_xpose[k]));
        const double* src;
        if (dense) {
            src = host + c0 * n;
        } else {
            double* pin = c->stage_host + (size_t)k * blk;
            for (int64_t cc = 0; cc < m; ++cc)
                for (int i = 0; i < n; ++i) pin[cc * n + i] = host[(c0 + cc) * cs + (int64_t)i * ls];
            src = pin;
        }
        LH_CUDA(c, cudaMemcpyAsync(c->stage_dev[k], src, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
        LH_CUDA(c, cudaEventRecord(c->ev_copy[k], c->copy_stream));
        LH_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copy[k], 0));
        LH_CUDA(c, lh_launch_to_soa(c->stage_dev[k], soa, c0, m, n, c->ncol_pad, c->stream));
        LH_CUDA(c, cudaEventRecord(c->ev_xpose[k], c->stream));
    }
    LH_CUDA(c, lh_launch_fill_padding(soa, c->ncol, c->ncol_pad, n, c->stream));
    // Host pointers are borrowed for the call only: wait for the LAST H2D COPY, not for the transposes that trail it on the
    // compute stream.  Everything that consumes the field afterwards is enqueued on that same stream (or waits for it: downloads,
    // snapshots, lh_soil_sync), and a staging block is only overwritten after its ev_xpose event — so returning here lets the
    // next upload (this ctx's next field, or another shard's) claim the PCIe link while the transposes are still queued behind
    // other contexts' stage kernels.  In a sharded end-to-end run that removes one full drain of the GPU queue per field.
    LH_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    return LH_OK;
}

int32_t download_field(lh_soil_ctx* c, const double* soa, double* host, int64_t cs, int64_t ls)
{
    if (!host) return fail(c, LH_ERR_INVALID_ARG, "host pointer is NULL");
    if (cs == 0 && c->ncol != 1) return fail(c, LH_ERR_INVALID_ARG, "col_stride 0 is only valid for uploads");
    LH_CUDA(c, cudaSetDevice(c->device));
    const int n = c->nlayer;
    if (cs == 1 && n == 1) {
        LH_CUDA(c, cudaMemcpyAsync(host, soa, c->ncol * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        return LH_OK;
    }
    if (cs == 1 || c->ncol == 1) {
        const int64_t lstride = ls;
        LH_CUDA(c, cudaMemcpy2DAsync(host, lstride * sizeof(double), soa, c->ncol_pad * sizeof(double),
                                     c->ncol * sizeof(double), n, cudaMemcpyDeviceToHost, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        return LH_OK;
    }
    int32_t st = ensure_staging(c);
    if (st != LH_OK) return st;
    const bool dense = (ls == 1 && cs == n);
    const size_t blk = (size_t)c->chunk_cols * n;
    // pipeline: transpose block j on `stream`, D2H on `copy_stream`, scatter on the host
    struct Pending { int64_t c0, m; bool active; } pend[2] = {{0, 0, false}, {0, 0, false}};
    auto finish = [&](int k) -> int32_t {
        if (!pend[k].active) return LH_OK;
        LH_CUDA(c, cudaEventSynchronize(c->ev_copy[k]));
        if (!dense) {
            const double* pin = c->stage_host + (size_t)k * blk;
            for (int64_t cc = 0; cc < pend[k].m; ++cc)
                for (int i = 0; i < n; ++i) host[(pend[k].c0 + cc) * cs + (int64_t)i * ls] = pin[cc * n + i];
        }
        pend[k].active = false;
        return LH_OK;
    };
    int k = 0;
    for (int64_t c0 = 0; c0 < c->ncol; c0 += c->chunk_cols, k ^= 1) {
        const int64_t m = std::min<int64_t>(c->chunk_cols, c->ncol - c0);
        if ((st = finish(k)) != LH_OK) return st;   // staging buffer k must be drained first
        LH_CUDA(c, lh_launch_from_soa(soa, c->stage_dev[k], c0, m, n, c->ncol_pad, c->stream));
        LH_CUDA(c, cudaEventRecord(c->ev_xpose[k], c->stream));
        LH_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_xpose[k], 0));
        double* dst = dense ? host + c0 * n : c->stage_host + (size_t)k * blk;
        LH_CUDA(c, cudaMemcpyAsync(dst, c->stage_dev[k], (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        LH_CUDA(c, cudaEventRecord(c->ev_copy[k], c->copy_stream));
        pend[k] = {c0, m, true};
    }
    if ((st = finish(0)) != LH_OK) return st;
    if ((st = finish(1)) != LH_OK) return st;
    LH_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    // later kernels on `stream` must not overwrite a staging buffer that copy_stream still reads:
    // both events are complete here.
    return LH_OK;
}

void update_kernel_flags(lh_soil_ctx* c)
{
    const lh_soil_params& q = c->cfg.params;
    // theta_r != 0 only matters where the Kersten number reuses log S (coupled model); Richards has no
    // Kersten number and the heat-only model no water closures.
    const bool het = c->colp_dev != nullptr;             // per-column parameters: general closures, per-lane values
    const bool gen = het || c->dp.visc_on || c->dp.imp_on || !c->dp.om_zero ||
                     (c->model == LH_MODEL_COUPLED && q.theta_r != 0.0);
    const bool vg2 = !het && q.vg_n == 2.0 && q.vg_m == 0.5;      // S^(1/m) = S^2, x^m = sqrt(x): no log/exp needed
    c->kernel_flags = (c->has_ice ? LH_FLAG_ICE : 0) | (gen ? LH_FLAG_GEN : 0) | (vg2 && !c->force_general_vg ? LH_FLAG_VG2 : 0) |
                      (het ? LH_FLAG_HET : 0) | (het && c->col_heat && has_heat(c->model) ? LH_FLAG_HETH : 0) |
                      (het && c->cellp_dev ? LH_FLAG_CELLP : 0);
    c->shape = lh_choose_shape(c->model, c->ncol_pad, c->nlayer, c->sm_count, het);   // the HET variants have a smaller warp budget
    c->budget_fresh = false;
    // The block -> column-group map may have changed: the next stage launch takes the full grid dependency.  (The flags
    // keep their values; a later chained launch waits for the value ITS predecessor publishes.)
    c->chain_break = true;
    if (c->chain_cap < c->shape.nblocks) {
        if (c->chain_dev) { cudaStreamSynchronize(c->stream); cudaFree(c->chain_dev); }
        c->chain_dev = nullptr;
        c->chain_cap = 0;
        if (cudaMalloc(&c->chain_dev, (size_t)c->shape.nblocks * sizeof(int32_t)) == cudaSuccess &&
            cudaMemset(c->chain_dev, 0, (size_t)c->shape.nblocks * sizeof(int32_t)) == cudaSuccess) c->chain_cap = c->shape.nblocks;
        else { if (c->chain_dev) cudaFree(c->chain_dev); c->chain_dev = nullptr; }     // stage launches then never chain
    }
    if (c->fused_nblocks < c->shape.nblocks) {        // per-block budget sums of the last-stage launches
        if (c->fused_partials) cudaFree(c->fused_partials);
        c->fused_partials = nullptr;
        c->fused_nblocks = 0;
        if (cudaMalloc(&c->fused_partials, (size_t)c->shape.nblocks * 2 * sizeof(double)) == cudaSuccess) c->fused_nblocks = c->shape.nblocks;
        else c->fused_partials = nullptr;              // the budgets then come from the full pass
    }
}

// The pending answer of the last θ_i scan, if any: wait for it and select the kernel variant.  Called at the top of every entry
// point that launches kernels, reports the variant or changes what selects it.
int32_t resolve_ice(lh_soil_ctx* c)
{
    if (!c->ice_pending) return LH_OK;
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaEventSynchronize(c->ev_ice));
    c->ice_pending = false;
    c->has_ice = *c->ice_flag_host != 0 || c->theta_i_ptr_out;   // a caller holding the raw θ_i pointer may write ice at any time
    update_kernel_flags(c);
    return LH_OK;
}

// θ_i was (re)written: is there any ice?  One pass over the field; θ_i never changes afterwards.  Only ENQUEUED here — an upload
// must not drain the ctx stream (the next field's H2D copy, or another shard's, is waiting for the PCIe link), and the 4-byte
// answer must not sit in front of this stream's kernels in a copy engine's queue, nor in front of the next field's H2D copies on
// the copy stream: it travels on a stream of its own.
int32_t detect_ice(lh_soil_ctx* c)
{
    int32_t st = resolve_ice(c);                 // two θ_i uploads in a row: the first answer is consumed before its slot is reused
    if (st != LH_OK) return st;
    LH_CUDA(c, cudaMemsetAsync(c->ice_flag_dev, 0, sizeof(int), c->stream));
    LH_CUDA(c, lh_launch_any_nonzero(c->U[1], (int64_t)c->ncol_pad * c->nlayer, c->ice_flag_dev, c->stream));
    LH_CUDA(c, cudaEventRecord(c->ev_ice_scan, c->stream));
    LH_CUDA(c, cudaStreamWaitEvent(c->flag_stream, c->ev_ice_scan, 0));
    LH_CUDA(c, cudaMemcpyAsync(c->ice_flag_host, c->ice_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, c->flag_stream));
    LH_CUDA(c, cudaEventRecord(c->ev_ice, c->flag_stream));
    c->ice_pending = true;
    c->budget_fresh = false;
    return LH_OK;
}

void fill_args(lh_soil_ctx* c, int stage, double dt, LhKernelArgs& a)
{
    a.p = c->dp;
    const bool from_V = stage >= 2;
    a.io.in_th = (from_V && has_water(c->model)) ? c->V[0] : c->U[0];
    a.io.in_ti = c->U[1];
    a.io.in_re = (from_V && has_heat(c->model)) ? c->V[2] : c->U[2];
    a.io.aux_T = c->U[3];
    a.io.u0_th = c->U[0];
    a.io.u0_re = c->U[2];
    if (stage == 0) { a.io.out_th = c->tend[0]; a.io.out_re = c->tend[2]; }
    else if (stage == 3) { a.io.out_th = c->U[0]; a.io.out_re = c->U[2]; }
    else { a.io.out_th = c->V[0]; a.io.out_re = c->V[2]; }
    a.zc = c->zc_dev;
    a.colp = c->colp_dev;
    a.cellp = c->cellp_dev;
    a.pow_tab = c->pow_tab_dev;
    a.budget_partials = c->fused_partials;
    a.ncol = c->ncol;
    a.ncol_pad = c->ncol_pad;
    a.nlayer = c->nlayer;
    a.Lc = c->shape.Lc;
    a.W = c->shape.W;
    a.top_e_kind = c->cfg.top.energy_kind;
    a.top_h_kind = c->cfg.top.hydrology_kind;
    a.bot_e_kind = c->cfg.bottom.energy_kind;
    a.bot_h_kind = c->cfg.bottom.hydrology_kind;
    memcpy(a.io.bcv, c->bcv, sizeof a.io.bcv);
    for (int k = 0; k < 4; ++k) a.io.flux_cols[k] = c->flux_cols_dev[k];
    a.io.dt = dt;
    a.io.out2_th = c->V[0];
    a.io.out2_re = c->V[2];
    a.io.sa = 0.0; a.io.sb = 1.0; a.io.sg = 1.0;
    a.io.first2n = 0;
    a.io.budget = 0;
    a.nsteps = 0;
    a.bc_dev = nullptr;
    a.chain_flags = nullptr;
    a.chain_wait = a.chain_set = 0;
}

// Every one-stage launch goes through here: block j waits for block j of the previous stage launch of this ctx (same
// shape since then, see update_kernel_flags) instead of the whole previous grid, and publishes its own completion.
// Anything else enqueued on the stream in between (copies, transposes, budget kernels: none of them triggers dependents
// early) is fully ordered before the launch by the stream itself.
cudaError_t launch_chained(lh_soil_ctx* c, int stage, LhKernelArgs& a)
{
    if (c->atmos_on) {
        // boundary_fluxes(X, bc::PrescribedAtmosForcing, :top, ...) (boundary_conditions.jl:516-536): one Monin-Obukhov solve per
        // column from the top cell of THIS stage's input, then the stage kernel takes the two arrays as per-column fluxes.
        const int64_t top = (int64_t)(c->nlayer - 1) * c->ncol_pad;
        cudaError_t e = lh_launch_atmos_fluxes(c->dp, c->pow_tab_dev, c->atmos, a.io.in_th + top, a.io.in_ti + top, a.io.in_re + top,
                                               c->atm_flux_dev[0], c->atm_flux_dev[1], c->ncol_pad, c->colp_dev, (c->col_heat || c->cellp_dev) ? 1 : 0,
                                               c->cellp_dev ? c->cellp_dev + top : nullptr, (int64_t)c->nlayer * c->ncol_pad, c->stream);
        if (e != cudaSuccess) return e;
        a.top_e_kind = LH_BC_FLUX;
        a.top_h_kind = LH_BC_FLUX;
        a.io.flux_cols[LH_BCV_TOP_ENERGY] = c->atm_flux_dev[0];
        a.io.flux_cols[LH_BCV_TOP_HYDROLOGY] = c->atm_flux_dev[1];
    }
    if (c->chain_dev && !(c->cfg.flags & LH_FLAG_NO_CHAIN)) {
        a.chain_flags = c->chain_dev;
        if (c->chain_seq == 0x7fffffff) {                // (2^31 launches: start over behind a full synchronisation)
            cudaStreamSynchronize(c->stream);
            cudaMemset(c->chain_dev, 0, (size_t)c->chain_cap * sizeof(int32_t));
            c->chain_seq = 0;
            c->chain_break = true;
        }
        a.chain_wait = c->chain_break ? 0 : c->chain_seq;       // 0: griddepcontrol.wait
        a.chain_set = ++c->chain_seq;
        c->chain_break = false;
    }
    return lh_launch_stage(c->model, stage, c->kernel_flags, a, c->shape, c->stream);
}

int32_t check_finite(lh_soil_ctx* c)
{
    LH_CUDA(c, cudaMemsetAsync(c->nonfinite_dev, 0, sizeof(unsigned long long), c->stream));
    const int64_t n = (int64_t)c->ncol_pad * c->nlayer;
    if (has_water(c->model)) LH_CUDA(c, lh_launch_count_nonfinite(c->U[0], n, c->nonfinite_dev, c->stream));
    if (has_heat(c->model)) LH_CUDA(c, lh_launch_count_nonfinite(c->U[2], n, c->nonfinite_dev, c->stream));
    unsigned long long cnt = 0;
    LH_CUDA(c, cudaMemcpyAsync(&cnt, c->nonfinite_dev, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (cnt) return fail(c, LH_ERR_NONFINITE, "%llu non-finite state values (the reference raises DomainError)", cnt);
    return LH_OK;
}

// HBM bytes per cell and SSPRK33 step the selected kernel variant moves (per-stage launches): every stage reads the
// stage input, stages 2 and 3 also u^n, every stage writes the prognostic fields; θ_i is read only by the ICE variants,
// the prescribed T only by the Richards variants with the viscosity factor on.
int lh_bytes_on_wire(const lh_soil_ctx* c)
{
    const int nprog = c->model == LH_MODEL_COUPLED ? 2 : 1;
    int per_stage_in = nprog;                                        // stage input of the prognostic fields
    if (c->model == LH_MODEL_HEAT) per_stage_in += 1;                // prescribed ϑ_l
    if (c->kernel_flags & LH_FLAG_ICE) per_stage_in += 1;            // θ_i
    if (c->model == LH_MODEL_RICHARDS && (c->kernel_flags & LH_FLAG_GEN)) per_stage_in += 1;   // prescribed T row
    if (c->kernel_flags & LH_FLAG_CELLP) per_stage_in += LHCELL_COUNT;   // per-cell parameter fields
    return 8 * (3 * per_stage_in + 2 * nprog + 3 * nprog);
}

}  // namespace

extern "C" {

static bool use_persistent(const lh_soil_ctx* c);
static int32_t local_budgets(lh_soil_ctx* c, double* out_dev = nullptr);
static int32_t apply_aux_tables(lh_soil_ctx* c);

int32_t lh_soil_abi_version(void) { return LH_SOIL_ABI_VERSION; }

const char* lh_soil_last_error(const lh_soil_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

int32_t lh_soil_create(const lh_soil_config* cfg, lh_soil_ctx** out)
{
    if (!cfg || !out) return fail(nullptr, LH_ERR_INVALID_ARG, "cfg/out is NULL");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(lh_soil_config))
        return fail(nullptr, LH_ERR_INVALID_ARG, "lh_soil_config.struct_size mismatch (%d != %zu)", cfg->struct_size, sizeof(lh_soil_config));
    if (cfg->ncol < 1 || cfg->nlayer < 1) return fail(nullptr, LH_ERR_INVALID_ARG, "ncol and nlayer must be >= 1");
    if (cfg->model < 0 || cfg->model > 2) return fail(nullptr, LH_ERR_INVALID_ARG, "unknown model kind %d", cfg->model);
    if (!(cfg->zmin < cfg->zmax)) return fail(nullptr, LH_ERR_DOMAIN, "zlim[1] < zlim[2] violated");   // domain.jl:30
    // vanGenuchten{FT}(; n, α, Ksat, θr) stores m = 1 - 1/n (SoilWaterParameterizations.jl:162-169); the closures use
    // 1/n = 1 - m (lh_closures.cuh)
    if (!(fabs(cfg->params.vg_m - (1.0 - 1.0 / cfg->params.vg_n)) <= 4.0 * LH_EPS))
        return fail(nullptr, LH_ERR_INVALID_ARG, "vg_m must be 1 - 1/vg_n (van Genuchten-Mualem, as the reference constructor stores it)");
    int32_t st;
    if ((st = validate_face(cfg->top, cfg->model, "top")) != LH_OK) return st;
    if ((st = validate_face(cfg->bottom, cfg->model, "bottom")) != LH_OK) return st;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, LH_ERR_NO_DEVICE, "no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, LH_ERR_NO_DEVICE, "device ordinal %d out of range (%d devices)", cfg->device, ndev);

    lh_soil_ctx* c = new (std::nothrow) lh_soil_ctx();
    if (!c) return fail(nullptr, LH_ERR_INVALID_ARG, "out of memory");
    c->cfg = *cfg;
    c->device = cfg->device;
    c->ncol = cfg->ncol;
    c->ncol_pad = (cfg->ncol + 31) / 32 * 32;
    c->nlayer = cfg->nlayer;
    c->model = cfg->model;
    c->dp = derive_params(*cfg);
    derive_pow(c->dp, c->pow_tab);
    c->force_general_vg = (cfg->flags & LH_FLAG_GENERAL_VG) != 0;
    c->bcv[LH_BCV_TOP_ENERGY] = cfg->top.energy_value;
    c->bcv[LH_BCV_TOP_HYDROLOGY] = cfg->top.hydrology_value;
    c->bcv[LH_BCV_BOTTOM_ENERGY] = cfg->bottom.energy_value;
    c->bcv[LH_BCV_BOTTOM_HYDROLOGY] = cfg->bottom.hydrology_value;

#define LH_CREATE_CUDA(expr)                                                                        \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            fail(nullptr, LH_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));             \
            free_all(c);                                                                            \
            return LH_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

    LH_CREATE_CUDA(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    LH_CREATE_CUDA(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    LH_CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    LH_CREATE_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LH_CREATE_CUDA(cudaStreamCreateWithFlags(&c->flag_stream, cudaStreamNonBlocking));
    LH_CREATE_CUDA(cudaEventCreate(&c->ev_start));
    LH_CREATE_CUDA(cudaEventCreate(&c->ev_stop));
    for (int k = 0; k < 2; ++k) {
        LH_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_copy[k], cudaEventDisableTiming));
        LH_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_xpose[k], cudaEventDisableTiming));
    }
    c->shape = lh_choose_shape(c->model, c->ncol_pad, c->nlayer, c->sm_count, false);
    update_kernel_flags(c);

    const size_t fb = field_bytes(c);
    // ϑ_l, θ_i always exist (prognostic or prescribed); ρe_int with an energy model; T aux only
    // for the Richards model (Ya.soil.T)
    LH_CREATE_CUDA(cudaMalloc(&c->U[0], fb));
    LH_CREATE_CUDA(cudaMalloc(&c->U[1], fb));
    LH_CREATE_CUDA(cudaMemsetAsync(c->U[0], 0, fb, c->stream));
    LH_CREATE_CUDA(cudaMemsetAsync(c->U[1], 0, fb, c->stream));
    if (has_heat(c->model)) {
        LH_CREATE_CUDA(cudaMalloc(&c->U[2], fb));
        LH_CREATE_CUDA(cudaMemsetAsync(c->U[2], 0, fb, c->stream));
        LH_CREATE_CUDA(cudaMalloc(&c->V[2], fb));
    }
    if (has_water(c->model)) LH_CREATE_CUDA(cudaMalloc(&c->V[0], fb));

    // z_c: faces zmin + j (zmax - zmin) / n, centres are face midpoints (domain.jl:58-69)
    c->zc.resize(c->nlayer);
    for (int i = 0; i < c->nlayer; ++i) {
        const double zf0 = cfg->zmin + (cfg->zmax - cfg->zmin) * (double)i / (double)c->nlayer;
        const double zf1 = cfg->zmin + (cfg->zmax - cfg->zmin) * (double)(i + 1) / (double)c->nlayer;
        c->zc[i] = (zf0 + zf1) / 2.0;
    }
    // zc_dev holds zc followed by an nlayer profile scratch
    LH_CREATE_CUDA(cudaMalloc(&c->zc_dev, 2 * c->nlayer * sizeof(double)));
    LH_CREATE_CUDA(cudaMemcpyAsync(c->zc_dev, c->zc.data(), c->nlayer * sizeof(double), cudaMemcpyHostToDevice, c->stream));

    if (c->model == LH_MODEL_RICHARDS) {
        // PrescribedTemperatureModel default T ≡ 288 (models.jl:51-54)
        LH_CREATE_CUDA(cudaMalloc(&c->U[3], fb));
        std::vector<double> prof(c->nlayer, 288.0);
        LH_CREATE_CUDA(cudaMemcpyAsync(c->zc_dev + c->nlayer, prof.data(), c->nlayer * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LH_CREATE_CUDA(cudaStreamSynchronize(c->stream));
        LH_CREATE_CUDA(lh_launch_fill_profile(c->zc_dev + c->nlayer, c->U[3], c->nlayer, c->ncol_pad, c->stream));
    }
    LH_CREATE_CUDA(cudaMalloc(&c->pow_tab_dev, c->pow_tab.size() * sizeof(double)));
    LH_CREATE_CUDA(cudaMemcpyAsync(c->pow_tab_dev, c->pow_tab.data(), c->pow_tab.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->npartials = (int32_t)std::min<int64_t>(1024, std::max<int64_t>(1, (c->ncol + 255) / 256));
    LH_CREATE_CUDA(cudaMalloc(&c->partials, 2 * c->npartials * sizeof(double)));
    LH_CREATE_CUDA(cudaMalloc(&c->budget_dev, 4 * sizeof(double)));
    LH_CREATE_CUDA(cudaMalloc(&c->nonfinite_dev, sizeof(unsigned long long)));
    LH_CREATE_CUDA(cudaMalloc(&c->ice_flag_dev, sizeof(int)));
    LH_CREATE_CUDA(cudaMallocHost(&c->ice_flag_host, sizeof(int)));
    LH_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_ice_scan, cudaEventDisableTiming));
    LH_CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_ice, cudaEventDisableTiming));
    LH_CREATE_CUDA(cudaStreamSynchronize(c->stream));
#undef LH_CREATE_CUDA
    *out = c;
    return LH_OK;
}

int32_t lh_soil_destroy(lh_soil_ctx* ctx)
{
    free_all(ctx);
    return LH_OK;
}

int32_t lh_soil_get_zc(const lh_soil_ctx* ctx, double* zc_out)
{
    if (!ctx || !zc_out) return LH_ERR_INVALID_ARG;
    memcpy(zc_out, ctx->zc.data(), ctx->nlayer * sizeof(double));
    return LH_OK;
}

int32_t lh_soil_set_state(lh_soil_ctx* c, int32_t field, const double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (!field_ok(field)) return fail(c, LH_ERR_INVALID_ARG, "bad field id %d", field);
    if (!c->U[field]) return fail(c, LH_ERR_INVALID_ARG, "field %d does not exist for model kind %d", field, c->model);
    c->budget_fresh = false;
    int32_t st = upload_field(c, c->U[field], host, cs, ls);
    if (st == LH_OK && field == LH_FIELD_THETA_I) st = detect_ice(c);
    return st;
}

int32_t lh_soil_set_aux(lh_soil_ctx* c, int32_t field, const double* host, int64_t cs, int64_t ls)
{
    return lh_soil_set_state(c, field, host, cs, ls);
}

int32_t lh_soil_get_state(lh_soil_ctx* c, int32_t field, double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (!field_ok(field)) return fail(c, LH_ERR_INVALID_ARG, "bad field id %d", field);
    if (!c->U[field]) return fail(c, LH_ERR_INVALID_ARG, "field %d does not exist for model kind %d", field, c->model);
    return download_field(c, c->U[field], host, cs, ls);
}

// Derive, per column, exactly what derive_phys derives for the model (same expressions, same rounding) from the host's
// per-column arrays, and upload the table the HET / HETH kernels read per lane.
static int32_t rebuild_column_params(lh_soil_ctx* c)
{
    LH_RESOLVE(c);
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    bool any = false, heat = false, cells = false;
    for (int k = 0; k < 12; ++k) { any |= !c->col_user[k].empty(); if (k >= 5) heat |= !c->col_user[k].empty(); }
    for (int k = 0; k < 5; ++k) cells |= !c->cell_user[k].empty();
    c->col_heat = heat;
    if (!cells && c->cellp_dev) { LH_CUDA(c, cudaFree(c->cellp_dev)); c->cellp_dev = nullptr; }
    if (!any && !cells) {                                          // back to the homogeneous kernels
        if (c->colp_dev) { LH_CUDA(c, cudaFree(c->colp_dev)); c->colp_dev = nullptr; }
        update_kernel_flags(c);
        return LH_OK;
    }
    std::vector<double> h((size_t)LHCP_COUNT * c->ncol_pad);
    for (int64_t j = 0; j < c->ncol_pad; ++j) {
        const int64_t col = std::min<int64_t>(j, c->ncol - 1);                  // padding replicates the last column
        lh_soil_config cfg = c->cfg;
        lh_soil_params& q = cfg.params;
        auto get = [&](int k, double& dst) { if (!c->col_user[k].empty()) dst = c->col_user[k][col]; };
        get(0, q.nu); get(1, q.theta_r);
        if (!c->col_user[2].empty()) { q.vg_n = c->col_user[2][col]; q.vg_m = 1.0 - 1.0 / q.vg_n; }   // the reference constructor's m
        get(3, q.vg_alpha); get(4, q.Ksat);
        get(5, q.rho_c_ds); get(6, q.kappa_sat_unfrozen); get(7, q.kappa_sat_frozen); get(8, q.kappa_solid);
        get(9, q.nu_ss_om); get(10, q.nu_ss_quartz); get(11, q.nu_ss_gravel);
        if (!(q.nu > q.theta_r) || !(q.vg_n > 1.0) || !(q.vg_alpha > 0.0) || !std::isfinite(q.Ksat))
            return fail(c, LH_ERR_INVALID_ARG, "column %lld: need nu > theta_r, vg_n > 1, vg_alpha > 0, finite Ksat", (long long)col);
        if (!(q.kappa_sat_unfrozen > 0.0) || !(q.kappa_sat_frozen > 0.0) || !std::isfinite(q.rho_c_ds))
            return fail(c, LH_ERR_INVALID_ARG, "column %lld: need kappa_sat_unfrozen > 0, kappa_sat_frozen > 0, finite rho_c_ds", (long long)col);
        LhPhys d;
        derive_phys(cfg, d);
        double* o = h.data() + j;
        const int64_t st = c->ncol_pad;
        o[LHCP_NU * st] = d.nu; o[LHCP_THETA_R * st] = d.theta_r; o[LHCP_THETA_R_EPS * st] = d.theta_r_eps;
        o[LHCP_INV_NU_THR * st] = d.inv_nu_thr; o[LHCP_NU_THR * st] = d.nu_thr;
        o[LHCP_VG_M * st] = d.vg_m; o[LHCP_VG_INV_M * st] = d.vg_inv_m; o[LHCP_VG_INV_N * st] = d.vg_inv_n;
        o[LHCP_NEG_INV_ALPHA * st] = d.neg_inv_alpha; o[LHCP_KSAT * st] = d.Ksat;
        o[LHCP_INV_NU * st] = d.inv_nu; o[LHCP_KAPPA_DRY * st] = d.kappa_dry;
        o[LHCP_RHO_C_DS * st] = d.rho_c_ds; o[LHCP_KERSTEN_P1 * st] = d.kersten_p1; o[LHCP_KERSTEN_P2 * st] = d.kersten_p2;
        o[LHCP_KERSTEN_P3 * st] = d.kersten_p3; o[LHCP_K_UNFROZEN * st] = d.k_unfrozen;
        o[LHCP_LOG2_K_UNFROZEN * st] = d.log2_k_unfrozen; o[LHCP_LOG2_K_FROZEN * st] = d.log2_k_frozen;
    }
    if (!c->colp_dev) LH_CUDA(c, cudaMalloc(&c->colp_dev, h.size() * sizeof(double)));
    LH_CUDA(c, cudaMemcpyAsync(c->colp_dev, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (cells) {
        // per-cell fields: the same derivation cell by cell (the column's / model's values for what is not given per cell)
        const int n = c->nlayer;
        const size_t fs = (size_t)n * c->ncol_pad;
        std::vector<double> hc((size_t)LHCELL_COUNT * fs);
        for (int64_t j = 0; j < c->ncol_pad; ++j) {
            const int64_t col = std::min<int64_t>(j, c->ncol - 1);
            lh_soil_config cfg = c->cfg;
            lh_soil_params& q = cfg.params;
            auto getc = [&](int k, double& dst) { if (!c->col_user[k].empty()) dst = c->col_user[k][col]; };
            getc(0, q.nu); getc(1, q.theta_r);
            if (!c->col_user[2].empty()) q.vg_n = c->col_user[2][col];
            getc(3, q.vg_alpha); getc(4, q.Ksat); getc(8, q.kappa_solid);
            const lh_soil_params qcol = q;
            for (int i = 0; i < n; ++i) {
                q = qcol;
                auto get = [&](int k, double& dst) { if (!c->cell_user[k].empty()) dst = c->cell_user[k][(size_t)col * n + i]; };
                get(0, q.nu); get(1, q.theta_r); get(2, q.vg_n); get(3, q.vg_alpha); get(4, q.Ksat);
                q.vg_m = 1.0 - 1.0 / q.vg_n;
                if (!(q.nu > q.theta_r) || !(q.vg_n > 1.0) || !(q.vg_alpha > 0.0) || !std::isfinite(q.Ksat))
                    return fail(c, LH_ERR_INVALID_ARG, "column %lld layer %d: need nu > theta_r, vg_n > 1, vg_alpha > 0, finite Ksat", (long long)col, i);
                LhPhys d;
                derive_phys(cfg, d);
                double* o = hc.data() + (size_t)i * c->ncol_pad + j;
                o[LHCELL_NU * fs] = d.nu; o[LHCELL_THETA_R * fs] = d.theta_r; o[LHCELL_INV_NU_THR * fs] = d.inv_nu_thr;
                o[LHCELL_VG_M * fs] = d.vg_m; o[LHCELL_VG_INV_M * fs] = d.vg_inv_m; o[LHCELL_NEG_INV_ALPHA * fs] = d.neg_inv_alpha;
                o[LHCELL_KSAT * fs] = d.Ksat; o[LHCELL_INV_NU * fs] = d.inv_nu; o[LHCELL_KAPPA_DRY * fs] = d.kappa_dry;
            }
        }
        if (!c->cellp_dev) LH_CUDA(c, cudaMalloc(&c->cellp_dev, hc.size() * sizeof(double)));
        LH_CUDA(c, cudaMemcpy(c->cellp_dev, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    update_kernel_flags(c);
    return LH_OK;
}

static void keep_user_columns(lh_soil_ctx* c, int first, int count, const double* const* src)
{
    for (int k = 0; k < count; ++k) {
        if (src[k]) c->col_user[first + k].assign(src[k], src[k] + c->ncol);
        else c->col_user[first + k].clear();
    }
}

int32_t lh_soil_set_column_params(lh_soil_ctx* c, const double* nu, const double* theta_r, const double* vg_n,
                                  const double* vg_alpha, const double* Ksat)
{
    if (!c) return LH_ERR_INVALID_ARG;
    const double* src[5] = {nu, theta_r, vg_n, vg_alpha, Ksat};
    keep_user_columns(c, 0, 5, src);
    return rebuild_column_params(c);
}

int32_t lh_soil_set_cell_params(lh_soil_ctx* c, const double* nu, const double* theta_r, const double* vg_n, const double* vg_alpha,
                                const double* Ksat, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    const double* src[5] = {nu, theta_r, vg_n, vg_alpha, Ksat};
    const int n = c->nlayer;
    for (int k = 0; k < 5; ++k) {
        if (!src[k]) { c->cell_user[k].clear(); continue; }
        c->cell_user[k].resize((size_t)c->ncol * n);
        for (int64_t col = 0; col < c->ncol; ++col)
            for (int i = 0; i < n; ++i) c->cell_user[k][(size_t)col * n + i] = src[k][col * cs + (int64_t)i * ls];
    }
    return rebuild_column_params(c);
}

int32_t lh_soil_set_column_heat_params(lh_soil_ctx* c, const double* rho_c_ds, const double* kappa_sat_unfrozen,
                                       const double* kappa_sat_frozen, const double* kappa_solid, const double* nu_ss_om,
                                       const double* nu_ss_quartz, const double* nu_ss_gravel)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (!has_heat(c->model) && (rho_c_ds || kappa_sat_unfrozen || kappa_sat_frozen || kappa_solid || nu_ss_om || nu_ss_quartz || nu_ss_gravel))
        return fail(c, LH_ERR_INVALID_ARG, "model kind %d has no energy equation: no heat parameters to set", c->model);
    const double* src[7] = {rho_c_ds, kappa_sat_unfrozen, kappa_sat_frozen, kappa_solid, nu_ss_om, nu_ss_quartz, nu_ss_gravel};
    keep_user_columns(c, 5, 7, src);
    return rebuild_column_params(c);
}

int32_t lh_soil_set_column_fluxes(lh_soil_ctx* c, const double* const values[4])
{
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    const int kinds[4] = {c->cfg.top.energy_kind, c->cfg.top.hydrology_kind, c->cfg.bottom.energy_kind, c->cfg.bottom.hydrology_kind};
    for (int k = 0; k < 4; ++k)
        if (values && values[k] && kinds[k] != LH_BC_FLUX) return fail(c, LH_ERR_INVALID_ARG, "per-column fluxes need a face of kind LH_BC_FLUX (boundary value %d is of kind %d)", k, kinds[k]);
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; ++k) {
        if (!(values && values[k])) {
            if (c->flux_cols_dev[k]) { LH_CUDA(c, cudaFree(c->flux_cols_dev[k])); c->flux_cols_dev[k] = nullptr; }
            continue;
        }
        if (!c->flux_cols_dev[k]) LH_CUDA(c, cudaMalloc(&c->flux_cols_dev[k], (size_t)c->ncol_pad * sizeof(double)));
        std::vector<double> h((size_t)c->ncol_pad);
        for (int64_t j = 0; j < c->ncol_pad; ++j) h[j] = values[k][std::min<int64_t>(j, c->ncol - 1)];
        LH_CUDA(c, cudaMemcpy(c->flux_cols_dev[k], h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return LH_OK;
}

int32_t lh_soil_set_atmos_forcing(lh_soil_ctx* c, const lh_soil_atmos* a)
{
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (!a) { c->atmos_on = false; return LH_OK; }
    if (a->struct_size != (int32_t)sizeof(lh_soil_atmos)) return fail(c, LH_ERR_INVALID_ARG, "lh_soil_atmos.struct_size mismatch");
    if (c->model != LH_MODEL_COUPLED)           // boundary_conditions.jl:103-112: both components must be prognostic
        return fail(c, LH_ERR_UNSUPPORTED_BC, "PrescribedAtmosForcing needs SoilEnergyModel + SoilHydrologyModel (model kind %d given)", c->model);
    const lh_soil_params& q = c->cfg.params;
    if (!(a->z_atm > 0.0) || !(a->rho_a_sfc > 0.0) || !(a->theta_scale > 0.0) || !(q.z_0m > 0.0) || !(q.z_0s > 0.0))
        return fail(c, LH_ERR_INVALID_ARG, "PrescribedAtmosForcing needs z_atm, rho_a_sfc, theta_scale, z_0m, z_0s > 0");
    LH_CUDA(c, cudaSetDevice(c->device));
    for (auto& p : c->atm_flux_dev) if (!p) LH_CUDA(c, cudaMalloc(&p, (size_t)c->ncol_pad * sizeof(double)));
    LhAtmos& d = c->atmos;
    d.u_atm = a->u_atm; d.theta_atm = a->theta_atm; d.z_atm = a->z_atm; d.theta_scale = a->theta_scale; d.rho_a_sfc = a->rho_a_sfc; d.q_atm = a->q_atm;
    d.R_v = a->R_v; d.R_d = a->R_d; d.grav = a->grav; d.cp_d = a->cp_d; d.cp_v = a->cp_v; d.LH_v0 = a->LH_v0;
    d.press_triple = a->press_triple; d.T_triple = a->T_triple; d.von_karman = a->von_karman;
    d.Pr_0 = a->Pr_0; d.a_m = a->a_m; d.a_h = a->a_h;
    d.cp_l = q.cp_l; d.T_0 = q.T_0; d.rho_l = q.rho_cloud_liq; d.z_0m = q.z_0m; d.z_0s = q.z_0s;
    c->atmos_on = true;
    return LH_OK;
}

int32_t lh_soil_atmos_fluxes(lh_soil_ctx* c, const double* th, const double* ti, const double* T, int64_t n, double* heat, double* water)
{
    if (!c || !th || !ti || !T || !heat || !water || n < 0) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (!c->atmos_on) return fail(c, LH_ERR_STATE, "lh_soil_set_atmos_forcing has not been called");
    if (n == 0) return LH_OK;
    LH_CUDA(c, cudaSetDevice(c->device));
    double* d = nullptr;
    LH_CUDA(c, cudaMalloc(&d, (size_t)n * 5 * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, th, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, ti, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + 2 * n, T, n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lh_launch_atmos_eval(c->dp, c->pow_tab_dev, c->atmos, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, n, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(heat, d + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(water, d + 4 * n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(c, LH_ERR_CUDA, "lh_soil_atmos_fluxes: %s", cudaGetErrorString(e));
    return LH_OK;
}

int32_t lh_soil_set_bc_values(lh_soil_ctx* c, const double values[4])
{
    if (!c || !values) return LH_ERR_INVALID_ARG;
    memcpy(c->bcv, values, sizeof c->bcv);
    return LH_OK;
}

static int32_t launch_stage(lh_soil_ctx* c, int stage, double dt)
{
    LhKernelArgs a;
    fill_args(c, stage, dt, a);
    LH_CUDA(c, launch_chained(c, stage, a));
    return LH_OK;
}

int32_t lh_soil_rhs(lh_soil_ctx* c, double t)
{
    (void)t;
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    LH_CUDA(c, cudaSetDevice(c->device));
    const size_t fb = field_bytes(c);
    if (has_water(c->model) && !c->tend[0]) LH_CUDA(c, cudaMalloc(&c->tend[0], fb));
    if (has_heat(c->model) && !c->tend[2]) LH_CUDA(c, cudaMalloc(&c->tend[2], fb));
    int32_t st = launch_stage(c, 0, 0.0);
    if (st != LH_OK) return st;
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) {
        LH_CUDA(c, cudaMemsetAsync(c->nonfinite_dev, 0, sizeof(unsigned long long), c->stream));
        const int64_t n = (int64_t)c->ncol_pad * c->nlayer;
        if (c->tend[0]) LH_CUDA(c, lh_launch_count_nonfinite(c->tend[0], n, c->nonfinite_dev, c->stream));
        if (c->tend[2]) LH_CUDA(c, lh_launch_count_nonfinite(c->tend[2], n, c->nonfinite_dev, c->stream));
        unsigned long long cnt = 0;
        LH_CUDA(c, cudaMemcpyAsync(&cnt, c->nonfinite_dev, sizeof cnt, cudaMemcpyDeviceToHost, c->stream));
        LH_CUDA(c, cudaStreamSynchronize(c->stream));
        if (cnt) return fail(c, LH_ERR_NONFINITE, "%llu non-finite tendency values (the reference raises DomainError)", cnt);
    }
    return LH_OK;
}

int32_t lh_soil_get_tendency(lh_soil_ctx* c, int32_t field, double* host, int64_t cs, int64_t ls)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (field < 0 || field > 2) return fail(c, LH_ERR_INVALID_ARG, "bad field id %d", field);
    if (!host) return fail(c, LH_ERR_INVALID_ARG, "host pointer is NULL");
    if (field == LH_FIELD_THETA_I || !c->tend[field]) {
        // dθ_i ≡ 0 (right_hand_side.jl:182,359); non-prognostic fields have no tendency
        if (field != LH_FIELD_THETA_I && !((field == 0 && has_water(c->model)) || (field == 2 && has_heat(c->model))))
            return fail(c, LH_ERR_INVALID_ARG, "field %d is not prognostic for model kind %d", field, c->model);
        if (field != LH_FIELD_THETA_I) return fail(c, LH_ERR_STATE, "lh_soil_rhs has not been called");
        for (int64_t col = 0; col < c->ncol; ++col)
            for (int i = 0; i < c->nlayer; ++i) host[col * cs + (int64_t)i * ls] = 0.0;
        return LH_OK;
    }
    return download_field(c, c->tend[field], host, cs, ls);
}

int32_t lh_soil_stage_ssprk33(lh_soil_ctx* c, int32_t stage, double dt)
{
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (stage < 1 || stage > 3) return fail(c, LH_ERR_INVALID_ARG, "stage must be 1, 2 or 3");
    LH_CUDA(c, cudaSetDevice(c->device));
    int32_t st = apply_aux_tables(c);
    if (st != LH_OK) return st;
    st = launch_stage(c, stage, dt);
    if (st == LH_OK && stage == 3) c->budget_fresh = c->fused_partials != nullptr;   // stages 1, 2 do not touch U
    return st;
}

// The persistent launch pays for grids of at most ~1.5 waves (small domains, launch-bound): measured
// +25..40 % there (profiles/r01_s_persistent_vs_stage.log).  From ~5 waves on it ties with, then loses to, the
// per-stage launches: it executes ~10 % more instructions per stage (its one run-time-stage body reads u^n in
// every stage) and the kernel is issue-bound.
static bool use_persistent(const lh_soil_ctx* c)
{
    if (c->atmos_on) return false;              // the atmospheric fluxes are re-evaluated by their own kernel before every stage
    if (c->cfg.flags & LH_FLAG_STAGE_LAUNCHES) return false;
    if (c->cfg.flags & LH_FLAG_PERSISTENT) return true;
    return c->shape.waves <= 1.5;
}

static bool has_aux_table(const lh_soil_ctx* c)
{
    for (int f = 0; f < LH_NUM_FIELDS; ++f) if (c->aux_tab_dev[f]) return true;
    return false;
}

// Before a stage launch: broadcast the next row of every prescribed-profile table into its field (device side, no host
// round trip: what update_aux! does in the reference before every rhs! call, right_hand_side.jl:54-81).
static int32_t apply_aux_tables(lh_soil_ctx* c)
{
    bool any = false;
    for (int f = 0; f < LH_NUM_FIELDS; ++f) {
        if (!c->aux_tab_dev[f]) continue;
        if (c->aux_row >= c->aux_tab_rows[f])
            return fail(c, LH_ERR_STATE, "prescribed-profile table of field %d exhausted (%lld rows): upload the rows of the next stages with lh_soil_set_aux_table",
                        f, (long long)c->aux_tab_rows[f]);
        LH_CUDA(c, lh_launch_fill_profile(c->aux_tab_dev[f] + c->aux_row * c->nlayer, c->U[f], c->nlayer, c->ncol_pad, c->stream));
        any = true;
    }
    if (any) ++c->aux_row;
    return LH_OK;
}

// nsteps SSPRK33 steps enqueued on the ctx stream (no timing events, no synchronisation).
// bc_dev_ready: the rows of bc_table for these steps are already in device memory there (lh_soil_run uploads the table of the
// whole run once; bc_table itself is still needed for the boundary values the ctx keeps after the call).
static int32_t advance_ssprk33(lh_soil_ctx* c, double dt, int64_t nsteps, const double* bc_table, int64_t* launches_out,
                               const double* bc_dev_ready = nullptr)
{
    const bool tables = has_aux_table(c);
    const bool persistent = use_persistent(c) && nsteps > 0 && !tables;
    if (persistent && bc_table && !bc_dev_ready) {
        if (c->bc_dev_steps < nsteps) {
            if (c->bc_dev) { LH_CUDA(c, cudaFree(c->bc_dev)); c->bc_dev = nullptr; c->bc_dev_steps = 0; }
            LH_CUDA(c, cudaMalloc(&c->bc_dev, (size_t)nsteps * 12 * sizeof(double)));
            c->bc_dev_steps = nsteps;
        }
        // pageable source: the call returns once the table has been staged, so the host buffer is only borrowed
        LH_CUDA(c, cudaMemcpyAsync(c->bc_dev, bc_table, (size_t)nsteps * 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    int64_t launches = 0;
    if (persistent) {
        const int64_t MAX_STEPS_PER_LAUNCH = 4096;
        for (int64_t s0 = 0; s0 < nsteps; s0 += MAX_STEPS_PER_LAUNCH) {
            LhKernelArgs a;
            fill_args(c, 1, dt, a);
            a.nsteps = std::min<int64_t>(MAX_STEPS_PER_LAUNCH, nsteps - s0);
            a.bc_dev = bc_table ? (bc_dev_ready ? bc_dev_ready : c->bc_dev) + s0 * 12 : nullptr;
            LH_CUDA(c, lh_launch_ssprk33_persistent(c->model, c->kernel_flags, a, c->shape, c->stream));
            ++launches;
        }
        if (bc_table) memcpy(c->bcv, bc_table + (nsteps * 3 - 1) * 4, sizeof c->bcv);   // what the per-stage path leaves behind
    } else {
        for (int64_t s = 0; s < nsteps; ++s) {
            for (int stage = 1; stage <= 3; ++stage) {
                if (bc_table) memcpy(c->bcv, bc_table + (s * 3 + (stage - 1)) * 4, sizeof c->bcv);
                int32_t st;
                if (tables && (st = apply_aux_tables(c)) != LH_OK) return st;
                if ((st = launch_stage(c, stage, dt)) != LH_OK) return st;
            }
        }
        launches = 3 * nsteps;
    }
    if (nsteps > 0) c->budget_fresh = c->fused_partials != nullptr;       // the last stage-3 launch summed what it wrote
    if (launches_out) *launches_out += launches;
    return LH_OK;
}

int32_t lh_soil_step_ssprk33(lh_soil_ctx* c, double t, double dt, int64_t nsteps, const double* bc_table)
{
    (void)t;
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (nsteps < 0) return fail(c, LH_ERR_INVALID_ARG, "nsteps < 0");
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    int64_t launches = 0;
    int32_t st = advance_ssprk33(c, dt, nsteps, bc_table, &launches);
    if (st != LH_OK) return st;
    LH_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    c->timing_valid = true;
    c->last_launches = launches;
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) return check_finite(c);
    return LH_OK;
}

int32_t lh_soil_set_aux_table(lh_soil_ctx* c, int32_t field, const double* table, int64_t nrows)
{
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (!field_ok(field) || !c->U[field]) return fail(c, LH_ERR_INVALID_ARG, "field %d does not exist for model kind %d", field, c->model);
    const bool prescribed = (c->model == LH_MODEL_RICHARDS && field == LH_FIELD_T) ||
                            (c->model == LH_MODEL_HEAT && (field == LH_FIELD_THETA_L || field == LH_FIELD_THETA_I));
    if (!prescribed) return fail(c, LH_ERR_INVALID_ARG, "field %d is not a prescribed profile of model kind %d", field, c->model);
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->aux_tab_dev[field]) { LH_CUDA(c, cudaFree(c->aux_tab_dev[field])); c->aux_tab_dev[field] = nullptr; c->aux_tab_rows[field] = 0; }
    c->aux_row = 0;                                     // every table restarts at its first row
    if (!table || nrows <= 0) return LH_OK;
    const size_t bytes = (size_t)nrows * c->nlayer * sizeof(double);
    LH_CUDA(c, cudaMalloc(&c->aux_tab_dev[field], bytes));
    LH_CUDA(c, cudaMemcpyAsync(c->aux_tab_dev[field], table, bytes, cudaMemcpyHostToDevice, c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    c->aux_tab_rows[field] = nrows;
    if (field == LH_FIELD_THETA_I) {                    // ice anywhere in the table selects the ICE kernels for the whole run
        bool ice = false;
        for (int64_t i = 0; i < nrows * c->nlayer && !ice; ++i) ice = !(table[i] == 0.0);
        if (ice && !c->has_ice) { c->has_ice = true; c->theta_i_ptr_out = true; update_kernel_flags(c); }
    }
    return LH_OK;
}

// ------------------------------------------------------------------------------------------------
// lh_soil_run: a whole run!() with saveat snapshots and per-step budgets inside ONE call
// ------------------------------------------------------------------------------------------------
namespace {
// Non-blocking download of one SoA field into a dense host block, everything on copy_stream.
int32_t enqueue_download(lh_soil_ctx* c, const double* soa, double* host, int64_t cs, int64_t ls)
{
    const int n = c->nlayer;
    if (cs == 1 && n == 1) {   // a single layer: the row of columns is contiguous whatever the layer stride says (a 2-D copy
                               // would be handed a destination pitch smaller than its width: the reference layout has ls = 1)
        LH_CUDA(c, cudaMemcpyAsync(host, soa, c->ncol * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        return LH_OK;
    }
    if (cs == 1 || c->ncol == 1) {
        LH_CUDA(c, cudaMemcpy2DAsync(host, ls * sizeof(double), soa, c->ncol_pad * sizeof(double), c->ncol * sizeof(double), n,
                                     cudaMemcpyDeviceToHost, c->copy_stream));
        return LH_OK;
    }
    int k = 0;
    for (int64_t c0 = 0; c0 < c->ncol; c0 += c->chunk_cols, k ^= 1) {
        const int64_t m = std::min<int64_t>(c->chunk_cols, c->ncol - c0);
        LH_CUDA(c, lh_launch_from_soa(soa, c->stage_dev[k], c0, m, n, c->ncol_pad, c->copy_stream));
        LH_CUDA(c, cudaMemcpyAsync(host + c0 * n, c->stage_dev[k], (size_t)m * n * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    }
    return LH_OK;
}
}  // namespace

int32_t lh_soil_run(lh_soil_ctx* c, double t0, double dt, int64_t nsteps, const lh_soil_run_opts* o)
{
    (void)t0;
    if (!c || !o) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (o->struct_size != (int32_t)sizeof(lh_soil_run_opts)) return fail(c, LH_ERR_INVALID_ARG, "lh_soil_run_opts.struct_size mismatch");
    if (nsteps < 0 || o->budget_every < 0 || o->save_every < 0) return fail(c, LH_ERR_INVALID_ARG, "negative step count or cadence");
    if (o->budget_every > 0 && !o->budgets_out) return fail(c, LH_ERR_INVALID_ARG, "budget_every > 0 needs budgets_out");
    const bool saving = o->save_every > 0 || o->save_first;
    if (saving) {
        if (!o->save_out || o->nsave_fields < 1 || o->nsave_fields > LH_NUM_FIELDS) return fail(c, LH_ERR_INVALID_ARG, "snapshots need save_out and 1..%d fields", LH_NUM_FIELDS);
        for (int k = 0; k < o->nsave_fields; ++k)
            if (!field_ok(o->save_fields[k]) || !c->U[o->save_fields[k]]) return fail(c, LH_ERR_INVALID_ARG, "snapshot field %d does not exist", o->save_fields[k]);
        const bool ref_layout = o->layer_stride == 1 && o->col_stride == c->nlayer;
        const bool soa_layout = o->col_stride == 1 && o->layer_stride >= c->ncol;
        if (!ref_layout && !soa_layout && c->ncol != 1)
            return fail(c, LH_ERR_INVALID_ARG, "snapshots are written densely: (col_stride, layer_stride) = (nlayer, 1) or (1, >= ncol)");
    }
    LH_CUDA(c, cudaSetDevice(c->device));
    int32_t st;
    const int64_t nb = o->budget_every > 0 ? nsteps / o->budget_every : 0;
    if (nb > c->hist_cap) {
        if (c->hist_dev) { LH_CUDA(c, cudaFree(c->hist_dev)); c->hist_dev = nullptr; }
        if (c->hist_host) { LH_CUDA(c, cudaFreeHost(c->hist_host)); c->hist_host = nullptr; }
        c->hist_cap = 0;
        LH_CUDA(c, cudaMalloc(&c->hist_dev, (size_t)nb * 2 * sizeof(double)));
        LH_CUDA(c, cudaMallocHost(&c->hist_host, (size_t)nb * 2 * sizeof(double)));
        c->hist_cap = nb;
    }
    if (nb > 0 && !c->ev_hist) LH_CUDA(c, cudaEventCreateWithFlags(&c->ev_hist, cudaEventDisableTiming));
    if (saving) {
        if ((st = ensure_staging(c)) != LH_OK) return st;
        for (int b = 0; b < 2; ++b) {
            if (!c->ev_snap_ready[b]) LH_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_ready[b], cudaEventDisableTiming));
            if (!c->ev_snap_done[b]) LH_CUDA(c, cudaEventCreateWithFlags(&c->ev_snap_done[b], cudaEventDisableTiming));
            for (int k = 0; k < o->nsave_fields; ++k) {
                const int f = o->save_fields[k];
                if (!c->snap_dev[b][f]) LH_CUDA(c, cudaMalloc(&c->snap_dev[b][f], field_bytes(c)));
            }
        }
    }
    int64_t nsnap = 0;
    // Snapshot: the state is copied device-to-device on the compute stream (0.2 ms per GB), which then goes on stepping;
    // the transposes and the PCIe transfer of the copy run on copy_stream, two snapshots deep.
    auto snapshot = [&]() -> int32_t {
        const int b = (int)(nsnap & 1);
        if (nsnap >= 2) LH_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_snap_done[b], 0));
        for (int k = 0; k < o->nsave_fields; ++k) {
            const int f = o->save_fields[k];
            LH_CUDA(c, cudaMemcpyAsync(c->snap_dev[b][f], c->U[f], field_bytes(c), cudaMemcpyDeviceToDevice, c->stream));
        }
        LH_CUDA(c, cudaEventRecord(c->ev_snap_ready[b], c->stream));
        LH_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_snap_ready[b], 0));
        for (int k = 0; k < o->nsave_fields; ++k) {
            const int f = o->save_fields[k];
            int32_t s2 = enqueue_download(c, c->snap_dev[b][f], o->save_out + nsnap * o->snapshot_stride + k * o->field_stride, o->col_stride, o->layer_stride);
            if (s2 != LH_OK) return s2;
        }
        LH_CUDA(c, cudaEventRecord(c->ev_snap_done[b], c->copy_stream));
        ++nsnap;
        return LH_OK;
    };

    // Persistent launches read their boundary values from device memory.  The table of the WHOLE run goes up once, before the
    // first step: uploaded per advance call (every step when the budgets are read every step) each cudaMemcpyAsync from the
    // caller's pageable table would first synchronise the ctx stream — the host would block once per step, and the copy would
    // queue behind other contexts' transfers.
    const bool table_on_device = o->bc_table && nsteps > 0 && use_persistent(c) && !has_aux_table(c);
    if (table_on_device) {
        if (c->bc_dev_steps < nsteps) {
            if (c->bc_dev) { LH_CUDA(c, cudaFree(c->bc_dev)); c->bc_dev = nullptr; c->bc_dev_steps = 0; }
            LH_CUDA(c, cudaMalloc(&c->bc_dev, (size_t)nsteps * 12 * sizeof(double)));
            c->bc_dev_steps = nsteps;
        }
        LH_CUDA(c, cudaMemcpyAsync(c->bc_dev, o->bc_table, (size_t)nsteps * 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    LH_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    int64_t launches = 0, done = 0, nbud = 0;
    if (o->save_first && (st = snapshot()) != LH_OK) return st;
    while (done < nsteps) {
        int64_t n = nsteps - done;
        if (o->budget_every > 0) n = std::min<int64_t>(n, o->budget_every - done % o->budget_every);
        if (o->save_every > 0) n = std::min<int64_t>(n, o->save_every - done % o->save_every);
        if ((st = advance_ssprk33(c, dt, n, o->bc_table ? o->bc_table + done * 12 : nullptr, &launches,
                                  table_on_device ? c->bc_dev + done * 12 : nullptr)) != LH_OK) return st;
        done += n;
        if (o->budget_every > 0 && done % o->budget_every == 0) {
            // The reduction writes this step's pair straight into its own slot of the history buffer, and the 16-byte D2H read
            // of it goes to the small-read-back stream behind an event: nothing the compute stream has to wait for ever sits in a copy
            // engine's queue — with other contexts' 32 MiB transfers ahead of it there, a per-step copy on the compute
            // stream would stall the next step's launches for the length of a transfer block.
            if ((st = local_budgets(c, c->hist_dev + 2 * nbud)) != LH_OK) return st;
            LH_CUDA(c, cudaEventRecord(c->ev_hist, c->stream));
            LH_CUDA(c, cudaStreamWaitEvent(c->flag_stream, c->ev_hist, 0));
            LH_CUDA(c, cudaMemcpyAsync(c->hist_host + 2 * nbud, c->hist_dev + 2 * nbud, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->flag_stream));
            ++nbud;
        }
        if (o->save_every > 0 && done % o->save_every == 0 && (st = snapshot()) != LH_OK) return st;
    }
    LH_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    c->timing_valid = true;
    c->last_launches = launches;
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    LH_CUDA(c, cudaStreamSynchronize(c->flag_stream));
    for (int64_t k = 0; k < nbud; ++k) {
        o->budgets_out[2 * k] = c->hist_host[2 * k];
        o->budgets_out[2 * k + 1] = c->U[2] ? c->hist_host[2 * k + 1] : 0.0;
    }
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) return check_finite(c);
    return LH_OK;
}

// ------------------------------------------------------------------------------------------------
// Checkpoint / restart: the device state as it is (padded SoA blocks), bit for bit
// ------------------------------------------------------------------------------------------------
namespace {
struct LhCheckpointHeader {
    char magic[8];            // "LHSOILCK"
    int32_t abi, model, nlayer, field_mask;
    int64_t ncol, ncol_pad;
    double bcv[4];
    int64_t aux_row;
};
int checkpoint_mask(const lh_soil_ctx* c) { int m = 0; for (int f = 0; f < LH_NUM_FIELDS; ++f) if (c->U[f]) m |= 1 << f; return m; }
}  // namespace

int64_t lh_soil_checkpoint_bytes(const lh_soil_ctx* c)
{
    if (!c) return LH_ERR_INVALID_ARG;
    int nf = 0;
    for (int f = 0; f < LH_NUM_FIELDS; ++f) nf += c->U[f] != nullptr;
    return (int64_t)sizeof(LhCheckpointHeader) + (int64_t)nf * (int64_t)field_bytes(c);
}

int32_t lh_soil_checkpoint_save(lh_soil_ctx* c, void* buf, int64_t cap)
{
    if (!c || !buf) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (cap < lh_soil_checkpoint_bytes(c)) return fail(c, LH_ERR_INVALID_ARG, "checkpoint buffer too small (%lld < %lld bytes)", (long long)cap, (long long)lh_soil_checkpoint_bytes(c));
    LH_CUDA(c, cudaSetDevice(c->device));
    LhCheckpointHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "LHSOILCK", 8);
    h.abi = LH_SOIL_ABI_VERSION; h.model = c->model; h.nlayer = c->nlayer; h.field_mask = checkpoint_mask(c);
    h.ncol = c->ncol; h.ncol_pad = c->ncol_pad; h.aux_row = c->aux_row;
    memcpy(h.bcv, c->bcv, sizeof h.bcv);
    memcpy(buf, &h, sizeof h);
    char* dst = (char*)buf + sizeof h;
    for (int f = 0; f < LH_NUM_FIELDS; ++f) {
        if (!c->U[f]) continue;
        LH_CUDA(c, cudaMemcpyAsync(dst, c->U[f], field_bytes(c), cudaMemcpyDeviceToHost, c->stream));
        dst += field_bytes(c);
    }
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    return LH_OK;
}

int32_t lh_soil_checkpoint_load(lh_soil_ctx* c, const void* buf, int64_t bytes)
{
    if (!c || !buf) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    LhCheckpointHeader h;
    if (bytes < (int64_t)sizeof h) return fail(c, LH_ERR_INVALID_ARG, "checkpoint truncated");
    memcpy(&h, buf, sizeof h);
    if (memcmp(h.magic, "LHSOILCK", 8) != 0 || h.abi != LH_SOIL_ABI_VERSION) return fail(c, LH_ERR_INVALID_ARG, "not a checkpoint of this ABI version");
    if (h.model != c->model || h.nlayer != c->nlayer || h.ncol != c->ncol || h.ncol_pad != c->ncol_pad || h.field_mask != checkpoint_mask(c))
        return fail(c, LH_ERR_INVALID_ARG, "checkpoint is of another problem (model %d, %lld x %d) than this ctx (model %d, %lld x %d)",
                    h.model, (long long)h.ncol, h.nlayer, c->model, (long long)c->ncol, c->nlayer);
    if (bytes < lh_soil_checkpoint_bytes(c)) return fail(c, LH_ERR_INVALID_ARG, "checkpoint truncated");
    LH_CUDA(c, cudaSetDevice(c->device));
    const char* src = (const char*)buf + sizeof h;
    for (int f = 0; f < LH_NUM_FIELDS; ++f) {
        if (!c->U[f]) continue;
        LH_CUDA(c, cudaMemcpyAsync(c->U[f], src, field_bytes(c), cudaMemcpyHostToDevice, c->stream));
        src += field_bytes(c);
    }
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    memcpy(c->bcv, h.bcv, sizeof c->bcv);
    c->aux_row = h.aux_row;
    c->budget_fresh = false;
    return detect_ice(c);
}

int32_t lh_soil_alloc_host(int64_t bytes, void** out)
{
    if (!out || bytes < 0) return LH_ERR_INVALID_ARG;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, (size_t)std::max<int64_t>(bytes, 1));
    if (e != cudaSuccess) return fail(nullptr, LH_ERR_CUDA, "cudaMallocHost(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
    return LH_OK;
}

int32_t lh_soil_free_host(void* p)
{
    if (p && cudaFreeHost(p) != cudaSuccess) return LH_ERR_CUDA;
    return LH_OK;
}

// Built-in tables.  Euler; SSPRK22 / SSPRK33 (Shu & Osher 1988); SSPRK43 (the 4-stage, third-order SSP
// method with SSP coefficient 2); CarpenterKennedy2N54 (Carpenter & Kennedy 1994, the (5,4) 2N scheme).
int32_t lh_soil_stepper_named(int32_t method, lh_soil_stepper* out)
{
    if (!out) return LH_ERR_INVALID_ARG;
    lh_soil_stepper s;
    memset(&s, 0, sizeof s);
    auto so = [&](int i, double a, double b, double g, double cc) { s.a[i] = a; s.b[i] = b; s.g[i] = g; s.c[i] = cc; };
    switch (method) {
    case LH_METHOD_EULER:
        s.kind = LH_STEPPER_SHU_OSHER; s.nstages = 1;
        so(0, 0.0, 1.0, 1.0, 0.0);
        break;
    case LH_METHOD_SSPRK22:
        s.kind = LH_STEPPER_SHU_OSHER; s.nstages = 2;
        so(0, 0.0, 1.0, 1.0, 0.0);
        so(1, 0.5, 0.5, 0.5, 1.0);
        break;
    case LH_METHOD_SSPRK33:
        s.kind = LH_STEPPER_SHU_OSHER; s.nstages = 3;
        so(0, 0.0, 1.0, 1.0, 0.0);
        so(1, 0.75, 0.25, 0.25, 1.0);
        so(2, 1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0, 0.5);
        break;
    case LH_METHOD_SSPRK43:
        s.kind = LH_STEPPER_SHU_OSHER; s.nstages = 4;
        so(0, 0.0, 1.0, 0.5, 0.0);
        so(1, 0.0, 1.0, 0.5, 0.5);
        so(2, 2.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, 1.0);
        so(3, 0.0, 1.0, 0.5, 0.5);
        break;
    case LH_METHOD_CK2N54: {
        static const double A[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                    -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
        static const double B[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                    1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                    2277821191437.0 / 14882151754819.0};
        static const double Cc[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                                     2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
        s.kind = LH_STEPPER_2N; s.nstages = 5;
        for (int i = 0; i < 5; ++i) { s.a[i] = A[i]; s.b[i] = B[i]; s.c[i] = Cc[i]; }
        break;
    }
    default:
        return fail(nullptr, LH_ERR_INVALID_ARG, "unknown stepper method %d", method);
    }
    *out = s;
    return LH_OK;
}

int32_t lh_soil_step(lh_soil_ctx* c, const lh_soil_stepper* sp, double t, double dt, int64_t nsteps, const double* bc_table)
{
    (void)t;
    if (!c || !sp) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (nsteps < 0) return fail(c, LH_ERR_INVALID_ARG, "nsteps < 0");
    if (sp->nstages < 1 || sp->nstages > LH_MAX_STAGES) return fail(c, LH_ERR_INVALID_ARG, "stepper.nstages must be 1..%d", LH_MAX_STAGES);
    if (sp->kind != LH_STEPPER_SHU_OSHER && sp->kind != LH_STEPPER_2N) return fail(c, LH_ERR_INVALID_ARG, "unknown stepper kind %d", sp->kind);
    if (sp->kind == LH_STEPPER_2N && sp->a[0] != 0.0) return fail(c, LH_ERR_INVALID_ARG, "a 2N scheme needs a[0] == 0");
    LH_CUDA(c, cudaSetDevice(c->device));
    const int ns = sp->nstages;
    LH_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    for (int64_t s = 0; s < nsteps; ++s) {
        for (int i = 0; i < ns; ++i) {
            if (bc_table) memcpy(c->bcv, bc_table + (s * ns + i) * 4, sizeof c->bcv);
            { int32_t st_ = apply_aux_tables(c); if (st_ != LH_OK) return st_; }
            LhKernelArgs a;
            int stage;
            if (sp->kind == LH_STEPPER_SHU_OSHER) {
                // u_i = a u^n + b u_{i-1} + g dt f(u_{i-1}); first stage reads U, the last writes U, the
                // stage register is V.  A stage with a == 0, b == 1 is a plain forward-Euler update and
                // takes the STAGE 1 kernel (no u^n read) with dt scaled by g.
                const bool first = i == 0, last = i == ns - 1;
                const bool euler = sp->a[i] == 0.0 && sp->b[i] == 1.0;
                stage = euler ? 1 : 4;
                fill_args(c, stage, euler ? sp->g[i] * dt : dt, a);
                a.io.in_th = (!first && has_water(c->model)) ? c->V[0] : c->U[0];
                a.io.in_re = (!first && has_heat(c->model)) ? c->V[2] : c->U[2];
                a.io.out_th = last ? c->U[0] : c->V[0];
                a.io.out_re = last ? c->U[2] : c->V[2];
                a.io.sa = sp->a[i]; a.io.sb = sp->b[i]; a.io.sg = sp->g[i];
            } else {
                // r = a r + dt f(u); u = u + b r: reads and writes U (u) and V (r) in place
                stage = 5;
                fill_args(c, stage, dt, a);
                a.io.in_th = c->U[0]; a.io.in_re = c->U[2];
                a.io.u0_th = c->V[0]; a.io.u0_re = c->V[2];
                a.io.out_th = c->U[0]; a.io.out_re = c->U[2];
                a.io.out2_th = c->V[0]; a.io.out2_re = c->V[2];
                a.io.sa = sp->a[i]; a.io.sb = sp->b[i];
                a.io.first2n = i == 0;
            }
            LH_CUDA(c, launch_chained(c, stage, a));
        }
    }
    LH_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    c->timing_valid = true;
    c->budget_fresh = false;
    c->last_launches = ns * nsteps;
    if (c->cfg.flags & LH_FLAG_CHECK_FINITE) return check_finite(c);
    return LH_OK;
}

// Local budgets into c->budget_dev[0..1].  Right after a step the last-stage launches have already summed, per block,
// the values they wrote (fused epilogue): only the fixed-shape reduction over the blocks is left.  Otherwise (fresh
// upload, generic stepper, heat-only model whose ϑ_l is prescribed, raw device pointer handed out) one pass over the state.
static int32_t local_budgets(lh_soil_ctx* c, double* out_dev)
{
    if (!out_dev) out_dev = c->budget_dev;           // lh_soil_run passes the step's own slot of its history buffer
    const bool fused = c->budget_fresh && !c->external_writes && c->fused_partials && has_water(c->model);
    if (fused) {
        LH_CUDA(c, lh_launch_budgets_from_partials(c->fused_partials, c->shape.nblocks, c->dp.dz, out_dev, c->stream));
        return LH_OK;
    }
    const double* re = c->U[2] ? c->U[2] : c->U[1];   // no energy model: E budget of θ_i slot is meaningless -> report 0
    LH_CUDA(c, lh_launch_budgets(c->U[0], re, c->ncol, c->ncol_pad, c->nlayer, c->dp.dz, c->partials,
                                 c->npartials, out_dev, c->stream));
    return LH_OK;
}

int32_t lh_soil_budgets(lh_soil_ctx* c, double out[2])
{
    if (!c || !out) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    LH_CUDA(c, cudaSetDevice(c->device));
    int32_t st = local_budgets(c);
    if (st != LH_OK) return st;
    LH_CUDA(c, cudaMemcpyAsync(out, c->budget_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!c->U[2]) out[1] = 0.0;
    return LH_OK;
}

int32_t lh_soil_budgets_async(lh_soil_ctx* c, int64_t* ticket_out)
{
    if (!c || !ticket_out) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    LH_CUDA(c, cudaSetDevice(c->device));
    // (each piece on its own: an earlier attempt may have run out of memory half way)
    if (!c->budget_ring_dev) LH_CUDA(c, cudaMalloc(&c->budget_ring_dev, lh_soil_ctx::BUDGET_SLOTS * 2 * sizeof(double)));
    if (!c->budget_ring_host) LH_CUDA(c, cudaMallocHost(&c->budget_ring_host, lh_soil_ctx::BUDGET_SLOTS * 2 * sizeof(double)));
    for (auto& e : c->budget_ev) if (!e) LH_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    const int64_t ticket = c->budget_next_ticket;
    const int slot = (int)(ticket % lh_soil_ctx::BUDGET_SLOTS);
    if (c->budget_ticket[slot] != 0)
        return fail(c, LH_ERR_STATE, "lh_soil_budgets_async: %d results outstanding, collect one with lh_soil_budgets_wait first", lh_soil_ctx::BUDGET_SLOTS);
    int32_t st = local_budgets(c);
    if (st != LH_OK) return st;
    // budget_dev is overwritten by the next budget call: park this result in its own slot before the copy to the host
    LH_CUDA(c, cudaMemcpyAsync(c->budget_ring_dev + 2 * slot, c->budget_dev, 2 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    LH_CUDA(c, cudaMemcpyAsync(c->budget_ring_host + 2 * slot, c->budget_ring_dev + 2 * slot, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LH_CUDA(c, cudaEventRecord(c->budget_ev[slot], c->stream));
    c->budget_ticket[slot] = ticket;
    ++c->budget_next_ticket;
    *ticket_out = ticket;
    return LH_OK;
}

int32_t lh_soil_budgets_wait(lh_soil_ctx* c, int64_t ticket, double out[2])
{
    if (!c || !out) return LH_ERR_INVALID_ARG;
    const int slot = (int)(ticket % lh_soil_ctx::BUDGET_SLOTS);
    if (ticket <= 0 || c->budget_ticket[slot] != ticket) return fail(c, LH_ERR_STATE, "lh_soil_budgets_wait: unknown or already collected ticket %lld", (long long)ticket);
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaEventSynchronize(c->budget_ev[slot]));
    out[0] = c->budget_ring_host[2 * slot];
    out[1] = c->U[2] ? c->budget_ring_host[2 * slot + 1] : 0.0;
    c->budget_ticket[slot] = 0;
    return LH_OK;
}

int32_t lh_soil_diagnostic(lh_soil_ctx* c, int32_t which, double* host, int64_t cs, int64_t ls)
{
    if (!c || !host) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (which < 0 || which >= LH_NUM_DIAGS) return fail(c, LH_ERR_INVALID_ARG, "bad diagnostic id %d", which);
    LH_CUDA(c, cudaSetDevice(c->device));
    if (!c->diag_dev) LH_CUDA(c, cudaMalloc(&c->diag_dev, field_bytes(c)));      // its own scratch: tendencies stay untouched
    LH_CUDA(c, lh_launch_diagnostic(c->model, which, c->dp, c->pow_tab_dev, c->U[0], c->U[1], c->U[2], c->U[3], c->diag_dev,
                                    (int64_t)c->ncol_pad * c->nlayer, c->colp_dev, c->ncol_pad, (c->col_heat || c->cellp_dev) ? 1 : 0, c->cellp_dev, c->stream));
    return download_field(c, c->diag_dev, host, cs, ls);
}

int32_t lh_soil_eval_math(lh_soil_ctx* c, int32_t fn, const double* x, double* y, int64_t n)
{
    if (!c || !x || !y || n < 0) return LH_ERR_INVALID_ARG;
    if (fn < LH_MATH_LOG2 || fn > LH_MATH_RSQRT_SEED) return fail(c, LH_ERR_INVALID_ARG, "bad math function id %d", fn);
    if (n == 0) return LH_OK;
    LH_CUDA(c, cudaSetDevice(c->device));
    const int64_t nin = fn == LH_MATH_DIV ? 2 * n : n;
    double *dx = nullptr, *dy = nullptr;
    LH_CUDA(c, cudaMalloc(&dx, nin * sizeof(double)));
    cudaError_t e = cudaMalloc(&dy, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, nin * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = lh_launch_eval_math(c->dp, fn, dx, dy, n, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dx);
    if (dy) cudaFree(dy);
    if (e != cudaSuccess) return fail(c, LH_ERR_CUDA, "lh_soil_eval_math: %s", cudaGetErrorString(e));
    return LH_OK;
}

int32_t lh_soil_sync(lh_soil_ctx* c)
{
    if (!c) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    return LH_OK;
}

int32_t lh_soil_last_step_timing(lh_soil_ctx* c, double* ms_out, int64_t* launches_out)
{
    if (!c) return LH_ERR_INVALID_ARG;
    if (!c->timing_valid) return fail(c, LH_ERR_STATE, "no lh_soil_step_ssprk33 call to time yet");
    LH_CUDA(c, cudaSetDevice(c->device));
    LH_CUDA(c, cudaEventSynchronize(c->ev_stop));
    float ms = 0.f;
    LH_CUDA(c, cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop));
    if (ms_out) *ms_out = (double)ms;
    if (launches_out) *launches_out = c->last_launches;
    return LH_OK;
}

int32_t lh_soil_kernel_info(lh_soil_ctx* c, char* buf, int64_t cap)
{
    if (!c || !buf || cap < 1) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    const int f = c->kernel_flags;
    const bool persistent = use_persistent(c);
    snprintf(buf, (size_t)cap,
             "%s<MODEL=%d,%sFLAGS=%d:%s%s%s%s> block=(32,W=%d,G=%d) layers/thread=%d blocks=%lld smem=%zu waves=%.2f "
             "launch=%s chain=%s bytes/cell-step(on wire)=%d",
             persistent ? "lh_soil_ssprk33_persistent_kernel" : "lh_soil_stage_kernel", c->model, persistent ? "" : "STAGE=1|2|3,", f,
             (f & LH_FLAG_ICE) ? "ICE" : "!ICE", (f & LH_FLAG_GEN) ? "+GEN" : "", (f & LH_FLAG_VG2) ? "+VG2" : "", (f & LH_FLAG_CELLP) ? "+HET+CELLP" : (f & LH_FLAG_HETH) ? "+HET+HETH" : (f & LH_FLAG_HET) ? "+HET" : "",
             c->shape.W, c->shape.G, c->shape.Lc, (long long)c->shape.nblocks, c->shape.smem_bytes, c->shape.waves,
             persistent ? "persistent(1 per call)" : "per-stage(3 per step)",
             (!persistent && c->chain_dev && !(c->cfg.flags & LH_FLAG_NO_CHAIN)) ? "block-to-block" : "whole-grid",
             lh_bytes_on_wire(c));
    return LH_OK;
}

int32_t lh_soil_device_ptr(lh_soil_ctx* c, int32_t field, void** dptr, int64_t* ncol_padded)
{
    if (!c || !dptr) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (!field_ok(field) || !c->U[field]) return fail(c, LH_ERR_INVALID_ARG, "field %d does not exist", field);
    LH_CUDA(c, cudaSetDevice(c->device));
    *dptr = c->U[field];
    c->external_writes = true;
    if (ncol_padded) *ncol_padded = c->ncol_pad;
    if (field == LH_FIELD_THETA_I) {   // the caller may write ice through the raw pointer: assume it does
        c->has_ice = true;
        c->theta_i_ptr_out = true;     // sticky: a later all-zero θ_i upload does not switch back to the !ICE kernels
        update_kernel_flags(c);
    }
    return LH_OK;
}

int32_t lh_soil_comm_unique_id(uint8_t id_out[128])
{
    if (!id_out) return LH_ERR_INVALID_ARG;
    NcclApi& n = nccl();
    if (!n.ok) return fail(nullptr, LH_ERR_NCCL, "libnccl.so.2 could not be loaded");
    NcclUniqueId_ id;
    int r = n.GetUniqueId(&id);
    if (r != 0) return fail(nullptr, LH_ERR_NCCL, "ncclGetUniqueId failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    memcpy(id_out, id.internal, 128);
    return LH_OK;
}

int32_t lh_soil_comm_init(lh_soil_ctx* c, int32_t nranks, int32_t rank, const uint8_t id[128])
{
    if (!c || !id) return LH_ERR_INVALID_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, LH_ERR_INVALID_ARG, "bad rank %d of %d", rank, nranks);
    NcclApi& n = nccl();
    if (!n.ok) return fail(c, LH_ERR_NCCL, "libnccl.so.2 could not be loaded");
    LH_CUDA(c, cudaSetDevice(c->device));
    if (c->comm) { n.CommDestroy(c->comm); c->comm = nullptr; }
    NcclUniqueId_ uid;
    memcpy(uid.internal, id, 128);
    int r = n.CommInitRank(&c->comm, nranks, uid, rank);
    if (r != 0) return fail(c, LH_ERR_NCCL, "ncclCommInitRank failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    c->nranks = nranks;
    c->rank = rank;
    return LH_OK;
}

int32_t lh_soil_budgets_allreduce(lh_soil_ctx* c, double out[2])
{
    if (!c || !out) return LH_ERR_INVALID_ARG;
    LH_RESOLVE(c);
    if (!c->comm) return fail(c, LH_ERR_STATE, "lh_soil_comm_init has not been called");
    NcclApi& n = nccl();
    LH_CUDA(c, cudaSetDevice(c->device));
    int32_t st = local_budgets(c);
    if (st != LH_OK) return st;
    int r = n.AllReduce(c->budget_dev, c->budget_dev + 2, 2, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
    if (r != 0) return fail(c, LH_ERR_NCCL, "ncclAllReduce failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    LH_CUDA(c, cudaMemcpyAsync(out, c->budget_dev + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LH_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!c->U[2]) out[1] = 0.0;
    return LH_OK;
}

}  // extern "C"
