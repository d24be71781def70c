// tests/support/hostemu/cuda_runtime.h — TEST INFRASTRUCTURE ONLY.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the PRODUCT's own sources (landhydrology.jl_b200/csrc/*.cu, *.cuh —
// the C ABI, the launch logic and every kernel template) into tests/support/hostemu/_build/liblh_soil_hostemu.so, so that the
// CPU-only test run executes the very code that runs on the GPU: each CUDA thread of a block is a fiber, __syncthreads /
// __syncwarp / __shfl_* are real rendezvous between fibers, cp.async groups are real queues (completed eagerly at issue or as
// late as wait_group allows), "device" allocations are guard-paged and poisoned.  Blocks of a grid run independently on a few
// host threads.  Streams and events: synchronous by default (an operation completes inside the call that enqueues it), or —
// lh_emu_set_async — deferred FIFO queues executed only when a synchronisation needs them (lazily, or in a seeded random
// interleaving), which exercises the host layer's stream / event dependencies.  Not modelled: timing, sm_100a code generation.
//
// Nothing in the product loads, links or falls back to this: the package opens csrc/liblh_soil.so (nvcc, sm_100a) and fails
// loudly without it; only tests/ builds and opens the emulated library, by explicit path (tests/hostemu.py).
#pragma once

#ifndef LH_HOSTEMU
#error "tests/support/hostemu/cuda_runtime.h is only for the host-emulation test build (-DLH_HOSTEMU)"
#endif

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <tuple>
#include <type_traits>

// ---------------------------------------------------------------- language surface
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__                              /* empty: libstdc++ spells __attribute__((__noinline__)), which must survive */
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local            // a host thread runs one block at a time: block-shared == thread-local static

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---------------------------------------------------------------- the emulator (hostemu.cpp)
struct LhEmuThread {               // the CUDA thread the calling fiber is
    uint3 tid, bid;
    dim3 bdim, gdim;
    int lane, warp;                // lane within the warp, warp within the block (linear thread id / 32)
    char* smem;                    // the block's dynamic shared memory
    size_t smem_bytes;
};
extern "C" {
LhEmuThread* lh_emu_self(void);
void lh_emu_syncthreads(void);
int lh_emu_syncthreads_or(int pred);
void lh_emu_syncwarp(void);
uint64_t lh_emu_shfl(uint64_t bits, int src_lane);          // value of `bits` held by lane src_lane of this warp
void lh_emu_yield(void);
void lh_emu_cp_async16(uint32_t dst_smem_off, const void* src);
void lh_emu_cp_commit(void);
void lh_emu_cp_wait(int keep_newest);
void lh_emu_fault(const char* what);
// controls (tests): schedule 0 = round robin, 1 = reverse, 2 = seeded random; cp.async completion 0 = at issue, 1 = as late as
// wait_group allows.  Return the previous value.
int lh_emu_set_schedule(int policy, uint64_t seed);
int lh_emu_set_cp_async_lazy(int lazy);
int lh_emu_set_async(int mode, uint64_t seed);   // 0 synchronous, 1 deferred (lazy), 2 deferred (seeded random interleaving)
int lh_emu_set_device_count(int n);
int lh_emu_set_sm_count(int n);
uint64_t lh_emu_launch_count(void);
int64_t lh_emu_live_handles(void);            // streams + events created and not yet destroyed
void lh_emu_fail_allocation_in(int64_t n);     // n more cudaMalloc / cudaMallocHost calls succeed, the next fails once (-1: off)
uint64_t lh_emu_live_allocations(void);      // cudaMalloc / cudaMallocHost blocks not yet freed
}
struct LhEmuStream;
void lh_emu_launch(LhEmuStream* stream, dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> thread_body);

// threadIdx.x etc.: objects whose members convert to the calling fiber's coordinate (not macros: cudaLaunchConfig_t has
// members called gridDim / blockDim)
template <int WHICH, int COMP> struct LhEmuCoord {
    operator unsigned() const
    {
        const LhEmuThread* t = lh_emu_self();
        if (WHICH == 0) return COMP == 0 ? t->tid.x : COMP == 1 ? t->tid.y : t->tid.z;
        if (WHICH == 1) return COMP == 0 ? t->bid.x : COMP == 1 ? t->bid.y : t->bid.z;
        if (WHICH == 2) return COMP == 0 ? t->bdim.x : COMP == 1 ? t->bdim.y : t->bdim.z;
        return COMP == 0 ? t->gdim.x : COMP == 1 ? t->gdim.y : t->gdim.z;
    }
};
template <int WHICH> struct LhEmuCoords { LhEmuCoord<WHICH, 0> x; LhEmuCoord<WHICH, 1> y; LhEmuCoord<WHICH, 2> z; };
static const LhEmuCoords<0> threadIdx = {};
static const LhEmuCoords<1> blockIdx = {};
static const LhEmuCoords<2> blockDim = {};
static const LhEmuCoords<3> gridDim = {};

// ---------------------------------------------------------------- device intrinsics
static inline void __syncthreads() { lh_emu_syncthreads(); }
static inline int __syncthreads_or(int p) { return lh_emu_syncthreads_or(p); }
static inline void __syncwarp(unsigned = 0xffffffffu) { lh_emu_syncwarp(); }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { lh_emu_yield(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    b = lh_emu_shfl(b, src & 31);
    T r;
    memcpy(&r, &b, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned delta)
{
    const int lane = lh_emu_self()->lane;
    const int src = lane + (int)delta;
    return __shfl_sync(m, v, src < 32 ? src : lane);          // out of range: the lane's own value, as the hardware does
}
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }   // blocks run on several host threads
static inline size_t __cvta_generic_to_shared(const void* p)
{
    LhEmuThread* t = lh_emu_self();
    const ptrdiff_t off = (const char*)p - t->smem;
    if (off < 0 || (size_t)off > t->smem_bytes) lh_emu_fault("__cvta_generic_to_shared: pointer outside the block's dynamic shared memory");
    return (size_t)off;
}
using std::max;
using std::min;

// ---------------------------------------------------------------- csrc/lh_ptx.cuh, emulated (same names, same meaning)
static inline void lh_cp16(uint32_t dst, const void* src, bool pred) { if (pred) lh_emu_cp_async16(dst, src); }
template <int OFF> static inline double lh_lds(uint32_t addr)
{
    LhEmuThread* t = lh_emu_self();
    if ((size_t)addr + OFF + 8 > t->smem_bytes) lh_emu_fault("ld.shared: address outside the block's dynamic shared memory");
    double v;
    memcpy(&v, t->smem + addr + OFF, 8);
    return v;
}
static inline void lh_cp_commit() { lh_emu_cp_commit(); }
template <int N> static inline void lh_cp_wait() { lh_emu_cp_wait(N); }
static inline void lh_pdl_launch_dependents() {}
static inline void lh_pdl_wait() {}               // launches are synchronous: the previous grid has completed
static inline int32_t lh_ld_acquire(const int32_t* p) { return *(const volatile int32_t*)p; }
static inline void lh_st_release(int32_t* p, int32_t v) { *(volatile int32_t*)p = v; }
#define LH_UNPAREN(...) __VA_ARGS__
// arguments are captured BY VALUE, as a CUDA launch copies them: an asynchronous-mode launch runs after the caller has returned
#define LH_LAUNCH(kernel, grid, block, smem, stream, ...) \
    lh_emu_launch(stream, dim3(grid), dim3(block), (size_t)(smem), [=]() { LH_UNPAREN kernel(__VA_ARGS__); })
#define LH_DYN_SMEM(type, name) type* const name = reinterpret_cast<type*>(lh_emu_self()->smem)

// ---------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInvalidConfiguration = 9,
       cudaErrorNoDevice = 100, cudaErrorInvalidDevice = 101 };
struct LhEmuStream;
struct LhEmuEvent;
typedef LhEmuStream* cudaStream_t;
typedef LhEmuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; size_t totalGlobalMem; };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; unsigned numAttrs; };

extern "C" {
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t lh_emu_malloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t lh_emu_malloc_host(void** p, size_t bytes);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
}
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return lh_emu_malloc((void**)p, bytes); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return lh_emu_malloc_host((void**)p, bytes); }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class... Params, class... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(Params...), Args&&... args)
{
    auto bound = std::make_tuple(std::decay_t<Args>(args)...);          // by value, as a CUDA launch copies its arguments
    lh_emu_launch(cfg->stream, cfg->gridDim, cfg->blockDim, cfg->dynamicSmemBytes, [kernel, bound]() { std::apply(kernel, bound); });
    return cudaSuccess;
}
