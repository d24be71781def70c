// lh_ptx.cuh — the inline PTX of the stage kernels and the kernel-launch spelling, in one place.
//
// Everything here is sm_100a device code.  The CPU-only test build (tests/support/hostemu: the product's own sources compiled
// by g++ against an emulated execution model, so that the kernels and the host logic are exercised without a GPU) supplies
// functions of the same names and meaning instead — test infrastructure; no product path uses it.
#pragma once

#ifndef LH_HOSTEMU

#include <cuda_runtime.h>
#include <stdint.h>

// 16-byte cp.async that bypasses L1 (.cg): global -> shared without allocating an L1 line while in flight.
__device__ __forceinline__ void lh_cp16(uint32_t dst, const void* src, bool pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                 ::"r"(dst), "l"(src), "r"((int)pred) : "memory");
}
template <int OFF>
__device__ __forceinline__ double lh_lds(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
    return v;
}
__device__ __forceinline__ void lh_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lh_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// programmatic dependent launch
__device__ __forceinline__ void lh_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void lh_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// block-to-block chaining flags
__device__ __forceinline__ int32_t lh_ld_acquire(const int32_t* p)
{
    int32_t v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lh_st_release(int32_t* p, int32_t v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// LH_LAUNCH((kernel<...>), grid, block, dynamic shared bytes, stream, arguments...): the kernel name goes in parentheses so
// that template commas survive the macro.
#define LH_UNPAREN(...) __VA_ARGS__
#define LH_LAUNCH(kernel, grid, block, smem, stream, ...) LH_UNPAREN kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
// the block's dynamic shared memory as `type name[]`
#define LH_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]

#endif  // !LH_HOSTEMU
