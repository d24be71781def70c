// lh_math.cuh — fp64 elementary functions used by the soil closures.
//
// Accuracy contract: <= ~2 ulp on the ranges the closures use; the parity gate downstream is
// 1e-12 in a cancellation-aware norm (tests/test_gpu_parity.py).
#pragma once

#include <cuda_runtime.h>

__device__ __forceinline__ double lh_log(double x) { return log(x); }
__device__ __forceinline__ double lh_exp(double x) { return exp(x); }
__device__ __forceinline__ double lh_expm1(double x) { return expm1(x); }
