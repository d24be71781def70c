// lh_kernels_m1_a.cu — stage-kernel variants of MODEL = 1 (heat-only): tendency and SSPRK33 stage 1.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m1_g0(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<1, 0>(stage, flags, args, shape, stream);
}
