// lh_kernels_m0_a.cu — stage-kernel variants of MODEL = 0 (Richards): tendency and SSPRK33 stage 1.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m0_g0(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<0, 0>(stage, flags, args, shape, stream);
}
