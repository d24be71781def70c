// lh_kernels_m2_a.cu — stage-kernel variants of MODEL = 2 (coupled): tendency and SSPRK33 stage 1.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m2_g0(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<2, 0>(stage, flags, args, shape, stream);
}
