// lh_kernels_m0_b.cu — stage-kernel variants of MODEL = 0 (Richards): SSPRK33 stages 2 and 3.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m0_g1(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<0, 1>(stage, flags, args, shape, stream);
}
