// lh_kernels_m2_b.cu — stage-kernel variants of MODEL = 2 (coupled): SSPRK33 stages 2 and 3.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m2_g1(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<2, 1>(stage, flags, args, shape, stream);
}
