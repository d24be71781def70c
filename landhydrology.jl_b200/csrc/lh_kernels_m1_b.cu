// lh_kernels_m1_b.cu — stage-kernel variants of MODEL = 1 (heat-only): SSPRK33 stages 2 and 3.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m1_g1(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<1, 1>(stage, flags, args, shape, stream);
}
