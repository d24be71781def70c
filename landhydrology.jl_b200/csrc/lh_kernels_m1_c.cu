// lh_kernels_m1_c.cu — stage-kernel variants of MODEL = 1 (heat-only): the generic Shu-Osher and 2N stages.
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m1_g2(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_model<1, 2>(stage, flags, args, shape, stream);
}
