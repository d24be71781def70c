// lh_kernels_m2.cu — stage-kernel variants of the coupled model (MODEL = 2).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m2(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape,
                                cudaStream_t stream)
{
    return launch_model<2>(stage, flags, args, shape, stream);
}

cudaError_t lh_launch_persistent_m2(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<2>(flags, args, shape, stream);
}
