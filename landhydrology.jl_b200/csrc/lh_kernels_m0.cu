// lh_kernels_m0.cu — stage-kernel variants of the richards model (MODEL = 0).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m0(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape,
                                cudaStream_t stream)
{
    return launch_model<0>(stage, flags, args, shape, stream);
}

cudaError_t lh_launch_persistent_m0(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<0>(flags, args, shape, stream);
}
