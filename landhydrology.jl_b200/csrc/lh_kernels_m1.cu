// lh_kernels_m1.cu — stage-kernel variants of the heat model (MODEL = 1).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_stage_m1(int stage, int flags, const LhKernelArgs& args, const LhLaunchShape& shape,
                                cudaStream_t stream)
{
    return launch_model<1>(stage, flags, args, shape, stream);
}

cudaError_t lh_launch_persistent_m1(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<1>(flags, args, shape, stream);
}
