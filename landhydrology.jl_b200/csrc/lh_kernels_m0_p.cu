// lh_kernels_m0_p.cu — persistent SSPRK33 kernel variants of MODEL = 0 (Richards).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_persistent_m0(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<0>(flags, args, shape, stream);
}
