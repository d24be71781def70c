// lh_kernels_m2_p.cu — persistent SSPRK33 kernel variants of MODEL = 2 (coupled).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_persistent_m2(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<2>(flags, args, shape, stream);
}
