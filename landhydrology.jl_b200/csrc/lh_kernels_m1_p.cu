// lh_kernels_m1_p.cu — persistent SSPRK33 kernel variants of MODEL = 1 (heat-only).
#include "lh_stage_kernel.cuh"

cudaError_t lh_launch_persistent_m1(int flags, const LhKernelArgs& args, const LhLaunchShape& shape, cudaStream_t stream)
{
    return launch_persistent_model<1>(flags, args, shape, stream);
}
