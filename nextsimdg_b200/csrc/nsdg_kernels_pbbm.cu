// translation unit of the parametric-mesh (factored operators) BBM subcycle kernels (see nsdg_fast_launch.cuh)
#include "nsdg_fast_launch.cuh"

namespace nsdg {

void prepareKernelsPBBM()
{
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_pbbm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pbbmSmemBytes<false>())));
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_pbbm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pbbmSmemBytes<true>())));
}
void launchStripPBBM(const UniformBBMArgs& a, bool spherical, unsigned nStrips, cudaStream_t s)
{
    const unsigned nb = (nStrips + kPbbmWarps - 1) / kPbbmWarps;
    if (spherical)
        subcycle_strip_pbbm<true><<<nb, 32 * kPbbmWarps, pbbmSmemBytes<true>(), s>>>(a);
    else
        subcycle_strip_pbbm<false><<<nb, 32 * kPbbmWarps, pbbmSmemBytes<false>(), s>>>(a);
}

} // namespace nsdg
