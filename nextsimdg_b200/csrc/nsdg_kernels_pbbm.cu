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
    if (spherical)
        subcycle_strip_pbbm<true><<<(nStrips + kPbbmWarps<true> - 1) / kPbbmWarps<true>, 32 * kPbbmWarps<true>, pbbmSmemBytes<true>(), s>>>(a);
    else
        subcycle_strip_pbbm<false><<<(nStrips + kPbbmWarps<false> - 1) / kPbbmWarps<false>, 32 * kPbbmWarps<false>, pbbmSmemBytes<false>(), s>>>(a);
}

} // namespace nsdg
