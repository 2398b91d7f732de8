// translation unit of the parametric-mesh (factored operators) mEVP subcycle kernels (see nsdg_fast_launch.cuh)
#include "nsdg_fast_launch.cuh"

namespace nsdg {

void prepareKernelsPMEVP()
{
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_pmevp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pmevpSmemBytes(false))));
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_pmevp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pmevpSmemBytes(true))));
}
void launchStripPMEVP(const UniformArgs& a, bool spherical, unsigned nStrips, cudaStream_t s)
{
    const unsigned nw = pmevpWarps(spherical), nb = (nStrips + nw - 1) / nw;
    if (spherical)
        subcycle_strip_pmevp<true><<<nb, 32 * nw, pmevpSmemBytes(true), s>>>(a);
    else
        subcycle_strip_pmevp<false><<<nb, 32 * nw, pmevpSmemBytes(false), s>>>(a);
}

} // namespace nsdg
