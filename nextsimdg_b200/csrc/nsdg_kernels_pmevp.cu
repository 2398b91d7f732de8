// translation unit of the parametric-mesh (factored operators) mEVP subcycle kernels (see nsdg_fast_launch.cuh)
// branch-free roots in this unit's law and node update (0.725 -> 0.699 ms distorted, 0.91 -> 0.85 ms spherical at 2048^2); the
// uniform mEVP kernel, which sits on the HBM roof, measured no gain and keeps the library calls
#ifndef NSDG_MEVP_SQRT
#define NSDG_MEVP_SQRT 3
#endif
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
