// translation unit of the uniform-mesh mEVP subcycle kernels (see nsdg_fast_launch.cuh)
#include "nsdg_fast_launch.cuh"

namespace nsdg {

void prepareKernelsUMEVP()
{
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_umevp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUmevpSmemBytes)));
}
void launchStripUMEVP(const UniformArgs& a, unsigned nStrips, cudaStream_t s)
{
    const unsigned nb = (nStrips + kUmevpWarps - 1) / kUmevpWarps;
    subcycle_strip_umevp<0><<<nb, 32 * kUmevpWarps, kUmevpSmemBytes, s>>>(a);
}
void launchLinesUMEVP(const UniformArgs& a, size_t nLine, cudaStream_t s)
{
    (void)nLine;
    subcycle_lines_umevp<2><<<linesGrid(a.g, a.nsx, a.nsy), 128, 0, s>>>(a);
}

} // namespace nsdg
