// translation unit of the uniform-mesh BBM subcycle kernels (see nsdg_fast_launch.cuh)
#include "nsdg_fast_launch.cuh"

namespace nsdg {

void prepareKernelsUBBM()
{
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_ubbm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUbbmSmemBytes)));
}
void launchStripUBBM(const UniformBBMArgs& a, unsigned nStrips, cudaStream_t s)
{
    const unsigned nb = (nStrips + kUbbmWarps - 1) / kUbbmWarps;
    subcycle_strip_ubbm<0><<<nb, 32 * kUbbmWarps, kUbbmSmemBytes, s>>>(a);
}
void launchLinesUBBM(const UniformBBMArgs& a, size_t nLine, cudaStream_t s)
{
    (void)nLine;
    subcycle_lines_ubbm<0><<<linesGrid(a.g, a.nsx, a.nsy), 128, 0, s>>>(a);
}

} // namespace nsdg
