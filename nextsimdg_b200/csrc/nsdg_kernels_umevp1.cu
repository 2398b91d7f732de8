// translation unit of the uniform-mesh mEVP subcycle kernel of the DG1 / CG1 build (see nsdg_fast_launch.cuh)
#include "nsdg_fast_launch.cuh"
#include "nsdg_momentum_uniform_cg1.cuh"

namespace nsdg {

void prepareKernelsUMEVP1()
{
    NSDG_CUDA_CHECK(cudaFuncSetAttribute(subcycle_strip_umevp1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kUmevp1SmemBytes)));
}
void launchStripUMEVP1(const UniformArgs& a, unsigned nStrips, cudaStream_t s)
{
    const unsigned nb = (nStrips + kUmevp1Warps - 1) / kUmevp1Warps;
    subcycle_strip_umevp1<0><<<nb, 32 * kUmevp1Warps, kUmevp1SmemBytes, s>>>(a);
}

void launchLinesUMEVP1(const UniformArgs& a, cudaStream_t s)
{
    subcycle_lines_umevp<1><<<linesGrid(a.g, a.nsx, a.nsy), 128, 0, s>>>(a);
}

} // namespace nsdg
