/*
 * nsdg_fast_launch.cuh -- host-side launchers of the fast subcycle kernels.
 *
 * Each fast strip kernel (uniform / parametric x mEVP / BBM) lives in ITS OWN translation unit
 * (nsdg_kernels_umevp.cu, _ubbm.cu, _pmevp.cu, _pbbm.cu).  These kernels are 35 - 65 KB loop bodies at 168 - 255 registers,
 * and their register allocation turned out to depend on what else is in the module: in one translation unit, unrelated
 * edits (a new transport kernel, another `-split-compile` partition) moved the headline kernel from 168 registers without
 * spills to 168 with 48 bytes of spills and the parametric mEVP kernel from 190 to 222 registers.  Compiled alone, a
 * kernel's code depends on its own headers only.  nsdg_cuda.cu (handle, C ABI, everything else) calls them through here.
 */
#pragma once
#include "nsdg_momentum_param.cuh"

namespace nsdg {

//! raises the dynamic shared-memory limit of the strip kernels on the CURRENT device (call once per handle, after cudaSetDevice)
void prepareKernelsUMEVP();
void prepareKernelsUBBM();
void prepareKernelsPMEVP();
void prepareKernelsPBBM();
void prepareKernelsUMEVP1();

//! nStrips = nsx * nsy warp strips; nLine = deferred-line nodes
void launchStripUMEVP(const UniformArgs& a, unsigned nStrips, cudaStream_t s);
void launchLinesUMEVP(const UniformArgs& a, size_t nLine, cudaStream_t s);
void launchStripPMEVP(const UniformArgs& a, bool spherical, unsigned nStrips, cudaStream_t s);
void launchStripUMEVP1(const UniformArgs& a, unsigned nStrips, cudaStream_t s); //!< DG1 / CG1 build
void launchLinesUMEVP1(const UniformArgs& a, cudaStream_t s);
void launchStripUBBM(const UniformBBMArgs& a, unsigned nStrips, cudaStream_t s);
void launchLinesUBBM(const UniformBBMArgs& a, size_t nLine, cudaStream_t s);
void launchStripPBBM(const UniformBBMArgs& a, bool spherical, unsigned nStrips, cudaStream_t s);

} // namespace nsdg
