/*
 * nsdg_state.cuh -- device-resident state of one dynamics handle and its memory layout.
 *
 * Layout (chosen for one-thread-per-element / one-warp-per-32-elements kernels):
 *  - element fields are structure-of-arrays "planes": component c of element e lives at
 *    f[c * Npad + e], e = ix + nxs*iy with the row stride nxs = nx rounded up to 32, so a warp
 *    reading 32 consecutive elements of one plane touches 256 contiguous, aligned bytes.  (The reference's DGVector is element-major AoS,
 *    dynamics/src/include/dgVector.hpp:89-91; the C ABI converts at the boundary.)
 *  - per-element operator matrices (general / parametric meshes) are planes too: matrix
 *    entry k of element e at op[k * Npad + e].
 *  - CG fields are row-major node arrays with the reference's node order
 *    (dynamics/src/include/cgVector.hpp:21-31) but a padded row stride `cgs`
 *    (multiple of 16 doubles = 128 B) so that every node row starts on a cache line.
 */
#pragma once
#include "../../include/nsdg.h"
#include "nsdg_basis.cuh"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace nsdg {

#define NSDG_CUDA_CHECK(call)                                                                                \
    do {                                                                                                     \
        cudaError_t err__ = (call);                                                                          \
        if (err__ != cudaSuccess)                                                                            \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " + __FILE__ \
                + ":" + std::to_string(__LINE__) + " in " #call);                                            \
    } while (0)

//! Physical parameters: DynamicsParameters.hpp:32-50, VPParameters.hpp:22-23, MEBParameters.hpp:40-65
struct PhysParams {
    double rho_ice = 900.0, rho_atm = 1.3, rho_ocean = 1026.0;
    double C_atm = 1.2e-3, C_ocean = 5.5e-3;
    double F_atm = 1.2e-3 * 1.3, F_ocean = 5.5e-3 * 1026.0;
    double fc = 1.45842e-4;
    double gravity = 9.81;
    double cosOceanAngle = 1.0, sinOceanAngle = 0.0; // turning angle forced to 0, DynamicsParameters.hpp:46-47
    double Pstar = 27500.0, DeltaMin = 2.e-9;
    double alpha = 1500.0, beta = 1500.0;
    double compaction_param = -20., nu0 = 1. / 3., young = 5.96e8, P0 = 10.e3;
    double undamaged_time_relaxation_sigma = 1e7;
    int exponent_relaxation_sigma = 5;
    double exponent_compression_factor = 1.5;
    double tan_phi = 0.7, compr_strength = 1e10, C_lab = 2.0e6;
};

//! Grid geometry shared by all kernels (passed by value)
struct GridDims {
    int nx, ny; //!< elements
    int nxs; //!< element row stride (nx rounded up to 32: every element row starts 256-B aligned)
    int N, Npad; //!< nx*ny (dense count) and plane pitch (>= nxs*ny); element (ix,iy) lives at ix + nxs*iy
    int CG; //!< CG degree
    int cgnx, cgny; //!< CG*nx+1, CG*ny+1
    int cgs; //!< padded CG row stride
    int spherical;
    int bnd; //!< bit s set: side s (0 bottom,1 right,2 top,3 left) is an edge of the GLOBAL domain (15 = single domain)
};

//! simple owning device buffer
//! wait for the legacy default stream (synchronous cudaMemcpy / cudaMemset work); the handles run on non-blocking streams
inline void legacySync() { NSDG_CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy)); }

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count, bool zero = true)
    {
        release();
        n = count;
        if (count == 0)
            return;
        NSDG_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
        if (zero) {
            // cudaMemset of device memory is asynchronous and runs on the legacy default stream, which the handles'
            // non-blocking streams do not wait for: without the synchronisation a kernel could read the buffer before it
            // is zeroed (or be overwritten by the late memset) whenever cudaMalloc hands back used memory
            NSDG_CUDA_CHECK(cudaMemsetAsync(p, 0, count * sizeof(T), cudaStreamLegacy));
            legacySync();
        }
    }
    operator T*() const { return p; }
};

//! Operator matrices of one element for the momentum equation (ParametricMap.hpp:72-113),
//! also the layout of the uniform-mesh copy in __constant__ memory.  Sized for CG2/DG8/DG6.
struct MomentumOps {
    double Gx[8 * 9], Gy[8 * 9], GM[8 * 9]; //!< iMgradX, iMgradY, iMM      (DGs x ND, row-major)
    double B[8 * 9]; //!< iMJwPSI                     (DGs x Q)
    double Bd[6 * 9]; //!< iMJwPSI_dam                 (DGadv x Q)
    double D1[9 * 8], D2[9 * 8], DM[9 * 8]; //!< divS1, divS2, divM          (ND x DGs)
};

//! Transport operators of one element (ParametricMap.hpp:22-27), uniform-mesh copy
template <int DG> struct TransportOps {
    double AdvX[DG * (gp1d(DG) * gp1d(DG))], AdvY[DG * (gp1d(DG) * gp1d(DG))], iMass[DG * DG];
};

} // namespace nsdg
