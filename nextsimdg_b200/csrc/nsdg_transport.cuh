/*
 * nsdg_transport.cuh -- DG advection on the parametric mesh and the DG<->CG transfers, as
 * one-thread-per-element (or per node / per edge) gather kernels on the plane layout.
 *
 *   cg2dg_kernel          Interpolations::CG2DG             dynamics/src/Interpolations.cpp:74-122
 *   dg2cg_kernel          Interpolations::DG2CG             dynamics/src/Interpolations.cpp:129-204
 *   normalvel_kernel      DGTransport::reinitnormalvelocity dynamics/src/DGTransport.cpp:158-252
 *   transport_stage_kernel DGTransport::DGTransportOperator dynamics/src/DGTransport.cpp:272-512
 *                         (cell_term + edge_term_X/Y + periodic + boundary_* + inverse mass), fused with the stage
 *                         update of DGTransport::step_rk1 / rk2 / rk3 (DGTransport.cpp:514-566) and, in the last
 *                         stage, LimitMax / LimitMin (dynamics/src/include/dgLimit.hpp:16-84); several fields per pass
 *
 * The reference scatters edge fluxes into both neighbours; here every element gathers the flux
 * through its own four edges (each interior edge flux is evaluated twice, identically), in the
 * reference's accumulation order: cell, left, right, bottom, top, then Dirichlet sides 0..3.
 * Periodic edges (DGTransport.cpp:466-481) are not reachable through IDynamics (DynamicsKernel.hpp:54 leaves them TODO);
 * they are set through nsdg_set_boundaries, the way the reference's own advection test assigns ParametricMesh::periodic.
 */
#pragma once
#include "nsdg_setup.cuh"

namespace nsdg {

//! element Dirichlet side bits (bit s set: element is in smesh.dirichlet[s]); bit 4: ice element
__device__ __forceinline__ bool isIce(const uint8_t* landmask, size_t e) { return __ldg(landmask + e) != 0; }

// ------------------------------------------------------------------------------------------
// CG -> DG L2 projection (per element).  Uses the stored inverse DG mass matrix of the transport
// map (Cartesian: exactly massMatrix<DG>().inverse(); spherical: that / EarthRadius, undone here).
// ------------------------------------------------------------------------------------------
template <int CG, int DG>
__global__ void cg2dg_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy,
    const double* __restrict__ cg, TransportOpPtrs op, double* __restrict__ dg)
{
    constexpr int G = gp1d(DG), Q = G * G, NR = CG + 1, ND = NR * NR;
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double loc[ND];
    for (int r = 0; r < NR; ++r)
        for (int c = 0; c < NR; ++c)
            loc[r * NR + c] = cg[size_t(CG * iy + r) * g.cgs + CG * ix + c];
    double crn[4][2], dx[2][Q], dy[2][Q], J[Q], lat[Q], gq[Q];
    elementCorners(vx, vy, g.nx, ix, iy, g.spherical, crn);
    elementMap<G>(crn, dx, dy, J, lat);
    for (int q = 0; q < Q; ++q) {
        double s = 0;
        for (int i = 0; i < ND; ++i)
            s += PHI(CG, G, i, q) * loc[i];
        double wj = J[q] * gaussweight2(G, q);
        if (g.spherical)
            wj *= cos(lat[q]);
        gq[q] = wj * s;
    }
    double rhs[DG];
    for (int j = 0; j < DG; ++j) {
        double s = 0;
        for (int q = 0; q < Q; ++q)
            s += PSI(G, j, q) * gq[q];
        rhs[j] = s;
    }
    const size_t eo = e * op.estride;
    for (int i = 0; i < DG; ++i) {
        double s = 0;
        for (int j = 0; j < DG; ++j)
            s += op.iMass[(i * DG + j) * op.pitch + eo] * rhs[j];
        dg[size_t(i) * g.Npad + e] = g.spherical ? s * EarthRadius : s;
    }
}

//! Both velocity components at once (DGTransport::prepareAdvection, DGTransport.cpp:262-268): the element map and the
//! inverse mass matrix are formed / read once for the pair.  UNI (congruent axis-aligned rectangles): J w = dx dy w_q and
//! M^-1 = diag(1 / (dx dy int psi_j^2)), so the element size cancels and neither geometry nor operators are read.
template <int CG, int DG, bool UNI>
__global__ void cg2dg_pair_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy,
    const double* __restrict__ cgA, const double* __restrict__ cgB, TransportOpPtrs op, double* __restrict__ dgA, double* __restrict__ dgB)
{
    constexpr int G = gp1d(DG), Q = G * G, NR = CG + 1, ND = NR * NR;
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double la[ND], lb[ND];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) {
            la[r * NR + c] = cgA[size_t(CG * iy + r) * g.cgs + CG * ix + c];
            lb[r * NR + c] = cgB[size_t(CG * iy + r) * g.cgs + CG * ix + c];
        }
    double wj[Q];
    if constexpr (UNI) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
            wj[q] = gaussweight2(G, q);
    } else {
        double crn[4][2], dx[2][Q], dy[2][Q], J[Q], lat[Q];
        elementCorners(vx, vy, g.nx, ix, iy, g.spherical, crn);
        elementMap<G>(crn, dx, dy, J, lat);
        for (int q = 0; q < Q; ++q) {
            wj[q] = J[q] * gaussweight2(G, q);
            if (g.spherical)
                wj[q] *= cos(lat[q]);
        }
    }
    double ra[DG], rb[DG];
#pragma unroll
    for (int j = 0; j < DG; ++j)
        ra[j] = rb[j] = 0.0;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double sa = 0, sb = 0;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            const double ph = PHI(CG, G, i, q);
            if (ph != 0.0) {
                sa = fma(ph, la[i], sa);
                sb = fma(ph, lb[i], sb);
            }
        }
        sa *= wj[q];
        sb *= wj[q];
#pragma unroll
        for (int j = 0; j < DG; ++j) {
            const double ps = PSI(G, j, q);
            if (ps != 0.0) {
                ra[j] = fma(ps, sa, ra[j]);
                rb[j] = fma(ps, sb, rb[j]);
            }
        }
    }
    if constexpr (UNI) {
        constexpr double minv[8] = { 1., 12., 12., 180., 180., 144., 2160., 2160. }; // 1 / int psi_j^2 on the unit square
#pragma unroll
        for (int i = 0; i < DG; ++i) {
            dgA[size_t(i) * g.Npad + e] = ra[i] * minv[i];
            dgB[size_t(i) * g.Npad + e] = rb[i] * minv[i];
        }
    } else {
        const size_t eo = e * op.estride;
        const double sc = g.spherical ? EarthRadius : 1.0; // the stored inverse carries 1 / EarthRadius on the sphere
        for (int i = 0; i < DG; ++i) {
            double sa = 0, sb = 0;
            for (int j = 0; j < DG; ++j) {
                const double m = op.iMass[(i * DG + j) * op.pitch + eo];
                sa += m * ra[j];
                sb += m * rb[j];
            }
            dgA[size_t(i) * g.Npad + e] = sa * sc;
            dgB[size_t(i) * g.Npad + e] = sb * sc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// DG -> CG nodal averaging (per node gather over <= 4 elements, reference order: odd element
// rows first, quirk Q7; weights 1/4 corner, 1/2 edge, 1 centre; domain-boundary nodes x2).
// Optional clamp [lo,hi] fuses prepareIteration's limits (CGDynamicsKernel.cpp:272-275).
// ------------------------------------------------------------------------------------------
template <int CG, int DG>
__global__ void dg2cg_kernel(GridDims g, const double* __restrict__ src, double* __restrict__ dest, double lo, double hi)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const int c = int(t % g.cgnx), r = int(t / g.cgnx);
    // the (up to) 2 x 2 elements touching the node: column A = the element to the left (only for nodes on an element
    // boundary), column B = the element the node lies in; rows likewise.  All four are LOADED unconditionally from clamped,
    // always valid indices (the loads are independent and in flight together); a missing one gets weight zero.
    const int jx = c % CG, jy = r % CG;
    const bool hasX[2] = { jx == 0 && c > 0, c < CG * g.nx }, hasY[2] = { jy == 0 && r > 0, r < CG * g.ny };
    const int ex[2] = { max(c / CG - 1, 0), min(c / CG, g.nx - 1) }, ey[2] = { max(r / CG - 1, 0), min(r / CG, g.ny - 1) };
    const int lx[2] = { CG, jx }, ly[2] = { CG, jy }; // local lattice index of the node in that element
    double At[2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const size_t e = size_t(ey[a]) * g.nxs + ex[b];
            // PSILagrange<DG, CG + 1>(j, q) = dgbasis(j, x, y) in the lattice point; the products below are exact there
            const double X = double(lx[b]) / CG - 0.5, Y = double(ly[a]) / CG - 0.5;
            double v = 0.0;
#pragma unroll
            for (int j = 0; j < DG; ++j) {
                const double w = j == 0 ? 1.0 : j == 1 ? X : j == 2 ? Y : j == 3 ? X * X - 1.0 / 12.0 : j == 4 ? Y * Y - 1.0 / 12.0 : j == 5 ? X * Y
                    : j == 6 ? Y * (X * X - 1.0 / 12.0) : X * (Y * Y - 1.0 / 12.0);
                v += src[size_t(j) * g.Npad + e] * w;
            }
            double wt = (CG == 1) ? 0.25 : ((lx[b] == 1) ? 1.0 : 0.5) * ((ly[a] == 1) ? 1.0 : 0.5);
            At[a][b] = (hasY[a] && hasX[b]) ? wt * v : 0.0;
        }
    // reference order of the accumulation (quirk Q7): odd element rows first, then even ones; within a row left to right.
    // Rows A and B are consecutive, so exactly one of them is odd; absent elements contribute +0.
    const int rowA = r / CG - 1;
    const bool aFirst = (rowA & 1) != 0;
    double sum = 0.0;
    sum += aFirst ? At[0][0] : At[1][0];
    sum += aFirst ? At[0][1] : At[1][1];
    sum += aFirst ? At[1][0] : At[0][0];
    sum += aFirst ? At[1][1] : At[0][1];
    // DG2CGBoundary, Interpolations.cpp:165-180: rows first, then columns (corners x4)
    // (only on edges of the global domain; a partition box's artificial edges are halo lines)
    if ((r == 0 && (g.bnd & 1)) || (r == g.cgny - 1 && (g.bnd & 4)))
        sum *= 2.0;
    if ((c == 0 && (g.bnd & 8)) || (c == g.cgnx - 1 && (g.bnd & 2)))
        sum *= 2.0;
    sum = fmin(fmax(sum, lo), hi);
    dest[size_t(r) * g.cgs + c] = sum;
}

// ------------------------------------------------------------------------------------------
// Normal velocity on the edges (per edge gather).  X-edges: nx*(ny+1), Y-edges: (nx+1)*ny.
// dirmask: per element, bit s = element has a Dirichlet edge on side s.
// Output planes: nvX[k*pitchX + edge], nvY[k*pitchY + edge].
// ------------------------------------------------------------------------------------------
template <int DG>
__global__ void normalvel_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy,
    const uint8_t* __restrict__ landmask, const uint8_t* __restrict__ dirmask, const double* __restrict__ velx,
    const double* __restrict__ vely, double* __restrict__ nvX, size_t pitchX, double* __restrict__ nvY, size_t pitchY)
{
    constexpr int ED = edgedofs(DG);
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    const long nXe = long(g.nx) * (g.ny + 1), nYe = long(g.nx + 1) * g.ny;
    if (t >= nXe + nYe)
        return;
    auto tangent = [&](size_t n1, size_t n2, double& tx, double& ty) { // ParametricMesh.hpp:381-397
        tx = vx[n2] - vx[n1];
        ty = vy[n2] - vy[n1];
        if (g.spherical) {
            if (tx > 0.5 * M_PI)
                tx -= 2.0 * M_PI;
            if (tx < -0.5 * M_PI)
                tx += 2.0 * M_PI;
        }
    };
    double acc[ED];
    for (int k = 0; k < ED; ++k)
        acc[k] = 0.0;
    bool twice = false;
    if (t < nXe) { // X-edge (ix, jy): between element rows jy-1 (below) and jy (above)
        const int ix = int(t % g.nx), jy = int(t / g.nx);
        double tx, ty;
        const size_t n = size_t(jy) * (g.nx + 1) + ix;
        tangent(n, n + 1, tx, ty);
        for (int side = 0; side < 2; ++side) { // 0: element below (its top edge), 1: element above (its bottom edge)
            const int ey = jy - 1 + side;
            if (ey < 0 || ey >= g.ny)
                continue;
            const size_t e = size_t(ey) * g.nxs + ix;
            if (!isIce(landmask, e))
                continue;
            double ex[ED], eyv[ED];
            edgeofcell<DG>([&](int k) { return velx[size_t(k) * g.Npad + e]; }, side == 0 ? 2 : 0, ex);
            edgeofcell<DG>([&](int k) { return vely[size_t(k) * g.Npad + e]; }, side == 0 ? 2 : 0, eyv);
            for (int k = 0; k < ED; ++k)
                acc[k] += 0.5 * (-ty * ex[k] + tx * eyv[k]);
            twice = twice || (dirmask[e] & (side == 0 ? 4 : 1));
        }
        for (int k = 0; k < ED; ++k)
            nvX[size_t(k) * pitchX + t] = twice ? acc[k] * 2.0 : acc[k];
    } else { // Y-edge (jx, iy): between element columns jx-1 (left) and jx (right)
        const long ty_ = t - nXe;
        const int jx = int(ty_ % (g.nx + 1)), iy = int(ty_ / (g.nx + 1));
        double tx, ty;
        const size_t n = size_t(iy) * (g.nx + 1) + jx;
        tangent(n, n + g.nx + 1, tx, ty);
        for (int side = 0; side < 2; ++side) { // 0: element on the left (its right edge), 1: on the right (its left edge)
            const int ex_ = jx - 1 + side;
            if (ex_ < 0 || ex_ >= g.nx)
                continue;
            const size_t e = size_t(iy) * g.nxs + ex_;
            if (!isIce(landmask, e))
                continue;
            double ex[ED], eyv[ED];
            edgeofcell<DG>([&](int k) { return velx[size_t(k) * g.Npad + e]; }, side == 0 ? 1 : 3, ex);
            edgeofcell<DG>([&](int k) { return vely[size_t(k) * g.Npad + e]; }, side == 0 ? 1 : 3, eyv);
            for (int k = 0; k < ED; ++k)
                acc[k] += 0.5 * (ty * ex[k] - tx * eyv[k]);
            twice = twice || (dirmask[e] & (side == 0 ? 2 : 8));
        }
        for (int k = 0; k < ED; ++k)
            nvY[size_t(k) * pitchY + ty_] = twice ? acc[k] * 2.0 : acc[k];
    }
}

// ------------------------------------------------------------------------------------------
// Limiters (dgLimit.hpp:16-84) on the coefficients of one element.  mode bits: 1 = LimitMax(maxv)
// then 2 = LimitMin(minv), in this order (the reference always calls LimitMax before LimitMin on
// the same field).
// ------------------------------------------------------------------------------------------
template <int DG> __device__ __forceinline__ void limitDG(double (&d)[DG], int mode, double maxv, double minv)
{
    if (mode & 1) {
        d[0] = fmin(maxv, d[0]);
        if constexpr (DG == 3) {
            const double l0 = 2.0 * fmax(fabs(d[1] + d[2]), fabs(d[1] - d[2]));
            if (l0 != 0 && d[0] + l0 - maxv > 0) {
                const double a = d[1] * ((maxv - d[0]) / l0), b = d[2] * ((maxv - d[0]) / l0);
                d[1] = a;
                d[2] = b;
            }
        } else if constexpr (DG == 6) {
            double mx = -1e300;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double v = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    v += d[j] * PSI(3, j, q);
                mx = fmax(mx, v);
            }
            const double l0 = mx - d[0];
            if (mx > maxv) {
                const double s = (maxv - d[0]) / l0;
#pragma unroll
                for (int j = 1; j < 6; ++j)
                    d[j] *= s;
            }
        }
    }
    if (mode & 2) {
        d[0] = fmax(minv, d[0]);
        if constexpr (DG == 3) {
            const double l0 = 2.0 * fmax(fabs(d[1] + d[2]), fabs(d[1] - d[2]));
            if (l0 != 0 && d[0] - l0 - minv < 0) {
                const double a = d[1] * ((d[0] - minv) / l0), b = d[2] * ((d[0] - minv) / l0);
                d[1] = a;
                d[2] = b;
            }
        } else if constexpr (DG == 6) {
            double mn = 1e300;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                double v = 0;
#pragma unroll
                for (int j = 0; j < 6; ++j)
                    v += d[j] * PSI(3, j, q);
                mn = fmin(mn, v);
            }
            const double l0 = mn - d[0];
            if (mn < minv) {
                const double s = (d[0] - minv) / l0;
#pragma unroll
                for (int j = 1; j < 6; ++j)
                    d[j] *= s;
            }
        }
    }
}

template <int DG> __global__ void limit_kernel(GridDims g, double* __restrict__ f, int mode, double maxv, double minv)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const size_t e = size_t(t_ / g.nx) * g.nxs + (t_ % g.nx);
    double d[DG];
#pragma unroll
    for (int j = 0; j < DG; ++j)
        d[j] = f[size_t(j) * g.Npad + e];
    limitDG<DG>(d, mode, maxv, minv);
#pragma unroll
    for (int j = 0; j < DG; ++j)
        f[size_t(j) * g.Npad + e] = d[j];
}

// ------------------------------------------------------------------------------------------
// One Runge-Kutta STAGE of the DG transport for up to three fields at once:
//     k   = M^-1 [ dt * cell term - dt * sum over edges of upwind flux ]        (DGTransportOperator, DGTransport.cpp:436-512)
//     out = epilogue(in, k, base)                                                 (step_rk1/rk2/rk3,    DGTransport.cpp:514-566)
//     out = limiter(out)   where asked                                            (LimitMax / LimitMin, dgLimit.hpp:16-84)
// The fields of one launch share everything that does not depend on them -- the velocity in the Gauss points, the four
// edge normal velocities, the element-map values (or AdvectionCellTermX/Y) and the inverse mass matrix -- which is most of
// the traffic: hice and cice (and the BBM damage) advance in one pass, the three stresses in another.  The sharing is
// through L2: neighbouring blocks take the same tile of elements for the different fields.
//
// One thread per element and field, one warp = 32 consecutive elements of a row.  The trace a neighbour needs from me is a linear map
// of my own coefficients (edgeofcell), so left / right neighbour traces travel by warp shuffle (lanes 0 / 31 and periodic
// seams load the neighbour's coefficients instead); bottom / top neighbours are read from the adjacent rows (L2 hits: the
// rows are in flight in neighbouring blocks).  Every interior edge flux is evaluated twice, identically, by its two
// elements; accumulation order per element as in the reference: cell, left, right, bottom, top, periodic, Dirichlet 0..3.
//
// Epilogues (epi):
//   0  EULER    out = in + k                                (rk1; first stage of rk2 and rk3: tmp1 = phi + k1)
//   1  HEUN     out = in + 0.5 (k - (in - base))            (rk2, second stage: phi1 + 0.5 (k2 - k1) with k1 = phi1 - phi0
//                                                            recomputed instead of stored: 1 ulp of phi, 6 doubles less traffic)
//   2  COMBINE  out = c1 (in + k) + c0 base                 (rk3: 0.25 (tmp1 + k) + 0.75 phi, then phi / 3 + 2/3 (tmp2 + k))
// `out` may alias `base` (each thread reads base[e] before it writes out[e]; neighbours read `in` only), never `in`.
//
// Periodic edges (DGTransport.cpp:466-481; lists of {type, c1, c2, edge}, ParametricMesh.hpp:79): per[side] holds, for the
// elements named in the lists, the element across the periodic edge and the edge whose normal velocity the list names (the
// reference's ring-mesh test names an INTERIOR edge of the same row, AdvectionPeriodicBC_test.cpp:244-248; kept verbatim).
// ------------------------------------------------------------------------------------------
constexpr int kTransportMaxFields = 3;
struct TransportStageArgs {
    GridDims g;
    double dt;
    const uint8_t *landmask, *dirmask;
    const double *velx, *vely, *nvX, *nvY;
    size_t pitchX, pitchY;
    TransportOpPtrs op;
    const double* geo; //!< element-map planes of the factored-operator path (G = 3 only) or nullptr
    double dxU, dyU, iAreaU; //!< uniform rectangular mesh (UNI kernels): cell size and 1 / (dx dy)
    const int* perNbr; //!< [4][Npad] element (plane index) across a periodic edge on that side, -1 = none; nullptr = no periodic edges
    const int* perEdge; //!< [4][Npad] index of the edge whose normal velocity that flux uses
    int nf; //!< fields advanced by this launch (grid.x = tiles * nf)
    const double* in[kTransportMaxFields];
    const double* base[kTransportMaxFields];
    double* out[kTransportMaxFields];
    int epi;
    double c0, c1;
    int limitMode[kTransportMaxFields];
    double maxv[kTransportMaxFields], minv[kTransportMaxFields];
};

#ifndef NSDG_TRANSPORT_MINB
#define NSDG_TRANSPORT_MINB 4 //!< resident blocks per SM asked of the compiler (128 registers): the kernel is latency bound, occupancy pays
#endif
#ifndef NSDG_TRANSPORT_HOIST
#define NSDG_TRANSPORT_HOIST 1 //!< uniform-mesh kernels: all neighbour / edge-velocity loads requested up front (more loads in flight)
#endif
#ifndef NSDG_TRANSPORT_SHFL
#define NSDG_TRANSPORT_SHFL 1 //!< left / right neighbour traces by warp shuffle (0: load the neighbour's coefficients)
#endif
//! UNI: all elements are congruent axis-aligned rectangles (dxU x dyU).  Then the element map is constant --
//! AdvectionCellTermX = w dy PSIx, AdvectionCellTermY = w dx PSIy (ParametricMap.cpp:13-66) -- and the DG mass matrix is
//! dx dy diag(int psi_j^2) (the Legendre-type basis is orthogonal on the rectangle), so neither the 2 x DG x Q cell-term
//! entries nor the DG x DG inverse mass matrix are read: 144 loads and ~250 FP64 operations less per element.
template <int DG, bool UNI>
__global__ void __launch_bounds__(128, NSDG_TRANSPORT_MINB) transport_stage_kernel(const __grid_constant__ TransportStageArgs a)
{
    constexpr int G = gp1d(DG), Q = G * G, ED = edgedofs(DG);
    constexpr unsigned FULL = 0xffffffffu;
    const GridDims& g = a.g;
    const int lane = threadIdx.x & 31;
    // blockIdx.x = tile * nf + field: the blocks that advance the different fields of one tile are scheduled together, so
    // what they share (velocities, edge velocities, element map, inverse mass) comes from DRAM once and from L2 after
    // that, while each thread carries the registers of ONE field (interleaving the fields in one thread spills)
    const int f = blockIdx.x % a.nf;
    const int ixRaw = (blockIdx.x / a.nf) * blockDim.x + threadIdx.x, iy = blockIdx.y;
    const bool active = ixRaw < g.nx;
    const int ix = active ? ixRaw : g.nx - 1;
    const size_t e = size_t(iy) * g.nxs + ix;
    const size_t Npad = g.Npad;
    const double dt = a.dt;
    const bool ice = isIce(a.landmask, e);
    const double* __restrict__ phi = a.in[f];
    const double* __restrict__ velx = a.velx;
    const double* __restrict__ vely = a.vely;
    const double* __restrict__ nvX = a.nvX;
    const double* __restrict__ nvY = a.nvY;
    double up[DG];
#pragma unroll
    for (int j = 0; j < DG; ++j)
        up[j] = 0.0;
    double ph[DG];
#pragma unroll
    for (int j = 0; j < DG; ++j)
        ph[j] = phi[size_t(j) * Npad + e];

    const size_t eo = e * a.op.estride;
    constexpr bool kHoist = UNI && (NSDG_TRANSPORT_HOIST != 0);
    // (kHoist) Everything the four edge fluxes read is requested HERE, unconditionally and from addresses that are always valid (a
    // missing neighbour is replaced by the element itself), so that ~40 independent loads per thread are in flight while
    // the cell term is computed.  Guarded loads inside the flux code cannot be hoisted by the compiler and cost two
    // dependent memory latencies per edge (ncu: 5.6 long-scoreboard stalls per issued instruction).
    // side: 0 bottom, 1 right, 2 top, 3 left
    // Measured at 2048^2 (profiles/r2_quickbench_transport_variants.txt): advection 2.23 -> 1.83 ms (mEVP), 6.2 -> 5.4 ms (BBM)
    // on uniform meshes; on parametric meshes, whose kernels also hold the element map and read the inverse mass matrix,
    // the extra registers spill and nothing is gained, so those keep the just-in-time form.
    bool nbOk[4] = { false, false, false, false };
    double nvs[4][ED], nbc[4][DG];
    if constexpr (kHoist) {
        const int nix[4] = { ix, ix + 1, ix, ix - 1 }, niy[4] = { iy - 1, iy, iy + 1, iy };
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const bool inside = nix[s] >= 0 && nix[s] < g.nx && niy[s] >= 0 && niy[s] < g.ny;
            const size_t en = inside ? size_t(niy[s]) * g.nxs + nix[s] : e;
            const size_t ie = (s == 0 || s == 2) ? size_t(s == 2 ? iy + 1 : iy) * g.nx + ix : size_t(iy) * (g.nx + 1) + (s == 1 ? ix + 1 : ix);
            nbOk[s] = inside && ice && isIce(a.landmask, en);
#pragma unroll
            for (int k = 0; k < ED; ++k)
                nvs[s][k] = (s == 0 || s == 2) ? nvX[size_t(k) * a.pitchX + ie] : nvY[size_t(k) * a.pitchY + ie];
            // left / right neighbours sit in the same cache lines as my own row: loading them costs no DRAM traffic
#pragma unroll
            for (int j = 0; j < DG; ++j)
                nbc[s][j] = phi[size_t(j) * Npad + en];
        }
    }
    // ---- cell term (DGTransport.cpp:278-303) ----
    if (DG > 1 && ice) {
        double vxg[Q], vyg[Q], pg[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q)
            vxg[q] = vyg[q] = pg[q] = 0.0;
#pragma unroll
        for (int j = 0; j < DG; ++j) {
            const double vxj = velx[size_t(j) * Npad + e], vyj = vely[size_t(j) * Npad + e];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const double w = PSI(G, j, q);
                if (w != 0.0) {
                    vxg[q] = fma(vxj, w, vxg[q]);
                    vyg[q] = fma(vyj, w, vyg[q]);
                    pg[q] = fma(ph[j], w, pg[q]);
                }
            }
        }
        if (UNI || (G == 3 && a.geo != nullptr)) {
            // AdvectionCellTermX/Y (ParametricMap.cpp:13-66) are w (PSIx dyT_1 - PSIy dxT_1) and w (PSIy dxT_0 - PSIx dyT_0):
            // formed from the 12 element-map values of the factored-operator path (nsdg_momentum_param.cuh, planes 0..11)
            // instead of streaming 2 x DG x 9 doubles per element; on a uniform rectangle they are just dy and dx
            double aq[Q], bq[Q];
            if constexpr (UNI) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double wp = (dt * gaussweight2(G, q)) * pg[q];
                    aq[q] = wp * (a.dyU * vxg[q]);
                    bq[q] = wp * (a.dxU * vyg[q]);
                }
            } else {
                double m[12];
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    m[k] = __ldg(a.geo + size_t(k) * Npad + e);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int qx = q % 3, qy = q / 3;
                    const double wp = (dt * gaussweight2(3, q)) * pg[q];
                    aq[q] = wp * (m[9 + qx] * vxg[q] - m[6 + qx] * vyg[q]); // yeta vx - xeta vy
                    bq[q] = wp * (m[0 + qy] * vyg[q] - m[3 + qy] * vxg[q]); // xxi vy - yxi vx
                }
            }
#pragma unroll
            for (int j = 0; j < DG; ++j) {
                double acc = 0;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double px = PSIx(G, j, q), py = PSIy(G, j, q);
                    if (px != 0.0)
                        acc = fma(px, aq[q], acc);
                    if (py != 0.0)
                        acc = fma(py, bq[q], acc);
                }
                up[j] += acc;
            }
        } else {
#pragma unroll
            for (int j = 0; j < DG; ++j) {
                double acc = 0;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const double ax = __ldg(a.op.AdvX + (j * Q + q) * a.op.pitch + eo);
                    const double ay = __ldg(a.op.AdvY + (j * Q + q) * a.op.pitch + eo);
                    acc += (dt * (ax * vxg[q] + ay * vyg[q])) * pg[q];
                }
                up[j] += acc;
            }
        }
    }
    // ---- edges (DGTransport.cpp:390-433, :466-481).  side: 0 bottom, 1 right, 2 top, 3 left; reference order of the
    //      accumulation: left, right, bottom, top, then the periodic ones.  `shfl`: the neighbour's trace arrives from the
    //      adjacent lane (all lanes take part in the shuffle, whether or not they have such an edge) ----
    auto edgeFlux = [&](int side, bool wantPeriodic) {
        const int nix = ix + (side == 1 ? 1 : (side == 3 ? -1 : 0));
        const int niy = iy + (side == 2 ? 1 : (side == 0 ? -1 : 0));
        const bool inside = nix >= 0 && nix < g.nx && niy >= 0 && niy < g.ny;
        long en = inside ? long(size_t(niy) * g.nxs + nix) : -1;
        long ie = (side == 0 || side == 2) ? long(size_t(side == 2 ? iy + 1 : iy) * g.nx + ix) : long(size_t(iy) * (g.nx + 1) + (side == 1 ? ix + 1 : ix));
        bool periodic = false;
        if (!inside && a.perNbr != nullptr) {
            const int pn = a.perNbr[size_t(side) * Npad + e];
            if (pn >= 0) {
                en = pn;
                ie = a.perEdge[size_t(side) * Npad + e];
                periodic = true;
            }
        }
        double tme[ED], tnb[ED];
        edgeofcell<DG>([&](int k) { return ph[k]; }, side, tme);
        bool viaShfl = false;
#if NSDG_TRANSPORT_SHFL
        if (!wantPeriodic && (side == 1 || side == 3)) { // uniform across the warp: every lane shuffles
            double mine[ED];
            edgeofcell<DG>([&](int k) { return ph[k]; }, (side + 2) & 3, mine); // what the neighbour on `side` wants from me is my trace on the opposite side
#pragma unroll
            for (int k = 0; k < ED; ++k)
                tnb[k] = side == 3 ? __shfl_up_sync(FULL, mine[k], 1) : __shfl_down_sync(FULL, mine[k], 1);
            viaShfl = side == 3 ? (lane > 0) : (lane < 31 && ixRaw + 1 < g.nx);
        }
#endif
        if (en < 0 || periodic != wantPeriodic || !ice || !isIce(a.landmask, size_t(en)))
            return; // edge_term_X/Y return when either element is land (DGTransport.cpp:395-398)
        // c1 = left/bottom element, c2 = right/top element of the edge
        const bool meFirst = (side == 1 || side == 2);
        double nv[ED];
#pragma unroll
        for (int k = 0; k < ED; ++k)
            nv[k] = (side == 0 || side == 2) ? nvX[size_t(k) * a.pitchX + ie] : nvY[size_t(k) * a.pitchY + ie];
        if (!viaShfl)
            edgeofcell<DG>([&](int k) { return phi[size_t(k) * Npad + size_t(en)]; }, (side + 2) & 3, tnb);
        double tmp[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            double vg = 0, g1 = 0, g2 = 0;
#pragma unroll
            for (int k = 0; k < ED; ++k) {
                vg += nv[k] * PSIe(G, k, q);
                g1 += (meFirst ? tme[k] : tnb[k]) * PSIe(G, k, q);
                g2 += (meFirst ? tnb[k] : tme[k]) * PSIe(G, k, q);
            }
            tmp[q] = fmax(vg, 0.) * g1 + fmin(vg, 0.) * g2;
        }
        const double sdt = meFirst ? -dt : dt;
#pragma unroll
        for (int j = 0; j < DG; ++j) {
            double acc = 0;
#pragma unroll
            for (int q = 0; q < G; ++q)
                acc += (sdt * tmp[q]) * PSIew(G, side, q, j);
            up[j] += acc;
        }
    };
    if constexpr (kHoist) {
        const int order[4] = { 3, 1, 0, 2 };
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int side = order[i];
            if (!nbOk[side])
                continue;
            const bool meFirst = (side == 1 || side == 2); // c1 = left/bottom element, c2 = right/top element of the edge
            double tme[ED], tnb[ED];
            edgeofcell<DG>([&](int k) { return ph[k]; }, side, tme);
            edgeofcell<DG>([&](int k) { return nbc[side][k]; }, (side + 2) & 3, tnb);
            double tmp[G];
#pragma unroll
            for (int q = 0; q < G; ++q) {
                double vg = 0, g1 = 0, g2 = 0;
#pragma unroll
                for (int k = 0; k < ED; ++k) {
                    vg += nvs[side][k] * PSIe(G, k, q);
                    g1 += (meFirst ? tme[k] : tnb[k]) * PSIe(G, k, q);
                    g2 += (meFirst ? tnb[k] : tme[k]) * PSIe(G, k, q);
                }
                tmp[q] = fmax(vg, 0.) * g1 + fmin(vg, 0.) * g2;
            }
            const double sdt = meFirst ? -dt : dt;
#pragma unroll
            for (int j = 0; j < DG; ++j) {
                double acc = 0;
#pragma unroll
                for (int q = 0; q < G; ++q)
                    acc += (sdt * tmp[q]) * (side == 0 ? PSIew(G, 0, q, j) : side == 1 ? PSIew(G, 1, q, j) : side == 2 ? PSIew(G, 2, q, j) : PSIew(G, 3, q, j));
                up[j] += acc;
            }
        }
    } else {
        edgeFlux(3, false);
        edgeFlux(1, false);
        edgeFlux(0, false);
        edgeFlux(2, false);
    }
    if (a.perNbr != nullptr) {
        edgeFlux(3, true);
        edgeFlux(1, true);
        edgeFlux(0, true);
        edgeFlux(2, true);
    }
    // ---- Dirichlet edges: outflow only (DGTransport.cpp:306-351), sides in list order 0,1,2,3 ----
    const uint8_t dm = a.dirmask[e];
    if (dm)
        for (int side = 0; side < 4; ++side) {
            if (!(dm & (1 << side)))
                continue;
            double nv[ED];
            if (side == 0 || side == 2) {
                const size_t ie = size_t(side == 2 ? iy + 1 : iy) * g.nx + ix;
                for (int k = 0; k < ED; ++k)
                    nv[k] = nvX[size_t(k) * a.pitchX + ie];
            } else {
                const size_t ie = size_t(iy) * (g.nx + 1) + (side == 1 ? ix + 1 : ix);
                for (int k = 0; k < ED; ++k)
                    nv[k] = nvY[size_t(k) * a.pitchY + ie];
            }
            double tme[ED];
            edgeofcell<DG>([&](int k) { return ph[k]; }, side, tme);
            const double sg = (side == 0 || side == 3) ? -1.0 : 1.0;
            double tmp[G];
            for (int q = 0; q < G; ++q) {
                double vg = 0, g1 = 0;
                for (int k = 0; k < ED; ++k) {
                    vg += nv[k] * PSIe(G, k, q);
                    g1 += tme[k] * PSIe(G, k, q);
                }
                tmp[q] = g1 * fmax(sg * vg, 0.);
            }
            for (int j = 0; j < DG; ++j) {
                double acc = 0;
                for (int q = 0; q < G; ++q) {
                    const double w = side == 0 ? PSIew(G, 0, q, j)
                        : side == 1            ? PSIew(G, 1, q, j)
                        : side == 2            ? PSIew(G, 2, q, j)
                                               : PSIew(G, 3, q, j);
                    acc += (-dt * tmp[q]) * w;
                }
                up[j] += acc;
            }
        }
    // ---- inverse mass (DGTransport.cpp:509-511), Runge-Kutta epilogue, limiter ----
    double res[DG];
#pragma unroll
    for (int i = 0; i < DG; ++i) {
        double k = 0;
        if constexpr (UNI) {
            constexpr double minv[8] = { 1., 12., 12., 180., 180., 144., 2160., 2160. }; // 1 / int psi_i^2 on the unit square
            k = up[i] * (minv[i] * a.iAreaU);
        } else {
#pragma unroll
            for (int j = 0; j < DG; ++j)
                k += __ldg(a.op.iMass + (i * DG + j) * a.op.pitch + eo) * up[j];
        }
        if (a.epi == 0)
            res[i] = ph[i] + k;
        else {
            const double b = a.base[f][size_t(i) * Npad + e];
            res[i] = a.epi == 1 ? ph[i] + 0.5 * (k - (ph[i] - b)) : a.c1 * (ph[i] + k) + a.c0 * b;
        }
    }
    if (a.limitMode[f])
        limitDG<DG>(res, a.limitMode[f], a.maxv[f], a.minv[f]);
    if (active) {
#pragma unroll
        for (int i = 0; i < DG; ++i)
            a.out[f][size_t(i) * Npad + e] = res[i];
    }
}

} // namespace nsdg
