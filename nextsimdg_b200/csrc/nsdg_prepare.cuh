/*
 * nsdg_prepare.cuh -- once-per-mesh and once-per-step kernels around the subcycle loop.
 *
 *   setup_*_kernel         operator planes per element      dynamics/src/ParametricMap.cpp:13-87, :209-356
 *   lumpedmass_kernel      lumpedcgmass / lumpedcg1mass     dynamics/src/ParametricMap.cpp:94-206
 *   nodemask_kernel        node set of dirichletZero        dynamics/src/CGDynamicsKernel.cpp:401-437
 *   sshgrad_*_kernel       ComputeGradientOfSeaSurfaceHeight dynamics/src/CGDynamicsKernel.cpp:126-258
 *   gaussconst_kernel      the h,a-dependent factors of the stress update, constant over a step:
 *                          mEVP P = P* h exp(-20(1-a))       MEVPStressUpdateStep.hpp:53-56,74-76
 *                          BBM  h, exp(C(1-a))               BBMStressUpdateStep.hpp:54-57,76-78
 *   iostress_kernel        getIceOceanStress                VPCGDynamicsKernel.hpp:96-110,
 *                                                           BrittleCGDynamicsKernel.hpp:168-182
 */
#pragma once
#include "nsdg_transport.cuh"

namespace nsdg {

template <int DG>
__global__ void setup_transport_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy,
    TransportOpPtrs op, int nelem)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(nelem))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double c[4][2];
    elementCorners(vx, vy, g.nx, ix, iy, g.spherical, c);
    transportOpsOfElement<DG>(c, g.spherical, op, e);
}

template <int CG, int DGA>
__global__ void setup_momentum_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy,
    MomentumOpPtrs op, int nelem, bool full)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(nelem))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double c[4][2];
    elementCorners(vx, vy, g.nx, ix, iy, g.spherical, c);
    momentumOpsOfElement<CG, DGA>(c, g.spherical, op, e, full);
}

//! element size h = sqrt(area) (ParametricMesh.hpp:276-299), used by the BBM stress update
__global__ void helem_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy, double* __restrict__ h)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    h[size_t(iy) * g.nxs + ix] = sqrt(elementArea(vx, vy, g.nx, ix, iy));
}

//! lumped mass per node of the CG(CGM) space, CGGP Gauss points per direction; gather in the
//! reference's scatter order (even element rows first, then odd; x ascending).
template <int CGM, int CGGP>
__global__ void lumpedmass_kernel(GridDims g, int cgnx, int cgny, int cgs, const double* __restrict__ vx,
    const double* __restrict__ vy, double* __restrict__ mass)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(cgnx) * cgny)
        return;
    const int c = int(t % cgnx), r = int(t / cgnx);
    const int jx = c % CGM, jy = r % CGM;
    double sum = 0.0;
    for (int pass = 0; pass < 2; ++pass)
        for (int a = 0; a < 2; ++a) {
            // a = 0: element row below the node row (local row CGM), a = 1: element row containing it
            int ey, ly;
            if (a == 0) {
                if (!(jy == 0 && r > 0))
                    continue;
                ey = r / CGM - 1;
                ly = CGM;
            } else {
                if (r >= CGM * g.ny)
                    continue;
                ey = r / CGM;
                ly = jy;
            }
            if ((ey % 2) != pass)
                continue;
            for (int b = 0; b < 2; ++b) {
                int ex, lx;
                if (b == 0) {
                    if (!(jx == 0 && c > 0))
                        continue;
                    ex = c / CGM - 1;
                    lx = CGM;
                } else {
                    if (c >= CGM * g.nx)
                        continue;
                    ex = c / CGM;
                    lx = jx;
                }
                double crn[4][2];
                elementCorners(vx, vy, g.nx, ex, ey, g.spherical, crn);
                sum += lumpedMassShare<CGM, CGGP>(crn, g.spherical, ly * (CGM + 1) + lx);
            }
        }
    mass[size_t(r) * cgs + c] = sum;
}

//! marks the CG nodes of every listed Dirichlet element edge (CGDynamicsKernel.cpp:401-437)
template <int CG>
__global__ void nodemask_kernel(GridDims g, const uint8_t* __restrict__ dirmask, uint8_t* __restrict__ nodemask)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const uint8_t dm = dirmask[size_t(iy) * g.nxs + ix];
    if (!dm)
        return;
    for (int j = 0; j <= CG; ++j) {
        if (dm & 1)
            nodemask[size_t(CG * iy) * g.cgs + CG * ix + j] = 1;
        if (dm & 2)
            nodemask[size_t(CG * iy + j) * g.cgs + CG * ix + CG] = 1;
        if (dm & 4)
            nodemask[size_t(CG * iy + CG) * g.cgs + CG * ix + j] = 1;
        if (dm & 8)
            nodemask[size_t(CG * iy + j) * g.cgs + CG * ix] = 1;
    }
}

//! raw CG1 gradient of the CG1 sea-surface height: per CG1 node gather, even element rows first
//! (CGDynamicsKernel.cpp:138-175).  g1 = CG1 node grid (cgnx = nx+1).
__global__ void sshgrad_cg1_kernel(GridDims g, int cg1s, const double* __restrict__ cgSSH, MomentumOpPtrs op,
    const double* __restrict__ mass1, double* __restrict__ gu, double* __restrict__ gv)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    const int n1x = g.nx + 1, n1y = g.ny + 1;
    if (t >= long(n1x) * n1y)
        return;
    const int c = int(t % n1x), r = int(t / n1x);
    double sx = 0.0, sy = 0.0;
    for (int pass = 0; pass < 2; ++pass)
        for (int a = 0; a < 2; ++a) {
            const int ey = r - 1 + a, ly = 1 - a;
            if (ey < 0 || ey >= g.ny || (ey % 2) != pass)
                continue;
            for (int b = 0; b < 2; ++b) {
                const int ex = c - 1 + b, lx = 1 - b;
                if (ex < 0 || ex >= g.nx)
                    continue;
                const size_t e = size_t(ey) * g.nxs + ex;
                const size_t n0 = size_t(ey) * cg1s + ex;
                const double loc[4] = { cgSSH[n0], cgSSH[n0 + 1], cgSSH[n0 + cg1s], cgSSH[n0 + cg1s + 1] };
                const int i = ly * 2 + lx;
                double tx = 0, ty = 0;
                for (int j = 0; j < 4; ++j) {
                    tx += op.dXssh[(i * 4 + j) * op.pitch + e * op.estride] * loc[j];
                    ty += op.dYssh[(i * 4 + j) * op.pitch + e * op.estride] * loc[j];
                }
                sx -= tx;
                sy -= ty;
            }
        }
    const size_t n = size_t(r) * cg1s + c;
    gu[n] = sx / mass1[n];
    gv[n] = sy / mass1[n];
}

//! boundary extension + CG1 -> CG interpolation of the SSH gradient (CGDynamicsKernel.cpp:177-257).
//! The boundary copies of the reference only ever read interior CG1 nodes, so the corrected CG1
//! value at (i,j) is the raw value at (clamp(i,1,nx-1), clamp(j,1,ny-1)).
template <int CG>
__global__ void sshgrad_cg_kernel(GridDims g, int cg1s, const double* __restrict__ gu1, const double* __restrict__ gv1,
    double* __restrict__ gu, double* __restrict__ gv)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const int c = int(t % g.cgnx), r = int(t / g.cgnx);
    auto at = [&](const double* f, int i, int j) {
        if (g.bnd & 8)
            i = max(i, 1);
        if (g.bnd & 2)
            i = min(i, g.nx - 1);
        if (g.bnd & 1)
            j = max(j, 1);
        if (g.bnd & 4)
            j = min(j, g.ny - 1);
        return f[size_t(j) * cg1s + i];
    };
    double u, v;
    if (CG == 1) {
        u = at(gu1, c, r);
        v = at(gv1, c, r);
    } else {
        const int i = c / 2, j = r / 2, mx = c % 2, my = r % 2;
        if (!mx && !my) {
            u = at(gu1, i, j);
            v = at(gv1, i, j);
        } else if (mx && !my) {
            u = 0.5 * (at(gu1, i, j) + at(gu1, i + 1, j));
            v = 0.5 * (at(gv1, i, j) + at(gv1, i + 1, j));
        } else if (!mx && my) {
            u = 0.5 * (at(gu1, i, j) + at(gu1, i, j + 1));
            v = 0.5 * (at(gv1, i, j) + at(gv1, i, j + 1));
        } else {
            u = 0.25 * (at(gu1, i, j) + at(gu1, i + 1, j) + at(gu1, i, j + 1) + at(gu1, i + 1, j + 1));
            v = 0.25 * (at(gv1, i, j) + at(gv1, i + 1, j) + at(gv1, i, j + 1) + at(gv1, i + 1, j + 1));
        }
    }
    gu[size_t(r) * g.cgs + c] = u;
    gv[size_t(r) * g.cgs + c] = v;
}

//! per-step Gauss-point constants of the stress update
template <int DGA, int GS, int RHEO>
__global__ void gaussconst_kernel(GridDims g, PhysParams p, const double* __restrict__ hice, const double* __restrict__ cice,
    double* __restrict__ outA, double* __restrict__ outB, double scaleA)
{
    constexpr int Q = GS * GS;
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const size_t e = size_t(t_ / g.nx) * g.nxs + (t_ % g.nx);
    double h[DGA], a[DGA];
#pragma unroll
    for (int j = 0; j < DGA; ++j) {
        h[j] = hice[size_t(j) * g.Npad + e];
        a[j] = cice[size_t(j) * g.Npad + e];
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        double hq = 0, aq = 0;
#pragma unroll
        for (int j = 0; j < DGA; ++j) {
            const double w = PSI(GS, j, q);
            if (w != 0.0) {
                hq = (j == 0) ? h[0] * w : fma(h[j], w, hq);
                aq = (j == 0) ? a[0] * w : fma(a[j], w, aq);
            }
        }
        hq = fmax(hq, 0.0);
        aq = fmin(fmax(aq, 0.0), 1.0);
        if constexpr (RHEO == NSDG_MEVP) {
            outA[size_t(q) * g.Npad + e] = scaleA * (p.Pstar * hq * exp(-20.0 * (1.0 - aq)));
        } else {
            outA[size_t(q) * g.Npad + e] = hq;
            outB[size_t(q) * g.Npad + e] = exp(p.compaction_param * (1.0 - aq));
        }
    }
}

//! ice-ocean stress on the CG nodes (quirk Q14)
template <int RHEO>
__global__ void iostress_kernel(GridDims g, PhysParams p, const double* __restrict__ u, const double* __restrict__ v,
    const double* __restrict__ uO, const double* __restrict__ vO, double* __restrict__ taux, double* __restrict__ tauy)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const size_t n = size_t(t / g.cgnx) * g.cgs + (t % g.cgnx);
    if constexpr (RHEO == NSDG_MEVP) {
        const double uR = u[n] - uO[n], vR = v[n] - vO[n];
        const double absocn = sqrt(uR * uR + vR * vR);
        taux[n] = p.F_ocean * absocn * uR;
        tauy[n] = p.F_ocean * absocn * vR;
    } else { // BBM: u, v = avgU, avgV ; free drift: u, v (FreeDriftDynamicsKernel.hpp:70-83)
        const double uR = uO[n] - u[n], vR = vO[n] - v[n];
        const double cPrime = p.F_ocean * hypot(uR, vR);
        taux[n] = cPrime * (uR * p.cosOceanAngle - vR * p.sinOceanAngle);
        tauy[n] = cPrime * (vR * p.cosOceanAngle + uR * p.sinOceanAngle);
    }
}

//! The four outputs that have to wait for the last subcycle -- cell means of u, v (CGDynamicsKernel::getDG0Data,
//! CGDynamicsKernel.cpp:92-118: CG2DG, component 0) and of the ice-ocean stress (getIceOceanStress + CG2DG) -- in ONE pass
//! on a uniform rectangular mesh, where component 0 of the L2 projection is the quadrature mean sum_q w_q f(q) (J w and the
//! (0,0) entry of the inverse mass matrix cancel).  The general path runs iostress_kernel and a full DG projection per
//! field.  (us, vs): the velocity the stress is taken with (mEVP: u, v; BBM: the running mean, quirk Q14).
template <int CG, int RHEO>
__global__ void export_dg0_uniform_kernel(GridDims g, PhysParams p, const double* __restrict__ u, const double* __restrict__ v,
    const double* __restrict__ us, const double* __restrict__ vs, const double* __restrict__ uO, const double* __restrict__ vO,
    double* __restrict__ out /* planes: u, v, taux, tauy */)
{
    constexpr int G = 3, Q = G * G, NR = CG + 1, ND = NR * NR;
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double f[4][ND];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int c = 0; c < NR; ++c) {
            const size_t n = size_t(CG * iy + r) * g.cgs + CG * ix + c;
            const int i = r * NR + c;
            f[0][i] = u[n];
            f[1][i] = v[n];
            const double a = us[n], b = vs[n], ao = uO[n], bo = vO[n];
            if constexpr (RHEO == NSDG_MEVP) {
                const double uR = a - ao, vR = b - bo;
                const double absocn = sqrt(uR * uR + vR * vR);
                f[2][i] = p.F_ocean * absocn * uR;
                f[3][i] = p.F_ocean * absocn * vR;
            } else {
                const double uR = ao - a, vR = bo - b;
                const double cPrime = p.F_ocean * hypot(uR, vR);
                f[2][i] = cPrime * (uR * p.cosOceanAngle - vR * p.sinOceanAngle);
                f[3][i] = cPrime * (vR * p.cosOceanAngle + uR * p.sinOceanAngle);
            }
        }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double mean = 0.0;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            double sq = 0.0;
#pragma unroll
            for (int i = 0; i < ND; ++i) {
                const double ph = PHI(CG, G, i, q);
                if (ph != 0.0)
                    sq = fma(ph, f[k][i], sq);
            }
            mean = fma(gaussweight2(G, q), sq, mean);
        }
        out[size_t(k) * g.Npad + e] = mean;
    }
}

//! FreeDriftDynamicsKernel::updateMomentum + applyBoundaries (FreeDriftDynamicsKernel.hpp:43-68)
__global__ void freedrift_kernel(GridDims g, PhysParams p, const double* __restrict__ uO, const double* __restrict__ vO,
    const double* __restrict__ uA, const double* __restrict__ vA, const uint8_t* __restrict__ nodemask, double* __restrict__ u,
    double* __restrict__ v)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const size_t n = size_t(t / g.cgnx) * g.cgs + (t % g.cgnx);
    const double NansenNumber = sqrt(p.F_atm / p.F_ocean);
    double un = uO[n] + NansenNumber * (uA[n] * p.cosOceanAngle - vA[n] * p.sinOceanAngle);
    double vn = vO[n] + NansenNumber * (-uA[n] * p.sinOceanAngle + vA[n] * p.cosOceanAngle);
    if (nodemask[n] & 1)
        un = vn = 0.0;
    u[n] = un;
    v[n] = vn;
}

/*
 * IDamageHealing::ConstantHealing::updateElement (physics/src/modules/DamageHealingModule/ConstantHealing.cpp:53-72) on
 * the DG0 damage and concentration:  lateral ice formation is undamaged, then linear healing with time scale tD.
 * deltaCi may be null (no thermodynamic growth, as with DummyIceThermodynamics).
 */
__global__ void healing_kernel(GridDims g, double step, double tD, const double* __restrict__ deltaCi, const double* __restrict__ cice,
    double* __restrict__ damage)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const size_t e = size_t(t_ / g.nx) * g.nxs + (t_ % g.nx);
    const double lateralGrowth = deltaCi ? fmax(0., deltaCi[e]) : 0.;
    double d = damage[e];
    d = (d * (cice[e] - lateralGrowth) + lateralGrowth) / cice[e];
    d += step / tD;
    damage[e] = fmin(1., d);
}

//! The reference's benchmark forcing per element (cell lower-left corner coordinates x = i dx, y = j dy of the
//! GLOBAL grid): BenchmarkAtmosphere.cpp:38-74 (cyclone centred at x0c, y0c) and BenchmarkOcean.cpp:27-36.
__global__ void benchforcing_kernel(GridDims g, int gi0, int gj0, double dx, double dy, double x0c, double y0c, double cosa,
    double sina, double Lx, double Ly, double* __restrict__ uw, double* __restrict__ vw, double* __restrict__ uo,
    double* __restrict__ vo)
{
    const size_t t_ = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t_ >= size_t(g.N))
        return;
    const int ix = int(t_ % g.nx), iy = int(t_ / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    const double x = double(gi0 + ix) * dx, y = double(gj0 + iy) * dy;
    const double xp = x - x0c, yp = y - y0c;
    const double A = 1e-5, k = 1e-5, vMax = 30.0, vMaxOcean = 0.01;
    const double scale = A * exp(-k * hypot(xp, yp));
    uw[e] = -scale * vMax * (cosa * xp + sina * yp);
    vw[e] = -scale * vMax * (-sina * xp + cosa * yp);
    uo[e] = vMaxOcean * (2 * (y / Ly) - 1);
    vo[e] = vMaxOcean * (1 - 2 * (x / Lx));
}

//! AoS (row-major N x ncomp, the ModelArray layout) <-> planes
__global__ void aos2planes_kernel(GridDims g, int ncomp, int nplanes, const double* __restrict__ aos, double* __restrict__ planes)
{
    const size_t d = size_t(blockIdx.x) * blockDim.x + threadIdx.x; // dense element index of the host array
    if (d >= size_t(g.N))
        return;
    const size_t e = size_t(d / g.nx) * g.nxs + (d % g.nx);
    for (int c = 0; c < nplanes; ++c)
        planes[size_t(c) * g.Npad + e] = (c < ncomp) ? aos[d * ncomp + c] : 0.0;
}
__global__ void planes2aos_kernel(GridDims g, int ncomp, const double* __restrict__ planes, double* __restrict__ aos)
{
    const size_t d = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (d >= size_t(g.N))
        return;
    const size_t e = size_t(d / g.nx) * g.nxs + (d % g.nx);
    for (int c = 0; c < ncomp; ++c)
        aos[d * ncomp + c] = planes[size_t(c) * g.Npad + e];
}

//! VectorManipulations::CGAveragePeriodic (dynamics/src/include/VectorManipulations.hpp:26-65): for every periodic entry
//! {type, c1 = right/top element id, c2 = left/bottom element id, edge} and j in [jFirst, jFirst + jCount) the node pair
//! (i1 on c2's left/bottom edge, i2 on c1's right/top edge) becomes the mean of the two.  One thread per (entry, j).
template <int CG>
__global__ void cg_average_periodic_kernel(GridDims g, const long* __restrict__ list, long count, int jFirst, int jCount, double* __restrict__ v)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= count * jCount)
        return;
    const long i = t / jCount;
    const int j = jFirst + int(t % jCount);
    const long type = list[4 * i], rt = list[4 * i + 1], lb = list[4 * i + 2];
    const size_t n0lb = size_t(CG * (lb / g.nx)) * g.cgs + size_t(CG * (lb % g.nx));
    const size_t n0rt = size_t(CG * (rt / g.nx)) * g.cgs + size_t(CG * (rt % g.nx));
    size_t i1, i2;
    if (type == 0) { // X-edge: bottom line of the lower-boundary element, top line of the upper-boundary element
        i1 = n0lb + j;
        i2 = n0rt + size_t(CG) * g.cgs + j;
    } else {
        i1 = n0lb + size_t(j) * g.cgs;
        i2 = n0rt + CG + size_t(j) * g.cgs;
    }
    const double m = 0.5 * (v[i1] + v[i2]);
    v[i1] = m;
    v[i2] = m;
}

__global__ void clamp_kernel(size_t n, double* __restrict__ v, double lo, double hi)
{
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n)
        v[i] = fmin(fmax(v[i], lo), hi);
}

} // namespace nsdg
