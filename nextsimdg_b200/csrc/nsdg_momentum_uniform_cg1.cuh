/*
 * nsdg_momentum_uniform_cg1.cuh -- the mEVP subcycle on a UNIFORM RECTANGULAR mesh for the reference's other compile-time
 * build, CG1 velocities / DG1 (3-component) stresses (CMakeLists.txt:112-118: -DDGCOMP=3 -DCGDEGREE=1; CG2DGSTRESS(1) = 3,
 * 2 x 2 Gauss points, NextsimDynamics.hpp:42-60).  Same organisation as the CG2 kernel of nsdg_momentum_uniform.cuh -- a warp
 * owns a strip of 32 elements x R rows, contributions to the node shared with the left neighbour travel by shuffle, those to
 * the upper node row are carried in registers, strip-boundary nodes go through the line buffers -- with what is special here:
 *
 *  - an element has 4 nodes and a lane advances ONE node per element row (its lower-left one);
 *  - grad(Q1 velocity) lies in the DG1 space {1, x, y}, so the L2 projection of projectVelocityToStrain followed by the
 *    Gauss-point evaluation of the stress update is the velocity gradient in the 2 x 2 Gauss points, taken directly;
 *  - the unit-square operators are a handful of numbers: iMJwPSI = [1/4 ; +-3g ; +-3g] (g = 1/(2 sqrt 3)),
 *    divS1 (c, d) = sgn_c [1/2, 0, -+1/12], divS2 (c, d) = sgn_d [1/2, -+1/12, 0]  (c, d = node column / row of the element);
 *  - 13 doubles of state per element are read and 9 written (3 x 3 stress coefficients, 4 x P / alpha), plus one node
 *    (6 constants, u, v): 256 B per element and subcycle against ~ 300 B and no staging in the generic kernel.
 *
 * The 3 stress fields and P travel by TMA (one tensor copy per field and row, nsdg_momentum_uniform.cuh), the node row by
 * per-lane cp.async.  The deferred-line nodes are advanced by subcycle_lines_umevp<1> (nsdg_momentum_uniform.cuh: the same
 * per-node functions as the CG2 path, instantiated for CG = 1), which reads the line buffers this kernel fills.
 *
 * Sweeps reproduced (results equal up to rounding; tests/test_gpu_parity.py, DG1/CG1 cases):
 *   projectVelocityToStrain  dynamics/src/CGDynamicsKernel.cpp:300-337
 *   stressUpdateHighOrder    dynamics/src/include/MEVPStressUpdateStep.hpp:30-118
 *   stressDivergence         dynamics/src/CGDynamicsKernel.cpp:340-398
 *   updateMomentum           dynamics/src/include/VPCGDynamicsKernel.hpp:132-172 (quirks Q1, Q2 kept)
 *   applyBoundaries          dynamics/src/CGDynamicsKernel.cpp:439-444
 */
#pragma once
#include "nsdg_momentum_uniform.cuh"

namespace nsdg {

struct Umevp1Stage {
    double P[4][32]; //!< P / alpha in the 2 x 2 Gauss points (q = 2 qy + qx)
    double S[9][32]; //!< s11[3], s12[3], s22[3]
    double ND[kNodeConsts][32]; //!< node constants of the lane's node in the row being completed
    double UV[2][32]; //!< u, v of the lane's node in the upper node row
    double UVr[2]; //!< ... and of the strip's right-most node in that row
    uint64_t bar[2]; //!< mbarriers of the P and S groups
    double pad[12];
};
static_assert(sizeof(Umevp1Stage) % 128 == 0 && offsetof(Umevp1Stage, S) % 128 == 0, "TMA destinations need 128-byte alignment");
constexpr int kUmevp1Warps = 4;
constexpr size_t kUmevp1SmemBytes = sizeof(Umevp1Stage) * kUmevp1Warps;

template <int DUMMY = 0> __global__ void __launch_bounds__(32 * kUmevp1Warps, 4) subcycle_strip_umevp1(const __grid_constant__ UniformArgs a)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr double g2 = 0.28867513459481288225457439025098; // 1 / (2 sqrt 3): Gauss points 1/2 -+ g2
    constexpr double w0 = 0.5 - g2, w1 = 0.5 + g2; //            values of the Q1 node functions there
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    Umevp1Stage& st = reinterpret_cast<Umevp1Stage*>(smemRaw)[threadIdx.x >> 5];
    const GridDims& g = a.g;
    const int sx = w % a.nsx, sy = w / a.nsx;
    const int exRaw = 32 * sx + lane;
    const bool active = exRaw < g.nx;
    const int ex = active ? exRaw : g.nx - 1;
    const bool lastLane = active && (lane == 31 || exRaw == g.nx - 1);
    const bool loadsRight = (lane == 31) || (exRaw >= g.nx - 1);
    const int ey0 = a.R * sy, ey1 = min(ey0 + a.R, g.ny);
    const size_t Npad = g.Npad;
    const double idx = 1.0 / a.dx, idy = 1.0 / a.dy;

    if (lane == 0) {
        mbarInit(&st.bar[0], 1);
        mbarInit(&st.bar[1], 1);
    }
    mbarInitFence();
    __syncwarp();
    unsigned phaseP = 0, phaseS = 0;

    // ---- staging: P and S of element row `row` by TMA; the upper node row (u, v) and the completed node row's constants by
    //      per-lane cp.async (every lane copies and reads its own slots only) ----
    auto issueP = [&](int row) {
        if (row < ey1) {
            __syncwarp(); // every lane has consumed the region that is refilled
            if (lane == 0) {
                mbarExpectTx(&st.bar[0], 4 * 256);
                tmaLoadTile(&st.P[0][0], &a.tm[3], row * g.nxs + 32 * sx, &st.bar[0]);
            }
        }
    };
    auto issueS = [&](int row) {
        if (row < ey1) {
            __syncwarp();
            if (lane == 0) {
                const int x = row * g.nxs + 32 * sx;
                mbarExpectTx(&st.bar[1], 9 * 256);
                tmaLoadTile(&st.S[0][0], &a.tm[0], x, &st.bar[1]);
                tmaLoadTile(&st.S[3][0], &a.tm[1], x, &st.bar[1]);
                tmaLoadTile(&st.S[6][0], &a.tm[2], x, &st.bar[1]);
            }
        }
    };
    auto issueUV = [&](int row) { // node row row + 1
        if (row < ey1) {
            const size_t n = size_t(row + 1) * g.cgs + ex;
            cpAsync8(&st.UV[0][lane], a.u + n);
            cpAsync8(&st.UV[1][lane], a.v + n);
            if (loadsRight) {
                cpAsync8(&st.UVr[0], a.u + n + 1);
                cpAsync8(&st.UVr[1], a.v + n + 1);
            }
        }
        cpAsyncCommit();
    };
    auto issueND = [&](int row) { // node row `row`
        if (row < ey1) {
            const size_t n = size_t(row) * g.cgs + ex;
            cpAsync8(&st.ND[0][lane], a.cA + n);
            cpAsync8(&st.ND[1][lane], a.rx + n);
            cpAsync8(&st.ND[2][lane], a.ry + n);
            cpAsync8(&st.ND[3][lane], a.uO + n);
            cpAsync8(&st.ND[4][lane], a.vO + n);
            cpAsync8(&st.ND[5][lane], a.ilm + n);
        }
        cpAsyncCommit();
    };

    issueUV(ey0);
    issueP(ey0);
    issueS(ey0);
    issueND(ey0);
    // the strip's bottom node row by plain loads; [0] = the lane's node, [1] = its right neighbour
    double ub[2], vb[2];
    {
        const size_t n = size_t(ey0) * g.cgs + ex;
        ub[0] = a.u[n];
        vb[0] = a.v[n];
        ub[1] = __shfl_down_sync(FULL, ub[0], 1);
        vb[1] = __shfl_down_sync(FULL, vb[0], 1);
        if (loadsRight) {
            ub[1] = a.u[n + 1];
            vb[1] = a.v[n + 1];
        }
    }
    double carryX = 0.0, carryY = 0.0;

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
        const bool ice = active && (a.landmask[e] != 0);
        const bool dirichlet = (a.nodemask[size_t(ey) * g.cgs + ex] & 1) != 0;

        // ---- the upper node row ----
        cpAsyncWait<1>(); // (groups in flight: UV of this row, ND of this row)
        double ut[2], vt[2];
        ut[0] = st.UV[0][lane];
        vt[0] = st.UV[1][lane];
        ut[1] = __shfl_down_sync(FULL, ut[0], 1);
        vt[1] = __shfl_down_sync(FULL, vt[0], 1);
        if (loadsRight) {
            ut[1] = st.UVr[0];
            vt[1] = st.UVr[1];
        }
        issueUV(ey + 1);

        // ---- velocity gradient in the 2 x 2 Gauss points (q = 2 qy + qx); zero on land (quirk Q8) ----
        double e11[4], e12[4], e22[4];
        {
            const double ix = ice ? idx : 0.0, iy = ice ? idy : 0.0;
            const double dxuB = ub[1] - ub[0], dxuT = ut[1] - ut[0], dxvB = vb[1] - vb[0], dxvT = vt[1] - vt[0]; // along x, bottom / top
            const double dyuL = ut[0] - ub[0], dyuR = ut[1] - ub[1], dyvL = vt[0] - vb[0], dyvR = vt[1] - vb[1]; // along y, left / right
#pragma unroll
            for (int qy = 0; qy < 2; ++qy)
#pragma unroll
                for (int qx = 0; qx < 2; ++qx) {
                    const double yb = qy == 0 ? w1 : w0, yt = qy == 0 ? w0 : w1; // weights of the bottom / top node row at y_q
                    const double xl = qx == 0 ? w1 : w0, xr = qx == 0 ? w0 : w1;
                    const double ux = fma(yb, dxuB, yt * dxuT), vx = fma(yb, dxvB, yt * dxvT);
                    const double uy = fma(xl, dyuL, xr * dyuR), vy = fma(xl, dyvL, xr * dyvR);
                    e11[2 * qy + qx] = ux * ix;
                    e22[2 * qy + qx] = vy * iy;
                    e12[2 * qy + qx] = 0.5 * fma(uy, iy, vx * ix);
                }
        }

        // ---- VP law in the Gauss points: e** become the integrands (MEVPStressUpdateStep.hpp:62-117) ----
        mbarWait(&st.bar[0], phaseP);
        phaseP ^= 1u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double Pa = st.P[q][lane];
            const double g11 = e11[q], g12 = e12[q], g22 = e22[q];
            const double iD = rsqrt(a.DeltaMin2 + 1.25 * (g11 * g11 + g22 * g22) + 1.50 * g11 * g22 + g12 * g12);
            const double pd = 0.125 * Pa * iD;
            e11[q] = fma(pd, 5.0 * g11 + 3.0 * g22, -0.5 * Pa);
            e22[q] = fma(pd, 5.0 * g22 + 3.0 * g11, -0.5 * Pa);
            e12[q] = 2.0 * pd * g12;
        }
        issueP(ey + 1);

        // ---- per stress component: s <- (1 - 1/alpha) s + iMJwPSI r, store, divergence contributions ----
        mbarWait(&st.bar[1], phaseS);
        phaseS ^= 1u;
        double Tx[4] = { 0.0, 0.0, 0.0, 0.0 }, Ty[4] = { 0.0, 0.0, 0.0, 0.0 }; // k = 2 jy + jx
        auto component = [&](double* plane, const double (&r)[4], auto COMP) {
            constexpr int comp = decltype(COMP)::value; // 0: s11, 1: s12, 2: s22
            // iMJwPSI = psi_j(q) w_q / int psi_j^2 = [1/4 ; 12 (x_q - 1/2) / 4 ; 12 (y_q - 1/2) / 4]
            const double s0 = fma(st.S[comp * 3 + 0][lane], a.keep, 0.25 * ((r[0] + r[1]) + (r[2] + r[3])));
            const double s1 = fma(st.S[comp * 3 + 1][lane], a.keep, (3.0 * g2) * ((r[1] - r[0]) + (r[3] - r[2])));
            const double s2 = fma(st.S[comp * 3 + 2][lane], a.keep, (3.0 * g2) * ((r[2] - r[0]) + (r[3] - r[1])));
            if (active) {
                plane[e] = s0;
                plane[Npad + e] = s1;
                plane[2 * Npad + e] = s2;
            }
            if (ice) {
                // divS1 s at node (c, d) = sgn_c (s0 / 2 -+ s2 / 12),  divS2 s = sgn_d (s0 / 2 -+ s1 / 12)
                const double h = 0.5 * s0, t2 = (1.0 / 12.0) * s2, t1 = (1.0 / 12.0) * s1;
                const double d1b = h - t2, d1t = h + t2; // bottom / top node row, x-derivative (sign by column)
                const double d2l = h - t1, d2r = h + t1; // left / right node column, y-derivative (sign by row)
                if constexpr (comp == 0) {
                    Tx[0] = fma(-d1b, a.dy, Tx[0]);
                    Tx[1] = fma(d1b, a.dy, Tx[1]);
                    Tx[2] = fma(-d1t, a.dy, Tx[2]);
                    Tx[3] = fma(d1t, a.dy, Tx[3]);
                }
                if constexpr (comp == 1) {
                    Tx[0] = fma(-d2l, a.dx, Tx[0]);
                    Tx[1] = fma(-d2r, a.dx, Tx[1]);
                    Tx[2] = fma(d2l, a.dx, Tx[2]);
                    Tx[3] = fma(d2r, a.dx, Tx[3]);
                    Ty[0] = fma(-d1b, a.dy, Ty[0]);
                    Ty[1] = fma(d1b, a.dy, Ty[1]);
                    Ty[2] = fma(-d1t, a.dy, Ty[2]);
                    Ty[3] = fma(d1t, a.dy, Ty[3]);
                }
                if constexpr (comp == 2) {
                    Ty[0] = fma(-d2l, a.dx, Ty[0]);
                    Ty[1] = fma(-d2r, a.dx, Ty[1]);
                    Ty[2] = fma(d2l, a.dx, Ty[2]);
                    Ty[3] = fma(d2r, a.dx, Ty[3]);
                }
            }
        };
        component(a.s11, e11, std::integral_constant<int, 0> {});
        component(a.s12, e12, std::integral_constant<int, 1> {});
        component(a.s22, e22, std::integral_constant<int, 2> {});
        issueS(ey + 1);

        // ---- raw contributions to the deferred lines (layout of the generic strip kernel, NR = 2) ----
        if (active && lane == 0 && sx > 0) {
            double* vbp = a.vbuf + ((size_t(sx - 1) * 2 + 1) * g.ny + ey) * 4;
            vbp[0] = Tx[0];
            vbp[1] = Ty[0];
            vbp[2] = Tx[2];
            vbp[3] = Ty[2];
        }
        if (lastLane) {
            double* vbp = a.vbuf + ((size_t(sx) * 2 + 0) * g.ny + ey) * 4;
            vbp[0] = Tx[1];
            vbp[1] = Ty[1];
            vbp[2] = Tx[3];
            vbp[3] = Ty[3];
        }
        const bool bottomDeferred = (ey == ey0) && (sy > 0);
        if (active && bottomDeferred) {
            double* hb = a.hbuf + ((size_t(sy - 1) * 2 + 1) * g.nx + ex) * 4;
            hb[0] = Tx[0];
            hb[1] = Ty[0];
            hb[2] = Tx[1];
            hb[3] = Ty[1];
        }
        if (active && ey == ey1 - 1) {
            double* hb = a.hbuf + ((size_t(sy) * 2 + 0) * g.nx + ex) * 4;
            hb[0] = Tx[2];
            hb[1] = Ty[2];
            hb[2] = Tx[3];
            hb[3] = Ty[3];
        }
        // ---- the left neighbour's right column by shuffle ----
#pragma unroll
        for (int jy = 0; jy < 2; ++jy) {
            const double lx = __shfl_up_sync(FULL, Tx[2 * jy + 1], 1);
            const double ly = __shfl_up_sync(FULL, Ty[2 * jy + 1], 1);
            if (lane > 0) {
                Tx[2 * jy] = lx + Tx[2 * jy];
                Ty[2 * jy] = ly + Ty[2 * jy];
            }
        }
        // ---- momentum update of the completed node (column ex, row ey) ----
        cpAsyncWait<1>(); // (in flight: ND of this row, UV of the next)
        {
            const double sumX = carryX + Tx[0], sumY = carryY + Ty[0];
            double un, vn;
            momentumNodeUniform(a, st.ND[0][lane], st.ND[1][lane], st.ND[2][lane], st.ND[3][lane], st.ND[4][lane], st.ND[5][lane], dirichlet,
                ub[0], vb[0], dirichlet ? 0.0 : -sumX, dirichlet ? 0.0 : -sumY, un, vn);
            const bool skip = !active || bottomDeferred || (lane == 0 && sx > 0);
            if (!skip) {
                const size_t n = size_t(ey) * g.cgs + ex;
                a.u[n] = un;
                a.v[n] = vn;
            }
        }
        issueND(ey + 1);
        carryX = Tx[2];
        carryY = Ty[2];
        ub[0] = ut[0];
        ub[1] = ut[1];
        vb[0] = vt[0];
        vb[1] = vt[1];
    }
    cpAsyncWait<0>();
}

} // namespace nsdg
