/*
 * nsdg_momentum_param.cuh -- the mEVP subcycle on a PARAMETRIC (non-uniform, Cartesian) mesh WITHOUT streaming the
 * per-element operator matrices (CG2 / DG8 build).
 *
 * The reference stores, per element, iMgradX/Y (2 x 72), iMJwPSI (72) and divS1/2 (2 x 72) = 360 doubles and
 * streams them every subcycle (ParametricMap.hpp:72-113; 2.88 kB per element and subcycle).  All of them are
 * products of three ingredients (ParametricMap.cpp:225-296):
 *      the bilinear element map at the 9 Gauss points (dxT, dyT, J),
 *      constant reference tables (PHIx, PHIy, PSI, weights), and
 *      the inverse of the DG8 mass matrix  M = PSI diag(w J) PSI^T.
 * The only expensive one is M^-1.  But DG8 is the 9-dimensional Q2 tensor space minus ONE mode,
 * psi_8 = (x^2 - 1/12)(y^2 - 1/12), and there are exactly 9 Gauss points.  For D = diag(w_q J_q) the
 * D-weighted L2 projection onto DG8, seen in Gauss-point values, is therefore a rank-one correction of the identity:
 *
 *      PSI^T M^-1 PSI D y  =  y - rho * gamma * (n . y),      n_q = w_q psi_8(q),  rho_q = psi_8(q) / J_q,
 *                                                             gamma = 1 / sum_q n_q rho_q
 *
 * (n spans the null space of PSI because 3-point Gauss quadrature makes the 9 tensor modes discretely orthogonal;
 * the residual of a D-weighted projection is D^-1 n times a scalar fixed by n . (projection) = 0.)
 * Consequences, all exact up to rounding:
 *   projectVelocityToStrain : strain at the Gauss points = pointwise physical velocity gradient
 *                             (reference derivatives x inverse Jacobian), minus the rank-one term;
 *   stressUpdateHighOrder   : iMJwPSI r = B^ (r - rho gamma n.r), B^ the constant table of the uniform kernel;
 *   stressDivergence        : divS = (w J grad phi) PSI^T applied as "evaluate the stress at the Gauss points,
 *                             multiply by the element map, contract with the 1-d Q2 tables".
 * Per element the kernel reads 22 doubles of geometry (dxT, dyT per Gauss row/column: 12, 1/J: 9, gamma: 1)
 * instead of 360 doubles of operators; algorithmic traffic per element and subcycle 816 + 176 = 992 B
 * instead of 3840 B.  Node side, staging and deferred lines are those of nsdg_momentum_uniform.cuh.
 *
 * Same sweeps as the reference (CGDynamicsKernel.cpp:300-398, MEVPStressUpdateStep.hpp:30-118,
 * VPCGDynamicsKernel.hpp:132-172); parity: tests/test_gpu_vs_reference.py, tests/test_gpu_parity.py.
 */
#pragma once
#include "nsdg_momentum_uniform.cuh"
#include "nsdg_setup.cuh"

namespace nsdg {

constexpr int kGeoPlanes = 22; //!< xxi[3] yxi[3] (per Gauss row qy), xeta[3] yeta[3] (per Gauss column qx), 1/J[9], gamma

//! psi_8 = (x^2 - 1/12)(y^2 - 1/12), the Q2 tensor mode missing from DG8, at the 3 x 3 Gauss points
NSDG_HD constexpr double psi8at(int q)
{
    const double x = gausspoint(3, q % 3) - 0.5, y = gausspoint(3, q / 3) - 0.5;
    return (x * x - 1.0 / 12.0) * (y * y - 1.0 / 12.0);
}

/*
 * Per-element geometry planes of the factored operators (one thread per element, once per mesh).
 * dxT/dyT/J exactly as ParametricTools::dxT/dyT/J (ParametricTools.hpp:73-103) through elementMap.
 */
__global__ void paramgeom_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy, double* __restrict__ geo)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.nx) * g.ny)
        return;
    const int ix = int(t % g.nx), iy = int(t / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double c[4][2], dx[2][9], dy[2][9], J[9], lat[9];
    elementCorners(vx, vy, g.nx, ix, iy, false, c);
    elementMap<3>(c, dx, dy, J, lat);
    for (int k = 0; k < 3; ++k) {
        geo[size_t(0 + k) * g.Npad + e] = dx[0][3 * k]; // dx/dxi   depends on the Gauss row only
        geo[size_t(3 + k) * g.Npad + e] = dx[1][3 * k]; // dy/dxi
        geo[size_t(6 + k) * g.Npad + e] = dy[0][k]; //     dx/deta  depends on the Gauss column only
        geo[size_t(9 + k) * g.Npad + e] = dy[1][k]; //     dy/deta
    }
    double den = 0.0;
    for (int q = 0; q < 9; ++q) {
        const double iJ = 1.0 / J[q];
        geo[size_t(12 + q) * g.Npad + e] = iJ;
        den += gaussweight2(3, q) * psi8at(q) * psi8at(q) * iJ;
    }
    geo[size_t(21) * g.Npad + e] = 1.0 / den;
}

struct PmevpStage {
    double P[9][32];
    double S[24][32];
    double GEO[kGeoPlanes][32];
    double2 ND[2][7][32];
    double2 UV[2][2][32];
    double UVr[2][2];
    double pad[2];
};
constexpr int kPmevpWarps = 3;
constexpr size_t kPmevpSmemBytes = sizeof(PmevpStage) * kPmevpWarps;

#ifndef NSDG_PMEVP_MINBLOCKS
#define NSDG_PMEVP_MINBLOCKS 3
#endif
template <int DUMMY = 0>
__global__ void __launch_bounds__(32 * kPmevpWarps, NSDG_PMEVP_MINBLOCKS) subcycle_strip_pmevp(const __grid_constant__ UniformArgs a)
{
    constexpr int CG = 2, NR = 3, DGs = 8;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    PmevpStage& st = reinterpret_cast<PmevpStage*>(smemRaw)[threadIdx.x >> 5];
    const GridDims& g = a.g;
    const int sx = w % a.nsx, sy = w / a.nsx;
    const int exRaw = 32 * sx + lane;
    const bool active = exRaw < g.nx;
    const int ex = active ? exRaw : g.nx - 1;
    const bool lastLane = active && (lane == 31 || exRaw == g.nx - 1);
    const bool loadsRight = (lane == 31) || (exRaw >= g.nx - 1);
    const int ey0 = a.R * sy, ey1 = min(ey0 + a.R, g.ny);
    const size_t Npad = g.Npad;
    const int col0 = CG * ex;

    // ---- the five staging groups of element row `row`, issued in the order UV, P, S, GEO, ND ----
    auto issueUV = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + 1 + k) * g.cgs + col0;
                cpAsync16(&st.UV[0][k][lane], a.u + n);
                cpAsync16(&st.UV[1][k][lane], a.v + n);
                if (loadsRight) {
                    cpAsync8(&st.UVr[0][k], a.u + n + CG);
                    cpAsync8(&st.UVr[1][k], a.v + n + CG);
                }
            }
        }
        cpAsyncCommit();
    };
    auto issueP = [&](int row) {
        if (row < ey1) {
            const size_t en = size_t(row) * g.nxs + ex;
#pragma unroll
            for (int q = 0; q < 9; ++q)
                cpAsync8(&st.P[q][lane], a.Pa + size_t(q) * Npad + en);
        }
        cpAsyncCommit();
    };
    auto issueS = [&](int row) {
        if (row < ey1) {
            const size_t en = size_t(row) * g.nxs + ex;
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                cpAsync8(&st.S[j][lane], a.s11 + size_t(j) * Npad + en);
                cpAsync8(&st.S[8 + j][lane], a.s12 + size_t(j) * Npad + en);
                cpAsync8(&st.S[16 + j][lane], a.s22 + size_t(j) * Npad + en);
            }
        }
        cpAsyncCommit();
    };
    auto issueGEO = [&](int row) {
        if (row < ey1) {
            const size_t en = size_t(row) * g.nxs + ex;
#pragma unroll
            for (int k = 0; k < kGeoPlanes; ++k)
                cpAsync8(&st.GEO[k][lane], a.geo + size_t(k) * Npad + en);
        }
        cpAsyncCommit();
    };
    auto issueND = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + k) * g.cgs + col0;
                cpAsync16(&st.ND[k][0][lane], a.c1 + n);
                cpAsync16(&st.ND[k][1][lane], a.cA + n);
                cpAsync16(&st.ND[k][2][lane], a.rx + n);
                cpAsync16(&st.ND[k][3][lane], a.ry + n);
                cpAsync16(&st.ND[k][4][lane], a.uO + n);
                cpAsync16(&st.ND[k][5][lane], a.vO + n);
                cpAsync16(&st.ND[k][6][lane], a.ilm + n);
            }
        }
        cpAsyncCommit();
    };

    double carryX[2] = { 0.0, 0.0 }, carryY[2] = { 0.0, 0.0 };
    double ul[9], vl[9];
    issueUV(ey0);
    issueP(ey0);
    issueS(ey0);
    issueGEO(ey0);
    issueND(ey0);
    {
        const size_t n = size_t(CG * ey0) * g.cgs + col0;
        const double2 tu = *reinterpret_cast<const double2*>(a.u + n), tv = *reinterpret_cast<const double2*>(a.v + n);
        ul[0] = tu.x;
        ul[1] = tu.y;
        vl[0] = tv.x;
        vl[1] = tv.y;
        double ru = __shfl_down_sync(FULL, tu.x, 1), rv = __shfl_down_sync(FULL, tv.x, 1);
        if (loadsRight) {
            ru = a.u[n + CG];
            rv = a.v[n + CG];
        }
        ul[2] = ru;
        vl[2] = rv;
    }
    uint8_t lmNext = __ldg(a.landmask + size_t(ey0) * g.nxs + ex);
    uchar2 nmNext[2];
#pragma unroll
    for (int k = 0; k < 2; ++k)
        nmNext[k] = __ldg(reinterpret_cast<const uchar2*>(a.nodemask + size_t(CG * ey0 + k) * g.cgs + col0));

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
        const bool ice = active && (lmNext != 0);
        const uchar2 nm[2] = { nmNext[0], nmNext[1] };
        if (ey + 1 < ey1) {
            lmNext = __ldg(a.landmask + e + g.nxs);
#pragma unroll
            for (int k = 0; k < 2; ++k)
                nmNext[k] = __ldg(reinterpret_cast<const uchar2*>(a.nodemask + size_t(CG * (ey + 1) + k) * g.cgs + col0));
        }
        // ---- the two upper node rows of u, v ----
        cpAsyncWait<4>();
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const double2 tu = st.UV[0][k][lane], tv = st.UV[1][k][lane];
            ul[3 * (k + 1)] = tu.x;
            ul[3 * (k + 1) + 1] = tu.y;
            vl[3 * (k + 1)] = tv.x;
            vl[3 * (k + 1) + 1] = tv.y;
            double ru = __shfl_down_sync(FULL, tu.x, 1), rv = __shfl_down_sync(FULL, tv.x, 1);
            if (loadsRight) {
                ru = st.UVr[0][k];
                rv = st.UVr[1][k];
            }
            ul[3 * (k + 1) + 2] = ru;
            vl[3 * (k + 1) + 2] = rv;
        }
        issueUV(ey + 1);

        // ---- pointwise physical velocity gradient in the 9 Gauss points ----
        cpAsyncWait<2>(); // P, S and GEO of this row have landed
        const double gam = st.GEO[21][lane];
        double e11[9], e12[9], e22[9];
        {
            double Au[3][3], Adu[3][3], Av[3][3], Adv[3][3]; // [jy][qx]: value / xi-derivative contracted in x
            static_for<3>([&](auto JY) {
                static_for<3>([&](auto QX) {
                    constexpr int jy = decltype(JY)::value, qx = decltype(QX)::value;
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    static_for<3>([&](auto JX) {
                        constexpr int jx = decltype(JX)::value;
                        constexpr double l = kUnitOps.L[jx][qx], lp = kUnitOps.Lp[jx][qx];
                        if constexpr (l != 0.0) {
                            s0 = fma(l, ul[jy * 3 + jx], s0);
                            s2 = fma(l, vl[jy * 3 + jx], s2);
                        }
                        if constexpr (lp != 0.0) {
                            s1 = fma(lp, ul[jy * 3 + jx], s1);
                            s3 = fma(lp, vl[jy * 3 + jx], s3);
                        }
                    });
                    Au[jy][qx] = s0;
                    Adu[jy][qx] = s1;
                    Av[jy][qx] = s2;
                    Adv[jy][qx] = s3;
                });
            });
            double l11 = 0.0, l12 = 0.0, l22 = 0.0; // n . y
            static_for<3>([&](auto QY) {
                constexpr int qy = decltype(QY)::value;
                const double xxi = st.GEO[0 + qy][lane], yxi = st.GEO[3 + qy][lane];
                static_for<3>([&](auto QX) {
                    constexpr int qx = decltype(QX)::value, q = qy * 3 + qx;
                    double uxi = 0, ueta = 0, vxi = 0, veta = 0; // reference derivatives
                    static_for<3>([&](auto JY) {
                        constexpr int jy = decltype(JY)::value;
                        constexpr double l = kUnitOps.L[jy][qy], lp = kUnitOps.Lp[jy][qy];
                        if constexpr (l != 0.0) {
                            uxi = fma(l, Adu[jy][qx], uxi);
                            vxi = fma(l, Adv[jy][qx], vxi);
                        }
                        if constexpr (lp != 0.0) {
                            ueta = fma(lp, Au[jy][qx], ueta);
                            veta = fma(lp, Av[jy][qx], veta);
                        }
                    });
                    const double xeta = st.GEO[6 + qx][lane], yeta = st.GEO[9 + qx][lane], iJ = st.GEO[12 + q][lane];
                    // J d/dx = yeta d/dxi - yxi d/deta ;  J d/dy = xxi d/deta - xeta d/dxi   (ParametricMap.cpp:248-254)
                    const double ux = (yeta * uxi - yxi * ueta) * iJ, uy = (xxi * ueta - xeta * uxi) * iJ;
                    const double vx = (yeta * vxi - yxi * veta) * iJ, vy = (xxi * veta - xeta * vxi) * iJ;
                    e11[q] = ux;
                    e22[q] = vy;
                    e12[q] = 0.5 * (uy + vx);
                    constexpr double n = gaussweight2(3, q) * psi8at(q);
                    l11 = fma(n, e11[q], l11);
                    l22 = fma(n, e22[q], l22);
                    l12 = fma(n, e12[q], l12);
                });
            });
            // remove the psi_8 component (rank-one part of the DG8 projection); land elements keep zero strain (Q8)
            l11 *= gam;
            l12 *= gam;
            l22 *= gam;
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                const double rho = psi8at(q) * st.GEO[12 + q][lane];
                e11[q] = ice ? fma(-l11, rho, e11[q]) : 0.0;
                e12[q] = ice ? fma(-l12, rho, e12[q]) : 0.0;
                e22[q] = ice ? fma(-l22, rho, e22[q]) : 0.0;
            });
        }

        // ---- VP law in the Gauss points (MEVPStressUpdateStep.hpp:62-117), then the same rank-one projection ----
        {
            double l11 = 0.0, l12 = 0.0, l22 = 0.0;
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                const double Pa = st.P[q][lane];
                const double g11 = e11[q], g12 = e12[q], g22 = e22[q];
                const double iD = rsqrt(a.DeltaMin2 + 1.25 * (g11 * g11 + g22 * g22) + 1.50 * g11 * g22 + g12 * g12);
                const double pd = 0.125 * Pa * iD;
                e11[q] = fma(pd, 5.0 * g11 + 3.0 * g22, -0.5 * Pa);
                e22[q] = fma(pd, 5.0 * g22 + 3.0 * g11, -0.5 * Pa);
                e12[q] = 2.0 * pd * g12;
                constexpr double n = gaussweight2(3, q) * psi8at(q);
                l11 = fma(n, e11[q], l11);
                l22 = fma(n, e22[q], l22);
                l12 = fma(n, e12[q], l12);
            });
            l11 *= gam;
            l12 *= gam;
            l22 *= gam;
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                const double rho = psi8at(q) * st.GEO[12 + q][lane];
                e11[q] = fma(-l11, rho, e11[q]);
                e12[q] = fma(-l12, rho, e12[q]);
                e22[q] = fma(-l22, rho, e22[q]);
            });
        }
        issueP(ey + 1);

        // ---- per stress component: coefficients of the projected increment, relax, store, value at the Gauss points ----
        auto component = [&](double* plane, double (&r)[9], auto COMP) {
            constexpr int comp = decltype(COMP)::value; // 0: s11, 1: s12, 2: s22
            double s[DGs];
            static_for<DGs>([&](auto J) {
                constexpr int j = decltype(J)::value;
                double acc = 0.0;
                static_for<9>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    constexpr double b = kUnitOps.B[j][q];
                    if constexpr (b != 0.0)
                        acc = fma(b, r[q], acc);
                });
                s[j] = fma(st.S[comp * 8 + j][lane], a.keep, acc);
                if (active)
                    plane[size_t(j) * Npad + e] = s[j];
            });
            // the new stress in the Gauss points (overwrites the increment)
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                double acc = 0.0;
                static_for<DGs>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    constexpr double p = UnitOps::clean(PSI(3, j, q));
                    if constexpr (p != 0.0)
                        acc = fma(p, s[j], acc);
                });
                r[q] = acc;
            });
        };
        component(a.s11, e11, std::integral_constant<int, 0> {});
        component(a.s12, e12, std::integral_constant<int, 1> {});
        component(a.s22, e22, std::integral_constant<int, 2> {});
        issueS(ey + 1);

        // ---- stress divergence: (w J grad phi_i, sigma) by contraction with the 1-d Q2 tables ----
        double Tx[9], Ty[9];
        if (ice) {
            double CAx[3][3], CBx[3][3], CAy[3][3], CBy[3][3]; // [ix][qy]
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    CAx[i][k] = CBx[i][k] = CAy[i][k] = CBy[i][k] = 0.0;
            static_for<3>([&](auto QY) {
                constexpr int qy = decltype(QY)::value;
                const double xxi = st.GEO[0 + qy][lane], yxi = st.GEO[3 + qy][lane];
                static_for<3>([&](auto QX) {
                    constexpr int qx = decltype(QX)::value, q = qy * 3 + qx;
                    constexpr double wq = gaussweight2(3, q);
                    const double xeta = st.GEO[6 + qx][lane], yeta = st.GEO[9 + qx][lane];
                    // tx_i = sum_q phi_i,xi (w (yeta s11 - xeta s12)) + phi_i,eta (w (xxi s12 - yxi s11)),  ty alike
                    const double ax = wq * (yeta * e11[q] - xeta * e12[q]), bx = wq * (xxi * e12[q] - yxi * e11[q]);
                    const double ay = wq * (yeta * e12[q] - xeta * e22[q]), by = wq * (xxi * e22[q] - yxi * e12[q]);
                    static_for<3>([&](auto IX) {
                        constexpr int ix = decltype(IX)::value;
                        constexpr double l = kUnitOps.L[ix][qx], lp = kUnitOps.Lp[ix][qx];
                        if constexpr (lp != 0.0) {
                            CAx[ix][qy] = fma(lp, ax, CAx[ix][qy]);
                            CAy[ix][qy] = fma(lp, ay, CAy[ix][qy]);
                        }
                        if constexpr (l != 0.0) {
                            CBx[ix][qy] = fma(l, bx, CBx[ix][qy]);
                            CBy[ix][qy] = fma(l, by, CBy[ix][qy]);
                        }
                    });
                });
            });
            static_for<3>([&](auto IY) {
                static_for<3>([&](auto IX) {
                    constexpr int iy = decltype(IY)::value, ix = decltype(IX)::value;
                    double tx = 0.0, ty = 0.0;
                    static_for<3>([&](auto QY) {
                        constexpr int qy = decltype(QY)::value;
                        constexpr double l = kUnitOps.L[iy][qy], lp = kUnitOps.Lp[iy][qy];
                        if constexpr (l != 0.0) {
                            tx = fma(l, CAx[ix][qy], tx);
                            ty = fma(l, CAy[ix][qy], ty);
                        }
                        if constexpr (lp != 0.0) {
                            tx = fma(lp, CBx[ix][qy], tx);
                            ty = fma(lp, CBy[ix][qy], ty);
                        }
                    });
                    Tx[iy * 3 + ix] = tx;
                    Ty[iy * 3 + ix] = ty;
                });
            });
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                Tx[k] = Ty[k] = 0.0;
        }
        issueGEO(ey + 1);

        // ---- raw contributions to the deferred lines ----
        if (active && lane == 0 && sx > 0) {
            double* vb = a.vbuf + ((size_t(sx - 1) * 2 + 1) * g.ny + ey) * (NR * 2);
#pragma unroll
            for (int jy = 0; jy < NR; ++jy) {
                vb[jy * 2 + 0] = Tx[jy * NR];
                vb[jy * 2 + 1] = Ty[jy * NR];
            }
        }
        if (lastLane) {
            double* vb = a.vbuf + ((size_t(sx) * 2 + 0) * g.ny + ey) * (NR * 2);
#pragma unroll
            for (int jy = 0; jy < NR; ++jy) {
                vb[jy * 2 + 0] = Tx[jy * NR + CG];
                vb[jy * 2 + 1] = Ty[jy * NR + CG];
            }
        }
        const bool bottomDeferred = (ey == ey0) && (sy > 0);
        if (active && bottomDeferred) {
            double* hb = a.hbuf + ((size_t(sy - 1) * 2 + 1) * g.nx + ex) * (NR * 2);
#pragma unroll
            for (int jx = 0; jx < NR; ++jx) {
                hb[jx * 2 + 0] = Tx[jx];
                hb[jx * 2 + 1] = Ty[jx];
            }
        }
        if (active && ey == ey1 - 1) {
            double* hb = a.hbuf + ((size_t(sy) * 2 + 0) * g.nx + ex) * (NR * 2);
#pragma unroll
            for (int jx = 0; jx < NR; ++jx) {
                hb[jx * 2 + 0] = Tx[CG * NR + jx];
                hb[jx * 2 + 1] = Ty[CG * NR + jx];
            }
        }
        // ---- left neighbour's right column by shuffle ----
#pragma unroll
        for (int jy = 0; jy < NR; ++jy) {
            const double lx = __shfl_up_sync(FULL, Tx[jy * NR + CG], 1);
            const double ly = __shfl_up_sync(FULL, Ty[jy * NR + CG], 1);
            if (lane > 0) {
                Tx[jy * NR] = lx + Tx[jy * NR];
                Ty[jy * NR] = ly + Ty[jy * NR];
            }
        }
        // ---- momentum update of the completed nodes (rows 2ey, 2ey+1; columns 2ex, 2ex+1) ----
        cpAsyncWait<4>();
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
            const double2 c1 = st.ND[jy][0][lane], cA = st.ND[jy][1][lane], rx = st.ND[jy][2][lane], ry = st.ND[jy][3][lane];
            const double2 uO = st.ND[jy][4][lane], vO = st.ND[jy][5][lane], ilm = st.ND[jy][6][lane];
            const uchar2 msk = nm[jy];
            double sx0 = Tx[jy * NR], sy0 = Ty[jy * NR], sx1 = Tx[jy * NR + 1], sy1 = Ty[jy * NR + 1];
            if (jy == 0) {
                sx0 += carryX[0];
                sy0 += carryY[0];
                sx1 += carryX[1];
                sy1 += carryY[1];
            }
            const bool d0 = msk.x & 1, d1 = msk.y & 1;
            double2 un, vn;
            momentumNodeUniform(a, c1.x, cA.x, rx.x, ry.x, uO.x, vO.x, ilm.x, d0, ul[jy * NR], vl[jy * NR], d0 ? 0.0 : -sx0,
                d0 ? 0.0 : -sy0, un.x, vn.x);
            momentumNodeUniform(a, c1.y, cA.y, rx.y, ry.y, uO.y, vO.y, ilm.y, d1, ul[jy * NR + 1], vl[jy * NR + 1],
                d1 ? 0.0 : -sx1, d1 ? 0.0 : -sy1, un.y, vn.y);
            const bool rowSkip = !active || (jy == 0 && bottomDeferred);
            const bool skip0 = rowSkip || (lane == 0 && sx > 0);
            if (!rowSkip) {
                if (!skip0) {
                    *reinterpret_cast<double2*>(a.u + n0) = un;
                    *reinterpret_cast<double2*>(a.v + n0) = vn;
                } else {
                    a.u[n0 + 1] = un.y;
                    a.v[n0 + 1] = vn.y;
                }
            }
        }
        issueND(ey + 1);
        carryX[0] = Tx[CG * NR];
        carryX[1] = Tx[CG * NR + 1];
        carryY[0] = Ty[CG * NR];
        carryY[1] = Ty[CG * NR + 1];
#pragma unroll
        for (int jx = 0; jx < NR; ++jx) {
            ul[jx] = ul[CG * NR + jx];
            vl[jx] = vl[CG * NR + jx];
        }
    }
    cpAsyncWait<0>();
}

} // namespace nsdg
