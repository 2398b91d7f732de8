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
 * SPHERICAL meshes (template flag SPH; ParametricMap.cpp:298-351): coordinates are (lon, lat) in radians, the mass
 * matrix is weighted with w J cos(lat) (SphericalTools::massMatrix), every derivative carries 1/R, the lat-derivative
 * a cos(lat), and the metric terms divM / iMM = (phi, J sin(lat) psi)/R couple u and v.  With D = diag(w J cos) the same
 * rank-one identity holds; the planes then hold 1/(J cos), cos, sin per Gauss point (40 planes):
 *      e11 = P[ (J d_lon u - J sin v) / (R J cos) ],   e22 = P[ J d_lat v / (R J) ],
 *      e12 = P[ ( J d_lon v + cos J d_lat u + J sin u ) / (2 R J cos) ],   stress increment = B^ P[ r / cos ].
 *
 * Same sweeps as the reference (CGDynamicsKernel.cpp:300-398, MEVPStressUpdateStep.hpp:30-118,
 * VPCGDynamicsKernel.hpp:132-172); parity: tests/test_gpu_vs_reference.py, tests/test_gpu_parity.py.
 */
#pragma once
#include "nsdg_momentum_uniform.cuh"
#include "nsdg_momentum_uniform_bbm.cuh"
#include "nsdg_setup.cuh"

namespace nsdg {

#ifndef NSDG_PARAM_TMA
#define NSDG_PARAM_TMA 1 //!< plane rows of the parametric kernels by TMA (one tensor copy per field and row) instead of cp.async per plane
#endif
constexpr bool kParamTma = NSDG_PARAM_TMA != 0;
#ifndef NSDG_PARAM_SEP
#define NSDG_PARAM_SEP 1 //!< the constant tables as 1-d passes (q2Values / q2Derivs, evalGaussSep, projectSep) in the parametric kernels
#endif

//! planes: xxi[3] yxi[3] (per Gauss row qy), xeta[3] yeta[3] (per Gauss column qx), 1/J[9] (spherical: 1/(J cos)), gamma,
//! spherical only: cos(lat)[9], sin(lat)[9]
constexpr int kGeoPlanes = 22, kGeoPlanesSph = 40;
constexpr int geoPlanes(bool sph) { return sph ? kGeoPlanesSph : kGeoPlanes; }

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
template <bool SPH>
__global__ void paramgeom_kernel(GridDims g, const double* __restrict__ vx, const double* __restrict__ vy, double* __restrict__ geo)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.nx) * g.ny)
        return;
    const int ix = int(t % g.nx), iy = int(t / g.nx);
    const size_t e = size_t(iy) * g.nxs + ix;
    double c[4][2], dx[2][9], dy[2][9], J[9], lat[9];
    elementCorners(vx, vy, g.nx, ix, iy, SPH, c);
    elementMap<3>(c, dx, dy, J, lat);
    for (int k = 0; k < 3; ++k) {
        geo[size_t(0 + k) * g.Npad + e] = dx[0][3 * k]; // dx/dxi   depends on the Gauss row only
        geo[size_t(3 + k) * g.Npad + e] = dx[1][3 * k]; // dy/dxi
        geo[size_t(6 + k) * g.Npad + e] = dy[0][k]; //     dx/deta  depends on the Gauss column only
        geo[size_t(9 + k) * g.Npad + e] = dy[1][k]; //     dy/deta
    }
    double den = 0.0;
    for (int q = 0; q < 9; ++q) {
        const double iJ = SPH ? 1.0 / (J[q] * cos(lat[q])) : 1.0 / J[q];
        geo[size_t(12 + q) * g.Npad + e] = iJ;
        den += gaussweight2(3, q) * psi8at(q) * psi8at(q) * iJ;
        if constexpr (SPH) {
            geo[size_t(22 + q) * g.Npad + e] = cos(lat[q]);
            geo[size_t(31 + q) * g.Npad + e] = sin(lat[q]);
        }
    }
    geo[size_t(21) * g.Npad + e] = 1.0 / den;
}

// =====================================================================================================================
// Gauss-point building blocks shared by the parametric mEVP and BBM kernels.  GEO(k) reads geometry plane k of the
// lane's element from the warp's staging buffer.
// =====================================================================================================================

//! n_q = w_q psi_8(q): the null vector of the DG8 Gauss-point evaluation matrix
NSDG_HD constexpr double null8(int q) { return gaussweight2(3, q) * psi8at(q); }
//! Q2 tensor modes missing from DG6: psi_6 = y (x^2 - 1/12), psi_7 = x (y^2 - 1/12), psi_8
NSDG_HD constexpr double missing6at(int k, int q) { return k < 2 ? PSI(3, 6 + k, q) : psi8at(q); }

//! D-weighted DG8 projection of three Gauss-point fields, in place:  y <- y - rho gamma (n . y)
template <class GEO> __device__ __forceinline__ void projectDG8x3(const GEO& geo, double (&a)[9], double (&b)[9], double (&c)[9])
{
    double la = 0.0, lb = 0.0, lc = 0.0;
    static_for<9>([&](auto QQ) {
        constexpr int q = decltype(QQ)::value;
        constexpr double n = null8(q);
        la = fma(n, a[q], la);
        lb = fma(n, b[q], lb);
        lc = fma(n, c[q], lc);
    });
    const double gam = geo(21);
    la *= gam;
    lb *= gam;
    lc *= gam;
    static_for<9>([&](auto QQ) {
        constexpr int q = decltype(QQ)::value;
        const double rho = psi8at(q) * geo(12 + q);
        a[q] = fma(-la, rho, a[q]);
        b[q] = fma(-lb, rho, b[q]);
        c[q] = fma(-lc, rho, c[q]);
    });
}

//! 1 / cos(lat) in Gauss point q of a spherical element = J / (J cos)
template <class GEO> __device__ __forceinline__ double invCos(const GEO& geo, int qx, int qy, int q)
{
    return geo(12 + q) * (geo(0 + qy) * geo(9 + qx) - geo(3 + qy) * geo(6 + qx));
}

/*
 * projectVelocityToStrain (CGDynamicsKernel.cpp:300-337) in Gauss-point form: pointwise physical gradient of the Q2
 * velocity, then the DG8 projection.  Land elements keep zero strain (quirk Q8).
 */
template <bool SPH, class GEO>
__device__ __forceinline__ void gaussStrain(const GEO& geo, const double (&ul)[9], const double (&vl)[9], bool ice, double (&e11)[9],
    double (&e12)[9], double (&e22)[9])
{
    constexpr double iR = 1.0 / EarthRadius;
    double Au[3][3], Adu[3][3], Av[3][3], Adv[3][3]; // [jy][qx]: value / xi-derivative contracted in x
#if NSDG_PARAM_SEP
    double Uxi[3][3], Ueta[3][3], Vxi[3][3], Veta[3][3]; // [qy][qx]: reference derivatives in the Gauss points
    [[maybe_unused]] double Uq[3][3], Vq[3][3];
#pragma unroll
    for (int jy = 0; jy < 3; ++jy) {
        q2Values(ul[3 * jy], ul[3 * jy + 1], ul[3 * jy + 2], Au[jy][0], Au[jy][1], Au[jy][2]);
        q2Derivs(ul[3 * jy], ul[3 * jy + 1], ul[3 * jy + 2], Adu[jy][0], Adu[jy][1], Adu[jy][2]);
        q2Values(vl[3 * jy], vl[3 * jy + 1], vl[3 * jy + 2], Av[jy][0], Av[jy][1], Av[jy][2]);
        q2Derivs(vl[3 * jy], vl[3 * jy + 1], vl[3 * jy + 2], Adv[jy][0], Adv[jy][1], Adv[jy][2]);
    }
#pragma unroll
    for (int qx = 0; qx < 3; ++qx) {
        q2Values(Adu[0][qx], Adu[1][qx], Adu[2][qx], Uxi[0][qx], Uxi[1][qx], Uxi[2][qx]);
        q2Derivs(Au[0][qx], Au[1][qx], Au[2][qx], Ueta[0][qx], Ueta[1][qx], Ueta[2][qx]);
        q2Values(Adv[0][qx], Adv[1][qx], Adv[2][qx], Vxi[0][qx], Vxi[1][qx], Vxi[2][qx]);
        q2Derivs(Av[0][qx], Av[1][qx], Av[2][qx], Veta[0][qx], Veta[1][qx], Veta[2][qx]);
        if constexpr (SPH) {
            q2Values(Au[0][qx], Au[1][qx], Au[2][qx], Uq[0][qx], Uq[1][qx], Uq[2][qx]);
            q2Values(Av[0][qx], Av[1][qx], Av[2][qx], Vq[0][qx], Vq[1][qx], Vq[2][qx]);
        }
    }
#else
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
#endif
    static_for<3>([&](auto QY) {
        constexpr int qy = decltype(QY)::value;
        const double xxi = geo(0 + qy), yxi = geo(3 + qy);
        static_for<3>([&](auto QX) {
            constexpr int qx = decltype(QX)::value, q = qy * 3 + qx;
#if NSDG_PARAM_SEP
            const double uxi = Uxi[qy][qx], ueta = Ueta[qy][qx], vxi = Vxi[qy][qx], veta = Veta[qy][qx];
#else
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
#endif
            const double xeta = geo(6 + qx), yeta = geo(9 + qx), iJ = geo(12 + q);
            // J d/dx = yeta d/dxi - yxi d/deta ;  J d/dy = xxi d/deta - xeta d/dxi   (ParametricMap.cpp:248-254)
            if constexpr (!SPH) {
                const double ux = (yeta * uxi - yxi * ueta) * iJ, uy = (xxi * ueta - xeta * uxi) * iJ;
                const double vx = (yeta * vxi - yxi * veta) * iJ, vy = (xxi * veta - xeta * vxi) * iJ;
                e11[q] = ux;
                e22[q] = vy;
                e12[q] = 0.5 * (uy + vx);
            } else {
                // values of u, v in the Gauss point for the metric terms (divM / iMM, ParametricMap.cpp:321-337)
#if NSDG_PARAM_SEP
                const double uq = Uq[qy][qx], vq = Vq[qy][qx];
#else
                double uq = 0, vq = 0;
                static_for<3>([&](auto JY) {
                    constexpr int jy = decltype(JY)::value;
                    constexpr double l = kUnitOps.L[jy][qy];
                    if constexpr (l != 0.0) {
                        uq = fma(l, Au[jy][qx], uq);
                        vq = fma(l, Av[jy][qx], vq);
                    }
                });
#endif
                const double cl = geo(22 + q), sl = geo(31 + q);
                const double Js = (xxi * yeta - yxi * xeta) * sl; // J sin(lat)
                const double k = iJ * iR; //                         1 / (R J cos)
                e11[q] = ((yeta * uxi - yxi * ueta) - Js * vq) * k;
                e22[q] = (cl * (xxi * veta - xeta * vxi)) * k;
                e12[q] = 0.5 * ((yeta * vxi - yxi * veta) + cl * (xxi * ueta - xeta * uxi) + Js * uq) * k;
            }
        });
    });
    projectDG8x3(geo, e11, e12, e22);
    if (!ice) {
#pragma unroll
        for (int q = 0; q < 9; ++q)
            e11[q] = e12[q] = e22[q] = 0.0;
    }
}

//! DG coefficients (first NC basis functions) of a field of the DG space given by its Gauss-point values
template <int NC> __device__ __forceinline__ void coeffFromGauss(const double (&r)[9], double (&s)[NC])
{
#if NSDG_PARAM_SEP
    projectSep<NC>(r, s);
    return;
#endif
    static_for<NC>([&](auto J) {
        constexpr int j = decltype(J)::value;
        double acc = 0.0;
        static_for<9>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            constexpr double b = kUnitOps.B[j][q];
            if constexpr (b != 0.0)
                acc = fma(b, r[q], acc);
        });
        s[j] = acc;
    });
}
template <int NC> __device__ __forceinline__ void gaussFromCoeff(const double (&s)[NC], double (&r)[9])
{
#if NSDG_PARAM_SEP
    evalGaussSep<NC>(s, r);
    return;
#endif
    static_for<9>([&](auto QQ) {
        constexpr int q = decltype(QQ)::value;
        double acc = 0.0;
        static_for<NC>([&](auto J) {
            constexpr int j = decltype(J)::value;
            constexpr double p = UnitOps::clean(PSI(3, j, q));
            if constexpr (p != 0.0)
                acc = fma(p, s[j], acc);
        });
        r[q] = acc;
    });
}

/*
 * stressDivergence / addStressTensorCell (CGDynamicsKernel.cpp:340-398) from the stress in the Gauss points:
 *   tx_i = sum_q phi_i,xi (w (yeta s11 - xeta s12)) + phi_i,eta (w (xxi s12 - yxi s11)) [+ metric terms],  ty alike.
 */
template <bool SPH, class GEO>
__device__ __forceinline__ void gaussDivergence(const GEO& geo, const double (&s11)[9], const double (&s12)[9], const double (&s22)[9],
    double (&Tx)[9], double (&Ty)[9])
{
    constexpr double iR = 1.0 / EarthRadius;
    double CAx[3][3], CBx[3][3], CAy[3][3], CBy[3][3]; // [ix][qy]
    double CMx[SPH ? 3 : 1][3], CMy[SPH ? 3 : 1][3]; // metric (value) terms, spherical only
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            CAx[i][k] = CBx[i][k] = CAy[i][k] = CBy[i][k] = 0.0;
            if constexpr (SPH)
                CMx[i][k] = CMy[i][k] = 0.0;
        }
    static_for<3>([&](auto QY) {
        constexpr int qy = decltype(QY)::value;
        const double xxi = geo(0 + qy), yxi = geo(3 + qy);
        static_for<3>([&](auto QX) {
            constexpr int qx = decltype(QX)::value, q = qy * 3 + qx;
            constexpr double wq = gaussweight2(3, q);
            const double xeta = geo(6 + qx), yeta = geo(9 + qx);
            double ax, bx, ay, by;
            [[maybe_unused]] double mx = 0.0, my = 0.0;
            if constexpr (!SPH) {
                ax = wq * (yeta * s11[q] - xeta * s12[q]), bx = wq * (xxi * s12[q] - yxi * s11[q]);
                ay = wq * (yeta * s12[q] - xeta * s22[q]), by = wq * (xxi * s22[q] - yxi * s12[q]);
            } else { // divS1 = dx_cg2 PSI^T / R, divS2 = cos dy_cg2 PSI^T / R, divM = (phi J sin w) PSI^T / R
                constexpr double wr = wq * iR;
                const double cl = geo(22 + q), sl = geo(31 + q);
                const double Js = (xxi * yeta - yxi * xeta) * sl;
                ax = wr * (yeta * s11[q] - cl * (xeta * s12[q])), bx = wr * (cl * (xxi * s12[q]) - yxi * s11[q]);
                ay = wr * (yeta * s12[q] - cl * (xeta * s22[q])), by = wr * (cl * (xxi * s22[q]) - yxi * s12[q]);
                mx = wr * Js * s12[q]; //  tx += divM s12
                my = -wr * Js * s11[q]; // ty -= divM s11   (CGDynamicsKernel.cpp:348-351)
            }
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
                    if constexpr (SPH) {
                        CMx[ix][qy] = fma(l, mx, CMx[ix][qy]);
                        CMy[ix][qy] = fma(l, my, CMy[ix][qy]);
                    }
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
                    if constexpr (SPH) {
                        tx = fma(l, CMx[ix][qy], tx);
                        ty = fma(l, CMy[ix][qy], ty);
                    }
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
}

//! the warp-strip bookkeeping after the element part: deferred-line buffers and the left-neighbour shuffle
struct StripPos {
    int lane, sx, sy, ex, ey, ey0, ey1;
    bool active, lastLane;
};
__device__ __forceinline__ bool stripScatter(const GridDims& g, const StripPos& s, double* hbuf, double* vbuf, double (&Tx)[9], double (&Ty)[9])
{
    constexpr int CG = 2, NR = 3;
    constexpr unsigned FULL = 0xffffffffu;
    if (s.active && s.lane == 0 && s.sx > 0) {
        double* vb = vbuf + ((size_t(s.sx - 1) * 2 + 1) * g.ny + s.ey) * (NR * 2);
#pragma unroll
        for (int jy = 0; jy < NR; ++jy) {
            vb[jy * 2 + 0] = Tx[jy * NR];
            vb[jy * 2 + 1] = Ty[jy * NR];
        }
    }
    if (s.lastLane) {
        double* vb = vbuf + ((size_t(s.sx) * 2 + 0) * g.ny + s.ey) * (NR * 2);
#pragma unroll
        for (int jy = 0; jy < NR; ++jy) {
            vb[jy * 2 + 0] = Tx[jy * NR + CG];
            vb[jy * 2 + 1] = Ty[jy * NR + CG];
        }
    }
    const bool bottomDeferred = (s.ey == s.ey0) && (s.sy > 0);
    if (s.active && bottomDeferred) {
        double* hb = hbuf + ((size_t(s.sy - 1) * 2 + 1) * g.nx + s.ex) * (NR * 2);
#pragma unroll
        for (int jx = 0; jx < NR; ++jx) {
            hb[jx * 2 + 0] = Tx[jx];
            hb[jx * 2 + 1] = Ty[jx];
        }
    }
    if (s.active && s.ey == s.ey1 - 1) {
        double* hb = hbuf + ((size_t(s.sy) * 2 + 0) * g.nx + s.ex) * (NR * 2);
#pragma unroll
        for (int jx = 0; jx < NR; ++jx) {
            hb[jx * 2 + 0] = Tx[CG * NR + jx];
            hb[jx * 2 + 1] = Ty[CG * NR + jx];
        }
    }
#pragma unroll
    for (int jy = 0; jy < NR; ++jy) {
        const double lx = __shfl_up_sync(FULL, Tx[jy * NR + CG], 1);
        const double ly = __shfl_up_sync(FULL, Ty[jy * NR + CG], 1);
        if (s.lane > 0) {
            Tx[jy * NR] = lx + Tx[jy * NR];
            Ty[jy * NR] = ly + Ty[jy * NR];
        }
    }
    return bottomDeferred;
}

// =====================================================================================================================
// mEVP
// =====================================================================================================================
#ifndef NSDG_PMEVP_DIRECT_ND
#define NSDG_PMEVP_DIRECT_ND 0 //!< see NSDG_UMEVP_DIRECT_ND
#endif
#ifndef NSDG_PBBM_DIRECT_ND
#define NSDG_PBBM_DIRECT_ND 1 //!< Cartesian BBM kernel only: 4 instead of 3 blocks per SM fit (1.70 -> 1.39 ms at 2048^2); spherical loses
#endif
//! node-constant staging rows of a stage struct, absent when the constants are loaded directly
template <bool DIRECT> struct NodeStage {
    double2 ND[2][kNodeConsts][32];
};
template <> struct NodeStage<true> { };
template <bool SPH> constexpr bool kPbbmDirectND = !SPH && (NSDG_PBBM_DIRECT_ND != 0);
template <bool SPH> struct PmevpStage {
    double P[9][32];
    double S[24][32];
    double GEO[geoPlanes(SPH)][32];
#if !NSDG_PMEVP_DIRECT_ND
    double2 ND[2][kNodeConsts][32];
#endif
    double2 UV[2][2][32];
    MaskStage M;
    double UVr[2][2];
    uint64_t bar[3]; //!< mbarriers of the P, S and GEO groups (TMA staging)
    double pad[5];
};
static_assert(sizeof(PmevpStage<false>) % 128 == 0 && sizeof(PmevpStage<true>) % 128 == 0 && offsetof(PmevpStage<false>, S) % 128 == 0
        && offsetof(PmevpStage<false>, GEO) % 128 == 0 && offsetof(PmevpStage<true>, GEO) % 128 == 0,
    "TMA destinations need 128-byte alignment");
#ifndef NSDG_COOP_PARAM_CART
#define NSDG_COOP_PARAM_CART 0
#endif
//! staging mode of the plane rows: cooperative cp.async.cg on spherical meshes, per lane on Cartesian ones
template <bool SPH> constexpr bool kCoopPmevp = SPH || (NSDG_COOP_PARAM_CART != 0);
template <bool SPH> constexpr bool kCoopPbbm = SPH || (NSDG_COOP_PARAM_CART != 0);
#ifndef NSDG_PMEVP_WARPS
#define NSDG_PMEVP_WARPS 2 // 8 warps/SM with up to 255 registers beat 9 warps at 168 (0.91 -> 0.80 ms at 2048^2)
#define NSDG_PMEVP_MINB 4
#endif
constexpr int pmevpWarps(bool sph) { return sph ? 2 : NSDG_PMEVP_WARPS; }
constexpr size_t pmevpSmemBytes(bool sph) { return sph ? sizeof(PmevpStage<true>) * 2 : sizeof(PmevpStage<false>) * NSDG_PMEVP_WARPS; }

template <bool SPH>
__global__ void __launch_bounds__(32 * pmevpWarps(SPH), SPH ? 4 : NSDG_PMEVP_MINB) subcycle_strip_pmevp(const __grid_constant__ UniformArgs a)
{
    constexpr int CG = 2, NR = 3, DGs = 8;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    PmevpStage<SPH>& st = reinterpret_cast<PmevpStage<SPH>*>(smemRaw)[threadIdx.x >> 5];
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
    auto geo = [&](int k) { return st.GEO[k][lane]; };

    // ---- the five staging groups of element row `row`, issued in the order UV, P, S, GEO, ND ----
    auto issueUV = [&](int row) {
        if (row < ey1) {
            stageMasks(st.M, a.landmask, a.nodemask, g, row, sx, lane);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + 1 + k) * g.cgs + col0;
                cpAsync16cg(&st.UV[0][k][lane], a.u + n);
                cpAsync16cg(&st.UV[1][k][lane], a.v + n);
                if (loadsRight) {
                    cpAsync8(&st.UVr[0][k], a.u + n + CG);
                    cpAsync8(&st.UVr[1][k], a.v + n + CG);
                }
            }
        }
        cpAsyncCommit();
    };
    unsigned phase = 0; // TMA staging: the three groups are waited for together, once per row
    if constexpr (kParamTma) {
        if (lane == 0)
            for (int i = 0; i < 3; ++i)
                mbarInit(&st.bar[i], 1);
        mbarInitFence();
        __syncwarp();
    }
    //! TMA: lane 0 posts the group's bytes and issues one tensor copy per field (after every lane has consumed the region)
    auto tmaGroup = [&](int row, int bar, unsigned bytes, auto&& copies) {
        if (row < ey1) {
            __syncwarp();
            if (lane == 0) {
                mbarExpectTx(&st.bar[bar], bytes);
                copies(row * g.nxs + 32 * sx, &st.bar[bar]);
            }
        }
    };
    auto issueP = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 0, 9 * 256, [&](int x, uint64_t* b) { tmaLoadTile(&st.P[0][0], &a.tm[3], x, b); });
            return;
        }
        stageBarrier<kCoopPmevp<SPH>>(); // every lane has consumed the region that is refilled
        if (row < ey1)
            stagePlanes<9, kCoopPmevp<SPH>>(st.P, a.Pa, Npad, size_t(row) * g.nxs + 32 * sx, lane);
        cpAsyncCommit();
    };
    auto issueS = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 1, 24 * 256, [&](int x, uint64_t* b) {
                tmaLoadTile(&st.S[0][0], &a.tm[0], x, b);
                tmaLoadTile(&st.S[8][0], &a.tm[1], x, b);
                tmaLoadTile(&st.S[16][0], &a.tm[2], x, b);
            });
            return;
        }
        stageBarrier<kCoopPmevp<SPH>>();
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<8, kCoopPmevp<SPH>>(st.S, a.s11, Npad, first, lane);
            stagePlanes<8, kCoopPmevp<SPH>>(st.S + 8, a.s12, Npad, first, lane);
            stagePlanes<8, kCoopPmevp<SPH>>(st.S + 16, a.s22, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueGEO = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 2, geoPlanes(SPH) * 256, [&](int x, uint64_t* b) { tmaLoadTile(&st.GEO[0][0], &a.tm[4], x, b); });
            return;
        }
        stageBarrier<kCoopPmevp<SPH>>();
        if (row < ey1)
            stagePlanes<geoPlanes(SPH), kCoopPmevp<SPH>>(st.GEO, a.geo, Npad, size_t(row) * g.nxs + 32 * sx, lane);
        cpAsyncCommit();
    };
    auto issueND = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + k) * g.cgs + col0;
#if NSDG_PMEVP_DIRECT_ND
                for (const double* p : { a.cA, a.rx, a.ry, a.uO, a.vO, a.ilm })
                    prefetchL2(p + n);
#else
                cpAsync16cg(&st.ND[k][0][lane], a.cA + n);
                cpAsync16cg(&st.ND[k][1][lane], a.rx + n);
                cpAsync16cg(&st.ND[k][2][lane], a.ry + n);
                cpAsync16cg(&st.ND[k][3][lane], a.uO + n);
                cpAsync16cg(&st.ND[k][4][lane], a.vO + n);
                cpAsync16cg(&st.ND[k][5][lane], a.ilm + n);
#endif
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

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
        // ---- the two upper node rows of u, v ----
        cpAsyncWait<kParamTma ? 1 : 4>(); // (TMA staging: only the UV and ND groups are cp.async groups)
        __syncwarp(); // the mask bytes were staged by other lanes
        const bool ice = active && (st.M.LM[lane] != 0);
        const unsigned nm[2] = { nodeMaskWord(st.M, 0, lane), nodeMaskWord(st.M, 1, lane) };
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
        __syncwarp(); // every lane has read its mask bytes
        issueUV(ey + 1);

        // ---- strain in the 9 Gauss points ----
        if constexpr (kParamTma) {
            for (int i = 0; i < 3; ++i)
                mbarWait(&st.bar[i], phase);
            phase ^= 1u;
        } else {
            cpAsyncWait<2>(); // P, S and GEO of this row have landed
            stageBarrier<kCoopPmevp<SPH>>(); //     (staged cooperatively: visible to every lane after the warp barrier)
        }
        double e11[9], e12[9], e22[9];
        gaussStrain<SPH>(geo, ul, vl, ice, e11, e12, e22);

        // ---- VP law in the Gauss points (MEVPStressUpdateStep.hpp:62-117), then the same projection ----
        static_for<9>([&](auto QQ) {
            constexpr int q = decltype(QQ)::value;
            const double Pa = st.P[q][lane];
            const double g11 = e11[q], g12 = e12[q], g22 = e22[q];
#if NSDG_MEVP_SQRT & 2
            const double iD = rsqrtBranchFree(a.DeltaMin2 + 1.25 * (g11 * g11 + g22 * g22) + 1.50 * g11 * g22 + g12 * g12);
#else
            const double iD = rsqrt(a.DeltaMin2 + 1.25 * (g11 * g11 + g22 * g22) + 1.50 * g11 * g22 + g12 * g12);
#endif
            const double pd = 0.125 * Pa * iD;
            e11[q] = fma(pd, 5.0 * g11 + 3.0 * g22, -0.5 * Pa);
            e22[q] = fma(pd, 5.0 * g22 + 3.0 * g11, -0.5 * Pa);
            e12[q] = 2.0 * pd * g12;
            if constexpr (SPH) { // iMJwPSI weights with w J, the mass with w J cos (ParametricMap.cpp:339-343)
                const double icos = invCos(geo, q % 3, q / 3, q);
                e11[q] *= icos;
                e22[q] *= icos;
                e12[q] *= icos;
            }
        });
        projectDG8x3(geo, e11, e12, e22);
        issueP(ey + 1);

        // ---- per stress component: coefficients of the projected increment, relax, store, value at the Gauss points ----
        auto component = [&](double* plane, double (&r)[9], auto COMP) {
            constexpr int comp = decltype(COMP)::value; // 0: s11, 1: s12, 2: s22
            double s[DGs];
            coeffFromGauss<DGs>(r, s);
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                s[j] = fma(st.S[comp * 8 + j][lane], a.keep, s[j]);
                if (active)
                    plane[size_t(j) * Npad + e] = s[j];
            }
            gaussFromCoeff<DGs>(s, r); // the new stress in the Gauss points (overwrites the increment)
        };
        component(a.s11, e11, std::integral_constant<int, 0> {});
        component(a.s12, e12, std::integral_constant<int, 1> {});
        component(a.s22, e22, std::integral_constant<int, 2> {});
        issueS(ey + 1);

        // ---- stress divergence ----
        double Tx[9], Ty[9];
        if (ice) {
            gaussDivergence<SPH>(geo, e11, e12, e22, Tx, Ty);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                Tx[k] = Ty[k] = 0.0;
        }
        issueGEO(ey + 1);

        const bool bottomDeferred = stripScatter(g, StripPos { lane, sx, sy, ex, ey, ey0, ey1, active, lastLane }, a.hbuf, a.vbuf, Tx, Ty);
        // ---- momentum update of the completed nodes (rows 2ey, 2ey+1; columns 2ex, 2ex+1) ----
        cpAsyncWait<kParamTma ? 1 : 4>();
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
#if NSDG_PMEVP_DIRECT_ND
            auto ld2 = [&](const double* p) { return __ldg(reinterpret_cast<const double2*>(p + n0)); };
            const double2 cA = ld2(a.cA), rx = ld2(a.rx), ry = ld2(a.ry), uO = ld2(a.uO), vO = ld2(a.vO), ilm = ld2(a.ilm);
#else
            const double2 cA = st.ND[jy][0][lane], rx = st.ND[jy][1][lane], ry = st.ND[jy][2][lane];
            const double2 uO = st.ND[jy][3][lane], vO = st.ND[jy][4][lane], ilm = st.ND[jy][5][lane];
#endif
            const unsigned msk = nm[jy];
            double sx0 = Tx[jy * NR], sy0 = Ty[jy * NR], sx1 = Tx[jy * NR + 1], sy1 = Ty[jy * NR + 1];
            if (jy == 0) {
                sx0 += carryX[0];
                sy0 += carryY[0];
                sx1 += carryX[1];
                sy1 += carryY[1];
            }
            const bool d0 = msk & 1u, d1 = msk & 0x100u;
            double2 un, vn;
            momentumNodeUniform(a, cA.x, rx.x, ry.x, uO.x, vO.x, ilm.x, d0, ul[jy * NR], vl[jy * NR], d0 ? 0.0 : -sx0,
                d0 ? 0.0 : -sy0, un.x, vn.x);
            momentumNodeUniform(a, cA.y, rx.y, ry.y, uO.y, vO.y, ilm.y, d1, ul[jy * NR + 1], vl[jy * NR + 1],
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

// =====================================================================================================================
// BBM
// =====================================================================================================================
/*
 * Extra planes of the BBM kernel, after the geoPlanes(SPH) common ones:
 *   +0..5  G3 = (N^T D^-1 N)^-1 (symmetric 3 x 3: 00 01 02 11 12 22), N = w psi_{6,7,8}: the DG6 projection of the
 *          damage (iMJwPSI_dam, ParametricMap.cpp:289-296) is  y - D^-1 N G3 N^T y  in Gauss-point values
 *   +6     scale_coef = sqrt(0.1 / h_el)                                   (BBMStressUpdateStep.hpp:134)
 *   +7     1 / (h_el sqrt(2 (1 + nu) rho_ice))                             (BBMStressUpdateStep.hpp:158)
 */
constexpr int kBbmExtraPlanes = 8;
constexpr int geoPlanesBBM(bool sph) { return geoPlanes(sph) + kBbmExtraPlanes; }

template <bool SPH>
__global__ void paramgeom_bbm_kernel(GridDims g, PhysParams p, const double* __restrict__ helem, double* __restrict__ geo)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.nx) * g.ny)
        return;
    const size_t e = size_t(t / g.nx) * g.nxs + (t % g.nx);
    constexpr int base = geoPlanes(SPH);
    double A[3][3];
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
            double s = 0.0;
            for (int q = 0; q < 9; ++q)
                s += gaussweight2(3, q) * missing6at(k, q) * missing6at(l, q) * geo[size_t(12 + q) * g.Npad + e];
            A[k][l] = s;
        }
    double inv[3][3];
    inverse<3>(A, inv);
    geo[size_t(base + 0) * g.Npad + e] = inv[0][0];
    geo[size_t(base + 1) * g.Npad + e] = 0.5 * (inv[0][1] + inv[1][0]);
    geo[size_t(base + 2) * g.Npad + e] = 0.5 * (inv[0][2] + inv[2][0]);
    geo[size_t(base + 3) * g.Npad + e] = inv[1][1];
    geo[size_t(base + 4) * g.Npad + e] = 0.5 * (inv[1][2] + inv[2][1]);
    geo[size_t(base + 5) * g.Npad + e] = inv[2][2];
    const double hel = helem[e];
    geo[size_t(base + 6) * g.Npad + e] = sqrt(0.1 / hel);
    geo[size_t(base + 7) * g.Npad + e] = 1.0 / (hel * sqrt(2. * (1. + p.nu0) * p.rho_ice));
}

template <bool SPH> struct PbbmStage : NodeStage<kPbbmDirectND<SPH>> {
    double G[27][32]; //!< h, expC, Pmax in the 9 Gauss points
    double S[24][32];
    double D[6][32];
    double GEO[geoPlanesBBM(SPH)][32];
    double2 UV[2][2][32];
    MaskStage M;
    double UVr[2][2];
    uint64_t bar[3]; //!< mbarriers of the S (+ damage), G and GEO groups (TMA staging)
    double pad[5];
};
static_assert(sizeof(PbbmStage<false>) % 128 == 0 && sizeof(PbbmStage<true>) % 128 == 0 && sizeof(NodeStage<false>) % 128 == 0,
    "TMA destinations need 128-byte alignment (G, S, D, GEO are multiples of 256 bytes behind the optional node-constant rows)");
#ifndef NSDG_PBBM_WARPS
#define NSDG_PBBM_WARPS 2 //!< warps per block (1 or 2): 8 warps per SM on Cartesian meshes, 6 on spherical ones (shared memory)
#endif
constexpr int kPbbmWarps = NSDG_PBBM_WARPS;
template <bool SPH> constexpr size_t pbbmSmemBytes() { return sizeof(PbbmStage<SPH>) * kPbbmWarps; }

template <bool SPH>
__global__ void __launch_bounds__(32 * kPbbmWarps, (SPH ? 6 : 8) / kPbbmWarps) subcycle_strip_pbbm(const __grid_constant__ UniformBBMArgs a)
{
    constexpr int CG = 2, NR = 3, DGs = 8, DGA = 6;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int GB = geoPlanes(SPH); // first BBM-specific plane
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    PbbmStage<SPH>& st = reinterpret_cast<PbbmStage<SPH>*>(smemRaw)[threadIdx.x >> 5];
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
    auto geo = [&](int k) { return st.GEO[k][lane]; };

    // staging groups, issued in the order UV, S(+D), G, GEO, ND
    auto issueUV = [&](int row) {
        if (row < ey1) {
            stageMasks(st.M, a.landmask, a.nodemask, g, row, sx, lane);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + 1 + k) * g.cgs + col0;
                cpAsync16cg(&st.UV[0][k][lane], a.u + n);
                cpAsync16cg(&st.UV[1][k][lane], a.v + n);
                if (loadsRight) {
                    cpAsync8(&st.UVr[0][k], a.u + n + CG);
                    cpAsync8(&st.UVr[1][k], a.v + n + CG);
                }
            }
        }
        cpAsyncCommit();
    };
    unsigned phase = 0; // TMA staging: the three groups are waited for together, once per row
    if constexpr (kParamTma) {
        if (lane == 0)
            for (int i = 0; i < 3; ++i)
                mbarInit(&st.bar[i], 1);
        mbarInitFence();
        __syncwarp();
    }
    auto tmaGroup = [&](int row, int bar, unsigned bytes, auto&& copies) {
        if (row < ey1) {
            __syncwarp(); // every lane has consumed the region that is refilled
            if (lane == 0) {
                mbarExpectTx(&st.bar[bar], bytes);
                copies(row * g.nxs + 32 * sx, &st.bar[bar]);
            }
        }
    };
    auto issueG = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 1, 27 * 256, [&](int x, uint64_t* b) {
                tmaLoadTile(&st.G[0][0], &a.tm[4], x, b);
                tmaLoadTile(&st.G[9][0], &a.tm[5], x, b);
                tmaLoadTile(&st.G[18][0], &a.tm[6], x, b);
            });
            return;
        }
        stageBarrier<kCoopPbbm<SPH>>();
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<9, kCoopPbbm<SPH>>(st.G, a.gH, Npad, first, lane);
            stagePlanes<9, kCoopPbbm<SPH>>(st.G + 9, a.gE, Npad, first, lane);
            stagePlanes<9, kCoopPbbm<SPH>>(st.G + 18, a.gP, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueS = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 0, 30 * 256, [&](int x, uint64_t* b) {
                tmaLoadTile(&st.S[0][0], &a.tm[0], x, b);
                tmaLoadTile(&st.S[8][0], &a.tm[1], x, b);
                tmaLoadTile(&st.S[16][0], &a.tm[2], x, b);
                tmaLoadTile(&st.D[0][0], &a.tm[3], x, b);
            });
            return;
        }
        stageBarrier<kCoopPbbm<SPH>>();
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<8, kCoopPbbm<SPH>>(st.S, a.s11, Npad, first, lane);
            stagePlanes<8, kCoopPbbm<SPH>>(st.S + 8, a.s12, Npad, first, lane);
            stagePlanes<8, kCoopPbbm<SPH>>(st.S + 16, a.s22, Npad, first, lane);
            stagePlanes<DGA, kCoopPbbm<SPH>>(st.D, a.damage, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueGEO = [&](int row) {
        if constexpr (kParamTma) {
            tmaGroup(row, 2, geoPlanesBBM(SPH) * 256, [&](int x, uint64_t* b) { tmaLoadTile(&st.GEO[0][0], &a.tm[7], x, b); });
            return;
        }
        stageBarrier<kCoopPbbm<SPH>>();
        if (row < ey1)
            stagePlanes<geoPlanesBBM(SPH), kCoopPbbm<SPH>>(st.GEO, a.geo, Npad, size_t(row) * g.nxs + 32 * sx, lane);
        cpAsyncCommit();
    };
    auto issueND = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + k) * g.cgs + col0;
                if constexpr (kPbbmDirectND<SPH>) {
                    for (const double* p : { a.cA, a.ax, a.ay, a.uO, a.vO, a.ilm })
                        prefetchL2(p + n);
                } else {
                    cpAsync16cg(&st.ND[k][0][lane], a.cA + n);
                    cpAsync16cg(&st.ND[k][1][lane], a.ax + n);
                    cpAsync16cg(&st.ND[k][2][lane], a.ay + n);
                    cpAsync16cg(&st.ND[k][3][lane], a.uO + n);
                    cpAsync16cg(&st.ND[k][4][lane], a.vO + n);
                    cpAsync16cg(&st.ND[k][5][lane], a.ilm + n);
                }
                prefetchL2(a.avgU + n); // read-modify-written at the end of the row
                prefetchL2(a.avgV + n);
            }
        }
        cpAsyncCommit();
    };

    double carryX[2] = { 0.0, 0.0 }, carryY[2] = { 0.0, 0.0 };
    double ul[9], vl[9];
    issueUV(ey0);
    issueS(ey0);
    issueG(ey0);
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

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
        cpAsyncWait<kParamTma ? 1 : 4>();
        __syncwarp(); // the mask bytes were staged by other lanes
        const bool ice = active && (st.M.LM[lane] != 0);
        const unsigned nm[2] = { nodeMaskWord(st.M, 0, lane), nodeMaskWord(st.M, 1, lane) };
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
        __syncwarp(); // every lane has read its mask bytes
        issueUV(ey + 1);

        // ---- strain in the 9 Gauss points ----
        if constexpr (kParamTma) {
            for (int i = 0; i < 3; ++i)
                mbarWait(&st.bar[i], phase);
            phase ^= 1u;
        } else {
            cpAsyncWait<2>(); // S, G and GEO of this row have landed
            stageBarrier<kCoopPbbm<SPH>>();
        }
        double e11[9], e12[9], e22[9];
        gaussStrain<SPH>(geo, ul, vl, ice, e11, e12, e22);

        // ---- the BBM law point by point (BBMStressUpdateStep.hpp:66-160): e** become the updated stresses ----
        double dG[9];
        {
            double s11c[DGs], s12c[DGs], s22c[DGs], dc[DGA];
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                s11c[j] = st.S[j][lane];
                s12c[j] = st.S[8 + j][lane];
                s22c[j] = st.S[16 + j][lane];
            }
#pragma unroll
            for (int j = 0; j < DGA; ++j)
                dc[j] = st.D[j][lane];
            const double scale = geo(GB + 6), invTdK = geo(GB + 7);
            const double cohScale = a.C_lab * scale, comprScale = a.compr_strength * scale;
#if NSDG_PARAM_SEP
            double t11q[9], t12q[9], t22q[9], dq[9];
            evalGaussSep<DGs>(s11c, t11q);
            evalGaussSep<DGs>(s12c, t12q);
            evalGaussSep<DGs>(s22c, t22q);
            evalGaussSep<DGA>(dc, dq);
#endif
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
#if NSDG_PARAM_SEP
                double t11 = t11q[q], t12 = t12q[q], t22 = t22q[q], d = dq[q];
#else
                double t11 = evalGauss<DGs, 3, q>(s11c), t12 = evalGauss<DGs, 3, q>(s12c), t22 = evalGauss<DGs, 3, q>(s22c);
                double d = evalGauss<DGA, 3, q>(dc);
#endif
                const double h = st.G[q][lane], expC = st.G[9 + q][lane], Pmax = st.G[18 + q][lane];
                const double g11 = e11[q], g12 = e12[q], g22 = e22[q];
#if NSDG_PARAM_SEP
                d = d < 1e-12 ? 1e-12 : d; // compare + select (fmin / fmax: five instructions each for their NaN rules)
                d = d > 1.0 ? 1.0 : d;
#else
                d = fmin(fmax(d, 1e-12), 1.0);
#endif
                double sigma_n = 0.5 * (t11 + t22);
                const double de = d * expC, de2 = de * de;
                const double tv = a.lambda0 * (de2 * de2);
                // multiplicator = tv / (tv + (1 - tildeP) dt), tildeP = min(1, -Pmax / sigma_n) under compression, else 0
                // (BBMStressUpdateStep.hpp:93-101), with ONE division: tildeP = 1 gives exactly 1; otherwise numerator and
                // denominator are scaled by sigma_n (same value up to rounding)
                const bool compress = sigma_n < 0.0;
                const bool full = compress && (Pmax >= -sigma_n);
                const double mnum = compress ? tv * sigma_n : tv;
                const double mden = compress ? fma(sigma_n + Pmax, a.deltaT, tv * sigma_n) : tv + a.deltaT;
                const double mult = full ? 1.0 : mnum * fastRcp(mden);
                const double elasticity = h * a.young * d * expC;
                const double Dunit = a.dunitK * elasticity;
                t11 = (t11 + Dunit * (g11 + a.nu0 * g22)) * mult;
                t22 = (t22 + Dunit * (a.nu0 * g11 + g22)) * mult;
                t12 = (t12 + Dunit * g12 * (1.0 - a.nu0)) * mult;
                sigma_n = 0.5 * (t11 + t22);
                const double tau2 = 0.25 * (t11 - t22) * (t11 - t22) + t12 * t12;
#if NSDG_BBM_SQRT & 2
                const double tau = sqrtBranchFree(tau2);
#else
                const double tau = fastSqrt(tau2);
#endif
                const double cohesion = cohScale * h, compr = comprScale * h;
                const double mc = tau + a.tan_phi * sigma_n;
                // one division: the compressive cap, when active, replaces the Mohr-Coulomb value (same operands, same result)
                const bool capped = sigma_n < -compr;
                double dcrit = (capped || mc > 0.0) ? (capped ? -compr : cohesion) * fastRcp(capped ? sigma_n : mc) : 1.0;
#if NSDG_PARAM_SEP
                dcrit = dcrit > 1.0 ? 1.0 : dcrit;
#else
                dcrit = fmin(dcrit, 1.0);
#endif
#if NSDG_BBM_SQRT & 4
                const double sqrtE = sqrtBranchFree(elasticity);
#else
                const double sqrtE = fastSqrt(elasticity);
#endif
                const double relax = (1.0 - dcrit) * a.deltaT * (sqrtE * invTdK);
                double dn = d - d * relax;
                t11 -= t11 * relax;
                t12 -= t12 * relax;
                t22 -= t22 * relax;
                if constexpr (SPH) { // the projections weight with w J, the mass matrices with w J cos
                    const double icos = invCos(geo, q % 3, q / 3, q);
                    dn *= icos;
                    t11 *= icos;
                    t12 *= icos;
                    t22 *= icos;
                }
                dG[q] = dn;
                e11[q] = t11;
                e12[q] = t12;
                e22[q] = t22;
            });
        }
        issueG(ey + 1);

        // ---- damage: DG6 projection = remove the three missing tensor modes, then the coefficients ----
        {
            double m0 = 0.0, m1 = 0.0, m2 = 0.0; // N^T y
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                constexpr double w = gaussweight2(3, q);
                constexpr double n0 = w * missing6at(0, q), n1 = w * missing6at(1, q), n2 = w * missing6at(2, q);
                if constexpr (n0 != 0.0)
                    m0 = fma(n0, dG[q], m0);
                if constexpr (n1 != 0.0)
                    m1 = fma(n1, dG[q], m1);
                m2 = fma(n2, dG[q], m2);
            });
            const double g00 = geo(GB + 0), g01 = geo(GB + 1), g02 = geo(GB + 2), g11 = geo(GB + 3), g12 = geo(GB + 4), g22 = geo(GB + 5);
            const double c0 = g00 * m0 + g01 * m1 + g02 * m2, c1 = g01 * m0 + g11 * m1 + g12 * m2, c2 = g02 * m0 + g12 * m1 + g22 * m2;
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                constexpr double p0 = missing6at(0, q), p1 = missing6at(1, q), p2 = missing6at(2, q);
                dG[q] = fma(-(p0 * c0 + p1 * c1 + p2 * c2), geo(12 + q), dG[q]);
            });
            double dc[DGA];
            coeffFromGauss<DGA>(dG, dc);
            if (active) {
#pragma unroll
                for (int j = 0; j < DGA; ++j)
                    a.damage[size_t(j) * Npad + e] = dc[j];
            }
        }

        // ---- stress: DG8 projection, coefficients, store, back to the Gauss points for the divergence ----
        projectDG8x3(geo, e11, e12, e22);
        auto component = [&](double* plane, double (&r)[9]) {
            double s[DGs];
            coeffFromGauss<DGs>(r, s);
            if (active) {
#pragma unroll
                for (int j = 0; j < DGs; ++j)
                    plane[size_t(j) * Npad + e] = s[j];
            }
#if !NSDG_PARAM_SEP
            gaussFromCoeff<DGs>(s, r);
#endif
            // r already holds the values of the stored coefficients in the Gauss points: after projectDG8x3 it lies in the
            // DG8 space, where (evaluate) o (B^) is the identity
        };
        component(a.s11, e11);
        component(a.s12, e12);
        component(a.s22, e22);
        // direct node constants and means: loads issued here, one stress-divergence evaluation ahead of the node update
        // (pinned: the compiler would sink them to their first use, and cannot move them across the cp.async statements)
        double2 ndc[2][kNodeConsts], avg0[2][2];
        if constexpr (kPbbmDirectND<SPH>) {
#pragma unroll
            for (int jy = 0; jy < CG; ++jy) {
                const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
                const double* src[kNodeConsts] = { a.cA, a.ax, a.ay, a.uO, a.vO, a.ilm };
#pragma unroll
                for (int i = 0; i < kNodeConsts; ++i)
                    ndc[jy][i] = ldPinned2(src[i] + n0);
                avg0[jy][0] = ldPinned2(a.avgU + n0);
                avg0[jy][1] = ldPinned2(a.avgV + n0);
            }
        }
        issueS(ey + 1);

        double Tx[9], Ty[9];
        if (ice) {
            gaussDivergence<SPH>(geo, e11, e12, e22, Tx, Ty);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k)
                Tx[k] = Ty[k] = 0.0;
        }
        issueGEO(ey + 1);

        const bool bottomDeferred = stripScatter(g, StripPos { lane, sx, sy, ex, ey, ey0, ey1, active, lastLane }, a.hbuf, a.vbuf, Tx, Ty);
        // ---- momentum update of the completed nodes ----
        cpAsyncWait<kParamTma ? 1 : 4>();
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
            double2 cA, ax, ay, uO, vO, ilm;
            if constexpr (kPbbmDirectND<SPH>) {
                cA = ndc[jy][0], ax = ndc[jy][1], ay = ndc[jy][2], uO = ndc[jy][3], vO = ndc[jy][4], ilm = ndc[jy][5];
            } else {
                cA = st.ND[jy][0][lane], ax = st.ND[jy][1][lane], ay = st.ND[jy][2][lane];
                uO = st.ND[jy][3][lane], vO = st.ND[jy][4][lane], ilm = st.ND[jy][5][lane];
            }
            const unsigned msk = nm[jy];
            double sx0 = Tx[jy * NR], sy0 = Ty[jy * NR], sx1 = Tx[jy * NR + 1], sy1 = Ty[jy * NR + 1];
            if (jy == 0) {
                sx0 += carryX[0];
                sy0 += carryY[0];
                sx1 += carryX[1];
                sy1 += carryY[1];
            }
            const bool d0 = msk & 1u, d1 = msk & 0x100u;
            double2 un, vn, ua, va;
            momentumNodeUniformBBM(a, cA.x, ax.x, ay.x, uO.x, vO.x, ilm.x, d0, ul[jy * NR], vl[jy * NR], d0 ? 0.0 : -sx0,
                d0 ? 0.0 : -sy0, un.x, vn.x, ua.x, va.x);
            momentumNodeUniformBBM(a, cA.y, ax.y, ay.y, uO.y, vO.y, ilm.y, d1, ul[jy * NR + 1], vl[jy * NR + 1],
                d1 ? 0.0 : -sx1, d1 ? 0.0 : -sy1, un.y, vn.y, ua.y, va.y);
            const bool rowSkip = !active || (jy == 0 && bottomDeferred);
            const bool skip0 = rowSkip || (lane == 0 && sx > 0);
            if (!rowSkip) {
                if (!skip0) {
                    *reinterpret_cast<double2*>(a.u + n0) = un;
                    *reinterpret_cast<double2*>(a.v + n0) = vn;
                    double2 au, av;
                    if constexpr (kPbbmDirectND<SPH>)
                        au = avg0[jy][0], av = avg0[jy][1];
                    else
                        au = *reinterpret_cast<const double2*>(a.avgU + n0), av = *reinterpret_cast<const double2*>(a.avgV + n0);
                    au.x += ua.x;
                    au.y += ua.y;
                    av.x += va.x;
                    av.y += va.y;
                    *reinterpret_cast<double2*>(a.avgU + n0) = au;
                    *reinterpret_cast<double2*>(a.avgV + n0) = av;
                } else {
                    a.u[n0 + 1] = un.y;
                    a.v[n0 + 1] = vn.y;
                    if constexpr (kPbbmDirectND<SPH>) {
                        a.avgU[n0 + 1] = avg0[jy][0].y + ua.y;
                        a.avgV[n0 + 1] = avg0[jy][1].y + va.y;
                    } else {
                        a.avgU[n0 + 1] += ua.y;
                        a.avgV[n0 + 1] += va.y;
                    }
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
