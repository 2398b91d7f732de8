/*
 * nsdg_setup.cuh -- per-element geometry and operator matrices on the parametric mesh,
 * computed ON THE DEVICE once per mesh (one thread per element / node).
 *
 * What is computed (values identical to the reference up to rounding):
 *   transport:  AdvectionCellTermX/Y, InverseDGMassMatrix      dynamics/src/ParametricMap.cpp:13-87
 *   momentum:   lumpedcgmass, lumpedcg1mass                    dynamics/src/ParametricMap.cpp:94-206
 *               divS1, divS2, divM, iMgradX, iMgradY, iMM,
 *               iMJwPSI, iMJwPSI_dam, dX_SSH, dY_SSH           dynamics/src/ParametricMap.cpp:209-356
 *   geometry:   dxT, dyT, J, Gauss points, mass matrices       dynamics/src/include/ParametricTools.hpp:73-202
 *               element corner coordinates + longitude unwrap  dynamics/src/include/ParametricMesh.hpp:196-232
 *               area / h                                       dynamics/src/include/ParametricMesh.hpp:276-299
 *
 * Operator storage: entry k of element e at op[k*pitch + e*estride]; (pitch,estride) =
 * (Npad,1) for general meshes, (1,0) for the single shared copy of a uniform mesh.
 */
#pragma once
#include "nsdg_state.cuh"

namespace nsdg {

//! corner coordinates of element (ix,iy): lower-left, lower-right, upper-left, upper-right
__host__ __device__ inline void elementCorners(const double* __restrict__ vx, const double* __restrict__ vy, int nx,
    int ix, int iy, bool spherical, double (&c)[4][2])
{
    const size_t nid = size_t(iy) * (nx + 1) + ix;
    const size_t ids[4] = { nid, nid + 1, nid + nx + 1, nid + nx + 2 };
    for (int i = 0; i < 4; ++i) {
        c[i][0] = vx[ids[i]];
        c[i][1] = vy[ids[i]];
    }
    if (spherical) { // correctlongitude, ParametricMesh.hpp:196-210
        bool problem = false;
        for (int i = 1; i < 4; ++i)
            if (fabs(c[0][0] - c[i][0]) > 2.0 / 3.0 * M_PI)
                problem = true;
        if (problem)
            for (int i = 0; i < 4; ++i)
                if (c[i][0] < 0)
                    c[i][0] += 2.0 * M_PI;
    }
}

//! weighted-free dxT, dyT (2 x G^2), J and (optionally) physical Gauss points of one element
template <int G>
__host__ __device__ inline void elementMap(const double (&c)[4][2], double (&dx)[2][G * G], double (&dy)[2][G * G],
    double (&J)[G * G], double (&lat)[G * G])
{
    for (int q = 0; q < G * G; ++q) {
        for (int k = 0; k < 2; ++k) {
            double a = 0, b = 0;
            for (int i = 0; i < 4; ++i) {
                a += c[i][k] * PHIx(1, G, i, q);
                b += c[i][k] * PHIy(1, G, i, q);
            }
            dx[k][q] = a;
            dy[k][q] = b;
        }
        J[q] = dx[0][q] * dy[1][q] - dx[1][q] * dy[0][q];
        double l = 0;
        for (int i = 0; i < 4; ++i)
            l += c[i][1] * PHI(1, G, i, q);
        lat[q] = l;
    }
}

//! massMatrix<DG> (ParametricTools.hpp:109-153) / SphericalTools::massMatrix<DG> (:189-202)
template <int DG> __host__ __device__ inline void massMatrix(const double (&c)[4][2], bool coslat, double (&M)[DG][DG])
{
    constexpr int G = gp1d(DG), Q = G * G;
    double dx[2][Q], dy[2][Q], J[Q], lat[Q], wj[Q];
    elementMap<G>(c, dx, dy, J, lat);
    for (int q = 0; q < Q; ++q)
        wj[q] = coslat ? gaussweight2(G, q) * J[q] * cos(lat[q]) : gaussweight2(G, q) * J[q];
    for (int i = 0; i < DG; ++i)
        for (int j = 0; j < DG; ++j) {
            double s = 0;
            for (int q = 0; q < Q; ++q)
                s += (PSI(G, i, q) * wj[q]) * PSI(G, j, q);
            M[i][j] = s;
        }
}

//! dense inverse, LU with partial pivoting (the algorithm behind Eigen's fixed-size .inverse() for N > 4)
template <int N> __host__ __device__ inline void inverse(const double (&Ain)[N][N], double (&inv)[N][N])
{
    double A[N][N];
    int piv[N];
    for (int i = 0; i < N; ++i) {
        piv[i] = i;
        for (int j = 0; j < N; ++j)
            A[i][j] = Ain[i][j];
    }
    for (int k = 0; k < N; ++k) {
        int p = k;
        for (int i = k + 1; i < N; ++i)
            if (fabs(A[i][k]) > fabs(A[p][k]))
                p = i;
        if (p != k) {
            for (int j = 0; j < N; ++j) {
                const double t = A[k][j];
                A[k][j] = A[p][j];
                A[p][j] = t;
            }
            const int t = piv[k];
            piv[k] = piv[p];
            piv[p] = t;
        }
        for (int i = k + 1; i < N; ++i) {
            A[i][k] /= A[k][k];
            for (int j = k + 1; j < N; ++j)
                A[i][j] -= A[i][k] * A[k][j];
        }
    }
    for (int col = 0; col < N; ++col) {
        double y[N];
        for (int i = 0; i < N; ++i) {
            double s = (piv[i] == col) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j)
                s -= A[i][j] * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < N; ++j)
                s -= A[i][j] * inv[j][col];
            inv[i][col] = s / A[i][i];
        }
    }
}

//! pointers to the operator planes of the momentum equation
struct MomentumOpPtrs {
    double *Gx, *Gy, *GM, *B, *Bd, *D1, *D2, *DM, *dXssh, *dYssh;
    size_t pitch; //!< distance between entries
    int estride; //!< 1 (per-element) or 0 (uniform)
};
//! pointers to the transport operator planes
struct TransportOpPtrs {
    double *AdvX, *AdvY, *iMass;
    size_t pitch;
    int estride;
};

//! ParametricMap.cpp:13-87 for one element
template <int DG>
__host__ __device__ inline void transportOpsOfElement(const double (&c)[4][2], bool spherical, TransportOpPtrs o, size_t e)
{
    constexpr int G = gp1d(DG), Q = G * G;
    double dx[2][Q], dy[2][Q], J[Q], lat[Q];
    elementMap<G>(c, dx, dy, J, lat);
    const size_t eo = e * o.estride;
    for (int j = 0; j < DG; ++j)
        for (int q = 0; q < Q; ++q) {
            const double w = gaussweight2(G, q);
            // quirk Q6: the spherical branch computes cos(lat) and ignores it -- same formula
            o.AdvX[(j * Q + q) * o.pitch + eo] = PSIx(G, j, q) * (dy[1][q] * w) - PSIy(G, j, q) * (dx[1][q] * w);
            o.AdvY[(j * Q + q) * o.pitch + eo] = PSIy(G, j, q) * (dx[0][q] * w) - PSIx(G, j, q) * (dy[0][q] * w);
        }
    double M[DG][DG], iM[DG][DG];
    massMatrix<DG>(c, spherical, M);
    inverse<DG>(M, iM);
    for (int i = 0; i < DG; ++i)
        for (int j = 0; j < DG; ++j)
            o.iMass[(i * DG + j) * o.pitch + eo] = spherical ? iM[i][j] / EarthRadius : iM[i][j];
}

//! ParametricMap.cpp:209-356 for one element
//! full = false: only the 4 x 4 sea-surface-height matrices (the factored-operator kernels need nothing else)
template <int CG, int DGA>
__host__ __device__ inline void momentumOpsOfElement(const double (&c)[4][2], bool sph, MomentumOpPtrs o, size_t e, bool full = true)
{
    constexpr int DGs = cg2dgstress(CG), GS = gp1d(DGs), Q = GS * GS, ND = cgdofs(CG);
    double Fx[2][Q], Fy[2][Q], J[Q], lat[Q], cl[Q], sl[Q];
    elementMap<GS>(c, Fx, Fy, J, lat);
    for (int q = 0; q < Q; ++q) {
        const double w = gaussweight2(GS, q);
        for (int k = 0; k < 2; ++k) {
            Fx[k][q] *= w;
            Fy[k][q] *= w;
        }
        cl[q] = sph ? cos(lat[q]) : 1.0;
        sl[q] = sph ? sin(lat[q]) : 0.0;
    }
    const size_t eo = e * o.estride;
    double d1[ND][DGs], d2[ND][DGs], dm[ND][DGs];
    for (int i = 0; i < ND; ++i)
        for (int j = 0; j < DGs; ++j) {
            double a = 0, b = 0, m = 0;
            for (int q = 0; q < Q; ++q) {
                const double dxc = PHIx(CG, GS, i, q) * Fy[1][q] - PHIy(CG, GS, i, q) * Fx[1][q];
                const double dyc = PHIy(CG, GS, i, q) * Fx[0][q] - PHIx(CG, GS, i, q) * Fy[0][q];
                a += dxc * PSI(GS, j, q);
                b += (sph ? dyc * cl[q] : dyc) * PSI(GS, j, q);
                if (sph)
                    m += (PHI(CG, GS, i, q) * (J[q] * sl[q] * gaussweight2(GS, q))) * PSI(GS, j, q);
            }
            d1[i][j] = sph ? a / EarthRadius : a;
            d2[i][j] = sph ? b / EarthRadius : b;
            dm[i][j] = sph ? m / EarthRadius : 0.0;
            if (full) {
                o.D1[(i * DGs + j) * o.pitch + eo] = d1[i][j];
                o.D2[(i * DGs + j) * o.pitch + eo] = d2[i][j];
                if (sph)
                    o.DM[(i * DGs + j) * o.pitch + eo] = dm[i][j];
            }
        }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double a = 0, b = 0;
            for (int q = 0; q < Q; ++q) {
                const double dx1 = PHIx(1, GS, i, q) * Fy[1][q] - PHIy(1, GS, i, q) * Fx[1][q];
                const double dy1 = PHIy(1, GS, i, q) * Fx[0][q] - PHIx(1, GS, i, q) * Fy[0][q];
                a += dx1 * PHI(1, GS, j, q);
                b += (sph ? dy1 * cl[q] : dy1) * PHI(1, GS, j, q);
            }
            o.dXssh[(i * 4 + j) * o.pitch + eo] = sph ? a / EarthRadius : a;
            o.dYssh[(i * 4 + j) * o.pitch + eo] = sph ? b / EarthRadius : b;
        }
    if (!full)
        return;
    double M[DGs][DGs], iM[DGs][DGs];
    massMatrix<DGs>(c, sph, M);
    inverse<DGs>(M, iM);
    for (int i = 0; i < DGs; ++i) {
        for (int k = 0; k < ND; ++k) {
            double a = 0, b = 0, m = 0;
            for (int j = 0; j < DGs; ++j) {
                a += iM[i][j] * d1[k][j];
                b += iM[i][j] * d2[k][j];
                m += iM[i][j] * dm[k][j];
            }
            o.Gx[(i * ND + k) * o.pitch + eo] = a;
            o.Gy[(i * ND + k) * o.pitch + eo] = b;
            if (sph)
                o.GM[(i * ND + k) * o.pitch + eo] = m;
        }
        for (int q = 0; q < Q; ++q) {
            double a = 0;
            for (int j = 0; j < DGs; ++j)
                a += iM[i][j] * (PSI(GS, j, q) * (gaussweight2(GS, q) * J[q]));
            o.B[(i * Q + q) * o.pitch + eo] = a;
        }
    }
    double Md[DGA][DGA], iMd[DGA][DGA];
    massMatrix<DGA>(c, sph, Md);
    inverse<DGA>(Md, iMd);
    for (int i = 0; i < DGA; ++i)
        for (int q = 0; q < Q; ++q) {
            double a = 0;
            for (int j = 0; j < DGA; ++j)
                a += iMd[i][j] * (PSI(GS, j, q) * (gaussweight2(GS, q) * J[q]));
            o.Bd[(i * Q + q) * o.pitch + eo] = a;
        }
}

//! element's share of the lumped CG mass at its local node `loc` (ParametricMap.cpp:111-132, :175-191)
template <int CG, int CGGP>
__host__ __device__ inline double lumpedMassShare(const double (&c)[4][2], bool sph, int loc)
{
    constexpr int QQ = CGGP * CGGP;
    double dx[2][QQ], dy[2][QQ], J[QQ], lat[QQ];
    elementMap<CGGP>(c, dx, dy, J, lat);
    double s = 0;
    for (int q = 0; q < QQ; ++q) {
        double wj = J[q] * gaussweight2(CGGP, q);
        if (sph)
            wj *= cos(lat[q]);
        s += PHI(CG, CGGP, loc, q) * wj;
    }
    return s;
}

//! area of an element from its 4 corners (ParametricMesh.hpp:276-295); corners as stored (no unwrap)
__host__ __device__ inline double elementArea(const double* vx, const double* vy, int nx, int ix, int iy)
{
    const size_t n = size_t(iy) * (nx + 1) + ix;
    auto d2 = [&](size_t a, size_t b) {
        return (vx[a] - vx[b]) * (vx[a] - vx[b]) + (vy[a] - vy[b]) * (vy[a] - vy[b]);
    };
    const double a = d2(n, n + 1), b = d2(n + 1, n + nx + 2), cc = d2(n + 1 + nx, n + 2 + nx), d = d2(n, n + nx + 1);
    const double e = d2(n, n + nx + 2), f = d2(n + 1, n + nx + 1);
    const double t = b + d - a - cc;
    return 0.25 * sqrt(4.0 * e * f - t * t);
}

} // namespace nsdg
