/*
 * nsdg_transport.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of the DG advection of neXtSIM_DG: ParametricTransportMap,
 * Interpolations (Function2DG, CG2DG, DG2CG, L2 error), DGTransport, Gauss-point limiters.
 *
 * Reference: dynamics/src/ParametricMap.cpp:13-87
 *            dynamics/src/Interpolations.cpp:29-247
 *            dynamics/src/DGTransport.cpp:31-566
 *            dynamics/src/include/dgLimit.hpp:16-84
 *            dynamics/src/include/Tools.hpp:24-46 (MeanValue)
 *
 * Vectors use the reference layouts: DG vector = row-major N x DG (dgVector.hpp:89-91),
 * CG vector = flat, node (jx,jy) at jx + (CG*nx+1)*jy (cgVector.hpp:21-31),
 * edge vectors nx(ny+1) x ED (X) and (nx+1)ny x ED (Y) (dgVector.hpp:139-196).
 */
#pragma once
#include "nsdg_mesh.hpp"

#include <functional>
#include <string>

namespace nso {

typedef std::function<double(double, double)> Function2;
typedef std::vector<double> Vec;

// ----------------------------------------------------------------------------------
// Interpolations
// ----------------------------------------------------------------------------------

//! Interpolations.cpp:29-72
template <int DG> void Function2DG(const Mesh& m, Vec& phi, const Function2& f)
{
    constexpr int G = gp1d(DG), Q = G * G;
    const auto& T = DGTab<DG, G>::get();
    phi.assign(m.nelements * DG, 0.0);
#pragma omp parallel for
    for (size_t eid = 0; eid < m.nelements; ++eid) {
        if (!m.landmask[eid])
            continue;
        double gp[2][Q], J[Q], M[DG][DG], iM[DG][DG], rhs[DG];
        gaussPointsInElement<G>(m, eid, gp);
        jacobian<G>(m, eid, J);
        massMatrix<DG>(m, eid, m.spherical, M);
        inverse<DG>(M, iM);
        for (int j = 0; j < DG; ++j) {
            double s = 0;
            for (int q = 0; q < Q; ++q) {
                double wj = J[q] * T.w[q];
                if (m.spherical)
                    wj *= cos(gp[1][q]);
                s += (T.psi[j][q] * wj) * f(gp[0][q], gp[1][q]);
            }
            rhs[j] = s;
        }
        for (int i = 0; i < DG; ++i) {
            double s = 0;
            for (int j = 0; j < DG; ++j)
                s += iM[i][j] * rhs[j];
            phi[eid * DG + i] = s;
        }
    }
}

//! The 4 / 9 local CG values of an element, x-fastest (CGDynamicsKernel.cpp:278-298)
template <int CG> inline void cgLocal(const Mesh& m, const Vec& cg, size_t eid, double* loc)
{
    const size_t cgshift = CG * m.nx + 1;
    const size_t cgi = CG * cgshift * (eid / m.nx) + CG * (eid % m.nx);
    for (int r = 0; r <= CG; ++r)
        for (int c = 0; c <= CG; ++c)
            loc[r * (CG + 1) + c] = cg[cgi + c + r * cgshift];
}

//! Interpolations.cpp:74-122  (L2 projection CG -> DG; ignores the land mask)
template <int CG, int DG> void CG2DG(const Mesh& m, Vec& dg, const Vec& cg)
{
    constexpr int G = gp1d(DG), Q = G * G, ND = cgdofs(CG);
    const auto& T = DGTab<DG, G>::get();
    const auto& P = CGTab<CG, G>::get();
    dg.resize(m.nelements * DG);
#pragma omp parallel for
    for (size_t e = 0; e < m.nelements; ++e) {
        double loc[ND], J[Q], gp[2][Q], M[DG][DG], iM[DG][DG], g[Q], rhs[DG];
        cgLocal<CG>(m, cg, e, loc);
        jacobian<G>(m, e, J);
        if (m.spherical)
            gaussPointsInElement<G>(m, e, gp);
        massMatrix<DG>(m, e, m.spherical, M);
        inverse<DG>(M, iM);
        for (int q = 0; q < Q; ++q) {
            double s = 0;
            for (int i = 0; i < ND; ++i)
                s += P.phi[i][q] * loc[i];
            double wj = J[q] * T.w[q];
            if (m.spherical)
                wj *= cos(gp[1][q]);
            g[q] = wj * s;
        }
        // (M^-1 * PSI) * g, evaluated left to right as in the reference expression
        for (int i = 0; i < DG; ++i) {
            double s = 0;
            for (int q = 0; q < Q; ++q) {
                double mp = 0;
                for (int j = 0; j < DG; ++j)
                    mp += iM[i][j] * T.psi[j][q];
                s += mp * g[q];
            }
            rhs[i] = s;
        }
        for (int i = 0; i < DG; ++i)
            dg[e * DG + i] = rhs[i];
    }
}

//! Interpolations.cpp:129-204 (nodal averaging; odd rows first, quirk Q7; boundary x2)
template <int CG, int DG> void DG2CG(const Mesh& m, Vec& dest, const Vec& src)
{
    constexpr int L = CG + 1;
    const auto& T = LagTab<DG, L>::get();
    const size_t row = CG * m.nx + 1;
    dest.assign(row * (CG * m.ny + 1), 0.0);
    static const double w1[4] = { 0.25, 0.25, 0.25, 0.25 };
    static const double w2[9] = { 0.25, 0.5, 0.25, 0.5, 1.0, 0.5, 0.25, 0.5, 0.25 };
    const double* wt = (CG == 1) ? w1 : w2;
    for (size_t p = 0; p < 2; ++p) {
#pragma omp parallel for
        for (size_t cy = 0; cy < m.ny; ++cy) {
            if (cy % 2 == p)
                continue;
            for (size_t cx = 0; cx < m.nx; ++cx) {
                const size_t c = cy * m.nx + cx;
                const size_t cgi = CG * row * cy + CG * cx;
                for (int r = 0; r < L; ++r)
                    for (int k = 0; k < L; ++k) {
                        const int q = r * L + k;
                        double At = 0;
                        for (int j = 0; j < DG; ++j)
                            At += src[c * DG + j] * T.psi[j][q];
                        dest[cgi + k + r * row] += wt[q] * At;
                    }
            }
        }
    }
    const size_t upperleft = CG * row * m.ny;
    for (size_t i = 0; i < CG * m.nx + 1; ++i) {
        dest[i] *= 2.0;
        dest[upperleft + i] *= 2.0;
    }
    for (size_t i = 0; i < CG * m.ny + 1; ++i) {
        dest[i * row] *= 2.0;
        dest[i * row + CG * m.nx] *= 2.0;
    }
}

//! Interpolations.cpp:206-247 (returns the squared error)
template <int DG> double L2ErrorFunctionDG(const Mesh& m, const Vec& src, const Function2& f)
{
    constexpr int G = gp1d(DG), Q = G * G;
    const auto& T = DGTab<DG, G>::get();
    double error = 0;
    for (size_t eid = 0; eid < m.nelements; ++eid) {
        if (!m.landmask[eid])
            continue;
        double gp[2][Q], J[Q];
        gaussPointsInElement<G>(m, eid, gp);
        jacobian<G>(m, eid, J);
        double s = 0;
        for (int q = 0; q < Q; ++q) {
            double v = 0;
            for (int j = 0; j < DG; ++j)
                v += src[eid * DG + j] * T.psi[j][q];
            double wj = J[q] * T.w[q];
            if (m.spherical)
                wj *= cos(gp[1][q]);
            s += wj * SQR(v - f(gp[0][q], gp[1][q]));
        }
        error += s;
    }
    return error;
}

//! Tools.hpp:24-46
template <int DG> double MeanValue(const Mesh& m, const Vec& phi)
{
    const auto& T = DGTab<DG, 3>::get();
    double mass = 0;
    for (size_t e = 0; e < m.nelements; ++e) {
        double J[9], gp[2][9];
        jacobian<3>(m, e, J);
        if (m.spherical)
            gaussPointsInElement<3>(m, e, gp);
        for (int q = 0; q < 9; ++q) {
            double v = 0;
            for (int j = 0; j < DG; ++j)
                v += phi[e * DG + j] * T.psi[j][q];
            mass += v * (m.spherical ? cos(gp[1][q]) : 1.0) * (J[q] * T.w[q]);
        }
    }
    return mass;
}

// ----------------------------------------------------------------------------------
// Limiters (dgLimit.hpp:16-84)
// ----------------------------------------------------------------------------------
template <int DG> void LimitMax(Vec& dg, double max)
{
    const long n = dg.size() / DG;
    if constexpr (DG == 1) {
        for (long i = 0; i < n; ++i)
            dg[i] = std::min(dg[i], max);
    } else if constexpr (DG == 3) {
#pragma omp parallel for
        for (long i = 0; i < n; ++i) {
            double* d = &dg[i * 3];
            d[0] = std::min(max, d[0]);
            const double l0 = 2.0 * std::max(fabs(d[1] + d[2]), fabs(d[1] - d[2]));
            if (l0 == 0)
                continue;
            const double ex = d[0] + l0 - max;
            if (ex > 0) {
                d[1] *= (max - d[0]) / l0;
                d[2] *= (max - d[0]) / l0;
            }
        }
    } else {
        static_assert(DG == 6, "LimitMax exists for DG 1,3,6 only");
        const auto& T = DGTab<6, 3>::get();
#pragma omp parallel for
        for (long i = 0; i < n; ++i) {
            double* d = &dg[i * 6];
            d[0] = std::min(max, d[0]);
            double maxvalue = -1e300;
            for (int q = 0; q < 9; ++q) {
                double v = 0;
                for (int j = 0; j < 6; ++j)
                    v += d[j] * T.psi[j][q];
                maxvalue = std::max(maxvalue, v);
            }
            const double l0 = maxvalue - d[0];
            if (maxvalue > max) {
                const double f = (max - d[0]) / l0;
                for (int j = 1; j < 6; ++j)
                    d[j] *= f;
            }
        }
    }
}
template <int DG> void LimitMin(Vec& dg, double min)
{
    const long n = dg.size() / DG;
    if constexpr (DG == 1) {
        for (long i = 0; i < n; ++i)
            dg[i] = std::max(dg[i], min);
    } else if constexpr (DG == 3) {
#pragma omp parallel for
        for (long i = 0; i < n; ++i) {
            double* d = &dg[i * 3];
            d[0] = std::max(min, d[0]);
            const double l0 = 2.0 * std::max(fabs(d[1] + d[2]), fabs(d[1] - d[2]));
            if (l0 == 0)
                continue;
            const double ex = d[0] - l0 - min;
            if (ex < 0) {
                d[1] *= (d[0] - min) / l0;
                d[2] *= (d[0] - min) / l0;
            }
        }
    } else {
        static_assert(DG == 6, "LimitMin exists for DG 1,3,6 only");
        const auto& T = DGTab<6, 3>::get();
#pragma omp parallel for
        for (long i = 0; i < n; ++i) {
            double* d = &dg[i * 6];
            d[0] = std::max(min, d[0]);
            double minvalue = 1e300;
            for (int q = 0; q < 9; ++q) {
                double v = 0;
                for (int j = 0; j < 6; ++j)
                    v += d[j] * T.psi[j][q];
                minvalue = std::min(minvalue, v);
            }
            const double l0 = minvalue - d[0];
            if (minvalue < min) {
                const double f = (d[0] - min) / l0;
                for (int j = 1; j < 6; ++j)
                    d[j] *= f;
            }
        }
    }
}

// ----------------------------------------------------------------------------------
// Edge traces of a cell vector (DGTransport.cpp:44-142)
// side: 0 bottom, 1 right, 2 top, 3 left
// ----------------------------------------------------------------------------------
template <int DG> inline void edgeofcell(const double* c, int side, double* out)
{
    const double sg = (side == 1 || side == 2) ? 0.5 : -0.5;
    const bool lr = (side == 1 || side == 3); // left/right edges run in y
    if constexpr (DG == 1) {
        out[0] = c[0];
    } else if constexpr (DG == 3) {
        if (lr) {
            out[0] = c[0] + sg * c[1];
            out[1] = c[2];
        } else {
            out[0] = c[0] + sg * c[2];
            out[1] = c[1];
        }
    } else if constexpr (DG == 6) {
        if (lr) {
            out[0] = c[0] + sg * c[1] + 1. / 6. * c[3];
            out[1] = c[2] + sg * c[5];
            out[2] = c[4];
        } else {
            out[0] = c[0] + sg * c[2] + 1. / 6. * c[4];
            out[1] = c[1] + sg * c[5];
            out[2] = c[3];
        }
    } else {
        static_assert(DG == 8, "DG must be 1,3,6,8");
        if (lr) {
            out[0] = c[0] + sg * c[1] + 1. / 6. * c[3];
            out[1] = c[2] + sg * c[5] + 1. / 6. * c[6];
            out[2] = c[4] + sg * c[7];
        } else {
            out[0] = c[0] + sg * c[2] + 1. / 6. * c[4];
            out[1] = c[1] + sg * c[5] + 1. / 6. * c[7];
            out[2] = c[3] + sg * c[6];
        }
    }
}

// ----------------------------------------------------------------------------------
// DGTransport (DGTransport.hpp:27-152, DGTransport.cpp:158-566)
// ----------------------------------------------------------------------------------
template <int DG> class Transport {
public:
    static constexpr int G = gp1d(DG), Q = G * G, ED = edgedofs(DG);
    const Mesh& m;
    std::string scheme = "rk2";
    Vec velx, vely; //!< N x DG
    Vec normalvel_X, normalvel_Y; //!< nx(ny+1) x ED, (nx+1)ny x ED
    Vec AdvX, AdvY; //!< N x DG x Q   (ParametricMap.cpp:13-66)
    Vec iMass; //!< N x DG x DG  (ParametricMap.cpp:68-87)
    Vec tmp1, tmp2, tmp3;

    explicit Transport(const Mesh& mesh)
        : m(mesh)
    {
        const size_t N = m.nelements;
        velx.assign(N * DG, 0.0);
        vely.assign(N * DG, 0.0);
        tmp1.assign(N * DG, 0.0);
        tmp2.assign(N * DG, 0.0);
        tmp3.assign(N * DG, 0.0);
        normalvel_X.assign(m.nx * (m.ny + 1) * ED, 0.0);
        normalvel_Y.assign((m.nx + 1) * m.ny * ED, 0.0);
        initCellTerms();
        initInverseMass();
    }

    //! ParametricMap.cpp:13-66 (spherical branch computes cos_lat and ignores it, quirk Q6)
    void initCellTerms()
    {
        const auto& T = DGTab<DG, G>::get();
        AdvX.resize(m.nelements * DG * Q);
        AdvY.resize(m.nelements * DG * Q);
#pragma omp parallel for
        for (size_t e = 0; e < m.nelements; ++e) {
            double a[2][Q], b[2][Q];
            dxT<G>(m, e, a);
            dyT<G>(m, e, b);
            for (int k = 0; k < 2; ++k)
                for (int q = 0; q < Q; ++q) {
                    a[k][q] *= T.w[q];
                    b[k][q] *= T.w[q];
                }
            for (int j = 0; j < DG; ++j)
                for (int q = 0; q < Q; ++q) {
                    AdvX[(e * DG + j) * Q + q] = T.psix[j][q] * b[1][q] - T.psiy[j][q] * a[1][q];
                    AdvY[(e * DG + j) * Q + q] = T.psiy[j][q] * a[0][q] - T.psix[j][q] * b[0][q];
                }
        }
    }
    //! ParametricMap.cpp:68-87
    void initInverseMass()
    {
        iMass.resize(m.nelements * DG * DG);
#pragma omp parallel for
        for (size_t e = 0; e < m.nelements; ++e) {
            double M[DG][DG], iM[DG][DG];
            massMatrix<DG>(m, e, m.spherical, M);
            inverse<DG>(M, iM);
            for (int i = 0; i < DG; ++i)
                for (int j = 0; j < DG; ++j)
                    iMass[(e * DG + i) * DG + j] = m.spherical ? iM[i][j] / EarthRadius : iM[i][j];
        }
    }

    //! DGTransport.cpp:158-252
    void reinitnormalvelocity()
    {
        std::fill(normalvel_X.begin(), normalvel_X.end(), 0.0);
        std::fill(normalvel_Y.begin(), normalvel_Y.end(), 0.0);
        const size_t nx = m.nx, ny = m.ny;
#pragma omp parallel for
        for (size_t iy = 0; iy < ny; ++iy) {
            size_t ey = iy * (nx + 1), cy = iy * nx;
            for (size_t ix = 0; ix < nx; ++ix, ++ey, ++cy) {
                if (m.landmask[cy] == 0)
                    continue;
                double tx, ty, ex[ED], eyv[ED];
                m.edgevector(ey, ey + nx + 1, tx, ty);
                edgeofcell<DG>(&velx[cy * DG], 3, ex);
                edgeofcell<DG>(&vely[cy * DG], 3, eyv);
                for (int k = 0; k < ED; ++k)
                    normalvel_Y[ey * ED + k] += 0.5 * (ty * ex[k] - tx * eyv[k]);
                m.edgevector(ey + 1, ey + nx + 2, tx, ty);
                edgeofcell<DG>(&velx[cy * DG], 1, ex);
                edgeofcell<DG>(&vely[cy * DG], 1, eyv);
                for (int k = 0; k < ED; ++k)
                    normalvel_Y[(ey + 1) * ED + k] += 0.5 * (ty * ex[k] - tx * eyv[k]);
            }
        }
#pragma omp parallel for
        for (size_t ix = 0; ix < nx; ++ix) {
            size_t cx = ix, nn = ix;
            for (size_t iy = 0; iy < ny; ++iy, cx += nx, nn += nx + 1) {
                if (m.landmask[cx] == 0)
                    continue;
                double tx, ty, ex[ED], eyv[ED];
                m.edgevector(nn, nn + 1, tx, ty);
                edgeofcell<DG>(&velx[cx * DG], 0, ex);
                edgeofcell<DG>(&vely[cx * DG], 0, eyv);
                for (int k = 0; k < ED; ++k)
                    normalvel_X[cx * ED + k] += 0.5 * (-ty * ex[k] + tx * eyv[k]);
                m.edgevector(nn + nx + 1, nn + nx + 2, tx, ty);
                edgeofcell<DG>(&velx[cx * DG], 2, ex);
                edgeofcell<DG>(&vely[cx * DG], 2, eyv);
                for (int k = 0; k < ED; ++k)
                    normalvel_X[(cx + nx) * ED + k] += 0.5 * (-ty * ex[k] + tx * eyv[k]);
            }
        }
        for (size_t seg = 0; seg < 4; ++seg)
            for (size_t i = 0; i < m.dirichlet[seg].size(); ++i) {
                const size_t eid = m.dirichlet[seg][i];
                const size_t ix = eid % nx, iy = eid / nx;
                double* r;
                if (seg == 0)
                    r = &normalvel_X[(nx * iy + ix) * ED];
                else if (seg == 1)
                    r = &normalvel_Y[((nx + 1) * iy + ix + 1) * ED];
                else if (seg == 2)
                    r = &normalvel_X[(nx * (iy + 1) + ix) * ED];
                else
                    r = &normalvel_Y[((nx + 1) * iy + ix) * ED];
                for (int k = 0; k < ED; ++k)
                    r[k] *= 2.0;
            }
    }

    //! DGTransport.cpp:261-268
    template <int CG> void prepareAdvection(const Vec& cg_vx, const Vec& cg_vy)
    {
        CG2DG<CG, DG>(m, velx, cg_vx);
        CG2DG<CG, DG>(m, vely, cg_vy);
        reinitnormalvelocity();
    }

    //! DGTransport.cpp:272-303
    void cell_term(double dt, Vec& phiup, const Vec& phi, size_t eid) const
    {
        if constexpr (DG == 1)
            return;
        if (m.landmask[eid] == 0)
            return;
        const auto& T = DGTab<DG, G>::get();
        double vxg[Q], vyg[Q], pg[Q];
        for (int q = 0; q < Q; ++q) {
            double a = 0, b = 0, c = 0;
            for (int j = 0; j < DG; ++j) {
                a += velx[eid * DG + j] * T.psi[j][q];
                b += vely[eid * DG + j] * T.psi[j][q];
                c += phi[eid * DG + j] * T.psi[j][q];
            }
            vxg[q] = a;
            vyg[q] = b;
            pg[q] = c;
        }
        for (int j = 0; j < DG; ++j) {
            const double* ax = &AdvX[(eid * DG + j) * Q];
            const double* ay = &AdvY[(eid * DG + j) * Q];
            double s = 0;
            for (int q = 0; q < Q; ++q)
                s += (dt * (ax[q] * vxg[q] + ay[q] * vyg[q])) * pg[q];
            phiup[eid * DG + j] += s;
        }
    }

    //! edge values in the edge Gauss points: row * PSIe<ED,G>
    static inline void toGauss(const double* row, double* out)
    {
        const auto& E = EdgeTab<DG, G>::get();
        for (int q = 0; q < G; ++q) {
            double s = 0;
            for (int k = 0; k < ED; ++k)
                s += row[k] * E.psie[k][q];
            out[q] = s;
        }
    }
    //! phiup.row(c) += sign * dt * tmp * PSIe_w<DG,G,side>
    static inline void applyEdge(Vec& phiup, size_t c, double sdt, const double* tmp, int side)
    {
        const auto& E = EdgeTab<DG, G>::get();
        for (int j = 0; j < DG; ++j) {
            double s = 0;
            for (int q = 0; q < G; ++q)
                s += (sdt * tmp[q]) * E.psiew[side][q][j];
            phiup[c * DG + j] += s;
        }
    }

    //! DGTransport.cpp:355-410 (c1 below, c2 above; ie = X-edge id)
    void edge_term_X(double dt, Vec& phiup, const Vec& phi, size_t c1, size_t c2, size_t ie) const
    {
        if (m.landmask[c1] == 0 || m.landmask[c2] == 0)
            return;
        double vg[G], t1[ED], t2[ED], g1[G], g2[G], tmp[G];
        toGauss(&normalvel_X[ie * ED], vg);
        edgeofcell<DG>(&phi[c1 * DG], 2, t1);
        edgeofcell<DG>(&phi[c2 * DG], 0, t2);
        toGauss(t1, g1);
        toGauss(t2, g2);
        for (int q = 0; q < G; ++q)
            tmp[q] = std::max(vg[q], 0.) * g1[q] + std::min(vg[q], 0.) * g2[q];
        applyEdge(phiup, c1, -dt, tmp, 2);
        applyEdge(phiup, c2, dt, tmp, 0);
    }
    //! DGTransport.cpp:372-433 (c1 left, c2 right; ie = Y-edge id)
    void edge_term_Y(double dt, Vec& phiup, const Vec& phi, size_t c1, size_t c2, size_t ie) const
    {
        if (m.landmask[c1] == 0 || m.landmask[c2] == 0)
            return;
        double vg[G], t1[ED], t2[ED], g1[G], g2[G], tmp[G];
        toGauss(&normalvel_Y[ie * ED], vg);
        edgeofcell<DG>(&phi[c1 * DG], 1, t1);
        edgeofcell<DG>(&phi[c2 * DG], 3, t2);
        toGauss(t1, g1);
        toGauss(t2, g2);
        for (int q = 0; q < G; ++q)
            tmp[q] = std::max(vg[q], 0.) * g1[q] + std::min(vg[q], 0.) * g2[q];
        applyEdge(phiup, c1, -dt, tmp, 1);
        applyEdge(phiup, c2, dt, tmp, 3);
    }
    //! DGTransport.cpp:306-351 (Dirichlet edges: outflow only). side as in dirichlet[]
    void boundary(double dt, Vec& phiup, const Vec& phi, int side, size_t c, size_t e) const
    {
        double vg[G], t[ED], g[G], tmp[G];
        const Vec& nv = (side == 0 || side == 2) ? normalvel_X : normalvel_Y;
        toGauss(&nv[e * ED], vg);
        edgeofcell<DG>(&phi[c * DG], side, t);
        toGauss(t, g);
        const double sg = (side == 0 || side == 3) ? -1.0 : 1.0; // outward normal vs edge normal
        for (int q = 0; q < G; ++q)
            tmp[q] = g[q] * std::max(sg * vg[q], 0.);
        applyEdge(phiup, c, -dt, tmp, side);
    }

    //! DGTransport.cpp:435-512
    void DGTransportOperator(double dt, const Vec& phi, Vec& phiup) const
    {
        const size_t nx = m.nx, ny = m.ny;
        std::fill(phiup.begin(), phiup.end(), 0.0);
#pragma omp parallel for
        for (size_t eid = 0; eid < m.nelements; ++eid)
            cell_term(dt, phiup, phi, eid);
#pragma omp parallel for
        for (size_t iy = 0; iy < ny; ++iy) {
            size_t ic = iy * nx, ie = iy * (nx + 1) + 1;
            for (size_t i = 0; i + 1 < nx; ++i, ++ic, ++ie)
                edge_term_Y(dt, phiup, phi, ic, ic + 1, ie);
        }
#pragma omp parallel for
        for (size_t ix = 0; ix < nx; ++ix) {
            size_t ic = ix, ie = ix + nx;
            for (size_t i = 0; i + 1 < ny; ++i, ic += nx, ie += nx)
                edge_term_X(dt, phiup, phi, ic, ic + nx, ie);
        }
        for (size_t pc = 0; pc < m.periodic.size(); ++pc)
            for (size_t i = 0; i < m.periodic[pc].size(); ++i) {
                const auto& p = m.periodic[pc][i];
                if (p[0] == 0)
                    edge_term_X(dt, phiup, phi, p[1], p[2], p[3]);
                else
                    edge_term_Y(dt, phiup, phi, p[1], p[2], p[3]);
            }
        for (size_t seg = 0; seg < 4; ++seg)
            for (size_t i = 0; i < m.dirichlet[seg].size(); ++i) {
                const size_t eid = m.dirichlet[seg][i];
                const size_t ix = eid % nx, iy = eid / nx;
                if (seg == 0)
                    boundary(dt, phiup, phi, 0, eid, nx * iy + ix);
                else if (seg == 1)
                    boundary(dt, phiup, phi, 1, eid, (nx + 1) * iy + ix + 1);
                else if (seg == 2)
                    boundary(dt, phiup, phi, 2, eid, nx * (iy + 1) + ix);
                else
                    boundary(dt, phiup, phi, 3, eid, (nx + 1) * iy + ix);
            }
#pragma omp parallel for
        for (size_t eid = 0; eid < m.nelements; ++eid) {
            double r[DG];
            for (int i = 0; i < DG; ++i) {
                double s = 0;
                for (int j = 0; j < DG; ++j)
                    s += iMass[(eid * DG + i) * DG + j] * phiup[eid * DG + j];
                r[i] = s;
            }
            for (int i = 0; i < DG; ++i)
                phiup[eid * DG + i] = r[i];
        }
    }

    //! DGTransport.cpp:514-566
    void step(double dt, Vec& phi)
    {
        const size_t n = phi.size();
        if (scheme == "rk1") {
            DGTransportOperator(dt, phi, tmp1);
            for (size_t i = 0; i < n; ++i)
                phi[i] += tmp1[i];
        } else if (scheme == "rk2") {
            DGTransportOperator(dt, phi, tmp1);
            for (size_t i = 0; i < n; ++i)
                phi[i] += tmp1[i];
            DGTransportOperator(dt, phi, tmp2);
            for (size_t i = 0; i < n; ++i)
                phi[i] += 0.5 * (tmp2[i] - tmp1[i]);
        } else if (scheme == "rk3") {
            DGTransportOperator(dt, phi, tmp1);
            for (size_t i = 0; i < n; ++i)
                tmp1[i] += phi[i];
            DGTransportOperator(dt, tmp1, tmp2);
            for (size_t i = 0; i < n; ++i) {
                tmp2[i] += tmp1[i];
                tmp2[i] *= 0.25;
                tmp2[i] += 0.75 * phi[i];
            }
            DGTransportOperator(dt, tmp2, tmp3);
            for (size_t i = 0; i < n; ++i) {
                tmp3[i] += tmp2[i];
                phi[i] *= 1.0 / 3.0;
                phi[i] += 2.0 / 3.0 * tmp3[i];
            }
        } else
            abort();
    }
};

} // namespace nso
