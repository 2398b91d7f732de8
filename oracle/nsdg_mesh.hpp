/*
 * nsdg_mesh.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of neXtSIM_DG's ParametricMesh and ParametricTools in plain C++17
 * (no Eigen).  Every function cites the reference lines it follows.
 *
 * Reference: dynamics/src/include/ParametricMesh.hpp:42-442
 *            dynamics/src/ParametricMesh.cpp:198-292
 *            dynamics/src/include/ParametricTools.hpp:73-202
 */
#pragma once
#include "nsdg_tables.hpp"

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <vector>

namespace nso {

inline double SQR(double x) { return x * x; }

struct Mesh {
    bool spherical = false;
    size_t nx = 0, ny = 0, nnodes = 0, nelements = 0;
    std::vector<double> vx, vy; //!< vertex coordinates, index ix + (nx+1)*iy
    std::vector<uint8_t> landmask; //!< 1 = ice/ocean, 0 = land (ParametricMesh.cpp:221, quirk Q10)
    std::array<std::vector<size_t>, 4> dirichlet; //!< bottom, right, top, left (ParametricMesh.hpp:70)
    //! periodic[seg][i] = {type(0 X-edge,1 Y-edge), elem left/bottom, elem right/top, edge id}
    std::vector<std::vector<std::array<size_t, 4>>> periodic;

    void reset()
    {
        nx = ny = nnodes = nelements = 0;
        for (auto& d : dirichlet)
            d.clear();
        periodic.clear();
        landmask.clear();
    }

    //! ParametricMesh.cpp:198-210. coords = interleaved (x,y) per vertex, vertex index x-fastest.
    void coordinatesFromArray(size_t nx_, size_t ny_, const double* coords)
    {
        nx = nx_;
        ny = ny_;
        nelements = nx * ny;
        nnodes = (nx + 1) * (ny + 1);
        vx.resize(nnodes);
        vy.resize(nnodes);
        for (size_t i = 0; i < nnodes; ++i) {
            vx[i] = coords[2 * i];
            vy[i] = coords[2 * i + 1];
        }
    }

    //! ParametricMesh.hpp:137-157
    void rotatePoleToGreenland()
    {
        for (size_t i = 0; i < nnodes; ++i) {
            const double x = cos(vy[i]) * cos(vx[i]);
            const double y = cos(vy[i]) * sin(vx[i]);
            const double z = sin(vy[i]);
            const double aw = 40.0 * M_PI / 180.0;
            const double x1 = cos(aw) * x - sin(aw) * y;
            const double y1 = sin(aw) * x + cos(aw) * y;
            const double z1 = z;
            const double bw = 15.0 * M_PI / 180.0;
            const double x2 = cos(bw) * x1 - sin(bw) * z1;
            const double y2 = y1;
            const double z2 = sin(bw) * x1 + cos(bw) * z1;
            vy[i] = asin(z2);
            vx[i] = atan2(y2, x2);
        }
    }

    //! ParametricMesh.cpp:217-223
    void landmaskFromArray(const double* mask)
    {
        landmask.resize(nelements);
        for (size_t i = 0; i < nelements; ++i)
            landmask[i] = (mask[i] == 1.);
    }

    //! ParametricMesh.cpp:228-252
    void dirichletFromMask()
    {
        const std::array<size_t, 4> startX = { 0, 0, 0, 1 };
        const std::array<size_t, 4> stopX = { nx, nx - 1, nx, nx };
        const std::array<size_t, 4> startY = { 1, 0, 0, 0 };
        const std::array<size_t, 4> stopY = { ny, ny, ny - 1, ny };
        const std::array<long, 4> deltaIdx = { -static_cast<long>(nx), 1, static_cast<long>(nx), -1 };
        for (int edge = 0; edge < 4; ++edge) {
            for (size_t j = startY[edge]; j < stopY[edge]; ++j)
                for (size_t i = startX[edge]; i < stopX[edge]; ++i) {
                    const size_t idx = i + nx * j;
                    if (!landmask[idx])
                        continue;
                    if (!landmask[idx + deltaIdx[edge]])
                        dirichlet[edge].push_back(idx);
                }
            std::sort(dirichlet[edge].begin(), dirichlet[edge].end());
        }
    }

    //! ParametricMesh.cpp:259-272
    void dirichletFromEdge(int edge)
    {
        const std::array<size_t, 4> start = { 0, nx - 1, nelements - nx, 0 };
        const std::array<size_t, 4> stop = { nx, nelements, nelements, nelements };
        const std::array<size_t, 4> stride = { 1, nx, 1, nx };
        for (size_t idx = start[edge]; idx < stop[edge]; idx += stride[edge])
            if (landmask[idx])
                dirichlet[edge].push_back(idx);
        std::sort(dirichlet[edge].begin(), dirichlet[edge].end());
    }

    //! The sequence of DynamicsKernel::initialise, DynamicsKernel.hpp:44-58
    void initFromArrays(size_t nx_, size_t ny_, const double* coords, const double* mask, bool sph)
    {
        reset();
        spherical = sph;
        coordinatesFromArray(nx_, ny_, coords);
        if (spherical)
            rotatePoleToGreenland();
        landmaskFromArray(mask);
        dirichletFromMask();
        for (int edge = 0; edge < 4; ++edge)
            dirichletFromEdge(edge);
    }

    size_t eid2nid(size_t eid) const { return (eid / nx) * (nx + 1) + (eid % nx); }

    //! ParametricMesh.hpp:196-210
    template <int N> void correctlongitude(double (&c)[N][2]) const
    {
        bool problem = false;
        for (int i = 1; i < N; ++i)
            if (fabs(c[0][0] - c[i][0]) > 2.0 / 3.0 * M_PI) {
                problem = true;
                break;
            }
        if (problem)
            for (int i = 0; i < N; ++i)
                if (c[i][0] < 0)
                    c[i][0] += 2.0 * M_PI;
    }

    //! ParametricMesh.hpp:218-232
    void coordinatesOfElement(size_t eid, double (&c)[4][2]) const
    {
        const size_t nid = eid2nid(eid);
        const size_t ids[4] = { nid, nid + 1, nid + nx + 1, nid + nx + 2 };
        for (int i = 0; i < 4; ++i) {
            c[i][0] = vx[ids[i]];
            c[i][1] = vy[ids[i]];
        }
        if (spherical)
            correctlongitude<4>(c);
    }

    //! ParametricMesh.hpp:276-299
    double area(size_t eid) const
    {
        const size_t n = eid2nid(eid);
        auto d2 = [&](size_t a, size_t b) { return SQR(vx[a] - vx[b]) + SQR(vy[a] - vy[b]); };
        const double a = d2(n, n + 1);
        const double b = d2(n + 1, n + nx + 2);
        const double c = d2(n + 1 + nx, n + 2 + nx);
        const double d = d2(n, n + nx + 1);
        const double e = d2(n, n + nx + 2);
        const double f = d2(n + 1, n + nx + 1);
        return 0.25 * sqrt(4.0 * e * f - SQR(b + d - a - c));
    }
    double h(size_t eid) const { return sqrt(area(eid)); }

    //! ParametricMesh.hpp:381-397
    void edgevector(size_t n1, size_t n2, double& dx, double& dy) const
    {
        dx = vx[n2] - vx[n1];
        dy = vy[n2] - vy[n1];
        if (spherical) {
            if (dx > 0.5 * M_PI)
                dx -= 2.0 * M_PI;
            if (dx < -0.5 * M_PI)
                dx += 2.0 * M_PI;
        }
    }
};

// ----------------------------------------------------------------------------------
// ParametricTools (ParametricTools.hpp:73-202)
// ----------------------------------------------------------------------------------

//! dxT<G>, dyT<G> (2 x G^2), ParametricTools.hpp:73-86
template <int G> void dxT(const Mesh& m, size_t eid, double (&out)[2][G * G])
{
    double c[4][2];
    m.coordinatesOfElement(eid, c);
    const auto& T = CGTab<1, G>::get();
    for (int k = 0; k < 2; ++k)
        for (int q = 0; q < G * G; ++q) {
            double s = 0;
            for (int i = 0; i < 4; ++i)
                s += c[i][k] * T.phix[i][q];
            out[k][q] = s;
        }
}
template <int G> void dyT(const Mesh& m, size_t eid, double (&out)[2][G * G])
{
    double c[4][2];
    m.coordinatesOfElement(eid, c);
    const auto& T = CGTab<1, G>::get();
    for (int k = 0; k < 2; ++k)
        for (int q = 0; q < G * G; ++q) {
            double s = 0;
            for (int i = 0; i < 4; ++i)
                s += c[i][k] * T.phiy[i][q];
            out[k][q] = s;
        }
}
//! J<G> (G^2), ParametricTools.hpp:92-105
template <int G> void jacobian(const Mesh& m, size_t eid, double (&J)[G * G])
{
    double a[2][G * G], b[2][G * G];
    dxT<G>(m, eid, a);
    dyT<G>(m, eid, b);
    for (int q = 0; q < G * G; ++q)
        J[q] = a[0][q] * b[1][q] - a[1][q] * b[0][q];
}
//! getGaussPointsInElement<G> (2 x G^2), ParametricTools.hpp:159-164
template <int G> void gaussPointsInElement(const Mesh& m, size_t eid, double (&gp)[2][G * G])
{
    double c[4][2];
    m.coordinatesOfElement(eid, c);
    const auto& T = CGTab<1, G>::get();
    for (int k = 0; k < 2; ++k)
        for (int q = 0; q < G * G; ++q) {
            double s = 0;
            for (int i = 0; i < 4; ++i)
                s += c[i][k] * T.phi[i][q];
            gp[k][q] = s;
        }
}

//! ParametricTools::massMatrix<DG> (:109-153) / SphericalTools::massMatrix<DG> (:189-202)
template <int DG> void massMatrix(const Mesh& m, size_t eid, bool coslat, double (&M)[DG][DG])
{
    constexpr int G = gp1d(DG), Q = G * G;
    const auto& T = DGTab<DG, G>::get();
    double J[Q], wj[Q];
    jacobian<G>(m, eid, J);
    if (coslat) {
        double gp[2][Q];
        gaussPointsInElement<G>(m, eid, gp);
        for (int q = 0; q < Q; ++q)
            wj[q] = T.w[q] * J[q] * cos(gp[1][q]);
    } else
        for (int q = 0; q < Q; ++q)
            wj[q] = T.w[q] * J[q];
    for (int i = 0; i < DG; ++i)
        for (int j = 0; j < DG; ++j) {
            double s = 0;
            for (int q = 0; q < Q; ++q)
                s += (T.psi[i][q] * wj[q]) * T.psi[j][q];
            M[i][j] = s;
        }
}

//! Dense inverse by LU with partial pivoting (what Eigen 3.4 `.inverse()` does for N > 4).
template <int N> void inverse(const double (&Ain)[N][N], double (&inv)[N][N])
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
            for (int j = 0; j < N; ++j)
                std::swap(A[k][j], A[p][j]);
            std::swap(piv[k], piv[p]);
        }
        for (int i = k + 1; i < N; ++i) {
            A[i][k] /= A[k][k];
            for (int j = k + 1; j < N; ++j)
                A[i][j] -= A[i][k] * A[k][j];
        }
    }
    for (int c = 0; c < N; ++c) {
        double y[N];
        for (int i = 0; i < N; ++i) {
            double s = (piv[i] == c) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j)
                s -= A[i][j] * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < N; ++j)
                s -= A[i][j] * inv[j][c];
            inv[i][c] = s / A[i][i];
        }
    }
}

} // namespace nso
