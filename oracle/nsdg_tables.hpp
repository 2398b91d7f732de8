/*
 * nsdg_tables.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Basis-function / quadrature constants of neXtSIM_DG's dynamics, regenerated from
 * their defining formulas instead of being copied as literal tables.
 *
 * Reference (formulas):  dynamics/codegeneration/basisfunctions.py:28-166
 *                        dynamics/codegeneration/gaussquadrature.py:6-22
 * Reference (layouts):   dynamics/src/include/codeGenerationDGinGauss.hpp
 *                          GAUSSWEIGHTS:47-74, PSIe_w:105-344, PSILagrange:348-420,
 *                          PSI:424-603, PSIx/PSIy:605-885, PSIe:954-981
 *                        dynamics/src/include/codeGenerationCGinGauss.hpp
 *                          PHI/PHIx/PHIy:14-407, PHI1d:410-461
 *
 * Index conventions (same as the reference): Gauss/Lagrange points and local CG dofs
 * are x-fastest, q = qy*G + qx.  PSI[j][q] = j-th DG basis function in point q.
 * The product (nextsimdg_b200/csrc) has its own independent tables; nothing outside
 * tests/, smoke() and bench.py's cpu_baseline leg may include this file.
 */
#pragma once
#include <cmath>
#include <cstddef>

namespace nso {

constexpr int gp1d(int DG) { return (DG == 8 || DG == 6) ? 3 : (DG == 3 ? 2 : (DG == 1 ? 1 : -1)); }
constexpr int ngp(int DG) { return gp1d(DG) * gp1d(DG); }
constexpr int edgedofs(int DG) { return DG == 1 ? 1 : (DG == 3 ? 2 : 3); }
constexpr int cgdofs(int CG) { return CG == 1 ? 4 : 9; }
constexpr int cg2dgstress(int CG) { return CG == 1 ? 3 : 8; }
constexpr double EarthRadius = 6371000.0; // NextsimDynamics.hpp:35

// ---- quadrature on [0,1] (gaussquadrature.py:6-22) ----
inline double gausspoint(int G, int k)
{
    switch (G) {
    case 1:
        return 0.5;
    case 2:
        return k == 0 ? 0.5 - std::sqrt(1. / 12.) : 0.5 + std::sqrt(1. / 12.);
    case 3:
        return k == 0 ? 0.5 - std::sqrt(3. / 20.) : (k == 1 ? 0.5 : 0.5 + std::sqrt(3. / 20.));
    default: {
        const double a = 0.5 * std::sqrt(3. / 7. + 2. / 7. * std::sqrt(6. / 5.));
        const double b = 0.5 * std::sqrt(3. / 7. - 2. / 7. * std::sqrt(6. / 5.));
        return k == 0 ? 0.5 - a : (k == 1 ? 0.5 - b : (k == 2 ? 0.5 + b : 0.5 + a));
    }
    }
}
inline double gaussweight(int G, int k)
{
    switch (G) {
    case 1:
        return 1.0;
    case 2:
        return 0.5;
    case 3:
        return k == 1 ? 8. / 18. : 5. / 18.;
    default:
        return (k == 0 || k == 3) ? (18. - std::sqrt(30.0)) / 72. : (18. + std::sqrt(30.0)) / 72.;
    }
}
// Lagrange points: L=2 -> {0,1}, L=3 -> {0,1/2,1}
inline double lagrangepoint(int L, int k) { return L == 2 ? (k == 0 ? 0.0 : 1.0) : 0.5 * k; }

// ---- DG basis on the unit square (basisfunctions.py:28-100) ----
inline double dgbasis(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    switch (j) {
    case 0:
        return 1.;
    case 1:
        return X;
    case 2:
        return Y;
    case 3:
        return X * X - 1.0 / 12.0;
    case 4:
        return Y * Y - 1.0 / 12.0;
    case 5:
        return X * Y;
    case 6:
        return Y * (X * X - 1.0 / 12.0);
    default:
        return X * (Y * Y - 1.0 / 12.0);
    }
}
inline double dx_dgbasis(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    switch (j) {
    case 1:
        return 1.;
    case 3:
        return 2.0 * X;
    case 5:
        return Y;
    case 6:
        return Y * (2. * X);
    case 7:
        return Y * Y - 1.0 / 12.0;
    default:
        return 0.;
    }
}
inline double dy_dgbasis(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    switch (j) {
    case 2:
        return 1.;
    case 4:
        return 2. * Y;
    case 5:
        return X;
    case 6:
        return X * X - 1.0 / 12.0;
    case 7:
        return X * (2. * Y);
    default:
        return 0.;
    }
}
inline double dgbasis_edge(int j, double t)
{
    const double T = t - 0.5;
    return j == 0 ? 1. : (j == 1 ? T : T * T - 1.0 / 12.0);
}

// ---- Lagrange CG basis (basisfunctions.py:107-166) ----
inline double cgbasis1d(int cg, int j, double x)
{
    if (cg == 1)
        return j == 0 ? 1.0 - x : x;
    return j == 0 ? 2.0 * (x - 0.5) * (x - 1.0) : (j == 1 ? 4.0 * x * (1.0 - x) : 2.0 * x * (x - 0.5));
}
inline double cgbasis1d_dx(int cg, int j, double x)
{
    if (cg == 1)
        return j == 0 ? -1. : 1.;
    return j == 0 ? 4.0 * x - 3.0 : (j == 1 ? 4.0 - 8.0 * x : 4.0 * x - 1.0);
}
inline double cgbasis(int cg, int j, double x, double y)
{
    return cgbasis1d(cg, j % (cg + 1), x) * cgbasis1d(cg, j / (cg + 1), y);
}
inline double cgbasis_dx(int cg, int j, double x, double y)
{
    return cgbasis1d_dx(cg, j % (cg + 1), x) * cgbasis1d(cg, j / (cg + 1), y);
}
inline double cgbasis_dy(int cg, int j, double x, double y)
{
    return cgbasis1d(cg, j % (cg + 1), x) * cgbasis1d_dx(cg, j / (cg + 1), y);
}

// ---- the tables, built once per instantiation ----

//! PSI<DG,G>, PSIx, PSIy: DG x G^2, and GAUSSWEIGHTS<G>: G^2
template <int DG, int G> struct DGTab {
    double psi[DG][G * G], psix[DG][G * G], psiy[DG][G * G], w[G * G];
    DGTab()
    {
        for (int qy = 0; qy < G; ++qy)
            for (int qx = 0; qx < G; ++qx) {
                const int q = qy * G + qx;
                const double x = gausspoint(G, qx), y = gausspoint(G, qy);
                w[q] = gaussweight(G, qx) * gaussweight(G, qy);
                for (int j = 0; j < DG; ++j) {
                    psi[j][q] = dgbasis(j, x, y);
                    psix[j][q] = dx_dgbasis(j, x, y);
                    psiy[j][q] = dy_dgbasis(j, x, y);
                }
            }
    }
    static const DGTab& get()
    {
        static const DGTab t;
        return t;
    }
};

//! PSILagrange<DG,L>: DG x L^2
template <int DG, int L> struct LagTab {
    double psi[DG][L * L];
    LagTab()
    {
        for (int qy = 0; qy < L; ++qy)
            for (int qx = 0; qx < L; ++qx)
                for (int j = 0; j < DG; ++j)
                    psi[j][qy * L + qx] = dgbasis(j, lagrangepoint(L, qx), lagrangepoint(L, qy));
    }
    static const LagTab& get()
    {
        static const LagTab t;
        return t;
    }
};

//! PSIe<ED,G> (ED x G) and PSIe_w<DG,G,E> (G x DG, E = 0 bottom, 1 right, 2 top, 3 left)
template <int DG, int G> struct EdgeTab {
    static constexpr int ED = edgedofs(DG);
    double psie[ED][G];
    double psiew[4][G][DG];
    EdgeTab()
    {
        for (int q = 0; q < G; ++q) {
            const double g = gausspoint(G, q), wq = gaussweight(G, q);
            for (int j = 0; j < ED; ++j)
                psie[j][q] = dgbasis_edge(j, g);
            for (int j = 0; j < DG; ++j) {
                psiew[0][q][j] = wq * dgbasis(j, g, 0.0);
                psiew[1][q][j] = wq * dgbasis(j, 1.0, g);
                psiew[2][q][j] = wq * dgbasis(j, g, 1.0);
                psiew[3][q][j] = wq * dgbasis(j, 0.0, g);
            }
        }
    }
    static const EdgeTab& get()
    {
        static const EdgeTab t;
        return t;
    }
};

//! PHI<CG,G>, PHIx, PHIy: cgdofs x G^2
template <int CG, int G> struct CGTab {
    static constexpr int ND = cgdofs(CG);
    double phi[ND][G * G], phix[ND][G * G], phiy[ND][G * G];
    CGTab()
    {
        for (int qy = 0; qy < G; ++qy)
            for (int qx = 0; qx < G; ++qx) {
                const int q = qy * G + qx;
                const double x = gausspoint(G, qx), y = gausspoint(G, qy);
                for (int j = 0; j < ND; ++j) {
                    phi[j][q] = cgbasis(CG, j, x, y);
                    phix[j][q] = cgbasis_dx(CG, j, x, y);
                    phiy[j][q] = cgbasis_dy(CG, j, x, y);
                }
            }
    }
    static const CGTab& get()
    {
        static const CGTab t;
        return t;
    }
};

} // namespace nso
