/*
 * nsdg_basis.cuh -- DG / CG basis functions and Gauss quadrature of the dynamics, as
 * constexpr __host__ __device__ functions.  Called with compile-time indices inside
 * fully unrolled loops they fold to immediates, so the per-element Gauss-point
 * contractions need no table memory and structurally-zero entries cost nothing.
 *
 * Same definitions as the reference generators (values, not code):
 *   DG basis {1, x-1/2, y-1/2, (x-1/2)^2-1/12, (y-1/2)^2-1/12, (x-1/2)(y-1/2),
 *             (y-1/2)((x-1/2)^2-1/12), (x-1/2)((y-1/2)^2-1/12)}   dynamics/codegeneration/basisfunctions.py:28-52
 *   Lagrange Q1/Q2 CG basis, x-fastest local numbering            basisfunctions.py:107-166
 *   Gauss rules on [0,1]                                          gaussquadrature.py:6-22
 *   table shapes PSI<DG,GP>, PSIe_w<DG,GP,E>                      dynamics/src/include/codeGenerationDGinGauss.hpp:105-981
 *                PHI<CG,GP>, PHIx, PHIy, PHI1d                      dynamics/src/include/codeGenerationCGinGauss.hpp:14-461
 */
#pragma once

#define NSDG_HD __host__ __device__ __forceinline__

namespace nsdg {

NSDG_HD constexpr int gp1d(int DG) { return (DG == 8 || DG == 6) ? 3 : (DG == 3 ? 2 : 1); }
NSDG_HD constexpr int edgedofs(int DG) { return DG == 1 ? 1 : (DG == 3 ? 2 : 3); }
NSDG_HD constexpr int cgdofs(int CG) { return CG == 1 ? 4 : 9; }
NSDG_HD constexpr int cg2dgstress(int CG) { return CG == 1 ? 3 : 8; }
constexpr double EarthRadius = 6371000.0; // dynamics/src/include/NextsimDynamics.hpp:35

// Gauss points / weights on [0,1]; literals are the correctly rounded values of
// 1/2 -+ sqrt(1/12), 1/2 -+ sqrt(3/20), 1/2 -+ 1/2 sqrt(3/7 -+ 2/7 sqrt(6/5)).
NSDG_HD constexpr double gausspoint(int G, int k)
{
    if (G == 1)
        return 0.5;
    if (G == 2)
        return k == 0 ? 0.21132486540518711774542560974902 : 0.78867513459481288225457439025098;
    if (G == 3)
        return k == 0 ? 0.11270166537925831148207346002176 : (k == 1 ? 0.5 : 0.88729833462074168851792653997824);
    return k == 0 ? 0.06943184420297371238802675555360
                  : (k == 1 ? 0.33000947820757186759866712044838
                            : (k == 2 ? 0.66999052179242813240133287955162 : 0.93056815579702628761197324444640));
}
NSDG_HD constexpr double gaussweight(int G, int k)
{
    if (G == 1)
        return 1.0;
    if (G == 2)
        return 0.5;
    if (G == 3)
        return k == 1 ? 8. / 18. : 5. / 18.;
    return (k == 0 || k == 3) ? 0.17392742256872692868653197461100 : 0.32607257743127307131346802538900;
}
//! weight of 2-d Gauss point q = qy*G + qx
NSDG_HD constexpr double gaussweight2(int G, int q) { return gaussweight(G, q % G) * gaussweight(G, q / G); }

NSDG_HD constexpr double dgbasis(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    return j == 0 ? 1.
        : j == 1  ? X
        : j == 2  ? Y
        : j == 3  ? X * X - 1.0 / 12.0
        : j == 4  ? Y * Y - 1.0 / 12.0
        : j == 5  ? X * Y
        : j == 6  ? Y * (X * X - 1.0 / 12.0)
                  : X * (Y * Y - 1.0 / 12.0);
}
NSDG_HD constexpr double dgbasis_dx(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    return j == 1 ? 1. : j == 3 ? 2.0 * X : j == 5 ? Y : j == 6 ? Y * (2. * X) : j == 7 ? Y * Y - 1.0 / 12.0 : 0.;
}
NSDG_HD constexpr double dgbasis_dy(int j, double x, double y)
{
    const double X = x - 0.5, Y = y - 0.5;
    return j == 2 ? 1. : j == 4 ? 2. * Y : j == 5 ? X : j == 6 ? X * X - 1.0 / 12.0 : j == 7 ? X * (2. * Y) : 0.;
}
NSDG_HD constexpr double dgbasis_edge(int j, double t)
{
    const double T = t - 0.5;
    return j == 0 ? 1. : (j == 1 ? T : T * T - 1.0 / 12.0);
}

NSDG_HD constexpr double cgbasis1d(int cg, int j, double x)
{
    return cg == 1 ? (j == 0 ? 1.0 - x : x)
                   : (j == 0 ? 2.0 * (x - 0.5) * (x - 1.0) : (j == 1 ? 4.0 * x * (1.0 - x) : 2.0 * x * (x - 0.5)));
}
NSDG_HD constexpr double cgbasis1d_dx(int cg, int j, double x)
{
    return cg == 1 ? (j == 0 ? -1. : 1.) : (j == 0 ? 4.0 * x - 3.0 : (j == 1 ? 4.0 - 8.0 * x : 4.0 * x - 1.0));
}
NSDG_HD constexpr double cgbasis(int cg, int j, double x, double y)
{
    return cgbasis1d(cg, j % (cg + 1), x) * cgbasis1d(cg, j / (cg + 1), y);
}
NSDG_HD constexpr double cgbasis_dx(int cg, int j, double x, double y)
{
    return cgbasis1d_dx(cg, j % (cg + 1), x) * cgbasis1d(cg, j / (cg + 1), y);
}
NSDG_HD constexpr double cgbasis_dy(int cg, int j, double x, double y)
{
    return cgbasis1d(cg, j % (cg + 1), x) * cgbasis1d_dx(cg, j / (cg + 1), y);
}

// ---- "table entries" (fold to immediates for constant arguments) ----
//! PSI<DG,G>(j,q): DG basis j in 2-d Gauss point q (x-fastest)
NSDG_HD constexpr double PSI(int G, int j, int q) { return dgbasis(j, gausspoint(G, q % G), gausspoint(G, q / G)); }
NSDG_HD constexpr double PSIx(int G, int j, int q) { return dgbasis_dx(j, gausspoint(G, q % G), gausspoint(G, q / G)); }
NSDG_HD constexpr double PSIy(int G, int j, int q) { return dgbasis_dy(j, gausspoint(G, q % G), gausspoint(G, q / G)); }
//! PSILagrange<DG,L>(j,q): DG basis j in Lagrange point q of the (L x L) lattice
NSDG_HD constexpr double PSILag(int L, int j, int q)
{
    return dgbasis(j, L == 2 ? double(q % 2) : 0.5 * (q % 3), L == 2 ? double(q / 2) : 0.5 * (q / 3));
}
//! PSIe<ED,G>(j,q): edge basis j in edge Gauss point q
NSDG_HD constexpr double PSIe(int G, int j, int q) { return dgbasis_edge(j, gausspoint(G, q)); }
//! PSIe_w<DG,G,E>(q,j) = w_q * psi_j(edge E point q); E: 0 bottom, 1 right, 2 top, 3 left
NSDG_HD constexpr double PSIew(int G, int E, int q, int j)
{
    const double g = gausspoint(G, q);
    return gaussweight(G, q) * (E == 0 ? dgbasis(j, g, 0.0) : E == 1 ? dgbasis(j, 1.0, g) : E == 2 ? dgbasis(j, g, 1.0) : dgbasis(j, 0.0, g));
}
NSDG_HD constexpr double PHI(int CG, int G, int j, int q) { return cgbasis(CG, j, gausspoint(G, q % G), gausspoint(G, q / G)); }
NSDG_HD constexpr double PHIx(int CG, int G, int j, int q) { return cgbasis_dx(CG, j, gausspoint(G, q % G), gausspoint(G, q / G)); }
NSDG_HD constexpr double PHIy(int CG, int G, int j, int q) { return cgbasis_dy(CG, j, gausspoint(G, q % G), gausspoint(G, q / G)); }

//! trace of a DG cell vector on one side in the edge basis (side: 0 bottom, 1 right, 2 top, 3 left).
//! Same linear maps as left/right/bottom/topedgeofcell, dynamics/src/DGTransport.cpp:44-142.
template <int DG, typename LOAD> NSDG_HD void edgeofcell(LOAD c, int side, double* out)
{
    const double sg = (side == 1 || side == 2) ? 0.5 : -0.5;
    const bool lr = (side == 1 || side == 3);
    if constexpr (DG == 1) {
        out[0] = c(0);
    } else if constexpr (DG == 3) {
        out[0] = c(0) + sg * (lr ? c(1) : c(2));
        out[1] = lr ? c(2) : c(1);
    } else if constexpr (DG == 6) {
        out[0] = c(0) + sg * (lr ? c(1) : c(2)) + 1. / 6. * (lr ? c(3) : c(4));
        out[1] = (lr ? c(2) : c(1)) + sg * c(5);
        out[2] = lr ? c(4) : c(3);
    } else {
        out[0] = c(0) + sg * (lr ? c(1) : c(2)) + 1. / 6. * (lr ? c(3) : c(4));
        out[1] = (lr ? c(2) : c(1)) + sg * c(5) + 1. / 6. * (lr ? c(6) : c(7));
        out[2] = (lr ? c(4) : c(3)) + sg * (lr ? c(7) : c(6));
    }
}

} // namespace nsdg
