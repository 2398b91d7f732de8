/*
 * nsdg_momentum_uniform.cuh -- the mEVP subcycle on a UNIFORM RECTANGULAR mesh (CG2 / DG8), the
 * configuration of BASELINE.json's headline metric, restructured around what is special there:
 *
 *  - all elements share one operator set, and on a rectangle dx x dy it factorises:
 *      iMgradX = (1/dx) Gx^, iMgradY = (1/dy) Gy^, iMJwPSI = B^ (mesh independent),
 *      divS1 = dy D1^, divS2 = dx D2^          (^ = unit-square matrices, compile-time constants)
 *    so the matrices live in the instruction stream as immediates and their structural zeros
 *    (32 of 72 entries in D1^/D2^, 17 of 72 in B^) cost nothing;
 *  - d(u,v)/d(x,y) of the Q2 velocity lies in the DG8 space (that is what DG8 = "grad Q2" is for),
 *    so the L2 projection of projectVelocityToStrain followed by the Gauss-point evaluation of
 *    stressUpdateHighOrder equals evaluating the velocity gradient directly in the 9 Gauss points
 *    (two 1-d contractions); the 3 x 8 strain coefficients are never formed;
 *  - everything in updateMomentum that does not change during the subcycles is folded into
 *    six per-node constants once per timestep (nodeconst_kernel), which removes 5 of 13 node
 *    reads and 3 of 4 divisions per node and subcycle.
 *
 * Same sweeps and same results up to rounding (re-association only; tests/test_gpu_parity.py):
 *   projectVelocityToStrain  dynamics/src/CGDynamicsKernel.cpp:300-337
 *   stressUpdateHighOrder    dynamics/src/include/MEVPStressUpdateStep.hpp:30-118
 *   stressDivergence         dynamics/src/CGDynamicsKernel.cpp:340-398
 *   updateMomentum           dynamics/src/include/VPCGDynamicsKernel.hpp:132-172 (quirks Q1, Q2 kept)
 *   applyBoundaries          dynamics/src/CGDynamicsKernel.cpp:439-444
 * The warp-strip / register-carry / deferred-line organisation is that of nsdg_momentum.cuh.
 */
#pragma once
#include "nsdg_momentum.cuh"

#include <cuda.h> // CUtensorMap (type only; the driver entry point is looked up at run time)
#include <utility>

namespace nsdg {

//! compile-time loop: f(std::integral_constant<int, 0>) ... f(<N-1>)
template <int N, typename F> __device__ __forceinline__ void static_for(F&& f)
{
    [&]<int... I>(std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I> {}), ...); }(
        std::make_integer_sequence<int, N> {});
}

//! unit-square operator tables of the CG2 / DG8 pair, evaluated at compile time
struct UnitOps {
    double L[3][3], Lp[3][3]; //!< Q2 1-d basis / derivative, [j][q] at the 3 Gauss points
    double B[8][9]; //!< psi_j(q) w_q / m_j           (= iMJwPSI on any rectangle)
    double D1[9][8], D2[9][8]; //!< int phi_i,xi psi_j ; int phi_i,eta psi_j (= divS1/dy, divS2/dx)
    constexpr UnitOps()
        : L {}
        , Lp {}
        , B {}
        , D1 {}
        , D2 {}
    {
        const double minv[8] = { 1., 12., 12., 180., 180., 144., 2160., 2160. }; // 1 / int psi_j^2
        for (int j = 0; j < 3; ++j)
            for (int q = 0; q < 3; ++q) {
                L[j][q] = cgbasis1d(2, j, gausspoint(3, q));
                Lp[j][q] = cgbasis1d_dx(2, j, gausspoint(3, q));
            }
        for (int j = 0; j < 8; ++j)
            for (int q = 0; q < 9; ++q)
                B[j][q] = clean(PSI(3, j, q) * gaussweight2(3, q) * minv[j]);
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 8; ++j) {
                double a = 0, b = 0;
                for (int q = 0; q < 9; ++q) {
                    a += gaussweight2(3, q) * PHIx(2, 3, i, q) * PSI(3, j, q);
                    b += gaussweight2(3, q) * PHIy(2, 3, i, q) * PSI(3, j, q);
                }
                D1[i][j] = clean(a);
                D2[i][j] = clean(b);
            }
    }
    //! entries that are zero in exact arithmetic come out as O(1e-17) from the quadrature sums
    static constexpr double clean(double x) { return (x < 1e-13 && x > -1e-13) ? 0.0 : x; }
};
inline constexpr UnitOps kUnitOps {};

/*
 * The same tables as two 1-d passes.  DG8 is the tensor product of {p0, p1, p2} = {1, t, t^2 - 1/12} (t = x - 1/2)
 * without the (2,2) mode -- coefficient j <-> (a, b): 0 (0,0), 1 (1,0), 2 (0,1), 3 (2,0), 4 (0,2), 5 (1,1), 6 (2,1),
 * 7 (1,2); DG6 is its first six -- and the 9 Gauss points are the product of the 1-d points t = -g, 0, +g.  So
 *   evaluation in the 9 points  (PSI<DG,3>, 55 / 43 multiply-adds as a table)  = 28 / 25 operations, 3 constants,
 *   the projection B (iMJwPSI on a rectangle, 55 / 43)                         = 40 / 36 operations, 4 constants,
 * using p1(+-g) = +-g, p2(+-g) = 1/15, p2(0) = -1/12 and 1 / int p_a^2 = 1, 12, 180.  Same linear maps, other grouping.
 */
namespace sep {
    inline constexpr double g = 0.5 - gausspoint(3, 0); //!< sqrt(3/20)
    inline constexpr double p2e = 1.0 / 15.0, p2c = -1.0 / 12.0;
    inline constexpr double we = 5.0 / 18.0, wc = 8.0 / 18.0;
    inline constexpr double o1 = 12.0 * we * g; //!< odd mode: 12 w_e g
    inline constexpr double q2 = 10.0 / 3.0; //!< quadratic mode: 180 w_e / 15 = 180 w_c / 24
}
//! 1-d pass: values in the points (-g, 0, +g) of c0 p0 + c1 p1 + c2 p2
NSDG_HD void sepEval1(double c0, double c1, double c2, double& vm, double& v0, double& vp)
{
    const double e = fma(sep::p2e, c2, c0), o = sep::g * c1;
    v0 = fma(sep::p2c, c2, c0);
    vm = e - o;
    vp = e + o;
}
NSDG_HD void sepEval1(double c0, double c1, double& vm, double& v0, double& vp)
{
    const double o = sep::g * c1;
    v0 = c0;
    vm = c0 - o;
    vp = c0 + o;
}
//! values of a DG8 (NC = 8) or DG6 (NC = 6) row in the 9 Gauss points, q = 3 qy + qx
template <int NC> NSDG_HD void evalGaussSep(const double (&c)[NC], double (&out)[9])
{
    static_assert(NC == 8 || NC == 6);
    double X0[3], X1[3], X2[3]; // x pass: coefficient of p_b(y) in the three x points
    sepEval1(c[0], c[1], c[3], X0[0], X0[1], X0[2]);
    if constexpr (NC == 8) {
        sepEval1(c[2], c[5], c[6], X1[0], X1[1], X1[2]);
        sepEval1(c[4], c[7], X2[0], X2[1], X2[2]);
    } else {
        sepEval1(c[2], c[5], X1[0], X1[1], X1[2]);
        X2[0] = X2[1] = X2[2] = c[4];
    }
#pragma unroll
    for (int qx = 0; qx < 3; ++qx)
        sepEval1(X0[qx], X1[qx], X2[qx], out[qx], out[3 + qx], out[6 + qx]);
}
//! 1-d pass of the projection: the three weighted moments of values in the points (-g, 0, +g)
NSDG_HD void sepProj1(double rm, double r0, double rp, double& m0, double& m1, double& m2)
{
    const double s = rm + rp, d = rp - rm;
    m0 = fma(sep::we, s, sep::wc * r0);
    m1 = sep::o1 * d;
    m2 = sep::q2 * fma(-2.0, r0, s);
}
//! DG8 / DG6 coefficients of the L2 projection of Gauss-point values on a rectangle (the table UnitOps::B)
template <int NC> NSDG_HD void projectSep(const double (&r)[9], double (&c)[NC])
{
    static_assert(NC == 8 || NC == 6);
    double Y0[3], Y1[3], Y2[3]; // y pass: moment b of the column qx
#pragma unroll
    for (int qx = 0; qx < 3; ++qx)
        sepProj1(r[qx], r[3 + qx], r[6 + qx], Y0[qx], Y1[qx], Y2[qx]);
    sepProj1(Y0[0], Y0[1], Y0[2], c[0], c[1], c[3]);
    if constexpr (NC == 8) {
        sepProj1(Y1[0], Y1[1], Y1[2], c[2], c[5], c[6]);
        const double s = Y2[0] + Y2[2];
        c[4] = fma(sep::we, s, sep::wc * Y2[1]);
        c[7] = sep::o1 * (Y2[2] - Y2[0]);
    } else {
        const double s = Y1[0] + Y1[2];
        c[2] = fma(sep::we, s, sep::wc * Y1[1]);
        c[5] = sep::o1 * (Y1[2] - Y1[0]);
        c[4] = fma(sep::we, Y2[0] + Y2[2], sep::wc * Y2[1]);
    }
}
//! Q2 interpolation in 1-d from the nodes t = -1/2, 0, 1/2 to the Gauss points -g, 0, +g: values (vm, v0, vp) and derivatives
//! (dm, d0, dp) of c0 L0 + c1 L1 + c2 L2, L0 = 2t^2 - t, L1 = 1 - 4t^2, L2 = 2t^2 + t: f = 2t^2 S + (1 - 4t^2) c1 + t D,
//! f' = 4t (S - 2 c1) + D with S = c0 + c2, D = c2 - c0 (the tables UnitOps::L, ::Lp; 10 operations instead of 14)
NSDG_HD void q2Values(double c0, double c1, double c2, double& vm, double& v0, double& vp)
{
    const double e = fma(2.0 * sep::g * sep::g, c0 + c2, (1.0 - 4.0 * sep::g * sep::g) * c1), o = sep::g * (c2 - c0);
    vm = e - o;
    v0 = c1;
    vp = e + o;
}
NSDG_HD void q2Derivs(double c0, double c1, double c2, double& dm, double& d0, double& dp)
{
    const double t = (4.0 * sep::g) * fma(-2.0, c1, c0 + c2);
    d0 = c2 - c0;
    dm = d0 - t;
    dp = d0 + t;
}

/*
 * The divergence tables likewise: D1[(c,d)][(a,b)] = A1[c][a] A0[d][b], D2 = A0[c][a] A1[d][b] with the 1-d integrals of the
 * Q2 node functions (nodes t = -1/2, 0, 1/2) against p_b, A0 = [[1/6, -1/12, 1/90], [2/3, 0, -1/45], [1/6, 1/12, 1/90]], and of
 * their derivatives, A1 = [[-1, 1/3, 0], [0, -2/3, 0], [1, 1/3, 0]] (p2 is orthogonal to the linear derivatives, so only the
 * modes a <= 1 of the differentiated direction contribute): 26 operations per table and stress component instead of 40.
 */
NSDG_HD void sepMass1(double c0, double c1, double c2, double& w0, double& w1, double& w2)
{
    const double e = fma(1.0 / 90.0, c2, (1.0 / 6.0) * c0), o = (1.0 / 12.0) * c1;
    w0 = e - o;
    w2 = e + o;
    w1 = fma(-1.0 / 45.0, c2, (2.0 / 3.0) * c0);
}
//! T (+)= scale * D s for one DG8 stress component: D = UnitOps::D1 (DIR = 0) or D2 (DIR = 1); scale3 = scale / 3;
//! T[k], k = 3 jy + jx as in the tables
template <int DIR, bool ACC> NSDG_HD void divergenceSep(const double (&s)[8], double scale, double scale3, double (&T)[9])
{
    double W0[3], W1[3]; // undifferentiated direction first: modes 0 and 1 of the differentiated one at the three nodes
    if constexpr (DIR == 0) {
        sepMass1(s[0], s[2], s[4], W0[0], W0[1], W0[2]);
        sepMass1(s[1], s[5], s[7], W1[0], W1[1], W1[2]);
    } else {
        sepMass1(s[0], s[1], s[3], W0[0], W0[1], W0[2]);
        sepMass1(s[2], s[5], s[6], W1[0], W1[1], W1[2]);
    }
#pragma unroll
    for (int n = 0; n < 3; ++n) {
        constexpr int st = (DIR == 0) ? 1 : 3; // stride of the differentiated direction in k
        const int k0 = (DIR == 0) ? 3 * n : n;
        if constexpr (ACC) {
            const double t = W1[n] * scale3;
            T[k0] = fma(-W0[n], scale, T[k0] + t);
            T[k0 + st] = fma(-2.0, t, T[k0 + st]);
            T[k0 + 2 * st] = fma(W0[n], scale, T[k0 + 2 * st] + t);
        } else {
            const double t = W1[n] * scale3;
            T[k0] = fma(-W0[n], scale, t);
            T[k0 + st] = -2.0 * t;
            T[k0 + 2 * st] = fma(W0[n], scale, t);
        }
    }
}

/*
 * Mask bytes (land mask of the element row, Dirichlet bytes of its two node lines) travel with the u, v staging group:
 * copied global -> shared by cp.async one element row ahead and read at the top of the row.  Typed register loads a row
 * ahead do not work: the compiler unpacks the bytes (PRMT) or evaluates `mask != 0` into a predicate right behind the
 * load, and every warp waits out the global latency at the top of each row (7 % of the stall samples of the BBM strip
 * kernel, 17 % of the parametric mEVP kernel).  An asynchronous copy has no register consumer.
 */
struct MaskStage {
    uint8_t LM[32]; //!< land mask of the strip's 32 elements
    uint8_t NM[2][64]; //!< node bytes of node lines 2 row, 2 row + 1, columns [64 sx, 64 sx + 64)
};
__device__ __forceinline__ void cpAsync4(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
//! lanes 0-15 / 16-31 copy the node bytes of the two node lines (4 bytes each), lanes 0-7 the land mask; rows are padded
//! (cgs to 16, nxs to 32), so a 4-byte chunk lies inside the row or is skipped as a whole
__device__ __forceinline__ void stageMasks(MaskStage& m, const uint8_t* landmask, const uint8_t* nodemask, const GridDims& g, int row, int sx, int lane)
{
    const int k = lane >> 4, j = 4 * (lane & 15), c = 64 * sx + j;
    if (c < g.cgs)
        cpAsync4(&m.NM[k][j], nodemask + size_t(2 * row + k) * g.cgs + c);
    if (lane < 8 && 32 * sx + 4 * lane < g.nxs)
        cpAsync4(&m.LM[4 * lane], landmask + size_t(row) * g.nxs + 32 * sx + 4 * lane);
}
//! Dirichlet bytes of the lane's two nodes of node line k: byte 0 = column 2 ex, byte 1 = column 2 ex + 1
__device__ __forceinline__ unsigned nodeMaskWord(const MaskStage& m, int k, int lane)
{
    return *reinterpret_cast<const unsigned short*>(&m.NM[k][2 * lane]);
}

//! 16-byte global load that stays where it is written (a plain or __ldg load is sunk to its first use by the compiler)
__device__ __forceinline__ double2 ldPinned2(const double* p)
{
    double2 v;
    asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

//! pull the line holding `p` into L2 (no register, no scoreboard): used one element row ahead
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

//! arguments of the uniform mEVP kernels
struct UniformArgs {
    GridDims g;
    int R, nsx, nsy;
    double *s11, *s12, *s22; //!< DG8 planes
    const double* Pa; //!< Gauss-point planes of P/alpha, P = P* h exp(-20(1-a))
    const uint8_t* landmask;
    double *u, *v;
    const double *cA, *rx, *ry, *uO, *vO, *ilm; //!< per-node constants (nodeconst_kernel)
    const double* vcon; //!< compact per-vertical-line copies of the node constants (vcon_kernel)
    const double* geo; //!< parametric fast path: kGeoPlanes geometry planes (nsdg_momentum_param.cuh)
    const uint8_t* nodemask;
    double *hbuf, *vbuf;
    double dx, dy; //!< element size
    double keep; //!< 1 - 1/alpha
    double beta, dtfc; //!< beta ; deltaT * fc
    double DeltaMin2;
    CUtensorMap tm[5]; //!< TMA staging: s11, s12, s22 (8 planes each), P / alpha (9), geometry (parametric meshes)
};

/*
 * Per-node constants of VPCGDynamicsKernel::updateMomentum (VPCGDynamicsKernel.hpp:147-170).  With
 *   c1 = rho_ice cgH / deltaT   (> 0: cgH is clamped to 1e-4 in prepareIteration)
 * the update  u_new = ( c1 beta u + R + cgA F_ocean |du_ocn| uO - c1 dt fc u + dStressX / lumpedmass ) / ( c1 (1+beta) + cgA F_ocean |du_ocn| ),
 * R = c1 u0 + cgA F_atm |ua| ua - rho_ice cgH g dSSH/dx, is divided through by c1, which leaves SIX constants per node:
 *   cA  = cgA F_ocean / c1,   rx = R_x / c1,   ry = R_y / c1,   ilm = 1 / (lumpedcgmass c1),   and uO, vO
 * (one array and 32 B per element and subcycle less than keeping c1).
 */
template <int DUMMY = 0>
__global__ void nodeconst_kernel(GridDims g, PhysParams p, double deltaT, const double* __restrict__ cgH,
    const double* __restrict__ cgA, const double* __restrict__ uA, const double* __restrict__ vA, const double* __restrict__ gx,
    const double* __restrict__ gy, const double* __restrict__ u0, const double* __restrict__ v0, const double* __restrict__ lm,
    double* __restrict__ cA, double* __restrict__ rx, double* __restrict__ ry, double* __restrict__ ilm)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const size_t n = size_t(t / g.cgnx) * g.cgs + (t % g.cgnx);
    const double H = cgH[n], A = cgA[n];
    const double k1 = p.rho_ice * H / deltaT, ik1 = 1.0 / k1;
    const double absatm = sqrt(uA[n] * uA[n] + vA[n] * vA[n]);
    cA[n] = A * p.F_ocean * ik1;
    rx[n] = (k1 * u0[n] + A * (p.F_atm * absatm * uA[n]) - p.rho_ice * H * p.gravity * gx[n]) * ik1;
    ry[n] = (k1 * v0[n] + A * (p.F_atm * absatm * vA[n]) - p.rho_ice * H * p.gravity * gy[n]) * ik1;
    ilm[n] = ik1 / lm[n];
}

/*
 * The nodes of the vertical deferred lines (every 64th node column) are visited column-wise by the lines kernels:
 * a strided walk through the row-major node arrays costs one 64-B DRAM atom per 8-B value.  Their six constants
 * and the Dirichlet flag are therefore gathered once per timestep into a compact array
 *     vcon[(k * nsx + line) * cgny + row],  k = 0..6,
 * so that only u and v remain strided (lines kernel: 394 -> ~190 MB per launch at 2048^2).
 */
constexpr int kNodeConsts = 6;
constexpr int kVconPlanes = kNodeConsts + 1;
template <int CG = 2>
__global__ void vcon_kernel(GridDims g, int nsx, const double* __restrict__ k0, const double* __restrict__ k1, const double* __restrict__ k2,
    const double* __restrict__ k3, const double* __restrict__ k4, const double* __restrict__ k5, const uint8_t* __restrict__ nodemask,
    double* __restrict__ vcon)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(nsx) * g.cgny)
        return;
    const int line = int(t / g.cgny), r = int(t % g.cgny);
    const int c = min(CG * 32 * (line + 1), CG * g.nx);
    const size_t n = size_t(r) * g.cgs + c;
    const double* src[kNodeConsts] = { k0, k1, k2, k3, k4, k5 };
#pragma unroll
    for (int k = 0; k < kNodeConsts; ++k)
        vcon[(size_t(k) * nsx + line) * g.cgny + r] = src[k][n];
    vcon[(size_t(kNodeConsts) * nsx + line) * g.cgny + r] = (nodemask[n] & 1) ? 1.0 : 0.0;
}

/*
 * Branch-free reciprocal and square root to <= 1-2 ulp: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps.
 * The IEEE-rounded `/` and sqrt() of CUDA carry a range check and a slow path per call; in the Gauss-point laws and
 * the node updates they cost more than the arithmetic around them (BBM strip kernel 1.58 -> 1.26 ms at 2048^2).
 * Callers guarantee x != 0 (or discard the result by a select).
 */
__device__ __forceinline__ double fastRcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
//! the same, pinned where it stands (volatile): the compiler cannot sink it into a branch around a select that discards it
__device__ __forceinline__ double fastRcpPinned(double x)
{
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fastSqrt(double x) { return x > 0.0 ? x * rsqrt(x) : 0.0; }
/*
 * Branch-free square root: hardware seed (MUFU.RSQ64H, ~2^-22), one coupled Newton step on (sqrt x, 1 / (2 sqrt x)) and a
 * residual correction -- 7 multiply-adds, no slow-path call, so the scheduler may interleave neighbouring Gauss points /
 * nodes across it (the library rsqrt() brings a BSSY / BRA / CALL / BSYNC group that fences them).  Faithful to < 1 ulp for
 * normal x; zero, negative and subnormal arguments give 0 (callers pass sums of squares and elasticities).
 */
__device__ __forceinline__ double sqrtBranchFree(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    g = fma(fma(-g, g, x), h, g);
    return x >= 2.2250738585072014e-308 ? g : 0.0;
}

//! branch-free reciprocal root for normal positive x: hardware seed and the third-order step of the library rsqrt(), without
//! its exponent check and slow-path call (5 multiply-adds)
__device__ __forceinline__ double rsqrtBranchFree(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(x, -(y * y), 1.0);
    return fma(fma(0.375, e, 0.5), y * e, y);
}
#ifndef NSDG_MEVP_SQRT
#define NSDG_MEVP_SQRT 0 //!< bit 0: branch-free root in the mEVP node update, bit 1: branch-free 1 / Delta in the mEVP law
#endif

//! mEVP momentum update of one node from the per-node constants (+ Dirichlet)
__device__ __forceinline__ void momentumNodeUniform(const UniformArgs& a, double cA, double rx, double ry, double uO, double vO,
    double ilm, bool dirichlet, double un, double vn, double dSx, double dSy, double& unew, double& vnew)
{
    const double uOcnRel = uO - un;
    const double vOcnRel = vn - vO;
#if NSDG_MEVP_SQRT & 1
    const double absocn = sqrtBranchFree(uOcnRel * uOcnRel + vOcnRel * vOcnRel);
#else
    const double absocn = fastSqrt(uOcnRel * uOcnRel + vOcnRel * vOcnRel);
#endif
    const double drag = cA * absocn;
    const double inv = fastRcp((1.0 + a.beta) + drag);
    unew = inv * (a.beta * un + rx + drag * uO - a.dtfc * un + dSx * ilm);
    vnew = inv * (a.beta * vn + ry + drag * vO + a.dtfc * vn + dSy * ilm);
    if (dirichlet) {
        unew = 0.0;
        vnew = 0.0;
    }
}

// ---- cp.async (LDGSTS): global -> shared without passing through registers ----
__device__ __forceinline__ void cpAsync8(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem));
}
__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem));
}
//! 16 B, L2 only (bypasses L1): the form that reaches the copy bandwidth of HBM; the 8-byte .ca form saturates at
//! ~5.1 TB/s whatever the occupancy (scripts/microbench/stage_bw.cu: 5.09 vs 6.35 TB/s read+write on a B200)
__device__ __forceinline__ void cpAsync16cg(void* smem, const void* gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(smem))), "l"(gmem));
}
/*
 * Warp-cooperative staging of NPL plane rows (32 consecutive doubles = 256 B of each plane) global -> shared:
 * lanes 0-15 copy plane p, lanes 16-31 plane p+1, 16 B each.  `first` is the element index of the strip's lane 0
 * in the row.  Consumers read slots copied by OTHER lanes: __syncwarp() after the wait, and before a refill.
 * COOP = false is the per-lane form (8 B, .ca, every lane copies and reads only its own slots, no barriers): the
 * register-bound BBM kernels and the Cartesian parametric kernels are faster with it (scripts/quickbench_all.sh).
 */
template <int NPL, bool COOP = true>
__device__ __forceinline__ void stagePlanes(double (*dst)[32], const double* __restrict__ src, size_t pitch, size_t first, int lane)
{
    if constexpr (COOP) {
        const int half = lane >> 4, k2 = 2 * (lane & 15);
#pragma unroll
        for (int p = 0; p < NPL; p += 2) {
            const int pp = p + half;
            if ((NPL % 2 == 0) || pp < NPL)
                cpAsync16cg(&dst[pp][k2], src + size_t(pp) * pitch + first + k2);
        }
    } else {
#pragma unroll
        for (int p = 0; p < NPL; ++p)
            cpAsync8(&dst[p][lane], src + size_t(p) * pitch + first + lane);
    }
}
template <bool COOP> __device__ __forceinline__ void stageBarrier()
{
    if constexpr (COOP)
        __syncwarp();
}
/*
 * TMA staging.  A staging group (the 8 planes of a stress component, the 9 Gauss-point planes of a coefficient field, ...)
 * is a 32-element x NC-plane tile of the 2-d tensor {Npad elements (contiguous), NC planes (pitch Npad)}: ONE
 * cp.async.bulk.tensor instruction (SASS: UTMALDG), issued by lane 0, moves the tile global -> shared as [NC][32] doubles and
 * reports its bytes to an mbarrier, where the per-lane cp.async form issues one copy and one 64-bit address computation per
 * PLANE and lane (57 per element row in the uniform BBM kernel).  The barrier has one arrival (lane 0's arrive.expect_tx with
 * the group's byte count); consumers spin on try_wait.parity, the phase flips once per element row.  The non-tensor form
 * (cp.async.bulk, one 256-byte row per instruction) was tried first and does NOT save instructions: its operands are uniform
 * registers, so per-lane addresses compile into an ELECT / R2UR / UBLKCP loop over the lanes (8 instructions per plane).
 * Tensor maps are built on the host (planeTensorMap, nsdg_cuda.cu) and travel inside the kernel's __grid_constant__ arguments.
 */
__device__ __forceinline__ unsigned smemAddr(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* b, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(b)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(b)), "r"(bytes) : "memory");
}
//! wait for the phase with the given parity.  Bounded: a tensor copy that never completes (a bad tensor map, a device fault)
//! traps after ~2^26 polls -- some seconds -- instead of hanging the GPU; the host then sees a launch failure
__device__ __forceinline__ void mbarWait(uint64_t* b, unsigned parity)
{
    unsigned done;
    asm volatile("{\n.reg .pred p;\n.reg .u32 n;\nmov.u32 n, 0;\n"
                 "W%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n@p bra D%=;\n"
                 "add.u32 n, n, 1;\nsetp.lt.u32 p, n, 0x4000000;\n@p bra W%=;\nsetp.ne.u32 p, n, n;\n"
                 "D%=: selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smemAddr(b)), "r"(parity)
                 : "memory");
    if (!done)
        __trap();
}
//! tile (x .. x + 31, all planes) of a plane tensor -> dst[NC][32]; dst 128-byte aligned
__device__ __forceinline__ void tmaLoadTile(void* dst, const CUtensorMap* map, int x, uint64_t* b)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smemAddr(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(0), "r"(smemAddr(b))
                 : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/*
 * Per-warp staging buffer in shared memory.  Every lane copies (cp.async) and later reads ONLY its own
 * slots, so no cross-lane synchronisation is needed: shared memory acts as an asynchronous extension
 * of the register file that is filled one element row ahead of use, region by region:
 *   P   9 x 32 doubles      Gauss-point P/alpha of the row          refilled right after the VP law
 *   S  24 x 32 doubles      the three DG8 stresses of the row       refilled after the projection
 *   ND  2 x 6 x 32 double2  the six node constants, 2 node rows     refilled after the momentum update
 *   UV  2 x 2 x 32 double2  u, v of the two upper node rows         refilled at the top of the row
 * Four groups are always in flight; cp.async groups retire in order, so "wait_group 3" before each
 * region's first read is exactly "the group issued one row ago has landed".
 */
/*
 * NSDG_*_DIRECT_ND = 1: the per-node constants are not staged but loaded where the node update uses them (their lines
 * are pulled into L2 one row ahead).  6-8 KB less shared memory per warp moves the SM's carve-out one step towards L1,
 * which is what bounds the bytes in flight of the per-lane 8-byte cp.async.ca copies (they allocate L1 lines while they
 * wait): BBM strip kernel 1.16 -> 1.06 ms at 2048^2.  Kernels whose planes travel by cp.async.cg do not gain.
 */
#ifndef NSDG_UMEVP_DIRECT_ND
#define NSDG_UMEVP_DIRECT_ND 0
#endif
struct UmevpStage {
    double P[9][32];
    double S[24][32];
#if !NSDG_UMEVP_DIRECT_ND
    double2 ND[2][kNodeConsts][32];
#endif
    double2 UV[2][2][32];
    MaskStage M;
    double UVr[2][2]; //!< right-most node column of the strip (lane 31 / last element of the row)
    uint64_t bar[2]; //!< mbarriers of the P and S groups (TMA staging)
    double pad[6];
};
static_assert(sizeof(UmevpStage) % 128 == 0 && offsetof(UmevpStage, S) % 128 == 0, "TMA destinations need 128-byte alignment");
#ifndef NSDG_UMEVP_TMA
#define NSDG_UMEVP_TMA 0 //!< P and S rows by TMA (one tensor copy per field and row) instead of cooperative 16-byte cp.async.cg
#endif
constexpr bool kUmevpTma = NSDG_UMEVP_TMA != 0;
#ifndef NSDG_UMEVP_WARPS
#define NSDG_UMEVP_WARPS 4
#endif
constexpr int kUmevpWarps = NSDG_UMEVP_WARPS;
constexpr size_t kUmevpSmemBytes = sizeof(UmevpStage) * kUmevpWarps;

#ifndef NSDG_UMEVP_MINBLOCKS
#define NSDG_UMEVP_MINBLOCKS 3
#endif
/*
 * Deferred-line nodes (strip boundaries: every 32nd element column, every R-th element row, the domain's top / right edge).
 * A node of horizontal line L (node row min(2 R L, 2 ny)) or vertical line L (node column min(64 L, 2 nx)) sums the raw
 * contributions its 2 - 4 elements left in hbuf / vbuf, always in the same order, and is advanced like any other node.
 */
//! contributions of one deferred node; horizontal: cr = node column c on line L, vertical: cr = node row r on line L
template <int CG = 2, class ARGS>
__device__ __forceinline__ void lineNodeSum(const ARGS& a, bool horizontal, int L, int cr, int& r, int& c, double& sumX, double& sumY)
{
    constexpr int NR = CG + 1;
    const GridDims& g = a.g;
    sumX = 0.0;
    sumY = 0.0;
    if (horizontal) {
        c = cr;
        r = min(CG * a.R * L, CG * g.ny);
        const int jx = c % CG, exr = c / CG;
        const bool above = r < CG * g.ny;
        auto add = [&](int side, int ex, int j) {
            const double2 t = __ldcg(reinterpret_cast<const double2*>(a.hbuf + ((size_t(L - 1) * 2 + side) * g.nx + ex) * (NR * 2) + j * 2));
            sumX += t.x;
            sumY += t.y;
        };
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !above)
                break;
            if (jx == 0 && exr > 0)
                add(side, exr - 1, CG);
            if (exr < g.nx)
                add(side, exr, jx);
        }
    } else {
        r = cr;
        c = min(CG * 32 * L, CG * g.nx);
        const int jy = r % CG, eyr = r / CG;
        const bool right = c < CG * g.nx;
        auto add = [&](int side, int ey, int j) {
            const double2 t = __ldcg(reinterpret_cast<const double2*>(a.vbuf + ((size_t(L - 1) * 2 + side) * g.ny + ey) * (NR * 2) + j * 2));
            sumX += t.x;
            sumY += t.y;
        };
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !right)
                break;
            if (jy == 0 && eyr > 0)
                add(side, eyr - 1, CG);
            add(side, eyr, jy);
        }
    }
}
//! node constants of a deferred node: rows of the node arrays (horizontal) or the compact per-line copies (vertical, vcon_kernel)
template <class ARGS>
__device__ __forceinline__ void lineNodeConsts(const ARGS& a, const double* const (&src)[kNodeConsts], bool horizontal, int L, int r, size_t n,
    double (&k)[kNodeConsts], bool& d)
{
    const GridDims& g = a.g;
    if (horizontal) {
#pragma unroll
        for (int i = 0; i < kNodeConsts; ++i)
            k[i] = __ldg(src[i] + n);
        d = __ldg(a.nodemask + n) & 1;
    } else {
        const size_t m = size_t(L - 1) * g.cgny + r, pitch = size_t(a.nsx) * g.cgny;
#pragma unroll
        for (int i = 0; i < kNodeConsts; ++i)
            k[i] = __ldg(a.vcon + i * pitch + m);
        d = __ldg(a.vcon + kNodeConsts * pitch + m) != 0.0;
    }
}
//! one deferred node of the mEVP paths (uniform and parametric)
template <int CG = 2> __device__ __forceinline__ void lineNodeMEVP(const UniformArgs& a, bool horizontal, int L, int cr)
{
    const GridDims& g = a.g;
    int r, c;
    if (horizontal) {
        c = cr;
        r = min(CG * a.R * L, CG * g.ny);
    } else {
        r = cr;
        c = min(CG * 32 * L, CG * g.nx);
    }
    const size_t n = size_t(r) * g.cgs + c;
    // request everything that does not depend on the contributions first (node constants, Dirichlet flag, u, v)
    double k[kNodeConsts];
    bool d;
    const double* const src[kNodeConsts] = { a.cA, a.rx, a.ry, a.uO, a.vO, a.ilm };
    lineNodeConsts(a, src, horizontal, L, r, n, k, d);
    const double uOld = a.u[n], vOld = a.v[n];
    double sumX, sumY;
    lineNodeSum<CG>(a, horizontal, L, cr, r, c, sumX, sumY);
    double un, vn;
    momentumNodeUniform(a, k[0], k[1], k[2], k[3], k[4], k[5], d, uOld, vOld, d ? 0.0 : -sumX, d ? 0.0 : -sumY, un, vn);
    a.u[n] = un;
    a.v[n] = vn;
}

//! a thread of a lines kernel -> its deferred node.  Grid: y = line (the nsy horizontal lines, then the nsx vertical ones),
//! x * blockDim + thread = position along the line (no integer divisions); false for positions beyond the line and for the
//! vertical-line slots that lie on a horizontal line
template <int CG = 2, class ARGS> __device__ __forceinline__ bool lineNodeOfThread(const ARGS& a, bool& horizontal, int& L, int& cr)
{
    const GridDims& g = a.g;
    cr = int(blockIdx.x * blockDim.x + threadIdx.x);
    horizontal = int(blockIdx.y) < a.nsy;
    if (horizontal) {
        L = int(blockIdx.y) + 1;
        return cr < g.cgnx;
    }
    L = int(blockIdx.y) - a.nsy + 1;
    if (cr >= g.cgny)
        return false;
    return !(cr > 0 && (cr % (CG * a.R) == 0 || cr == CG * g.ny));
}
//! launch geometry of a lines kernel
inline dim3 linesGrid(const GridDims& g, int nsx, int nsy) { return dim3(unsigned((max(g.cgnx, g.cgny) + 127) / 128), unsigned(nsx + nsy)); }

//! deferred-line nodes for the uniform mEVP path as a kernel of their own (see subcycle_lines in nsdg_momentum.cuh)
template <int CG = 2>
__global__ void __launch_bounds__(128) subcycle_lines_umevp(const __grid_constant__ UniformArgs a)
{
    bool horizontal;
    int L, cr;
    if (lineNodeOfThread<CG>(a, horizontal, L, cr))
        lineNodeMEVP<CG>(a, horizontal, L, cr);
}

template <int DUMMY = 0>
__global__ void __launch_bounds__(32 * kUmevpWarps, NSDG_UMEVP_MINBLOCKS) subcycle_strip_umevp(const __grid_constant__ UniformArgs a)
{
    constexpr int CG = 2, NR = 3, DGs = 8;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    UmevpStage& st = reinterpret_cast<UmevpStage*>(smemRaw)[threadIdx.x >> 5];
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
    const double idx = 1.0 / a.dx, idy = 1.0 / a.dy;

    // ---- issue functions of the four staging groups for element row `row` (empty group past the strip) ----
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
    unsigned phaseP = 0, phaseS = 0;
    if constexpr (kUmevpTma) {
        if (lane == 0) {
            mbarInit(&st.bar[0], 1);
            mbarInit(&st.bar[1], 1);
        }
        mbarInitFence();
        __syncwarp();
    }
    auto issueP = [&](int row) {
        __syncwarp(); // every lane has consumed the region that is refilled
        if constexpr (kUmevpTma) {
            if (row < ey1 && lane == 0) {
                mbarExpectTx(&st.bar[0], 9 * 256);
                tmaLoadTile(&st.P[0][0], &a.tm[3], row * g.nxs + 32 * sx, &st.bar[0]);
            }
            return;
        }
        if (row < ey1)
            stagePlanes<9>(st.P, a.Pa, Npad, size_t(row) * g.nxs + 32 * sx, lane);
        cpAsyncCommit();
    };
    auto issueS = [&](int row) {
        __syncwarp();
        if constexpr (kUmevpTma) {
            if (row < ey1 && lane == 0) {
                const int x = row * g.nxs + 32 * sx;
                mbarExpectTx(&st.bar[1], 24 * 256);
                tmaLoadTile(&st.S[0][0], &a.tm[0], x, &st.bar[1]);
                tmaLoadTile(&st.S[8][0], &a.tm[1], x, &st.bar[1]);
                tmaLoadTile(&st.S[16][0], &a.tm[2], x, &st.bar[1]);
            }
            return;
        }
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<8>(st.S, a.s11, Npad, first, lane);
            stagePlanes<8>(st.S + 8, a.s12, Npad, first, lane);
            stagePlanes<8>(st.S + 16, a.s22, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueND = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + k) * g.cgs + col0;
#if NSDG_UMEVP_DIRECT_ND
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
    // prologue: the four groups of the first row; the bottom node row comes by plain loads
    issueUV(ey0);
    issueP(ey0);
    issueS(ey0);
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

    // mask bytes travel one element row ahead in registers (their use right after the load cost 18 % of the
    // stall samples, profiles/r1_strip_v3_cpasync.txt)

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
        // ---- the two upper node rows of u, v from the staging buffer ----
        cpAsyncWait<kUmevpTma ? 1 : 3>(); // (TMA staging: only the UV and ND groups are cp.async groups)
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

        // ---- velocity gradient in the 9 Gauss points by two 1-d contractions ----
        double e11[9], e12[9], e22[9];
        {
            double Au[3][3], Adu[3][3], Av[3][3], Adv[3][3]; // [jy][qx]
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
            static_for<3>([&](auto QY) {
                static_for<3>([&](auto QX) {
                    constexpr int qy = decltype(QY)::value, qx = decltype(QX)::value, q = qy * 3 + qx;
                    double ux = 0, uy = 0, vx = 0, vy = 0;
                    static_for<3>([&](auto JY) {
                        constexpr int jy = decltype(JY)::value;
                        constexpr double l = kUnitOps.L[jy][qy], lp = kUnitOps.Lp[jy][qy];
                        if constexpr (l != 0.0) {
                            ux = fma(l, Adu[jy][qx], ux);
                            vx = fma(l, Adv[jy][qx], vx);
                        }
                        if constexpr (lp != 0.0) {
                            uy = fma(lp, Au[jy][qx], uy);
                            vy = fma(lp, Av[jy][qx], vy);
                        }
                    });
                    // land elements keep zero strain (quirk Q8)
                    e11[q] = ice ? ux * idx : 0.0;
                    e22[q] = ice ? vy * idy : 0.0;
                    e12[q] = ice ? 0.5 * (uy * idy + vx * idx) : 0.0;
                });
            });
        }

        // ---- VP law in the Gauss points: e** become the integrands r** (MEVPStressUpdateStep.hpp:62-117) ----
        if constexpr (kUmevpTma) {
            mbarWait(&st.bar[0], phaseP);
            phaseP ^= 1u;
        } else {
            cpAsyncWait<3>();
            __syncwarp(); // P was staged cooperatively
        }
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
        });
        issueP(ey + 1);

        // ---- per stress component: project, relax, store, accumulate the divergence contributions ----
        double Tx[9], Ty[9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
            Tx[k] = Ty[k] = 0.0;
        if constexpr (kUmevpTma) {
            mbarWait(&st.bar[1], phaseS);
            phaseS ^= 1u;
        } else {
            cpAsyncWait<3>();
            __syncwarp(); // S was staged cooperatively
        }
        auto component = [&](double* plane, const double (&r)[9], auto COMP) {
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
            if (ice) {
                static_for<9>([&](auto K) {
                    constexpr int k = decltype(K)::value;
                    double d1 = 0.0, d2 = 0.0;
                    static_for<DGs>([&](auto J) {
                        constexpr int j = decltype(J)::value;
                        constexpr double c1 = kUnitOps.D1[k][j], c2 = kUnitOps.D2[k][j];
                        if constexpr (c1 != 0.0 && comp != 2)
                            d1 = fma(c1, s[j], d1);
                        if constexpr (c2 != 0.0 && comp != 0)
                            d2 = fma(c2, s[j], d2);
                    });
                    // tx = divS1 s11 + divS2 s12 ; ty = divS1 s12 + divS2 s22
                    if constexpr (comp == 0)
                        Tx[k] = fma(d1, a.dy, Tx[k]);
                    if constexpr (comp == 1) {
                        Tx[k] = fma(d2, a.dx, Tx[k]);
                        Ty[k] = fma(d1, a.dy, Ty[k]);
                    }
                    if constexpr (comp == 2)
                        Ty[k] = fma(d2, a.dx, Ty[k]);
                });
            }
        };
        component(a.s11, e11, std::integral_constant<int, 0> {});
        component(a.s12, e12, std::integral_constant<int, 1> {});
        component(a.s22, e22, std::integral_constant<int, 2> {});
        issueS(ey + 1);

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
        cpAsyncWait<kUmevpTma ? 1 : 3>();
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
#if NSDG_UMEVP_DIRECT_ND
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

} // namespace nsdg
