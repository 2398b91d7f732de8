/*
 * nsdg_momentum_uniform_bbm.cuh -- the BBM subcycle on a UNIFORM RECTANGULAR mesh (CG2 / DG8 / DG6),
 * organised like the mEVP kernel of nsdg_momentum_uniform.cuh: warp strips, register carry, deferred
 * lines, cp.async staging one element row ahead, compile-time unit-square operators, direct
 * Gauss-point evaluation of the velocity gradient, per-node constants.
 *
 * Sweeps reproduced (results equal to rounding; re-association + x/y -> x * rcp(y) only):
 *   projectVelocityToStrain  dynamics/src/CGDynamicsKernel.cpp:300-337
 *   stressUpdateHighOrder    dynamics/src/include/BBMStressUpdateStep.hpp:31-196
 *   stressDivergence         dynamics/src/CGDynamicsKernel.cpp:340-398
 *   updateMomentum           dynamics/src/include/BrittleCGDynamicsKernel.hpp:206-254 (quirk Q3 kept)
 *   applyBoundaries          dynamics/src/CGDynamicsKernel.cpp:439-444
 * Once per timestep (constant over the subcycles):
 *   Gauss points: h = max(h_q,0), expC = exp(C(1-a)), Pmax = P0 h^2.5 expC     (BBMStressUpdateStep.hpp:54-57,76-90)
 *   nodes: dte = deltaT/(rho cgH), cA = cgA F_ocean, ax = cgA F_atm |ua| ua - rho cgH g dSSH/dx, ay, 1/lumped mass
 */
#pragma once
#include "nsdg_momentum_uniform.cuh"

namespace nsdg {

struct UniformBBMArgs {
    GridDims g;
    int R, nsx, nsy;
    double *s11, *s12, *s22; //!< DG8 planes
    double* damage; //!< DG6 planes
    const double *gH, *gE, *gP; //!< Gauss-point planes: h, expC, Pmax
    const uint8_t* landmask;
    double *u, *v, *avgU, *avgV;
    const double *cA, *ax, *ay, *uO, *vO, *ilm; //!< per-node constants (all pre-multiplied by dte = deltaT / (rho cgH))
    const uint8_t* nodemask;
    double *hbuf, *vbuf;
    double dx, dy;
    double deltaT, dtfc, invNSteps, cosA, sinA;
    // rheology constants (MEBParameters.hpp:40-65) and their per-mesh products
    double young, nu0, lambda0, tan_phi;
    double cohScale, comprScale; //!< C_lab * sqrt(0.1/h_el), compr_strength * sqrt(0.1/h_el)
    double invTdK; //!< 1 / (h_el sqrt(2 (1+nu) rho_ice)):  1/td = sqrt(elasticity) * invTdK
    double dunitK; //!< deltaT / (1 - nu^2)
    // parametric fast path (nsdg_momentum_param.cuh): per-element geometry planes; h_el varies per element
    const double* geo;
    const double* vcon; //!< compact per-vertical-line node constants (vcon_kernel)
    double* vavg; //!< compact per-vertical-line mean velocities during the subcycle loop: avgU [nsx][cgny], then avgV (vavg_kernel)
    double C_lab, compr_strength;
    CUtensorMap tm[8]; //!< TMA staging: s11, s12, s22 (8 planes each), damage (6), h, expC, Pmax (9 each), geometry (parametric)
};

//! per-node constants of BrittleCGDynamicsKernel::updateMomentum (BrittleCGDynamicsKernel.hpp:209-240); the factor
//! dte = deltaT / (rho_ice cgH) that multiplies every one of them in the update is folded in (six arrays instead of seven)
template <int DUMMY = 0>
__global__ void nodeconst_bbm_kernel(GridDims g, PhysParams p, double deltaT, const double* __restrict__ cgH,
    const double* __restrict__ cgA, const double* __restrict__ uA, const double* __restrict__ vA, const double* __restrict__ gx,
    const double* __restrict__ gy, const double* __restrict__ lm, double* __restrict__ cA, double* __restrict__ ax,
    double* __restrict__ ay, double* __restrict__ ilm)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= long(g.cgnx) * g.cgny)
        return;
    const size_t n = size_t(t / g.cgnx) * g.cgs + (t % g.cgnx);
    const double H = cgH[n], A = cgA[n];
    const double dragAtm = A * p.F_atm * hypot(uA[n], vA[n]);
    const double dte = deltaT / (p.rho_ice * H);
    cA[n] = dte * (A * p.F_ocean);
    ax[n] = dte * (dragAtm * uA[n] - p.rho_ice * H * p.gravity * gx[n]);
    ay[n] = dte * (dragAtm * vA[n] - p.rho_ice * H * p.gravity * gy[n]);
    ilm[n] = dte / lm[n];
}

//! Gauss-point constants of the BBM law: h, exp(C(1-a)), Pmax (uniform path: one extra plane saves the pow)
template <int DGA, int GS>
__global__ void gaussconst_bbm3_kernel(GridDims g, PhysParams p, const double* __restrict__ hice, const double* __restrict__ cice,
    double* __restrict__ outH, double* __restrict__ outE, double* __restrict__ outP)
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
        const double expC = exp(p.compaction_param * (1.0 - aq));
        outH[size_t(q) * g.Npad + e] = hq;
        outE[size_t(q) * g.Npad + e] = expC;
        outP[size_t(q) * g.Npad + e] = p.P0 * pow(hq, p.exponent_compression_factor + 1.) * expC;
    }
}

//! brittle momentum update of one node from the per-node constants; returns the pre-boundary value for the mean
__device__ __forceinline__ void momentumNodeUniformBBM(const UniformBBMArgs& a, double cA, double ax, double ay, double uO, double vO,
    double ilm, bool dirichlet, double un, double vn, double dSx, double dSy, double& unew, double& vnew, double& uAvg, double& vAvg)
{
    const double du = uO - un, dv = vO - vn;
#if NSDG_BBM_SQRT & 1
    const double cPrime = cA * sqrtBranchFree(du * du + dv * dv); // dte * cPrime of the reference
#else
    const double cPrime = cA * fastSqrt(du * du + dv * dv); // dte * cPrime of the reference
#endif
    const double alpha = 1.0 + cPrime * a.cosA;
    const double beta = a.dtfc + cPrime * a.sinA;
    const double rDenom = fastRcp(alpha * alpha + beta * beta);
    const double X = dSx * ilm + ax + cPrime * (uO * a.cosA - vO * a.sinA); // dte (gradX + tauX)
    const double Y = dSy * ilm + ay + cPrime * (vO * a.cosA + uO * a.sinA); // dte (gradY + tauY)
    unew = (alpha * un + beta * vn + (alpha * X + beta * Y)) * rDenom;
    vnew = (alpha * vn - beta * un + (alpha * Y + beta * X)) * rDenom; // quirk Q3: "+ beta X" as in the reference
    uAvg = unew * a.invNSteps; // taken before applyBoundaries, as in the reference
    vAvg = vnew * a.invNSteps;
    if (dirichlet) {
        unew = 0.0;
        vnew = 0.0;
    }
}

#ifndef NSDG_UBBM_SEP
#define NSDG_UBBM_SEP 4 //!< Gauss-point evaluation and projection as two 1-d passes (evalGaussSep / projectSep) instead of tables
#endif
#ifndef NSDG_BBM_SQRT
#define NSDG_BBM_SQRT 7 //!< bit 0: branch-free root in the BBM node update, bit 1: in the law (tau), bit 2: in the law (sqrt E); bit 3 / 4: the two reciprocals of the law unconditional
#endif
#ifndef NSDG_UBBM_DIRECT_ND
#define NSDG_UBBM_DIRECT_ND 1 //!< node constants and means loaded where they are used (see NSDG_UMEVP_DIRECT_ND): 1.16 -> 1.06 ms
#endif
struct UbbmStage {
    double G[27][32]; //!< h, expC, Pmax in the 9 Gauss points
    double S[24][32];
    double D[6][32];
#if !NSDG_UBBM_DIRECT_ND
    double2 ND[2][kNodeConsts][32];
    double2 AVG[2][2][32]; //!< avgU, avgV of the row's two node lines (read-modify-written by the node update)
#endif
    double2 UV[2][2][32];
    MaskStage M;
    double UVr[2][2];
    uint64_t bar[2]; //!< mbarriers of the S (+ damage) and G groups (TMA staging)
    double pad[6]; //!< sizeof % 128 == 0: TMA destinations are 128-byte aligned in every warp's stage
};
#ifndef NSDG_UBBM_BULK
#define NSDG_UBBM_BULK 1 //!< plane rows by TMA (one tensor copy per field and row, lane 0) instead of one cp.async per plane and lane
#endif
constexpr bool kUbbmBulk = NSDG_UBBM_BULK != 0;
static_assert(sizeof(UbbmStage) % 128 == 0 && offsetof(UbbmStage, G) % 128 == 0 && offsetof(UbbmStage, S) % 128 == 0 && offsetof(UbbmStage, D) % 128 == 0
        && offsetof(UbbmStage, bar) % 8 == 0,
    "TMA destinations need 128-byte alignment");
#ifndef NSDG_UBBM_WARPS
#define NSDG_UBBM_WARPS 1 // one-warp blocks: the staging buffer sits at a compile-time shared-memory address (240 instructions
#define NSDG_UBBM_MINB 8 //  fewer per element row than with 4 x 2; 0.887 against 0.902 ms).  8 warps per SM either way: 255 registers
#endif
constexpr int kUbbmWarps = NSDG_UBBM_WARPS;
#ifndef NSDG_COOP_UBBM
#define NSDG_COOP_UBBM 0
#endif
constexpr bool kCoopUbbm = NSDG_COOP_UBBM != 0; //!< plane rows staged cooperatively (cp.async.cg) or per lane
#ifndef NSDG_UBBM_ND_HOIST
#define NSDG_UBBM_ND_HOIST 1 //!< (0: 1.075, 1: 1.018, 2: 1.010 with a spill, 3: 1.031 ms) how many stress projections ahead of the node update the direct node loads are issued (0..3)
#endif
#ifndef NSDG_COOP_UBBM_G
#define NSDG_COOP_UBBM_G 0
#endif
constexpr bool kCoopUbbmG = kCoopUbbm || NSDG_COOP_UBBM_G != 0; //!< the read-only Gauss-point planes alone
constexpr size_t kUbbmSmemBytes = sizeof(UbbmStage) * kUbbmWarps;

//! one deferred node of the BBM paths (uniform and parametric), see lineNodeMEVP
__device__ __forceinline__ void lineNodeBBM(const UniformBBMArgs& a, bool horizontal, int L, int cr)
{
    const GridDims& g = a.g;
    int r, c;
    if (horizontal) {
        c = cr;
        r = min(2 * a.R * L, 2 * g.ny);
    } else {
        r = cr;
        c = min(64 * L, 2 * g.nx);
    }
    const size_t n = size_t(r) * g.cgs + c;
    double k[kNodeConsts];
    bool d;
    const double* const src[kNodeConsts] = { a.cA, a.ax, a.ay, a.uO, a.vO, a.ilm };
    lineNodeConsts(a, src, horizontal, L, r, n, k, d);
    // the running means of a vertical line's nodes live in a compact per-line array while the subcycles run: a node column
    // of the row-major CG arrays costs a 32-byte sector per 8-byte value, read and written, every subcycle
    double* const mU = horizontal ? a.avgU + n : a.vavg + size_t(L - 1) * g.cgny + r;
    double* const mV = horizontal ? a.avgV + n : mU + size_t(a.nsx) * g.cgny;
    const double uOld = a.u[n], vOld = a.v[n], aU = *mU, aV = *mV;
    double sumX, sumY;
    lineNodeSum(a, horizontal, L, cr, r, c, sumX, sumY);
    double un, vn, ua, va;
    momentumNodeUniformBBM(a, k[0], k[1], k[2], k[3], k[4], k[5], d, uOld, vOld, d ? 0.0 : -sumX, d ? 0.0 : -sumY, un, vn, ua, va);
    a.u[n] = un;
    a.v[n] = vn;
    *mU = aU + ua;
    *mV = aV + va;
}

//! mean velocities of the vertical deferred lines: CG arrays -> compact per-line arrays before the subcycle loop (GATHER) and
//! back after it.  One thread per (line, node row); the nodes that belong to a horizontal line stay in the CG arrays.
template <bool GATHER>
__global__ void vavg_kernel(GridDims g, int nsx, int R, double* __restrict__ avgU, double* __restrict__ avgV, double* __restrict__ vavg)
{
    const int r = int(blockIdx.x * blockDim.x + threadIdx.x), L = int(blockIdx.y) + 1;
    if (r >= g.cgny || (r > 0 && (r % (2 * R) == 0 || r == 2 * g.ny)))
        return;
    const size_t n = size_t(r) * g.cgs + min(64 * L, 2 * g.nx), m = size_t(L - 1) * g.cgny + r, pitch = size_t(nsx) * g.cgny;
    if constexpr (GATHER) {
        vavg[m] = avgU[n];
        vavg[pitch + m] = avgV[n];
    } else {
        avgU[n] = vavg[m];
        avgV[n] = vavg[pitch + m];
    }
}

//! deferred-line nodes for the uniform BBM path as a kernel of their own
template <int DUMMY = 0>
__global__ void __launch_bounds__(128) subcycle_lines_ubbm(const __grid_constant__ UniformBBMArgs a)
{
    bool horizontal;
    int L, cr;
    if (lineNodeOfThread(a, horizontal, L, cr))
        lineNodeBBM(a, horizontal, L, cr);
}

template <int DUMMY = 0>
__global__ void __launch_bounds__(32 * kUbbmWarps, NSDG_UBBM_MINB) subcycle_strip_ubbm(const __grid_constant__ UniformBBMArgs a)
{
    constexpr int CG = 2, NR = 3, DGs = 8, DGA = 6;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
    UbbmStage& st = reinterpret_cast<UbbmStage*>(smemRaw)[threadIdx.x >> 5];
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
    [[maybe_unused]] const double dx3 = a.dx * (1.0 / 3.0), dy3 = a.dy * (1.0 / 3.0);

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
    unsigned phaseS = 0, phaseG = 0;
    if constexpr (kUbbmBulk) {
        if (lane == 0) {
            mbarInit(&st.bar[0], 1);
            mbarInit(&st.bar[1], 1);
        }
        mbarInitFence();
        __syncwarp();
    }
    auto issueG = [&](int row) {
        if constexpr (kUbbmBulk) {
            if (row < ey1) {
                __syncwarp(); // every lane has consumed the region that is refilled
                if (lane == 0) {
                    const int x = row * g.nxs + 32 * sx;
                    mbarExpectTx(&st.bar[1], 27 * 256);
                    tmaLoadTile(&st.G[0][0], &a.tm[4], x, &st.bar[1]);
                    tmaLoadTile(&st.G[9][0], &a.tm[5], x, &st.bar[1]);
                    tmaLoadTile(&st.G[18][0], &a.tm[6], x, &st.bar[1]);
                }
            }
            return;
        }
        stageBarrier<kCoopUbbmG>();
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<9, kCoopUbbmG>(st.G, a.gH, Npad, first, lane);
            stagePlanes<9, kCoopUbbmG>(st.G + 9, a.gE, Npad, first, lane);
            stagePlanes<9, kCoopUbbmG>(st.G + 18, a.gP, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueS = [&](int row) {
        if constexpr (kUbbmBulk) {
            if (row < ey1) {
                __syncwarp();
                if (lane == 0) {
                    const int x = row * g.nxs + 32 * sx;
                    mbarExpectTx(&st.bar[0], 30 * 256);
                    tmaLoadTile(&st.S[0][0], &a.tm[0], x, &st.bar[0]);
                    tmaLoadTile(&st.S[8][0], &a.tm[1], x, &st.bar[0]);
                    tmaLoadTile(&st.S[16][0], &a.tm[2], x, &st.bar[0]);
                    tmaLoadTile(&st.D[0][0], &a.tm[3], x, &st.bar[0]);
                }
            }
            return;
        }
        stageBarrier<kCoopUbbm>();
        if (row < ey1) {
            const size_t first = size_t(row) * g.nxs + 32 * sx;
            stagePlanes<8, kCoopUbbm>(st.S, a.s11, Npad, first, lane);
            stagePlanes<8, kCoopUbbm>(st.S + 8, a.s12, Npad, first, lane);
            stagePlanes<8, kCoopUbbm>(st.S + 16, a.s22, Npad, first, lane);
            stagePlanes<DGA, kCoopUbbm>(st.D, a.damage, Npad, first, lane);
        }
        cpAsyncCommit();
    };
    auto issueND = [&](int row) {
        if (row < ey1) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const size_t n = size_t(CG * row + k) * g.cgs + col0;
#if NSDG_UBBM_DIRECT_ND
                prefetchL2(a.cA + n);
                prefetchL2(a.ax + n);
                prefetchL2(a.ay + n);
                prefetchL2(a.uO + n);
                prefetchL2(a.vO + n);
                prefetchL2(a.ilm + n);
                prefetchL2(a.avgU + n);
                prefetchL2(a.avgV + n);
                continue;
#else
                cpAsync16cg(&st.ND[k][0][lane], a.cA + n);
                cpAsync16cg(&st.ND[k][1][lane], a.ax + n);
                cpAsync16cg(&st.ND[k][2][lane], a.ay + n);
                cpAsync16cg(&st.ND[k][3][lane], a.uO + n);
                cpAsync16cg(&st.ND[k][4][lane], a.vO + n);
                cpAsync16cg(&st.ND[k][5][lane], a.ilm + n);
                cpAsync16cg(&st.AVG[k][0][lane], a.avgU + n); // each node's mean is touched by exactly one lane and
                cpAsync16cg(&st.AVG[k][1][lane], a.avgV + n); // iteration of this launch, so staging a row ahead is safe
#endif
            }
        }
        cpAsyncCommit();
    };

    double carryX[2] = { 0.0, 0.0 }, carryY[2] = { 0.0, 0.0 };
    double ul[9], vl[9];
    issueUV(ey0);
    issueS(ey0); // order of first use in the row: UV, S (+damage), G, ND
    issueG(ey0);
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
        cpAsyncWait<kUbbmBulk ? 1 : 3>(); // u, v of this row (bulk staging: only the UV and ND groups are cp.async groups)
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

        // ---- velocity gradient in the 9 Gauss points (as in the mEVP kernel) ----
        double e11[9], e12[9], e22[9];
#if NSDG_UBBM_SEP >= 3
        {
            double Au[3][3], Adu[3][3], Av[3][3], Adv[3][3]; // [jy][qx]: value / x derivative along the node row jy
#pragma unroll
            for (int jy = 0; jy < 3; ++jy) {
                q2Values(ul[3 * jy], ul[3 * jy + 1], ul[3 * jy + 2], Au[jy][0], Au[jy][1], Au[jy][2]);
                q2Derivs(ul[3 * jy], ul[3 * jy + 1], ul[3 * jy + 2], Adu[jy][0], Adu[jy][1], Adu[jy][2]);
                q2Values(vl[3 * jy], vl[3 * jy + 1], vl[3 * jy + 2], Av[jy][0], Av[jy][1], Av[jy][2]);
                q2Derivs(vl[3 * jy], vl[3 * jy + 1], vl[3 * jy + 2], Adv[jy][0], Adv[jy][1], Adv[jy][2]);
            }
            const double ix = ice ? idx : 0.0, iy = ice ? idy : 0.0, hx = 0.5 * ix, hy = 0.5 * iy; // quirk Q8: no strain on land
#pragma unroll
            for (int qx = 0; qx < 3; ++qx) {
                double ux[3], uy[3], vx[3], vy[3]; // [qy]
                q2Values(Adu[0][qx], Adu[1][qx], Adu[2][qx], ux[0], ux[1], ux[2]);
                q2Derivs(Au[0][qx], Au[1][qx], Au[2][qx], uy[0], uy[1], uy[2]);
                q2Values(Adv[0][qx], Adv[1][qx], Adv[2][qx], vx[0], vx[1], vx[2]);
                q2Derivs(Av[0][qx], Av[1][qx], Av[2][qx], vy[0], vy[1], vy[2]);
#pragma unroll
                for (int qy = 0; qy < 3; ++qy) {
                    e11[3 * qy + qx] = ux[qy] * ix;
                    e22[3 * qy + qx] = vy[qy] * iy;
                    e12[3 * qy + qx] = fma(uy[qy], hy, vx[qy] * hx);
                }
            }
        }
#else
        {
            double Au[3][3], Adu[3][3], Av[3][3], Adv[3][3];
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
                    e11[q] = ice ? ux * idx : 0.0;
                    e22[q] = ice ? vy * idy : 0.0;
                    e12[q] = ice ? 0.5 * (uy * idy + vx * idx) : 0.0;
                });
            });
        }
#endif

        // ---- stress and damage coefficients of the row, then the BBM law point by point:
        //      e** become the updated Gauss-point stresses, dG the updated damage ----
        if constexpr (kUbbmBulk) {
            mbarWait(&st.bar[0], phaseS);
            phaseS ^= 1u;
        } else {
            cpAsyncWait<3>();
            stageBarrier<kCoopUbbm>(); // S and D were staged cooperatively
        }
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
#if NSDG_UBBM_SEP
            double t11q[9], t12q[9], t22q[9], dq[9];
            evalGaussSep<DGs>(s11c, t11q);
            evalGaussSep<DGs>(s12c, t12q);
            evalGaussSep<DGs>(s22c, t22q);
            evalGaussSep<DGA>(dc, dq);
#endif
            if constexpr (kUbbmBulk) {
                mbarWait(&st.bar[1], phaseG);
                phaseG ^= 1u;
            } else {
                cpAsyncWait<2>(); // Gauss constants (issued right after the S group one row ago)
                stageBarrier<kCoopUbbmG>();
            }
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
#if NSDG_UBBM_SEP
                double t11 = t11q[q], t12 = t12q[q], t22 = t22q[q], d = dq[q];
#else
                double t11 = evalGauss<DGs, 3, q>(s11c), t12 = evalGauss<DGs, 3, q>(s12c), t22 = evalGauss<DGs, 3, q>(s22c);
                double d = evalGauss<DGA, 3, q>(dc);
#endif
                const double h = st.G[q][lane], expC = st.G[9 + q][lane], Pmax = st.G[18 + q][lane];
                const double g11 = e11[q], g12 = e12[q], g22 = e22[q];
#if NSDG_UBBM_SEP >= 4
                d = d < 1e-12 ? 1e-12 : d; // compare + select: fmin / fmax cost five instructions each for their NaN rules
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
#if NSDG_BBM_SQRT & 8
                const double mrcp = mnum * fastRcpPinned(mden);
                const double mult = full ? 1.0 : mrcp;
#else
                const double mult = full ? 1.0 : mnum * fastRcp(mden);
#endif
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
                const double cohesion = a.cohScale * h, compr = a.comprScale * h;
                const double mc = tau + a.tan_phi * sigma_n;
                // one division: the compressive cap, when active, replaces the Mohr-Coulomb value (same operands, same result)
                const bool capped = sigma_n < -compr;
#if NSDG_BBM_SQRT & 16
                const double drcp = (capped ? -compr : cohesion) * fastRcpPinned(capped ? sigma_n : mc);
                double dcrit = (capped || mc > 0.0) ? drcp : 1.0;
#else
                double dcrit = (capped || mc > 0.0) ? (capped ? -compr : cohesion) * fastRcp(capped ? sigma_n : mc) : 1.0;
#endif
#if NSDG_UBBM_SEP >= 4
                dcrit = dcrit > 1.0 ? 1.0 : dcrit;
#else
                dcrit = fmin(dcrit, 1.0);
#endif
#if NSDG_BBM_SQRT & 4
                const double sqrtE = sqrtBranchFree(elasticity);
#else
                const double sqrtE = fastSqrt(elasticity);
#endif
                const double relax = (1.0 - dcrit) * a.deltaT * (sqrtE * a.invTdK);
                dG[q] = d - d * relax;
                e11[q] = t11 - t11 * relax;
                e12[q] = t12 - t12 * relax;
                e22[q] = t22 - t22 * relax;
            });
        }

        // ---- damage projection (iMJwPSI_dam = psi_j w / m_j on a rectangle, j < 6) ----
#if NSDG_UBBM_SEP
        {
            double dnew[DGA];
            projectSep<DGA>(dG, dnew);
#pragma unroll
            for (int j = 0; j < DGA; ++j)
                if (active)
                    a.damage[size_t(j) * Npad + e] = dnew[j];
        }
#else
        static_for<DGA>([&](auto J) {
            constexpr int j = decltype(J)::value;
            double acc = 0.0;
            static_for<9>([&](auto QQ) {
                constexpr int q = decltype(QQ)::value;
                constexpr double b = kUnitOps.B[j][q];
                if constexpr (b != 0.0)
                    acc = fma(b, dG[q], acc);
            });
            if (active)
                a.damage[size_t(j) * Npad + e] = acc;
        });
#endif

        // ---- per stress component: project, store, accumulate the divergence contributions ----
        double Tx[9], Ty[9];
#pragma unroll
        for (int k = 0; k < 9; ++k)
            Tx[k] = Ty[k] = 0.0;
        auto component = [&](double* plane, const double (&r)[9], auto COMP) {
            constexpr int comp = decltype(COMP)::value;
            double s[DGs];
#if NSDG_UBBM_SEP
            projectSep<DGs>(r, s);
#pragma unroll
            for (int j = 0; j < DGs; ++j)
                if (active)
                    plane[size_t(j) * Npad + e] = s[j];
#else
            static_for<DGs>([&](auto J) {
                constexpr int j = decltype(J)::value;
                double acc = 0.0;
                static_for<9>([&](auto QQ) {
                    constexpr int q = decltype(QQ)::value;
                    constexpr double b = kUnitOps.B[j][q];
                    if constexpr (b != 0.0)
                        acc = fma(b, r[q], acc);
                });
                s[j] = acc;
                if (active)
                    plane[size_t(j) * Npad + e] = acc;
            });
#endif
#if NSDG_UBBM_SEP >= 2
            if (ice) { // comp 0 comes first: Tx is assigned; s12 assigns Ty; the others accumulate
                if constexpr (comp == 0)
                    divergenceSep<0, false>(s, a.dy, dy3, Tx);
                if constexpr (comp == 1) {
                    divergenceSep<1, true>(s, a.dx, dx3, Tx);
                    divergenceSep<0, false>(s, a.dy, dy3, Ty);
                }
                if constexpr (comp == 2)
                    divergenceSep<1, true>(s, a.dx, dx3, Ty);
            }
#else
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
#endif
        };
#if NSDG_UBBM_DIRECT_ND
        // node constants and means of the row's two node lines: the loads are issued NSDG_UBBM_ND_HOIST projections
        // ahead of the node update (they cannot be moved across the cp.async statements by the compiler)
        double2 ndc[2][kNodeConsts], avg0[2][2];
        auto loadND = [&]() {
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
        };
        if constexpr (NSDG_UBBM_ND_HOIST == 3)
            loadND();
#endif
        component(a.s11, e11, std::integral_constant<int, 0> {});
#if NSDG_UBBM_DIRECT_ND
        if constexpr (NSDG_UBBM_ND_HOIST == 2)
            loadND();
#endif
        component(a.s12, e12, std::integral_constant<int, 1> {});
#if NSDG_UBBM_DIRECT_ND
        if constexpr (NSDG_UBBM_ND_HOIST == 1)
            loadND();
#endif
        component(a.s22, e22, std::integral_constant<int, 2> {});
#if NSDG_UBBM_DIRECT_ND
        if constexpr (NSDG_UBBM_ND_HOIST == 0)
            loadND();
#endif
        issueS(ey + 1);
        issueG(ey + 1);

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
#pragma unroll
        for (int jy = 0; jy < NR; ++jy) {
            const double lx = __shfl_up_sync(FULL, Tx[jy * NR + CG], 1);
            const double ly = __shfl_up_sync(FULL, Ty[jy * NR + CG], 1);
            if (lane > 0) {
                Tx[jy * NR] = lx + Tx[jy * NR];
                Ty[jy * NR] = ly + Ty[jy * NR];
            }
        }
        // ---- momentum update of the completed nodes ----
        cpAsyncWait<kUbbmBulk ? 1 : 3>();
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
#if NSDG_UBBM_DIRECT_ND
            const double2 cA = ndc[jy][0], ax = ndc[jy][1], ay = ndc[jy][2], uO = ndc[jy][3], vO = ndc[jy][4], ilm = ndc[jy][5];
            const double2 avgU0 = avg0[jy][0], avgV0 = avg0[jy][1];
#else
            const double2 cA = st.ND[jy][0][lane], ax = st.ND[jy][1][lane], ay = st.ND[jy][2][lane];
            const double2 uO = st.ND[jy][3][lane], vO = st.ND[jy][4][lane], ilm = st.ND[jy][5][lane];
            const double2 avgU0 = st.AVG[jy][0][lane], avgV0 = st.AVG[jy][1][lane];
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
                    double2 au = avgU0, av = avgV0;
                    au.x += ua.x;
                    au.y += ua.y;
                    av.x += va.x;
                    av.y += va.y;
                    *reinterpret_cast<double2*>(a.avgU + n0) = au;
                    *reinterpret_cast<double2*>(a.avgV + n0) = av;
                } else {
                    a.u[n0 + 1] = un.y;
                    a.v[n0 + 1] = vn.y;
                    a.avgU[n0 + 1] = avgU0.y + ua.y;
                    a.avgV[n0 + 1] = avgV0.y + va.y;
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
