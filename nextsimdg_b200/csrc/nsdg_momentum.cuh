/*
 * nsdg_momentum.cuh -- the subcycle hot loop as two kernels per subcycle.
 *
 * One subcycle of the reference is five whole-grid sweeps
 *   projectVelocityToStrain   dynamics/src/CGDynamicsKernel.cpp:300-337
 *   stressUpdateHighOrder     dynamics/src/include/MEVPStressUpdateStep.hpp:30-118  (mEVP)
 *                             dynamics/src/include/BBMStressUpdateStep.hpp:31-196   (BBM)
 *   stressDivergence          dynamics/src/CGDynamicsKernel.cpp:340-398
 *   updateMomentum            dynamics/src/include/VPCGDynamicsKernel.hpp:132-172    (mEVP)
 *                             dynamics/src/include/BrittleCGDynamicsKernel.hpp:206-254 (BBM)
 *   applyBoundaries           dynamics/src/CGDynamicsKernel.cpp:439-444
 * which stream strain, stress divergence and all operators through memory.  Here:
 *
 *  subcycle_strip  : one WARP owns a strip of 32 elements x R element rows and walks it bottom
 *                    to top.  Per element row each lane (= element) gets its 9 (4) CG velocities
 *                    from coalesced row loads + one shuffle, forms the strain in registers,
 *                    evaluates the rheology in the Gauss points, updates the stress in place,
 *                    and forms the element's stress-divergence contributions.  Contributions to
 *                    CG nodes shared with the left neighbour element travel by warp shuffle,
 *                    those shared with the element row above are carried in registers to the next
 *                    iteration.  A node whose contributions are complete is advanced by the
 *                    momentum equation by the lane that owns it, in the same iteration.
 *                    Strain and stress divergence never touch memory.
 *  subcycle_lines  : the nodes on strip boundaries (every 32nd element column, every R-th element
 *                    row, plus the domain's top/right edge) collect contributions from two warps;
 *                    strips write their raw per-element contributions for those lines to small
 *                    line buffers and this kernel sums them in a fixed order and does the same
 *                    momentum update.  No atomics anywhere: results are run-to-run reproducible.
 *
 * Dirichlet nodes (dirichletZero, CGDynamicsKernel.cpp:401-437) are applied through a per-node
 * byte mask built once from the sorted Dirichlet lists: the stress divergence is zeroed there
 * before the momentum update and u,v after it, as in the reference's sweep order.
 * Quirk Q8: strain and divergence skip land elements, the stress update does not.
 * The periodic averaging at the end of stressDivergence (CGAveragePeriodic, dynamics/src/include/VectorManipulations.hpp:26-65;
 * CGDynamicsKernel.cpp:395-397): nodes that appear in a periodic list carry bit 1 of the node mask.  Strips and lines do not
 * advance them; they leave the node's (Dirichlet-zeroed) stress divergence in seamX / seamY, the host averages those arrays
 * across the seam segment by segment like the reference (cg_average_periodic_kernel) and seam_update_kernel advances every
 * seam node once.  IDynamics never marks periodic edges (DynamicsKernel.hpp:54): the lists are set through
 * nsdg_set_boundaries, which also switches the handle to these generic kernels.
 */
#pragma once
#include "nsdg_state.cuh"

namespace nsdg {

//! Everything the subcycle kernels need (passed by value as a __grid_constant__)
struct SubcycleArgs {
    GridDims g;
    int R; //!< element rows per strip
    int nsx, nsy; //!< strips in x and y
    // element state (planes)
    double *s11, *s12, *s22; //!< DGs planes each
    double* damage; //!< DGadv planes (BBM)
    const double* gaussA; //!< per-step Gauss-point constants, plane q:  mEVP: P_q ; BBM: h_q
    const double* gaussB; //!< BBM: expC_q
    const double* helem; //!< BBM: element size h (smesh.h(i))
    const uint8_t* landmask;
    // operators (general path; the uniform path reads c_mops)
    const double *Gx, *Gy, *GM, *B, *Bd, *D1, *D2, *DM;
    // node state
    double *u, *v, *avgU, *avgV;
    const double *u0, *v0, *cgH, *cgA, *uAtm, *vAtm, *uOcn, *vOcn, *gradX, *gradY, *lmass;
    const uint8_t* nodemask; //!< bit0: Dirichlet node, bit1: node of a periodic seam
    double *seamX, *seamY; //!< stress divergence of the periodic-seam nodes (CG arrays; null without periodic edges)
    // deferred line buffers: raw contributions
    double* hbuf; //!< [line 1..nsy][side 0 below,1 above][ex][jx 0..CG][comp]
    double* vbuf; //!< [line 1..nsx][side 0 left,1 right][ey][jy 0..CG][comp]
    // parameters
    double deltaT; //!< mEVP: dt ; BBM: dt/nSteps
    double nSteps; //!< BBM: running-mean divisor
    PhysParams p;
};

static __constant__ MomentumOps c_mops; //!< the single operator set of a uniform rectangular mesh (one copy per translation unit; only nsdg_cuda.cu uses it)

//! node-constant inputs of the momentum equation
struct NodeIn {
    double cgH, cgA, uA, vA, uO, vO, gx, gy, lm, u0, v0;
    bool dirichlet, seam;
};

template <int RHEO> __device__ __forceinline__ NodeIn loadNode(const SubcycleArgs& a, size_t n)
{
    NodeIn in;
    in.cgH = __ldg(a.cgH + n);
    in.cgA = __ldg(a.cgA + n);
    in.uA = __ldg(a.uAtm + n);
    in.vA = __ldg(a.vAtm + n);
    in.uO = __ldg(a.uOcn + n);
    in.vO = __ldg(a.vOcn + n);
    in.gx = __ldg(a.gradX + n);
    in.gy = __ldg(a.gradY + n);
    in.lm = __ldg(a.lmass + n);
    if constexpr (RHEO == NSDG_MEVP) {
        in.u0 = __ldg(a.u0 + n);
        in.v0 = __ldg(a.v0 + n);
    } else {
        in.u0 = in.v0 = 0;
    }
    const uint8_t m = __ldg(a.nodemask + n);
    in.dirichlet = m & 1;
    in.seam = m & 2;
    return in;
}

//! momentum update of one node; VPCGDynamicsKernel.hpp:140-171 (quirks Q1, Q2 verbatim) or
//! BrittleCGDynamicsKernel.hpp:209-253 (quirk Q3 verbatim), followed by applyBoundaries.
//! (un,vn) = current velocity, (dSx,dSy) = stress divergence sums (dStressX/Y).  avgInc = u/nSteps
//! of the BBM running mean, which in the reference is taken BEFORE the Dirichlet zeroing.
//! ZERO_DS = false: (dSx, dSy) has been through dirichletZero and the periodic average already (seam_update_kernel)
template <int RHEO, bool ZERO_DS = true>
__device__ __forceinline__ void momentumNode(const SubcycleArgs& a, const NodeIn& in, double un, double vn, double dSx,
    double dSy, double& unew, double& vnew, double& avgIncU, double& avgIncV)
{
    const PhysParams& p = a.p;
    if (ZERO_DS && in.dirichlet) { // dirichletZero(dStress), CGDynamicsKernel.cpp:393-394
        dSx = 0.0;
        dSy = 0.0;
    }
    if constexpr (RHEO == NSDG_MEVP) {
        const double beta = p.beta, SC = 1.0;
        const double uOcnRel = in.uO - un;
        const double vOcnRel = vn - in.vO;
        const double absatm = sqrt(in.uA * in.uA + in.vA * in.vA);
        const double absocn = sqrt(uOcnRel * uOcnRel + vOcnRel * vOcnRel);
        const double denom
            = 1.0 / (p.rho_ice * in.cgH / a.deltaT * (1.0 + beta) + in.cgA * p.F_ocean * absocn);
        unew = denom
            * (p.rho_ice * in.cgH / a.deltaT * (beta * un + in.u0)
                + in.cgA * (p.F_atm * absatm * in.uA + p.F_ocean * absocn * SC * in.uO) - p.rho_ice * in.cgH * p.fc * un
                - p.rho_ice * in.cgH * p.gravity * in.gx + dSx / in.lm);
        vnew = denom
            * (p.rho_ice * in.cgH / a.deltaT * (beta * vn + in.v0)
                + in.cgA * (p.F_atm * absatm * in.vA + p.F_ocean * absocn * SC * in.vO) + p.rho_ice * in.cgH * p.fc * vn
                - p.rho_ice * in.cgH * p.gravity * in.gy + dSy / in.lm);
        avgIncU = avgIncV = 0.0;
    } else {
        const double dteOverMass = a.deltaT / (p.rho_ice * in.cgH);
        const double cPrime = in.cgA * p.F_ocean * hypot(in.uO - un, in.vO - vn);
        const double tauB = 0.;
        const double alpha = 1 + dteOverMass * (cPrime * p.cosOceanAngle + tauB);
        const double beta = a.deltaT * p.fc + dteOverMass * cPrime * p.sinOceanAngle;
        const double rDenom = 1 / (alpha * alpha + beta * beta);
        const double dragAtm = in.cgA * p.F_atm * hypot(in.uA, in.vA);
        const double tauX = dragAtm * in.uA + cPrime * (in.uO * p.cosOceanAngle - in.vO * p.sinOceanAngle);
        const double tauY = dragAtm * in.vA + cPrime * (in.vO * p.cosOceanAngle + in.uO * p.sinOceanAngle);
        const double gradX = dSx / in.lm - p.rho_ice * in.cgH * p.gravity * in.gx;
        const double gradY = dSy / in.lm - p.rho_ice * in.cgH * p.gravity * in.gy;
        unew = alpha * un + beta * vn + dteOverMass * (alpha * (gradX + tauX) + beta * (gradY + tauY));
        unew *= rDenom;
        vnew = alpha * vn - beta * un + dteOverMass * (alpha * (gradY + tauY) + beta * (gradX + tauX));
        vnew *= rDenom;
        avgIncU = unew / a.nSteps;
        avgIncV = vnew / a.nSteps;
    }
    if (in.dirichlet) { // applyBoundaries, CGDynamicsKernel.cpp:439-444
        unew = 0.0;
        vnew = 0.0;
    }
}

//! value of a DGs / DGadv row in Gauss point q, structural zeros of the basis skipped at compile time
template <int NC, int GS, int q> __device__ __forceinline__ double evalGauss(const double (&c)[NC])
{
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        constexpr double dummy = 0.0;
        (void)dummy;
        const double w = PSI(GS, j, q);
        if (w != 0.0)
            s = (j == 0) ? c[0] * w : fma(c[j], w, s);
    }
    return s;
}

/*
 * The per-Gauss-point constitutive update.
 *  mEVP (MEVPStressUpdateStep.hpp:62-117): returns the stress increment integrand r = 1/alpha (...)
 *  BBM  (BBMStressUpdateStep.hpp:66-160): advances stress and damage in the point.
 */
template <int RHEO> struct GaussLaw;

template <> struct GaussLaw<NSDG_MEVP> {
    //! P = Pstar * h * exp(-20 (1-a)) precomputed once per step
    __device__ __forceinline__ static void apply(const PhysParams& p, double P, double g11, double g12, double g22,
        double& r11, double& r12, double& r22)
    {
        const double DELTA = sqrt(p.DeltaMin * p.DeltaMin + 1.25 * (g11 * g11 + g22 * g22) + 1.50 * g11 * g22 + g12 * g12);
        r11 = 1.0 / p.alpha * (P / 8.0 / DELTA * (5.0 * g11 + 3.0 * g22) - 0.5 * P);
        r12 = 1.0 / p.alpha * (P / 4.0 / DELTA * g12);
        r22 = 1.0 / p.alpha * (P / 8.0 / DELTA * (5.0 * g22 + 3.0 * g11) - 0.5 * P);
    }
};

template <> struct GaussLaw<NSDG_BBM> {
    //! h (clamped >= 0) and expC = exp(C (1-a)) precomputed once per step; s**, d are updated in place
    __device__ __forceinline__ static void apply(const PhysParams& p, double dt, double hel, double scale_coef, double h,
        double expC, double e11, double e12, double e22, double& s11, double& s12, double& s22, double& d)
    {
        d = fmin(fmax(d, 1e-12), 1.0);
        double sigma_n = 0.5 * (s11 + s22);
        // (d expC)^(n-1) and h^(exponent+1): with the reference's constants n = 5 and exponent = 1.5 these are
        // x^4 and h^2.5; two multiplications / one sqrt instead of the generic pow (identical to ~1 ulp)
        const double de = d * expC;
        const double powalphaexpC = (p.exponent_relaxation_sigma == 5) ? (de * de) * (de * de)
                                                                       : pow(de, double(p.exponent_relaxation_sigma - 1));
        const double time_viscous = p.undamaged_time_relaxation_sigma * powalphaexpC;
        const double hpow = (p.exponent_compression_factor == 1.5) ? h * h * sqrt(h) : pow(h, p.exponent_compression_factor + 1.);
        const double Pmax = p.P0 * hpow * expC;
        const double tildeP = (sigma_n < 0.0) ? fmin(-Pmax / sigma_n, 1.0) : 0.;
        const double multiplicator = time_viscous / (time_viscous + (1. - tildeP) * dt);
        const double elasticity = h * p.young * d * expC;
        const double Dunit = dt * elasticity / (1. - (p.nu0 * p.nu0));
        s11 += Dunit * (e11 + p.nu0 * e22);
        s22 += Dunit * (p.nu0 * e11 + e22);
        s12 += Dunit * e12 * (1. - p.nu0);
        s11 *= multiplicator;
        s22 *= multiplicator;
        s12 *= multiplicator;
        sigma_n = 0.5 * (s11 + s22);
        const double tau = sqrt(0.25 * (s11 - s22) * (s11 - s22) + s12 * s12);
        const double cohesion = p.C_lab * scale_coef * h;
        const double compr = p.compr_strength * scale_coef * h;
        double dcrit = (tau + p.tan_phi * sigma_n > 0.) ? cohesion / (tau + p.tan_phi * sigma_n) : 1.;
        if (sigma_n < -compr)
            dcrit = -compr / sigma_n;
        dcrit = fmin(dcrit, 1.0);
        const double td = hel * sqrt(2. * (1. + p.nu0) * p.rho_ice) / sqrt(elasticity);
        const double relax = (1. - dcrit) * dt / td; // the reference evaluates x * (1-dcrit) * dt / td four times
        d -= d * relax;
        s11 -= s11 * relax;
        s12 -= s12 * relax;
        s22 -= s22 * relax;
    }
};

//! operator entry k of matrix M for element e: constant memory (uniform) or plane (general)
#define NSDG_OP(M, k) (UNIFORM ? c_mops.M[(k)] : __ldg(a.M + size_t(k) * Npad + e))

template <int CG, int DGA, int RHEO, bool UNIFORM, bool SPH>
__global__ void __launch_bounds__(128) subcycle_strip(const __grid_constant__ SubcycleArgs a)
{
    constexpr int DGs = cg2dgstress(CG), GS = gp1d(DGs), Q = GS * GS, NR = CG + 1, ND = NR * NR;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.nsx * a.nsy)
        return;
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

    auto loadRow = [&](const double* f, int r, double* out) {
        const double* ptr = f + size_t(r) * g.cgs + col0;
        if constexpr (CG == 2) {
            const double2 t = *reinterpret_cast<const double2*>(ptr);
            out[0] = t.x;
            out[1] = t.y;
        } else {
            out[0] = ptr[0];
        }
        double right = __shfl_down_sync(FULL, out[0], 1);
        if (loadsRight)
            right = ptr[CG];
        out[CG] = right;
    };

    double carryX[CG], carryY[CG];
#pragma unroll
    for (int i = 0; i < CG; ++i)
        carryX[i] = carryY[i] = 0.0;

    double ul[ND], vl[ND]; // local CG velocities, index jy*NR + jx
    loadRow(a.u, CG * ey0, ul);
    loadRow(a.v, CG * ey0, vl);

    for (int ey = ey0; ey < ey1; ++ey) {
        const size_t e = size_t(ey) * g.nxs + ex;
#pragma unroll
        for (int jy = 1; jy <= CG; ++jy) {
            loadRow(a.u, CG * ey + jy, ul + jy * NR);
            loadRow(a.v, CG * ey + jy, vl + jy * NR);
        }
        const bool ice = active && (__ldg(a.landmask + e) != 0);

        // ---- stress state (all elements, quirk Q8) ----
        double s11[DGs], s12[DGs], s22[DGs];
#pragma unroll
        for (int j = 0; j < DGs; ++j) {
            s11[j] = a.s11[size_t(j) * Npad + e];
            s12[j] = a.s12[size_t(j) * Npad + e];
            s22[j] = a.s22[size_t(j) * Npad + e];
        }

        // ---- projectVelocityToStrain (ice elements only; land strain stays zero) ----
        double e11[DGs], e12[DGs], e22[DGs];
        if (ice) {
#pragma unroll
            for (int i = 0; i < DGs; ++i) {
                double xu = 0, yv = 0, xv = 0, yu = 0;
#pragma unroll
                for (int k = 0; k < ND; ++k) {
                    const double gx = NSDG_OP(Gx, i * ND + k), gy = NSDG_OP(Gy, i * ND + k);
                    xu = fma(gx, ul[k], xu);
                    yv = fma(gy, vl[k], yv);
                    xv = fma(gx, vl[k], xv);
                    yu = fma(gy, ul[k], yu);
                }
                e11[i] = xu;
                e22[i] = yv;
                e12[i] = 0.5 * (xv + yu);
                if constexpr (SPH) {
                    double mv = 0, mu = 0;
#pragma unroll
                    for (int k = 0; k < ND; ++k) {
                        const double gm = NSDG_OP(GM, i * ND + k);
                        mv = fma(gm, vl[k], mv);
                        mu = fma(gm, ul[k], mu);
                    }
                    e11[i] -= mv;
                    e12[i] += 0.5 * mu;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < DGs; ++i)
                e11[i] = e12[i] = e22[i] = 0.0;
        }

        // ---- stress update in the Gauss points ----
        if constexpr (RHEO == NSDG_MEVP) {
            double a11[DGs], a12[DGs], a22[DGs];
#pragma unroll
            for (int j = 0; j < DGs; ++j)
                a11[j] = a12[j] = a22[j] = 0.0;
            auto gaussPoint = [&](auto qc) {
                constexpr int q = decltype(qc)::value;
                const double g11 = evalGauss<DGs, GS, q>(e11), g12 = evalGauss<DGs, GS, q>(e12),
                             g22 = evalGauss<DGs, GS, q>(e22);
                const double P = __ldg(a.gaussA + size_t(q) * Npad + e);
                double r11, r12, r22;
                GaussLaw<NSDG_MEVP>::apply(a.p, P, g11, g12, g22, r11, r12, r22);
#pragma unroll
                for (int j = 0; j < DGs; ++j) {
                    const double b = NSDG_OP(B, j * Q + q);
                    a11[j] = fma(b, r11, a11[j]);
                    a12[j] = fma(b, r12, a12[j]);
                    a22[j] = fma(b, r22, a22[j]);
                }
            };
            [&]<int... qs>(std::integer_sequence<int, qs...>) { (gaussPoint(std::integral_constant<int, qs> {}), ...); }(
                std::make_integer_sequence<int, Q> {});
            const double keep = 1.0 - 1.0 / a.p.alpha;
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                s11[j] = s11[j] * keep + a11[j];
                s12[j] = s12[j] * keep + a12[j];
                s22[j] = s22[j] * keep + a22[j];
            }
        } else {
            double dam[DGA];
#pragma unroll
            for (int j = 0; j < DGA; ++j)
                dam[j] = a.damage[size_t(j) * Npad + e];
            double a11[DGs], a12[DGs], a22[DGs], ad[DGA];
#pragma unroll
            for (int j = 0; j < DGs; ++j)
                a11[j] = a12[j] = a22[j] = 0.0;
#pragma unroll
            for (int j = 0; j < DGA; ++j)
                ad[j] = 0.0;
            const double hel = __ldg(a.helem + e);
            const double scale_coef = sqrt(0.1 / hel);
            auto gaussPoint = [&](auto qc) {
                constexpr int q = decltype(qc)::value;
                const double g11 = evalGauss<DGs, GS, q>(e11), g12 = evalGauss<DGs, GS, q>(e12),
                             g22 = evalGauss<DGs, GS, q>(e22);
                double t11 = evalGauss<DGs, GS, q>(s11), t12 = evalGauss<DGs, GS, q>(s12), t22 = evalGauss<DGs, GS, q>(s22);
                double d = evalGauss<DGA, GS, q>(dam);
                const double h = __ldg(a.gaussA + size_t(q) * Npad + e);
                const double expC = __ldg(a.gaussB + size_t(q) * Npad + e);
                GaussLaw<NSDG_BBM>::apply(a.p, a.deltaT, hel, scale_coef, h, expC, g11, g12, g22, t11, t12, t22, d);
#pragma unroll
                for (int j = 0; j < DGs; ++j) {
                    const double b = NSDG_OP(B, j * Q + q);
                    a11[j] = fma(b, t11, a11[j]);
                    a12[j] = fma(b, t12, a12[j]);
                    a22[j] = fma(b, t22, a22[j]);
                }
#pragma unroll
                for (int j = 0; j < DGA; ++j)
                    ad[j] = fma(NSDG_OP(Bd, j * Q + q), d, ad[j]);
            };
            [&]<int... qs>(std::integer_sequence<int, qs...>) { (gaussPoint(std::integral_constant<int, qs> {}), ...); }(
                std::make_integer_sequence<int, Q> {});
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                s11[j] = a11[j];
                s12[j] = a12[j];
                s22[j] = a22[j];
            }
            if (active) {
#pragma unroll
                for (int j = 0; j < DGA; ++j)
                    a.damage[size_t(j) * Npad + e] = ad[j];
            }
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < DGs; ++j) {
                a.s11[size_t(j) * Npad + e] = s11[j];
                a.s12[size_t(j) * Npad + e] = s12[j];
                a.s22[size_t(j) * Npad + e] = s22[j];
            }
        }

        // ---- stressDivergence: raw contributions of this element to its ND nodes ----
        double Tx[ND], Ty[ND];
        if (ice) {
#pragma unroll
            for (int k = 0; k < ND; ++k) {
                double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
#pragma unroll
                for (int j = 0; j < DGs; ++j) {
                    const double d1 = NSDG_OP(D1, k * DGs + j), d2 = NSDG_OP(D2, k * DGs + j);
                    a1 = fma(d1, s11[j], a1);
                    a2 = fma(d2, s12[j], a2);
                    b1 = fma(d1, s12[j], b1);
                    b2 = fma(d2, s22[j], b2);
                }
                Tx[k] = a1 + a2;
                Ty[k] = b1 + b2;
                if constexpr (SPH) {
                    double m12 = 0, m11 = 0;
#pragma unroll
                    for (int j = 0; j < DGs; ++j) {
                        const double dm = NSDG_OP(DM, k * DGs + j);
                        m12 = fma(dm, s12[j], m12);
                        m11 = fma(dm, s11[j], m11);
                    }
                    Tx[k] += m12;
                    Ty[k] -= m11;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < ND; ++k)
                Tx[k] = Ty[k] = 0.0;
        }

        // ---- raw contributions to the deferred vertical lines ----
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
        // ---- raw contributions to the deferred horizontal lines ----
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

        // ---- combine: left neighbour's right column by shuffle, row below by register carry ----
#pragma unroll
        for (int jy = 0; jy < NR; ++jy) {
            const double lx = __shfl_up_sync(FULL, Tx[jy * NR + CG], 1);
            const double ly = __shfl_up_sync(FULL, Ty[jy * NR + CG], 1);
            if (lane > 0) {
                Tx[jy * NR] = lx + Tx[jy * NR];
                Ty[jy * NR] = ly + Ty[jy * NR];
            }
        }
        // ---- momentum update of the completed nodes: rows jy < CG, columns jx < CG ----
#pragma unroll
        for (int jy = 0; jy < CG; ++jy) {
            const size_t n0 = size_t(CG * ey + jy) * g.cgs + col0;
            double un[CG], vn[CG], aiu[CG], aiv[CG];
            bool skip[CG];
#pragma unroll
            for (int jx = 0; jx < CG; ++jx) {
                skip[jx] = !active || (jx == 0 && lane == 0 && sx > 0) || (jy == 0 && bottomDeferred);
                double sumX = Tx[jy * NR + jx], sumY = Ty[jy * NR + jx];
                if (jy == 0) {
                    sumX = carryX[jx] + sumX;
                    sumY = carryY[jx] + sumY;
                }
                const NodeIn in = loadNode<RHEO>(a, n0 + jx);
                momentumNode<RHEO>(a, in, ul[jy * NR + jx], vl[jy * NR + jx], -sumX, -sumY, un[jx], vn[jx], aiu[jx], aiv[jx]);
                if (in.seam && !skip[jx]) { // periodic seam: leave the divergence for the average, seam_update_kernel advances the node
                    a.seamX[n0 + jx] = in.dirichlet ? 0.0 : -sumX;
                    a.seamY[n0 + jx] = in.dirichlet ? 0.0 : -sumY;
                    skip[jx] = true;
                }
            }
#pragma unroll
            for (int jx = 0; jx < CG; ++jx)
                if (!skip[jx]) {
                    a.u[n0 + jx] = un[jx];
                    a.v[n0 + jx] = vn[jx];
                    if constexpr (RHEO == NSDG_BBM) {
                        a.avgU[n0 + jx] += aiu[jx];
                        a.avgV[n0 + jx] += aiv[jx];
                    }
                }
        }
#pragma unroll
        for (int jx = 0; jx < CG; ++jx) {
            carryX[jx] = Tx[CG * NR + jx];
            carryY[jx] = Ty[CG * NR + jx];
        }
#pragma unroll
        for (int jx = 0; jx < NR; ++jx) {
            ul[jx] = ul[CG * NR + jx];
            vl[jx] = vl[CG * NR + jx];
        }
    }
}

/*
 * Nodes on the deferred lines.  One thread per node:
 *   t <  nsy*cgnx : node (c, row of horizontal line L = t / cgnx + 1)
 *   else          : node (col of vertical line L, r), skipped if r lies on a horizontal line
 */
template <int CG, int RHEO> __global__ void __launch_bounds__(128) subcycle_lines(const __grid_constant__ SubcycleArgs a)
{
    constexpr int NR = CG + 1;
    const GridDims& g = a.g;
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    const long nH = long(a.nsy) * g.cgnx;
    const long nV = long(a.nsx) * g.cgny;
    if (t >= nH + nV)
        return;
    int c, r;
    double sumX = 0.0, sumY = 0.0;
    if (t < nH) {
        const int L = int(t / g.cgnx) + 1;
        c = int(t % g.cgnx);
        r = min(CG * a.R * L, CG * g.ny);
        const int jx = c % CG, exr = c / CG;
        const bool above = r < CG * g.ny;
        auto add = [&](int side, int ex, int j) {
            const double* hb = a.hbuf + ((size_t(L - 1) * 2 + side) * g.nx + ex) * (NR * 2) + j * 2;
            sumX += hb[0];
            sumY += hb[1];
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
        const long tv = t - nH;
        const int L = int(tv / g.cgny) + 1;
        r = int(tv % g.cgny);
        c = min(CG * 32 * L, CG * g.nx);
        if (r > 0 && (r % (CG * a.R) == 0 || r == CG * g.ny))
            return; // handled as part of a horizontal line
        const int jy = r % CG, eyr = r / CG;
        const bool right = c < CG * g.nx;
        auto add = [&](int side, int ey, int j) {
            const double* vb = a.vbuf + ((size_t(L - 1) * 2 + side) * g.ny + ey) * (NR * 2) + j * 2;
            sumX += vb[0];
            sumY += vb[1];
        };
        for (int side = 0; side < 2; ++side) {
            if (side == 1 && !right)
                break;
            if (jy == 0 && eyr > 0)
                add(side, eyr - 1, CG);
            add(side, eyr, jy);
        }
    }
    const size_t n = size_t(r) * g.cgs + c;
    const NodeIn in = loadNode<RHEO>(a, n);
    if (in.seam) {
        a.seamX[n] = in.dirichlet ? 0.0 : -sumX;
        a.seamY[n] = in.dirichlet ? 0.0 : -sumY;
        return;
    }
    double un, vn, aiu, aiv;
    momentumNode<RHEO>(a, in, a.u[n], a.v[n], -sumX, -sumY, un, vn, aiu, aiv);
    a.u[n] = un;
    a.v[n] = vn;
    if constexpr (RHEO == NSDG_BBM) {
        a.avgU[n] += aiu;
        a.avgV[n] += aiv;
    }
}

//! the nodes of the periodic seams (each once), after their stress divergence has been averaged across the seam
template <int RHEO> __global__ void seam_update_kernel(const __grid_constant__ SubcycleArgs a, const long* __restrict__ nodes, long count)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= count)
        return;
    const size_t n = size_t(nodes[t]);
    const NodeIn in = loadNode<RHEO>(a, n);
    double un, vn, aiu, aiv;
    momentumNode<RHEO, false>(a, in, a.u[n], a.v[n], a.seamX[n], a.seamY[n], un, vn, aiu, aiv);
    a.u[n] = un;
    a.v[n] = vn;
    if constexpr (RHEO == NSDG_BBM) {
        a.avgU[n] += aiu;
        a.avgV[n] += aiv;
    }
}
//! bit 1 of the node mask for the listed nodes
template <int DUMMY = 0> __global__ void seam_flag_kernel(const long* __restrict__ nodes, long count, uint8_t* __restrict__ nodemask)
{
    const long t = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < count)
        nodemask[nodes[t]] |= 2; // the list holds every node once
}

#undef NSDG_OP

} // namespace nsdg
