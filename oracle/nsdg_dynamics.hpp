/*
 * nsdg_dynamics.hpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * CPU restatement of the subcycled CG momentum / rheology solve of neXtSIM_DG:
 * ParametricMomentumMap, CGDynamicsKernel, VPCGDynamicsKernel (+MEVPStressUpdateStep),
 * BrittleCGDynamicsKernel (+BBMStressUpdateStep) and the DynamicsKernel driver.
 *
 * Reference: dynamics/src/ParametricMap.cpp:94-356
 *            dynamics/src/CGDynamicsKernel.cpp:23-449
 *            dynamics/src/include/DynamicsKernel.hpp:44-172
 *            dynamics/src/include/VPCGDynamicsKernel.hpp:63-172
 *            dynamics/src/include/MEVPStressUpdateStep.hpp:30-118
 *            dynamics/src/include/BrittleCGDynamicsKernel.hpp:72-254
 *            dynamics/src/include/BBMStressUpdateStep.hpp:31-196
 *            dynamics/src/include/DGModelArray.hpp:20-48
 *            dynamics/src/include/VectorManipulations.hpp:26-65
 *            dynamics/src/include/{DynamicsParameters,VPParameters,MEBParameters}.hpp
 *
 * PARITY STATUS: the advection half is pinned by the reference's KATs (see nsdg_kat.cpp);
 * the momentum half has no golden vector in the reference's tests; it is pinned
 * against OUTPUTS OF THE REFERENCE ITSELF: oracle/_ref (the reference's kernels compiled from
 * /root/reference, `make -C oracle ref`) run on the same inputs, live in
 * tests/test_oracle_vs_reference.py and as committed fixtures tests/golden/ref_outputs.npz
 * (agreement 1e-13 after 100 subcycles), plus the analytic self-consistency tests.
 *
 * Unlike the reference (quirk Q9) all state is explicitly zero-initialised.
 */
#pragma once
#include "nsdg_transport.hpp"

#include <map>
#include <omp.h>
#include <stdexcept>

namespace nso {

enum Rheology { MEVP = 0, BBM = 1, FREEDRIFT = 2 };

//! DynamicsParameters.hpp:32-50, VPParameters.hpp:22-23, MEBParameters.hpp:40-65
struct Params {
    double rho_ice = 900.0, rho_atm = 1.3, rho_ocean = 1026.0;
    double C_atm = 1.2e-3, C_ocean = 5.5e-3;
    double F_atm = C_atm * rho_atm, F_ocean = C_ocean * rho_ocean;
    double fc = 1.45842e-4;
    double ocean_turning_angle = 0.0;
    double gravity = 9.81;
    // VP
    double Pstar = 27500.0, DeltaMin = 2.e-9;
    double alpha = 1500.0, beta = 1500.0; // MEVPStressUpdateStep.hpp:126-127, VPCGDynamicsKernel.hpp:125-126
    // MEB / BBM
    double compaction_param = -20., nu0 = 1. / 3., young = 5.96e8, P0 = 10.e3;
    double undamaged_time_relaxation_sigma = 1e7;
    int exponent_relaxation_sigma = 5;
    double exponent_compression_factor = 1.5;
    double tan_phi = 0.7, compr_strength = 1e10, C_lab = 2.0e6;
};

// ----------------------------------------------------------------------------------
// ParametricMomentumMap<CG,DG>  (ParametricMap.cpp:94-356)
// ----------------------------------------------------------------------------------
template <int CG, int DG> struct MomentumMap {
    static constexpr int DGs = cg2dgstress(CG), GS = gp1d(DGs), Q = GS * GS, ND = cgdofs(CG);
    const Mesh& m;
    Vec lumpedcgmass, lumpedcg1mass;
    Vec divS1, divS2, divM; //!< N x ND x DGs
    Vec iMgradX, iMgradY, iMM; //!< N x DGs x ND
    Vec iMJwPSI; //!< N x DGs x Q
    Vec iMJwPSI_dam; //!< N x DG x Q
    Vec dX_SSH, dY_SSH; //!< N x 4 x 4

    explicit MomentumMap(const Mesh& mesh)
        : m(mesh)
    {
    }

    //! ParametricMap.cpp:94-206
    void InitializeLumpedCGMassMatrix()
    {
        constexpr int CGGP = (CG == 1 ? 1 : 4), QQ = CGGP * CGGP;
        const size_t sy = CG * m.nx + 1;
        lumpedcgmass.assign(sy * (CG * m.ny + 1), 0.0);
        const auto& P = CGTab<CG, CGGP>::get();
        const auto& W = DGTab<1, CGGP>::get();
        for (size_t p = 0; p < 2; ++p)
            for (size_t iy = p; iy < m.ny; iy += 2)
                for (size_t ix = 0; ix < m.nx; ++ix) {
                    const size_t eid = m.nx * iy + ix;
                    double J[QQ], gp[2][QQ];
                    jacobian<CGGP>(m, eid, J);
                    if (m.spherical)
                        gaussPointsInElement<CGGP>(m, eid, gp);
                    for (int q = 0; q < QQ; ++q) {
                        J[q] = J[q] * W.w[q];
                        if (m.spherical)
                            J[q] *= cos(gp[1][q]);
                    }
                    const size_t n0 = CG * iy * sy + CG * ix;
                    for (int r = 0; r <= CG; ++r)
                        for (int c = 0; c <= CG; ++c) {
                            const int i = r * (CG + 1) + c;
                            double s = 0;
                            for (int q = 0; q < QQ; ++q)
                                s += P.phi[i][q] * J[q];
                            lumpedcgmass[n0 + c + r * sy] += s;
                        }
                }
        const size_t s1 = m.nx + 1;
        lumpedcg1mass.assign(s1 * (m.ny + 1), 0.0);
        const auto& P1 = CGTab<1, 2>::get();
        const auto& W2 = DGTab<1, 2>::get();
        for (size_t p = 0; p < 2; ++p)
            for (size_t iy = p; iy < m.ny; iy += 2)
                for (size_t ix = 0; ix < m.nx; ++ix) {
                    const size_t eid = m.nx * iy + ix;
                    double J[4], gp[2][4];
                    jacobian<2>(m, eid, J);
                    if (m.spherical)
                        gaussPointsInElement<2>(m, eid, gp);
                    for (int q = 0; q < 4; ++q) {
                        J[q] = J[q] * W2.w[q];
                        if (m.spherical)
                            J[q] *= cos(gp[1][q]);
                    }
                    const size_t n0 = iy * s1 + ix;
                    for (int r = 0; r < 2; ++r)
                        for (int c = 0; c < 2; ++c) {
                            double s = 0;
                            for (int q = 0; q < 4; ++q)
                                s += P1.phi[r * 2 + c][q] * J[q];
                            lumpedcg1mass[n0 + c + r * s1] += s;
                        }
                }
    }

    //! ParametricMap.cpp:209-356
    void InitializeDivSMatrices()
    {
        const size_t N = m.nelements;
        divS1.assign(N * ND * DGs, 0.0);
        divS2.assign(N * ND * DGs, 0.0);
        iMgradX.assign(N * DGs * ND, 0.0);
        iMgradY.assign(N * DGs * ND, 0.0);
        iMJwPSI.assign(N * DGs * Q, 0.0);
        iMJwPSI_dam.assign(N * DG * Q, 0.0);
        dX_SSH.assign(N * 16, 0.0);
        dY_SSH.assign(N * 16, 0.0);
        if (m.spherical) {
            divM.assign(N * ND * DGs, 0.0);
            iMM.assign(N * DGs * ND, 0.0);
        }
        const auto& S = DGTab<DGs, GS>::get(); // PSI<DGs,GS>, weights
        const auto& A = DGTab<DG, GS>::get(); // PSI<DG,GS> (damage)
        const auto& P = CGTab<CG, GS>::get();
        const auto& P1 = CGTab<1, GS>::get();
        const bool sph = m.spherical;
#pragma omp parallel for
        for (size_t e = 0; e < N; ++e) {
            double Fx[2][Q], Fy[2][Q], J[Q], gp[2][Q], cl[Q], sl[Q];
            dxT<GS>(m, e, Fx);
            dyT<GS>(m, e, Fy);
            for (int k = 0; k < 2; ++k)
                for (int q = 0; q < Q; ++q) {
                    Fx[k][q] *= S.w[q];
                    Fy[k][q] *= S.w[q];
                }
            jacobian<GS>(m, e, J);
            if (sph) {
                gaussPointsInElement<GS>(m, e, gp);
                for (int q = 0; q < Q; ++q) {
                    cl[q] = cos(gp[1][q]);
                    sl[q] = sin(gp[1][q]);
                }
            }
            double dxc[ND][Q], dyc[ND][Q], dx1[4][Q], dy1[4][Q];
            for (int i = 0; i < ND; ++i)
                for (int q = 0; q < Q; ++q) {
                    dxc[i][q] = P.phix[i][q] * Fy[1][q] - P.phiy[i][q] * Fx[1][q];
                    dyc[i][q] = P.phiy[i][q] * Fx[0][q] - P.phix[i][q] * Fy[0][q];
                }
            for (int i = 0; i < 4; ++i)
                for (int q = 0; q < Q; ++q) {
                    dx1[i][q] = P1.phix[i][q] * Fy[1][q] - P1.phiy[i][q] * Fx[1][q];
                    dy1[i][q] = P1.phiy[i][q] * Fx[0][q] - P1.phix[i][q] * Fy[0][q];
                }
            double* d1 = &divS1[e * ND * DGs];
            double* d2 = &divS2[e * ND * DGs];
            for (int i = 0; i < ND; ++i)
                for (int j = 0; j < DGs; ++j) {
                    double a = 0, b = 0, c = 0;
                    for (int q = 0; q < Q; ++q) {
                        a += dxc[i][q] * S.psi[j][q];
                        b += (sph ? dyc[i][q] * cl[q] : dyc[i][q]) * S.psi[j][q];
                        if (sph)
                            c += (P.phi[i][q] * (J[q] * sl[q] * S.w[q])) * S.psi[j][q];
                    }
                    d1[i * DGs + j] = sph ? a / EarthRadius : a;
                    d2[i * DGs + j] = sph ? b / EarthRadius : b;
                    if (sph)
                        divM[(e * ND + i) * DGs + j] = c / EarthRadius;
                }
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) {
                    double a = 0, b = 0;
                    for (int q = 0; q < Q; ++q) {
                        a += dx1[i][q] * P1.phi[j][q];
                        b += (sph ? dy1[i][q] * cl[q] : dy1[i][q]) * P1.phi[j][q];
                    }
                    dX_SSH[e * 16 + i * 4 + j] = sph ? a / EarthRadius : a;
                    dY_SSH[e * 16 + i * 4 + j] = sph ? b / EarthRadius : b;
                }
            double M[DGs][DGs], iM[DGs][DGs];
            massMatrix<DGs>(m, e, sph, M);
            inverse<DGs>(M, iM);
            for (int i = 0; i < DGs; ++i) {
                for (int k = 0; k < ND; ++k) {
                    double a = 0, b = 0, c = 0;
                    for (int j = 0; j < DGs; ++j) {
                        a += iM[i][j] * d1[k * DGs + j];
                        b += iM[i][j] * d2[k * DGs + j];
                        if (sph)
                            c += iM[i][j] * divM[(e * ND + k) * DGs + j];
                    }
                    iMgradX[(e * DGs + i) * ND + k] = a;
                    iMgradY[(e * DGs + i) * ND + k] = b;
                    if (sph)
                        iMM[(e * DGs + i) * ND + k] = c;
                }
                for (int q = 0; q < Q; ++q) {
                    double a = 0;
                    for (int j = 0; j < DGs; ++j)
                        a += iM[i][j] * (S.psi[j][q] * (S.w[q] * J[q]));
                    iMJwPSI[(e * DGs + i) * Q + q] = a;
                }
            }
            double Md[DG][DG], iMd[DG][DG];
            massMatrix<DG>(m, e, sph, Md);
            inverse<DG>(Md, iMd);
            for (int i = 0; i < DG; ++i)
                for (int q = 0; q < Q; ++q) {
                    double a = 0;
                    for (int j = 0; j < DG; ++j)
                        a += iMd[i][j] * (A.psi[j][q] * (S.w[q] * J[q]));
                    iMJwPSI_dam[(e * DG + i) * Q + q] = a;
                }
        }
    }
};

//! DGModelArray::ma2dg (DGModelArray.hpp:20-32, quirk Q4). data = row-major N x ncomp.
template <int N> void ma2dg(const double* data, int ncomp, size_t nel, Vec& dg)
{
    dg.assign(nel * N, 0.0);
    if (ncomp == N) {
        std::copy(data, data + nel * N, dg.begin());
    } else if (ncomp == 1) {
        for (size_t i = 0; i < nel; ++i)
            dg[i * N] = data[i];
    } else
        throw std::runtime_error("ma2dg: component count must be 1 or the DG size");
}

//! Type-erased interface for the C API
class IDynOracle {
public:
    virtual ~IDynOracle() = default;
    virtual void initialise(size_t nx, size_t ny, const double* coords, const double* mask, bool sph) = 0;
    virtual void setData(const std::string& name, const double* data, int ncomp) = 0;
    virtual void update(double dt) = 0;
    virtual void getDG0Data(const std::string& name, double* out) = 0;
    virtual int getDGData(const std::string& name, double* out) = 0;
    virtual const Vec* raw(const std::string& name) = 0;
    virtual Vec* rawMutable(const std::string& name) = 0;
    virtual const Mesh& mesh() const = 0;
    virtual Mesh& meshMutable() = 0;
    virtual void subcycle() = 0;
    virtual void sweep(const std::string& which) = 0;
    virtual void setDeltaT(double) = 0;
    size_t nSteps = 100; //!< DynamicsKernel.hpp:187 (hard-coded 100 in the reference, quirk Q5)
    Params params;
    double subcycleSeconds = 0; //!< wall time spent in the subcycle loop of the last update()
};

// ----------------------------------------------------------------------------------
// The dynamics kernel: DynamicsKernel + CGDynamicsKernel + VP / Brittle specialisations
// ----------------------------------------------------------------------------------
template <int DGadv, int CG> class DynOracle : public IDynOracle {
public:
    static constexpr int DGs = cg2dgstress(CG), GS = gp1d(DGs), Q = GS * GS, ND = cgdofs(CG);
    const Rheology rheology;
    Mesh m;
    Transport<DGadv>* dgtransport = nullptr;
    Transport<DGs>* stresstransport = nullptr;
    MomentumMap<CG, DGadv>* pmap = nullptr;

    Vec hice, cice, damage; //!< N x DGadv
    Vec ssh; //!< N (DG0)
    Vec e11, e12, e22, s11, s12, s22; //!< N x DGs
    Vec u, v, cgA, cgH, uGradSSH, vGradSSH, dStressX, dStressY, uOcean, vOcean, uAtmos, vAtmos;
    Vec u0, v0, avgU, avgV;
    double deltaT = 0;
    double cosOceanAngle = 1, sinOceanAngle = 0;
    std::map<std::string, Vec> advectedFields;

    explicit DynOracle(Rheology r)
        : rheology(r)
    {
    }
    ~DynOracle()
    {
        delete dgtransport;
        delete stresstransport;
        delete pmap;
    }
    const Mesh& mesh() const override { return m; }
    Mesh& meshMutable() override { return m; }

    //! DynamicsKernel.hpp:44-75, CGDynamicsKernel.cpp:24-51, BrittleCGDynamicsKernel.hpp:72-86
    void initialise(size_t nx, size_t ny, const double* coords, const double* mask, bool sph) override
    {
        m.initFromArrays(nx, ny, coords, mask, sph);
        delete dgtransport;
        delete stresstransport;
        delete pmap;
        stresstransport = nullptr;
        dgtransport = new Transport<DGadv>(m);
        dgtransport->scheme = "rk2";
        const size_t N = m.nelements, NCG = (CG * nx + 1) * (CG * ny + 1);
        for (Vec* x : { &hice, &cice, &damage })
            x->assign(N * DGadv, 0.0);
        ssh.assign(N, 0.0);
        for (Vec* x : { &e11, &e12, &e22, &s11, &s12, &s22 })
            x->assign(N * DGs, 0.0);
        pmap = new MomentumMap<CG, DGadv>(m);
        pmap->InitializeLumpedCGMassMatrix();
        pmap->InitializeDivSMatrices();
        for (Vec* x : { &u, &v, &cgA, &cgH, &uGradSSH, &vGradSSH, &dStressX, &dStressY, &uOcean, &vOcean,
                 &uAtmos, &vAtmos, &u0, &v0, &avgU, &avgV })
            x->assign(NCG, 0.0);
        if (rheology == BBM) {
            stresstransport = new Transport<DGs>(m);
            stresstransport->scheme = "rk2";
            const double radians = 0x1.1df46a2529d39p-6;
            cosOceanAngle = cos(radians * params.ocean_turning_angle);
            sinOceanAngle = sin(radians * params.ocean_turning_angle);
        }
    }

    //! DynamicsKernel.hpp:92-112, CGDynamicsKernel.cpp:53-90, BrittleCGDynamicsKernel.hpp:138-145
    void setData(const std::string& name, const double* data, int ncomp) override
    {
        const size_t N = m.nelements;
        auto viaDG2CG = [&](Vec& dest) {
            Vec tmp;
            ma2dg<DGadv>(data, ncomp, N, tmp);
            DG2CG<CG, DGadv>(m, dest, tmp);
        };
        if (name == "u")
            viaDG2CG(u);
        else if (name == "v")
            viaDG2CG(v);
        else if (name == "uwind")
            viaDG2CG(uAtmos);
        else if (name == "vwind")
            viaDG2CG(vAtmos);
        else if (name == "uocean")
            viaDG2CG(uOcean);
        else if (name == "vocean")
            viaDG2CG(vOcean);
        else if (name == "damage" && rheology == BBM)
            ma2dg<DGadv>(data, ncomp, N, damage);
        else if (name == "hice")
            ma2dg<DGadv>(data, ncomp, N, hice);
        else if (name == "cice")
            ma2dg<DGadv>(data, ncomp, N, cice);
        else if (name == "ssh")
            ma2dg<1>(data, ncomp, N, ssh);
        else
            ma2dg<DGadv>(data, ncomp, N, advectedFields[name]);
    }

    //! CGDynamicsKernel.cpp:126-258
    void ComputeGradientOfSeaSurfaceHeight()
    {
        const size_t nx = m.nx, ny = m.ny, cg1row = nx + 1;
        Vec cgSSH;
        DG2CG<1, 1>(m, cgSSH, ssh);
        Vec uG(cg1row * (ny + 1), 0.0), vG(cg1row * (ny + 1), 0.0);
        for (size_t p = 0; p < 2; ++p)
            for (size_t cy = 0; cy < ny; ++cy) {
                if (cy % 2 != p)
                    continue;
                size_t eid = nx * cy, id = cy * cg1row;
                for (size_t cx = 0; cx < nx; ++cx, ++eid, ++id) {
                    const size_t n[4] = { id, id + 1, id + cg1row, id + cg1row + 1 };
                    const double loc[4] = { cgSSH[n[0]], cgSSH[n[1]], cgSSH[n[2]], cgSSH[n[3]] };
                    for (int i = 0; i < 4; ++i) {
                        double tx = 0, ty = 0;
                        for (int j = 0; j < 4; ++j) {
                            tx += pmap->dX_SSH[eid * 16 + i * 4 + j] * loc[j];
                            ty += pmap->dY_SSH[eid * 16 + i * 4 + j] * loc[j];
                        }
                        uG[n[i]] -= tx;
                        vG[n[i]] -= ty;
                    }
                }
            }
        for (size_t i = 0; i < uG.size(); ++i) {
            uG[i] /= pmap->lumpedcg1mass[i];
            vG[i] /= pmap->lumpedcg1mass[i];
        }
        const size_t topleft = ny * cg1row;
        for (size_t i = 1; i < nx; ++i) {
            uG[i] = uG[i + cg1row];
            vG[i] = vG[i + cg1row];
            uG[topleft + i] = uG[topleft + i - cg1row];
            vG[topleft + i] = vG[topleft + i - cg1row];
        }
        for (size_t i = 1; i < ny; ++i) {
            uG[i * cg1row] = uG[i * cg1row + 1];
            vG[i * cg1row] = vG[i * cg1row + 1];
            uG[i * cg1row + cg1row - 1] = uG[i * cg1row + cg1row - 2];
            vG[i * cg1row + cg1row - 1] = vG[i * cg1row + cg1row - 2];
        }
        uG[0] = uG[cg1row + 1];
        vG[0] = vG[cg1row + 1];
        uG[nx] = uG[nx + cg1row - 1];
        vG[nx] = vG[nx + cg1row - 1];
        uG[ny * cg1row] = uG[(ny - 1) * cg1row + 1];
        vG[ny * cg1row] = vG[(ny - 1) * cg1row + 1];
        uG[(ny + 1) * cg1row - 1] = uG[ny * cg1row - 2];
        vG[(ny + 1) * cg1row - 1] = vG[ny * cg1row - 2];

        if (CG == 1) {
            uGradSSH = uG;
            vGradSSH = vG;
            return;
        }
        const size_t r2 = 2 * nx + 1;
        for (size_t iy = 0; iy <= ny; ++iy)
            for (size_t ix = 0; ix <= nx; ++ix) {
                uGradSSH[r2 * 2 * iy + 2 * ix] = uG[cg1row * iy + ix];
                vGradSSH[r2 * 2 * iy + 2 * ix] = vG[cg1row * iy + ix];
            }
        for (size_t iy = 0; iy <= ny; ++iy)
            for (size_t ix = 0; ix < nx; ++ix) {
                const size_t i1 = cg1row * iy + ix, i2 = r2 * 2 * iy + 1 + 2 * ix;
                uGradSSH[i2] = 0.5 * (uG[i1] + uG[i1 + 1]);
                vGradSSH[i2] = 0.5 * (vG[i1] + vG[i1 + 1]);
            }
        for (size_t iy = 0; iy < ny; ++iy)
            for (size_t ix = 0; ix <= nx; ++ix) {
                const size_t i1 = cg1row * iy + ix, i2 = r2 * (2 * iy + 1) + 2 * ix;
                uGradSSH[i2] = 0.5 * (uG[i1] + uG[i1 + cg1row]);
                vGradSSH[i2] = 0.5 * (vG[i1] + vG[i1 + cg1row]);
            }
        for (size_t iy = 0; iy < ny; ++iy)
            for (size_t ix = 0; ix < nx; ++ix) {
                const size_t i1 = cg1row * iy + ix, i2 = r2 * (2 * iy + 1) + 1 + 2 * ix;
                uGradSSH[i2] = 0.25 * (uG[i1] + uG[i1 + 1] + uG[i1 + cg1row] + uG[i1 + cg1row + 1]);
                vGradSSH[i2] = 0.25 * (vG[i1] + vG[i1 + 1] + vG[i1 + cg1row] + vG[i1 + cg1row + 1]);
            }
    }

    //! VectorManipulations.hpp:26-65
    void CGAveragePeriodic(Vec& x) const
    {
        const size_t row = CG * m.nx + 1;
        for (const auto& seg : m.periodic)
            for (const auto& p : seg) {
                const size_t lb = p[2], rt = p[1];
                const size_t i0lb = row * CG * (lb / m.nx) + CG * (lb % m.nx);
                const size_t i0rt = row * CG * (rt / m.nx) + CG * (rt % m.nx);
                for (size_t j = 0; j <= CG; ++j) {
                    size_t i1, i2;
                    if (p[0] == 0) {
                        i1 = i0lb + j;
                        i2 = i0rt + CG * row + j;
                    } else {
                        i1 = i0lb + j * row;
                        i2 = i0rt + CG + j * row;
                    }
                    x[i1] = 0.5 * (x[i1] + x[i2]);
                    x[i2] = x[i1];
                }
            }
    }

    //! CGDynamicsKernel.cpp:260-276
    void prepareIteration()
    {
        DG2CG<CG, DGadv>(m, cgH, hice);
        CGAveragePeriodic(cgH);
        DG2CG<CG, DGadv>(m, cgA, cice);
        CGAveragePeriodic(cgA);
        ComputeGradientOfSeaSurfaceHeight();
        for (auto& a : cgA)
            a = std::max(std::min(a, 1.0), 1.e-4);
        for (auto& h : cgH)
            h = std::max(h, 1.e-4);
    }

    //! CGDynamicsKernel.cpp:300-337
    void projectVelocityToStrain()
    {
#pragma omp parallel for
        for (size_t row = 0; row < m.ny; ++row)
            for (size_t col = 0; col < m.nx; ++col) {
                const size_t e = m.nx * row + col;
                if (m.landmask[e] == 0)
                    continue;
                double ul[ND], vl[ND];
                cgLocal<CG>(m, u, e, ul);
                cgLocal<CG>(m, v, e, vl);
                const double* gx = &pmap->iMgradX[e * DGs * ND];
                const double* gy = &pmap->iMgradY[e * DGs * ND];
                for (int i = 0; i < DGs; ++i) {
                    double xu = 0, yv = 0, xv = 0, yu = 0;
                    for (int k = 0; k < ND; ++k) {
                        xu += gx[i * ND + k] * ul[k];
                        yv += gy[i * ND + k] * vl[k];
                        xv += gx[i * ND + k] * vl[k];
                        yu += gy[i * ND + k] * ul[k];
                    }
                    e11[e * DGs + i] = xu;
                    e22[e * DGs + i] = yv;
                    e12[e * DGs + i] = 0.5 * (xv + yu);
                }
                if (m.spherical) {
                    const double* mm = &pmap->iMM[e * DGs * ND];
                    for (int i = 0; i < DGs; ++i) {
                        double mv = 0, mu = 0;
                        for (int k = 0; k < ND; ++k) {
                            mv += mm[i * ND + k] * vl[k];
                            mu += mm[i * ND + k] * ul[k];
                        }
                        e11[e * DGs + i] -= mv;
                        e12[e * DGs + i] += 0.5 * mu;
                    }
                }
            }
    }

    //! field row (n comps) evaluated in the Q stress Gauss points
    template <int NC> static inline void toGaussPts(const double* row, double* out)
    {
        const auto& T = DGTab<NC, GS>::get();
        for (int q = 0; q < Q; ++q) {
            double s = 0;
            for (int j = 0; j < NC; ++j)
                s += row[j] * T.psi[j][q];
            out[q] = s;
        }
    }

    //! MEVPStressUpdateStep.hpp:30-118 (runs on ALL elements, quirk Q8)
    void stressUpdateMEVP()
    {
        const double alpha = params.alpha;
#pragma omp parallel for
        for (size_t i = 0; i < m.nelements; ++i) {
            double hg[Q], ag[Q], g11[Q], g12[Q], g22[Q];
            toGaussPts<DGadv>(&hice[i * DGadv], hg);
            toGaussPts<DGadv>(&cice[i * DGadv], ag);
            toGaussPts<DGs>(&e11[i * DGs], g11);
            toGaussPts<DGs>(&e12[i * DGs], g12);
            toGaussPts<DGs>(&e22[i * DGs], g22);
            double r11[Q], r12[Q], r22[Q];
            for (int q = 0; q < Q; ++q) {
                const double h = std::max(hg[q], 0.0);
                const double a = std::min(std::max(ag[q], 0.0), 1.0);
                const double DELTA = sqrt(SQR(params.DeltaMin) + 1.25 * (SQR(g11[q]) + SQR(g22[q]))
                    + 1.50 * g11[q] * g22[q] + SQR(g12[q]));
                const double P = params.Pstar * h * exp(-20.0 * (1.0 - a));
                r11[q] = 1.0 / alpha * (P / 8.0 / DELTA * (5.0 * g11[q] + 3.0 * g22[q]) - 0.5 * P);
                r12[q] = 1.0 / alpha * (P / 4.0 / DELTA * g12[q]);
                r22[q] = 1.0 / alpha * (P / 8.0 / DELTA * (5.0 * g22[q] + 3.0 * g11[q]) - 0.5 * P);
            }
            const double* im = &pmap->iMJwPSI[i * DGs * Q];
            for (int j = 0; j < DGs; ++j) {
                double a = 0, b = 0, c = 0;
                for (int q = 0; q < Q; ++q) {
                    a += im[j * Q + q] * r11[q];
                    b += im[j * Q + q] * r12[q];
                    c += im[j * Q + q] * r22[q];
                }
                s11[i * DGs + j] = s11[i * DGs + j] * (1.0 - 1.0 / alpha) + a;
                s12[i * DGs + j] = s12[i * DGs + j] * (1.0 - 1.0 / alpha) + b;
                s22[i * DGs + j] = s22[i * DGs + j] * (1.0 - 1.0 / alpha) + c;
            }
        }
    }

    //! BBMStressUpdateStep.hpp:31-196 (runs on ALL elements; select() semantics, quirk Q13)
    void stressUpdateBBM()
    {
        const Params& p = params;
        const double dt = deltaT;
#pragma omp parallel for
        for (size_t i = 0; i < m.nelements; ++i) {
            double hG[Q], aG[Q], dG[Q], e11G[Q], e12G[Q], e22G[Q], s11G[Q], s12G[Q], s22G[Q];
            toGaussPts<DGadv>(&hice[i * DGadv], hG);
            toGaussPts<DGadv>(&cice[i * DGadv], aG);
            toGaussPts<DGadv>(&damage[i * DGadv], dG);
            toGaussPts<DGs>(&e11[i * DGs], e11G);
            toGaussPts<DGs>(&e12[i * DGs], e12G);
            toGaussPts<DGs>(&e22[i * DGs], e22G);
            toGaussPts<DGs>(&s11[i * DGs], s11G);
            toGaussPts<DGs>(&s12[i * DGs], s12G);
            toGaussPts<DGs>(&s22[i * DGs], s22G);
            const double hel = m.h(i);
            const double scale_coef = std::sqrt(0.1 / hel);
            for (int q = 0; q < Q; ++q) {
                const double h = std::max(hG[q], 0.0);
                const double a = std::min(std::max(aG[q], 0.0), 1.0);
                double d = std::min(std::max(dG[q], 1e-12), 1.0);
                double sigma_n = 0.5 * (s11G[q] + s22G[q]);
                const double expC = exp(p.compaction_param * (1.0 - a));
                const double powalphaexpC = pow(d * expC, p.exponent_relaxation_sigma - 1);
                const double time_viscous = p.undamaged_time_relaxation_sigma * powalphaexpC;
                const double Pmax = p.P0 * pow(h, p.exponent_compression_factor + 1.) * expC;
                const double tildeP = (sigma_n < 0.0) ? std::min(-Pmax / sigma_n, 1.0) : 0.;
                const double multiplicator = time_viscous / (time_viscous + (1. - tildeP) * dt);
                const double elasticity = h * p.young * d * expC;
                const double Dunit = dt * elasticity / (1. - (p.nu0 * p.nu0));
                s11G[q] += Dunit * (e11G[q] + p.nu0 * e22G[q]);
                s22G[q] += Dunit * (p.nu0 * e11G[q] + e22G[q]);
                s12G[q] += Dunit * e12G[q] * (1. - p.nu0);
                s11G[q] *= multiplicator;
                s22G[q] *= multiplicator;
                s12G[q] *= multiplicator;
                sigma_n = 0.5 * (s11G[q] + s22G[q]);
                const double tau = sqrt(0.25 * SQR(s11G[q] - s22G[q]) + SQR(s12G[q]));
                const double cohesion = p.C_lab * scale_coef * h;
                const double compr = p.compr_strength * scale_coef * h;
                double dcrit = (tau + p.tan_phi * sigma_n > 0.) ? cohesion / (tau + p.tan_phi * sigma_n) : 1.;
                if (sigma_n < -compr)
                    dcrit = -compr / sigma_n;
                dcrit = std::min(dcrit, 1.0);
                const double td = hel * std::sqrt(2. * (1. + p.nu0) * p.rho_ice) / sqrt(elasticity);
                d -= d * (1. - dcrit) * dt / td;
                s11G[q] -= s11G[q] * (1. - dcrit) * dt / td;
                s12G[q] -= s12G[q] * (1. - dcrit) * dt / td;
                s22G[q] -= s22G[q] * (1. - dcrit) * dt / td;
                dG[q] = d;
            }
            const double* im = &pmap->iMJwPSI[i * DGs * Q];
            for (int j = 0; j < DGs; ++j) {
                double a = 0, b = 0, c = 0;
                for (int q = 0; q < Q; ++q) {
                    a += im[j * Q + q] * s11G[q];
                    b += im[j * Q + q] * s12G[q];
                    c += im[j * Q + q] * s22G[q];
                }
                s11[i * DGs + j] = a;
                s12[i * DGs + j] = b;
                s22[i * DGs + j] = c;
            }
            const double* imd = &pmap->iMJwPSI_dam[i * DGadv * Q];
            for (int j = 0; j < DGadv; ++j) {
                double a = 0;
                for (int q = 0; q < Q; ++q)
                    a += imd[j * Q + q] * dG[q];
                damage[i * DGadv + j] = a;
            }
        }
    }

    //! CGDynamicsKernel.cpp:401-437
    void dirichletZero(Vec& x) const
    {
        const size_t row = CG * m.nx + 1;
        for (size_t seg = 0; seg < 4; ++seg)
            for (size_t i = 0; i < m.dirichlet[seg].size(); ++i) {
                const size_t eid = m.dirichlet[seg][i];
                const size_t ix = eid % m.nx, iy = eid / m.nx;
                for (size_t j = 0; j < CG + 1; ++j) {
                    if (seg == 0)
                        x[iy * CG * row + CG * ix + j] = 0.0;
                    else if (seg == 1)
                        x[iy * CG * row + CG * ix + CG + row * j] = 0.0;
                    else if (seg == 2)
                        x[(iy + 1) * CG * row + CG * ix + j] = 0.0;
                    else
                        x[iy * CG * row + CG * ix + row * j] = 0.0;
                }
            }
    }

    //! CGDynamicsKernel.cpp:340-398 (even rows first, quirk Q7)
    void stressDivergence()
    {
        std::fill(dStressX.begin(), dStressX.end(), 0.0);
        std::fill(dStressY.begin(), dStressY.end(), 0.0);
        const size_t cgRow = CG * m.nx + 1;
        for (size_t p = 0; p < 2; ++p) {
#pragma omp parallel for
            for (size_t cy = 0; cy < m.ny; ++cy) {
                if (cy % 2 != p)
                    continue;
                for (size_t cx = 0; cx < m.nx; ++cx) {
                    const size_t c = m.nx * cy + cx;
                    if (m.landmask[c] != 1)
                        continue;
                    const double* d1 = &pmap->divS1[c * ND * DGs];
                    const double* d2 = &pmap->divS2[c * ND * DGs];
                    const double* dm = m.spherical ? &pmap->divM[c * ND * DGs] : nullptr;
                    const size_t cg_i = CG * cgRow * cy + CG * cx;
                    for (int r = 0; r <= CG; ++r)
                        for (int k = 0; k <= CG; ++k) {
                            const int i = k + (CG + 1) * r;
                            double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
                            for (int j = 0; j < DGs; ++j) {
                                a1 += d1[i * DGs + j] * s11[c * DGs + j];
                                a2 += d2[i * DGs + j] * s12[c * DGs + j];
                                b1 += d1[i * DGs + j] * s12[c * DGs + j];
                                b2 += d2[i * DGs + j] * s22[c * DGs + j];
                            }
                            double tx = a1 + a2, ty = b1 + b2;
                            if (dm) {
                                double m12 = 0, m11 = 0;
                                for (int j = 0; j < DGs; ++j) {
                                    m12 += dm[i * DGs + j] * s12[c * DGs + j];
                                    m11 += dm[i * DGs + j] * s11[c * DGs + j];
                                }
                                tx += m12;
                                ty -= m11;
                            }
                            dStressX[cg_i + k + r * cgRow] -= tx;
                            dStressY[cg_i + k + r * cgRow] -= ty;
                        }
                }
            }
        }
        dirichletZero(dStressX);
        dirichletZero(dStressY);
        CGAveragePeriodic(dStressX);
        CGAveragePeriodic(dStressY);
    }

    //! VPCGDynamicsKernel.hpp:132-172 (quirks Q1, Q2 reproduced verbatim)
    void updateMomentumVP()
    {
        const Params& p = params;
        const double beta = p.beta, SC = 1.0;
#pragma omp parallel for
        for (size_t i = 0; i < u.size(); ++i) {
            const double uOcnRel = uOcean[i] - u[i];
            const double vOcnRel = v[i] - vOcean[i];
            const double absatm = sqrt(SQR(uAtmos[i]) + SQR(vAtmos[i]));
            const double absocn = sqrt(SQR(uOcnRel) + SQR(vOcnRel));
            u[i] = (1.0 / (p.rho_ice * cgH[i] / deltaT * (1.0 + beta) + cgA[i] * p.F_ocean * absocn)
                * (p.rho_ice * cgH[i] / deltaT * (beta * u[i] + u0[i])
                    + cgA[i] * (p.F_atm * absatm * uAtmos[i] + p.F_ocean * absocn * SC * uOcean[i])
                    - p.rho_ice * cgH[i] * p.fc * u[i] - p.rho_ice * cgH[i] * p.gravity * uGradSSH[i]
                    + dStressX[i] / pmap->lumpedcgmass[i]));
            v[i] = (1.0 / (p.rho_ice * cgH[i] / deltaT * (1.0 + beta) + cgA[i] * p.F_ocean * absocn)
                * (p.rho_ice * cgH[i] / deltaT * (beta * v[i] + v0[i])
                    + cgA[i] * (p.F_atm * absatm * vAtmos[i] + p.F_ocean * absocn * SC * vOcean[i])
                    + p.rho_ice * cgH[i] * p.fc * v[i] - p.rho_ice * cgH[i] * p.gravity * vGradSSH[i]
                    + dStressY[i] / pmap->lumpedcgmass[i]));
        }
    }

    //! BrittleCGDynamicsKernel.hpp:206-254 (quirk Q3 reproduced verbatim)
    void updateMomentumBrittle()
    {
        const Params& p = params;
#pragma omp parallel for
        for (size_t i = 0; i < u.size(); ++i) {
            const double dteOverMass = deltaT / (p.rho_ice * cgH[i]);
            const double uIce = u[i], vIce = v[i];
            const double cPrime = cgA[i] * p.F_ocean * std::hypot(uOcean[i] - uIce, vOcean[i] - vIce);
            const double tauB = 0.;
            const double alpha = 1 + dteOverMass * (cPrime * cosOceanAngle + tauB);
            const double beta = deltaT * p.fc + dteOverMass * cPrime * sinOceanAngle;
            const double rDenom = 1 / (SQR(alpha) + SQR(beta));
            const double dragAtm = cgA[i] * p.F_atm * std::hypot(uAtmos[i], vAtmos[i]);
            const double tauX = dragAtm * uAtmos[i] + cPrime * (uOcean[i] * cosOceanAngle - vOcean[i] * sinOceanAngle);
            const double tauY = dragAtm * vAtmos[i] + cPrime * (vOcean[i] * cosOceanAngle + uOcean[i] * sinOceanAngle);
            const double gradX = dStressX[i] / pmap->lumpedcgmass[i] - p.rho_ice * cgH[i] * p.gravity * uGradSSH[i];
            const double gradY = dStressY[i] / pmap->lumpedcgmass[i] - p.rho_ice * cgH[i] * p.gravity * vGradSSH[i];
            u[i] = alpha * uIce + beta * vIce + dteOverMass * (alpha * (gradX + tauX) + beta * (gradY + tauY));
            u[i] *= rDenom;
            v[i] = alpha * vIce - beta * uIce + dteOverMass * (alpha * (gradY + tauY) + beta * (gradX + tauX));
            v[i] *= rDenom;
            avgU[i] += u[i] / nSteps;
            avgV[i] += v[i] / nSteps;
        }
    }

    //! One subcycle body: VPCGDynamicsKernel.hpp:76-91 / BrittleCGDynamicsKernel.hpp:116-133
    void subcycle() override
    {
        projectVelocityToStrain();
        if (rheology == MEVP)
            stressUpdateMEVP();
        else
            stressUpdateBBM();
        stressDivergence();
        if (rheology == MEVP)
            updateMomentumVP();
        else
            updateMomentumBrittle();
        dirichletZero(u); // applyBoundaries, CGDynamicsKernel.cpp:439-444
        dirichletZero(v);
    }

    void setDeltaT(double d) override { deltaT = d; }

    //! a single sweep of the subcycle (or prepareIteration), for the analytic tests
    void sweep(const std::string& which) override
    {
        if (which == "strain")
            projectVelocityToStrain();
        else if (which == "stress")
            rheology == MEVP ? stressUpdateMEVP() : stressUpdateBBM();
        else if (which == "divergence")
            stressDivergence();
        else if (which == "momentum")
            rheology == MEVP ? updateMomentumVP() : updateMomentumBrittle();
        else if (which == "boundaries") {
            dirichletZero(u);
            dirichletZero(v);
        } else if (which == "prepare")
            prepareIteration();
        else
            throw std::runtime_error("unknown sweep " + which);
    }

    //! FreeDriftDynamicsKernel::updateMomentum, FreeDriftDynamicsKernel.hpp:57-68
    void updateMomentumFreeDrift()
    {
        const double NansenNumber = sqrt(params.F_atm / params.F_ocean);
        for (size_t i = 0; i < u.size(); ++i) {
            u[i] = uOcean[i] + NansenNumber * (uAtmos[i] * cosOceanAngle - vAtmos[i] * sinOceanAngle);
            v[i] = vOcean[i] + NansenNumber * (-uAtmos[i] * sinOceanAngle + vAtmos[i] * cosOceanAngle);
        }
    }

    //! VPCGDynamicsKernel.hpp:63-94 / BrittleCGDynamicsKernel.hpp:91-136 / FreeDriftDynamicsKernel.hpp:43-50
    void update(double dt) override
    {
        if (rheology == FREEDRIFT) {
            updateMomentumFreeDrift();
            dirichletZero(u);
            dirichletZero(v);
            dgtransport->template prepareAdvection<CG>(u, v);
            advect(dt);
            return;
        }
        if (rheology == MEVP) {
            dgtransport->template prepareAdvection<CG>(u, v);
            advect(dt);
            prepareIteration();
            u0 = u;
            v0 = v;
            deltaT = dt;
        } else {
            dgtransport->template prepareAdvection<CG>(avgU, avgV);
            advect(dt);
            stresstransport->template prepareAdvection<CG>(avgU, avgV);
            stresstransport->step(dt, s11);
            stresstransport->step(dt, s12);
            stresstransport->step(dt, s22);
            dgtransport->step(dt, damage);
            LimitMax<DGadv>(damage, 1.0);
            LimitMin<DGadv>(damage, 1e-12);
            prepareIteration();
            deltaT = dt / nSteps;
            std::fill(avgU.begin(), avgU.end(), 0.0);
            std::fill(avgV.begin(), avgV.end(), 0.0);
        }
        const double t0 = omp_get_wtime();
        for (size_t s = 0; s < nSteps; ++s)
            subcycle();
        subcycleSeconds = omp_get_wtime() - t0;
    }

    //! DynamicsKernel::advectionAndLimits after prepareAdvection, DynamicsKernel.hpp:160-172
    void advect(double dt)
    {
        dgtransport->step(dt, cice);
        dgtransport->step(dt, hice);
        LimitMax<DGadv>(cice, 1.0);
        LimitMin<DGadv>(cice, 0.0);
        LimitMin<DGadv>(hice, 0.0);
    }

    //! VPCGDynamicsKernel.hpp:96-120 / BrittleCGDynamicsKernel.hpp:168-192 (quirk Q14)
    void iceOceanStress(Vec& taux, Vec& tauy) const
    {
        taux.resize(u.size());
        tauy.resize(u.size());
        for (size_t i = 0; i < u.size(); ++i) {
            if (rheology == FREEDRIFT) { // FreeDriftDynamicsKernel.hpp:70-83
                const double uR = uOcean[i] - u[i], vR = vOcean[i] - v[i];
                const double cPrime = params.F_ocean * std::hypot(uR, vR);
                taux[i] = cPrime * (uR * cosOceanAngle - vR * sinOceanAngle);
                tauy[i] = cPrime * (vR * cosOceanAngle + uR * sinOceanAngle);
            } else if (rheology == MEVP) {
                const double uR = u[i] - uOcean[i], vR = v[i] - vOcean[i];
                const double absocn = sqrt(SQR(uR) + SQR(vR));
                taux[i] = params.F_ocean * absocn * uR;
                tauy[i] = params.F_ocean * absocn * vR;
            } else {
                const double uR = uOcean[i] - avgU[i], vR = vOcean[i] - avgV[i];
                const double cPrime = params.F_ocean * std::hypot(uR, vR);
                taux[i] = cPrime * (uR * cosOceanAngle - vR * sinOceanAngle);
                tauy[i] = cPrime * (vR * cosOceanAngle + uR * sinOceanAngle);
            }
        }
    }

    //! DynamicsKernel.hpp:114-126, CGDynamicsKernel.cpp:92-118, BrittleCGDynamicsKernel.hpp:147-156
    void getDG0Data(const std::string& name, double* out) override
    {
        const size_t N = m.nelements;
        auto col0 = [&](const Vec& dg, int nc) {
            for (size_t i = 0; i < N; ++i)
                out[i] = dg[i * nc];
        };
        Vec tmp;
        if (name == "hice")
            col0(hice, DGadv);
        else if (name == "cice")
            col0(cice, DGadv);
        else if (name == "damage" && rheology == BBM)
            col0(damage, DGadv);
        else if (name == "u") {
            CG2DG<CG, DGadv>(m, tmp, u);
            col0(tmp, DGadv);
        } else if (name == "v") {
            CG2DG<CG, DGadv>(m, tmp, v);
            col0(tmp, DGadv);
        } else if (name == "uiostress" || name == "viostress") {
            Vec tx, ty;
            iceOceanStress(tx, ty);
            CG2DG<CG, DGadv>(m, tmp, name == "uiostress" ? tx : ty);
            col0(tmp, DGadv);
        } else
            col0(advectedFields.at(name), DGadv);
    }

    //! DynamicsKernel.hpp:134-156, BrittleCGDynamicsKernel.hpp:158-166. Returns ncomp.
    int getDGData(const std::string& name, double* out) override
    {
        const Vec* src = (name == "hice") ? &hice
            : (name == "cice")            ? &cice
            : (name == "damage" && rheology == BBM) ? &damage
                                          : &advectedFields.at(name);
        std::copy(src->begin(), src->end(), out);
        return DGadv;
    }

    Vec* rawMutable(const std::string& n) override
    {
        std::map<std::string, Vec*> t = { { "hice", &hice }, { "cice", &cice }, { "damage", &damage },
            { "ssh", &ssh }, { "e11", &e11 }, { "e12", &e12 }, { "e22", &e22 }, { "s11", &s11 },
            { "s12", &s12 }, { "s22", &s22 }, { "cg_u", &u }, { "cg_v", &v }, { "cgA", &cgA }, { "cgH", &cgH },
            { "uGradSSH", &uGradSSH }, { "vGradSSH", &vGradSSH }, { "dStressX", &dStressX },
            { "dStressY", &dStressY }, { "uOcean", &uOcean }, { "vOcean", &vOcean }, { "uAtmos", &uAtmos },
            { "vAtmos", &vAtmos }, { "u0", &u0 }, { "v0", &v0 }, { "avgU", &avgU }, { "avgV", &avgV },
            { "lumpedcgmass", &pmap->lumpedcgmass }, { "lumpedcg1mass", &pmap->lumpedcg1mass },
            { "divS1", &pmap->divS1 }, { "divS2", &pmap->divS2 }, { "divM", &pmap->divM },
            { "iMgradX", &pmap->iMgradX }, { "iMgradY", &pmap->iMgradY }, { "iMM", &pmap->iMM },
            { "iMJwPSI", &pmap->iMJwPSI }, { "iMJwPSI_dam", &pmap->iMJwPSI_dam }, { "dX_SSH", &pmap->dX_SSH },
            { "dY_SSH", &pmap->dY_SSH }, { "velx", &dgtransport->velx }, { "vely", &dgtransport->vely },
            { "normalvel_X", &dgtransport->normalvel_X }, { "normalvel_Y", &dgtransport->normalvel_Y },
            { "AdvX", &dgtransport->AdvX }, { "AdvY", &dgtransport->AdvY }, { "iMass", &dgtransport->iMass } };
        auto it = t.find(n);
        return it == t.end() ? nullptr : it->second;
    }
    const Vec* raw(const std::string& n) override { return rawMutable(n); }
};

} // namespace nso
