/*
 * nsdg_cuda.cu -- libnsdg_cuda.so: handle, host orchestration and the C ABI of include/nsdg.h.
 *
 * The per-timestep sequence is that of the reference kernels
 *   VPCGDynamicsKernel::update       dynamics/src/include/VPCGDynamicsKernel.hpp:63-94
 *   BrittleCGDynamicsKernel::update  dynamics/src/include/BrittleCGDynamicsKernel.hpp:91-136
 *   DynamicsKernel::advectionAndLimits dynamics/src/include/DynamicsKernel.hpp:160-172
 *   CGDynamicsKernel::prepareIteration dynamics/src/CGDynamicsKernel.cpp:260-276
 * with all state resident on the device between calls.  Mesh-derived integer state (land mask,
 * sorted Dirichlet lists) is built on the host exactly as ParametricMesh does
 * (dynamics/src/ParametricMesh.cpp:217-292) and is bit-exact by construction.
 *
 * There is no CPU fallback: without a usable CUDA device every entry point fails.
 */
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <utility>

#include "nsdg_fast_launch.cuh" // argument structs + launchers of the fast strip kernels (compiled in their own translation units)
#include "nsdg_halo.cuh"
#include "nsdg_prepare.cuh"

namespace nsdg {

/*
 * Tensor map of a plane field for the TMA staging of the strip kernels (nsdg_momentum_uniform.cuh): the 2-d tensor
 * {Npad doubles (contiguous), ncomp planes (pitch Npad doubles)}, tile = 32 elements x all planes.  cuTensorMapEncodeTiled is a
 * driver entry point; it is looked up through the runtime so that the library keeps linking against cudart only.
 */
inline CUtensorMap planeTensorMap(const double* base, size_t Npad, int ncomp)
{
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q {};
        NSDG_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn || q != cudaDriverEntryPointSuccess)
            throw std::runtime_error("cuTensorMapEncodeTiled is not available in this driver");
        return reinterpret_cast<Encode>(fn);
    }();
    CUtensorMap m {};
    const cuuint64_t dims[2] = { cuuint64_t(Npad), cuuint64_t(ncomp) };
    const cuuint64_t strides[1] = { cuuint64_t(Npad) * sizeof(double) };
    const cuuint32_t box[2] = { 32u, cuuint32_t(ncomp) };
    const cuuint32_t estr[2] = { 1u, 1u };
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw std::runtime_error("cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
    return m;
}

static thread_local std::string g_lastError;

static inline unsigned blocksFor(size_t n, unsigned bs = 128) { return unsigned((n + bs - 1) / bs); }

//! Which handle's operator set currently sits in the per-device __constant__ symbol c_mops.  ONE table for every
//! Handle<CG, DGA> instantiation: the (6,2) and the (3,1) build upload into the same symbol.
static const void*& constOpsOwner(int device)
{
    static const void* owner[64] = {};
    return owner[device >= 0 && device < 64 ? device : 0];
}
static inline size_t alignUp(size_t n, size_t a) { return (n + a - 1) / a * a; }

//! type-erased handle
class HandleBase {
public:
    virtual ~HandleBase() = default;
    virtual void setMesh(int nx, int ny, const double* coords, const double* mask, int spherical) = 0;
    virtual void setField(int field, const double* host, int ncomp) = 0;
    virtual void getField(int field, double* host, int ncomp) = 0;
    virtual void step(double dt) = 0;
    virtual void update(const nsdg_update_io* io, double dt) = 0;
    virtual void subcycles(int n, float* ms) = 0;
    virtual void timeKernels(int n, float* stripMs, float* linesMs) = 0;
    virtual void getInternal(const std::string& name, double* host, size_t cap, size_t* count) = 0;
    virtual void setInternal(const std::string& name, const double* host, size_t count) = 0;
    virtual void benchmarkForcing(double elapsed, double Lx, double Ly) = 0;
    virtual void dims(int* nx, int* ny) = 0;
    virtual void healDamage(double dt, double td, const double* deltaCi) = 0;
    virtual void setBoundaries(const long* const* dir, const size_t* ndir, const long* per, const size_t* segSizes, size_t nseg) = 0;
    virtual void advectField(int field, double dt, int order, int nsteps, int limitMode, double maxv, double minv) = 0;
    virtual void haloExport(unsigned char* handle) = 0;
    virtual void haloConnect(int side, const unsigned char* handle) = 0;
    virtual void haloReady() = 0;
    nsdg_config cfg {};
    nsdg_timing timing {};
    std::vector<uint8_t> landmask;
    std::vector<long> dirichlet[4];
    bool meshSet = false;
};

template <int CG, int DGA> class Handle : public HandleBase {
public:
    static constexpr int DGs = cg2dgstress(CG), GS = gp1d(DGs), Q = GS * GS, NR = CG + 1, ND = NR * NR;
    static constexpr int EDA = edgedofs(DGA), EDS = edgedofs(DGs);
    static constexpr int QA = gp1d(DGA) * gp1d(DGA), QS = Q;

    PhysParams p;
    GridDims g {};
    bool uniform = false;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr; //!< update(): forcing uploads / early downloads beside the compute stream
    cudaEvent_t evForcing = nullptr, evCopyDone = nullptr;
    bool forcingPending = false;
    cudaEvent_t ev[5] {};
    int cg1s = 0; //!< CG1 row stride
    size_t ncg = 0, ncg1 = 0; //!< allocated CG / CG1 doubles

    // mesh
    std::vector<double> hvx, hvy;
    std::vector<uint8_t> hdirmask;
    DevBuf<double> vx, vy;
    DevBuf<uint8_t> d_landmask, d_dirmask, d_nodemask;
    // element fields
    DevBuf<double> hice, cice, damage, ssh, s11, s12, s22, gaussA, gaussB, helem;
    DevBuf<double> velx, vely, tmp1, tmp2, tmp3, nvX, nvY; // DGA transport (tmp3: third field of a pass, or rk3)
    DevBuf<double> velxS, velyS, tmp1S, tmp2S, tmp3S, nvXS, nvYS; // DGs transport (BBM)
    DevBuf<int> perNbr, perEdge; //!< periodic edges (nsdg_set_boundaries): [4][Npad] neighbour element / edge index, else empty
    std::vector<std::array<long, 4>> periodic; //!< ParametricMesh::periodic, flattened: {type, c1, c2, edge}
    std::vector<std::pair<size_t, size_t>> periodicSegs; //!< (first entry, count) of every periodic segment
    DevBuf<long> d_periodic; //!< the flattened list on the device (CGAveragePeriodic)
    DevBuf<long> d_seamNodes; //!< the CG nodes of the periodic seams, each once (seam_update_kernel)
    DevBuf<double> seamX, seamY; //!< their stress divergence between the lines kernel and the seam update
    DevBuf<double> scratchDG; // DGA planes scratch (set/get via DG2CG / CG2DG)
    // operators
    DevBuf<double> tAdvX, tAdvY, tiMass, sAdvX, sAdvY, siMass;
    DevBuf<double> oGx, oGy, oGM, oB, oBd, oD1, oD2, oDM, odX, odY;
    TransportOpPtrs topA {}, topS {};
    MomentumOpPtrs mop {};
    // CG fields
    DevBuf<double> u, v, u0, v0, cgH, cgA, gradX, gradY, uO, vO, uA, vA, lmass, avgU, avgV, taux, tauy;
    DevBuf<double> cgSSH, mass1, gu1, gv1;
    DevBuf<double> hbuf, vbuf;
    CUtensorMap mevpTm[5] {}; //!< TMA staging of the fast mEVP kernels: s11, s12, s22, P / alpha, geometry
    CUtensorMap bbmTm[8] {}; //!< TMA staging of the fast BBM kernels: s11, s12, s22, damage, h, expC, Pmax, geometry
    DevBuf<double> ncCA, ncRx, ncRy, ncIlm; // per-node constants of the fast paths (plus uO, vO)
    DevBuf<double> geo; // per-element geometry planes of the parametric fast path
    DevBuf<double> vcon; // compact node constants of the vertical deferred lines (vcon_kernel)
    DevBuf<double> vavg; // fast BBM paths: mean velocities of the vertical deferred lines during the subcycle loop (vavg_kernel)
    bool fastUniformMEVP = false, fastUniformBBM = false;
    bool fastUniformMEVP1 = false; //!< DG1 / CG1 build on a uniform mesh: subcycle_strip_umevp1 + the generic lines kernel
    bool fastParamMEVP = false; //!< factored-operator kernel on non-uniform Cartesian meshes (nsdg_momentum_param.cuh)
    bool fastParamBBM = false;
    bool fastMEVP() const { return fastUniformMEVP || fastParamMEVP; }
    bool fastBBM() const { return fastUniformBBM || fastParamBBM; }
    DevBuf<double> gaussC; //!< uniform BBM: Pmax in the Gauss points
    // halo exchange (partitioned domain)
    DevBuf<unsigned char> arena; //!< my receive arena: [side][parity] payload slots + flags
    HaloArenaLayout arenaLayout {};
    unsigned char* peerArena[kHaloSides] = { nullptr, nullptr, nullptr, nullptr }; //!< IPC-mapped neighbour arenas
    DevBuf<int> haloError;
    DevBuf<HaloDevState> haloState; //!< exchange epochs and block counters, advanced on the device
    bool haloActive = false;
    // staging
    DevBuf<double> staging;
    std::vector<std::pair<const void*, size_t>> registered;
    // strips
    int R = 16, nsx = 0, nsy = 0;
    // graph cache
    cudaGraphExec_t graphExec = nullptr;
    int graphN = 0;
    double graphDeltaT = 0;
    long launches = 0;

    explicit Handle(const nsdg_config& c)
    {
        cfg = c;
        if (cfg.device >= 0)
            NSDG_CUDA_CHECK(cudaSetDevice(cfg.device));
        int dev = 0;
        NSDG_CUDA_CHECK(cudaGetDevice(&dev)); // fails loudly when there is no GPU: no CPU fallback
        cfg.device = dev;
        NSDG_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (auto& e : ev)
            NSDG_CUDA_CHECK(cudaEventCreate(&e));
        p.alpha = cfg.alpha;
        p.beta = cfg.beta;
    }
    ~Handle() override
    {
        if (constOpsOwner(cfg.device) == this)
            constOpsOwner(cfg.device) = nullptr; // a later handle may be allocated at the same address
        if (graphExec)
            cudaGraphExecDestroy(graphExec);
        closePeers();
        for (auto& r : registered)
            cudaHostUnregister(const_cast<void*>(r.first));
        for (auto& e : ev)
            if (e)
                cudaEventDestroy(e);
        for (cudaEvent_t e : { evForcing, evCopyDone })
            if (e)
                cudaEventDestroy(e);
        if (copyStream)
            cudaStreamDestroy(copyStream);
        if (stream)
            cudaStreamDestroy(stream);
    }

    // ------------------------------------------------------------------------------------
    // mesh
    // ------------------------------------------------------------------------------------
    void buildDirichlet()
    {
        const size_t nx = g.nx, ny = g.ny, N = g.N;
        for (auto& d : dirichlet)
            d.clear();
        // dirichletFromMask, ParametricMesh.cpp:228-252
        const size_t startX[4] = { 0, 0, 0, 1 }, stopX[4] = { nx, nx - 1, nx, nx };
        const size_t startY[4] = { 1, 0, 0, 0 }, stopY[4] = { ny, ny, ny - 1, ny };
        const long delta[4] = { -long(nx), 1, long(nx), -1 };
        for (int edge = 0; edge < 4; ++edge) {
            for (size_t j = startY[edge]; j < stopY[edge]; ++j)
                for (size_t i = startX[edge]; i < stopX[edge]; ++i) {
                    const size_t idx = i + nx * j;
                    if (landmask[idx] && !landmask[idx + delta[edge]])
                        dirichlet[edge].push_back(long(idx));
                }
            std::sort(dirichlet[edge].begin(), dirichlet[edge].end());
        }
        // dirichletFromEdge for BOTTOM, RIGHT, TOP, LEFT, ParametricMesh.cpp:259-272.  With a partition,
        // only sides on the edge of the GLOBAL domain are closed.
        const size_t start[4] = { 0, nx - 1, N - nx, 0 }, stop[4] = { nx, N, N, N }, stride[4] = { 1, nx, 1, nx };
        for (int edge = 0; edge < 4; ++edge) {
            if (cfg.global_nx > 0 && cfg.neighbour[edge] >= 0)
                continue;
            for (size_t idx = start[edge]; idx < stop[edge]; idx += stride[edge])
                if (landmask[idx])
                    dirichlet[edge].push_back(long(idx));
            std::sort(dirichlet[edge].begin(), dirichlet[edge].end());
        }
        hdirmask.assign(N, 0);
        for (int edge = 0; edge < 4; ++edge)
            for (long e : dirichlet[edge])
                hdirmask[e] |= uint8_t(1 << edge);
    }

    //! all elements are congruent axis-aligned rectangles (Cartesian): one operator set suffices
    bool detectUniform() const
    {
        if (g.spherical || cfg.force_general)
            return false;
        const int nx = g.nx, ny = g.ny;
        const double dx = hvx[1] - hvx[0], dy = hvy[nx + 1] - hvy[0];
        if (!(dx > 0) || !(dy > 0))
            return false;
        const double tolx = 1e-9 * dx, toly = 1e-9 * dy;
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) {
                const size_t n = size_t(j) * (nx + 1) + i;
                if (std::fabs(hvx[n] - (hvx[0] + i * dx)) > tolx * (1 + i) || std::fabs(hvy[n] - (hvy[0] + j * dy)) > toly * (1 + j))
                    return false;
            }
        return true;
    }

    void setMesh(int nx, int ny, const double* coords, const double* mask, int spherical) override
    {
        if (nx < 2 || ny < 2)
            throw std::runtime_error("nsdg_set_mesh: nx and ny must be >= 2");
        meshSet = false;
        if (graphExec) {
            cudaGraphExecDestroy(graphExec);
            graphExec = nullptr;
        }
        // a new mesh frees the halo arena the neighbours have mapped and restarts the exchange epochs: drop my own peer
        // mappings; every box of the partition has to go through export / connect / ready again
        haloActive = false;
        closePeers();
        g.nx = nx;
        g.ny = ny;
        g.N = nx * ny;
        g.nxs = int(alignUp(size_t(nx), 32));
        // Plane pitch: a multiple of 32 elements (256 B), skewed by 8 KiB + 256 B so that the 33+ planes a
        // warp streams never sit a power-of-two apart (2048^2 * 8 B = 32 MiB exactly would alias every plane
        // onto the same L2 sets and DRAM banks).
        g.Npad = g.nxs * ny + (g.N >= 4096 ? 1056 : 0);
        if (const char* env = std::getenv("NSDG_PLANE_SKEW")) // tuning knob (elements, multiple of 32)
            g.Npad = g.nxs * ny + std::atoi(env);
        g.CG = CG;
        g.cgnx = CG * nx + 1;
        g.cgny = CG * ny + 1;
        g.cgs = int(alignUp(size_t(g.cgnx), 16));
        g.spherical = spherical ? 1 : 0;
        g.bnd = 0;
        for (int s = 0; s < 4; ++s)
            if (!(cfg.global_nx > 0 && cfg.neighbour[s] >= 0))
                g.bnd |= 1 << s;
        cg1s = int(alignUp(size_t(nx + 1), 16));
        ncg = size_t(g.cgs) * g.cgny;
        ncg1 = size_t(cg1s) * (ny + 1);
        const size_t N = g.N, Npad = g.Npad, nnodes = size_t(nx + 1) * (ny + 1);

        // ---- host mesh: coordinates, pole rotation, land mask, Dirichlet lists ----
        hvx.resize(nnodes);
        hvy.resize(nnodes);
        for (size_t i = 0; i < nnodes; ++i) { // coordinatesFromModelArray, ParametricMesh.cpp:198-210
            hvx[i] = coords[2 * i];
            hvy[i] = coords[2 * i + 1];
        }
        if (spherical) // RotatePoleToGreenland, ParametricMesh.hpp:137-157
            for (size_t i = 0; i < nnodes; ++i) {
                const double x = cos(hvy[i]) * cos(hvx[i]), y = cos(hvy[i]) * sin(hvx[i]), z = sin(hvy[i]);
                const double aw = 40.0 * M_PI / 180.0, bw = 15.0 * M_PI / 180.0;
                const double x1 = cos(aw) * x - sin(aw) * y, y1 = sin(aw) * x + cos(aw) * y, z1 = z;
                const double x2 = cos(bw) * x1 - sin(bw) * z1, y2 = y1, z2 = sin(bw) * x1 + cos(bw) * z1;
                hvy[i] = asin(z2);
                hvx[i] = atan2(y2, x2);
            }
        landmask.resize(N);
        for (size_t i = 0; i < N; ++i) // landmaskFromModelArray, ParametricMesh.cpp:217-223 (quirk Q10)
            landmask[i] = (mask[i] == 1.) ? 1 : 0;
        buildDirichlet();
        periodic.clear();
        periodicSegs.clear();
        perNbr.release();
        perEdge.release();
        d_periodic.release();
        d_seamNodes.release();
        seamX.release();
        seamY.release();
        uniform = detectUniform();

        vx.alloc(nnodes);
        vy.alloc(nnodes);
        NSDG_CUDA_CHECK(cudaMemcpy(vx, hvx.data(), nnodes * 8, cudaMemcpyHostToDevice));
        NSDG_CUDA_CHECK(cudaMemcpy(vy, hvy.data(), nnodes * 8, cudaMemcpyHostToDevice));
        d_landmask.alloc(Npad);
        d_dirmask.alloc(Npad);
        NSDG_CUDA_CHECK(cudaMemcpy2D(d_landmask, g.nxs, landmask.data(), nx, nx, ny, cudaMemcpyHostToDevice));
        NSDG_CUDA_CHECK(cudaMemcpy2D(d_dirmask, g.nxs, hdirmask.data(), nx, nx, ny, cudaMemcpyHostToDevice));
        legacySync(); // a pageable H2D copy may return before the DMA has landed; `stream` does not wait for the legacy stream
        d_nodemask.alloc(ncg);
        nodemask_kernel<CG><<<blocksFor(N), 128, 0, stream>>>(g, d_dirmask, d_nodemask);

        // ---- state (zero-initialised; the reference relies on fresh pages being zero, quirk Q9) ----
        const bool bbm = cfg.rheology == NSDG_BBM;
        for (auto* f : { &hice, &cice, &velx, &vely, &tmp1, &tmp2, &scratchDG })
            f->alloc(size_t(DGA) * Npad);
        damage.alloc(bbm ? size_t(DGA) * Npad : 0);
        ssh.alloc(Npad);
        for (auto* f : { &s11, &s12, &s22 })
            f->alloc(size_t(DGs) * Npad);
        gaussA.alloc(size_t(Q) * Npad);
        gaussB.alloc(bbm ? size_t(Q) * Npad : 0);
        helem.alloc(bbm ? Npad : 0);
        const size_t pX = alignUp(size_t(nx) * (ny + 1), 32), pY = alignUp(size_t(nx + 1) * ny, 32);
        nvX.alloc(EDA * pX);
        nvY.alloc(EDA * pY);
        if (bbm) {
            for (auto* f : { &velxS, &velyS, &tmp1S, &tmp2S, &tmp3S })
                f->alloc(size_t(DGs) * Npad);
            tmp3.alloc(size_t(DGA) * Npad);
            nvXS.alloc(EDS * pX);
            nvYS.alloc(EDS * pY);
        }
        for (auto* f : { &u, &v, &u0, &v0, &cgH, &cgA, &gradX, &gradY, &uO, &vO, &uA, &vA, &lmass, &taux, &tauy })
            f->alloc(ncg);
        avgU.alloc(bbm ? ncg : 0);
        avgV.alloc(bbm ? ncg : 0);
        for (auto* f : { &cgSSH, &mass1, &gu1, &gv1 })
            f->alloc(ncg1);
        staging.alloc(size_t(std::max(DGA, DGs)) * Npad, false);

        // ---- operators ----
        const size_t opN = uniform ? 1 : Npad;
        const int nel = uniform ? 1 : g.N;
        tAdvX.alloc(size_t(DGA) * QA * opN);
        tAdvY.alloc(size_t(DGA) * QA * opN);
        tiMass.alloc(size_t(DGA) * DGA * opN);
        topA = { tAdvX, tAdvY, tiMass, opN, uniform ? 0 : 1 };
        setup_transport_kernel<DGA><<<blocksFor(nel, 64), 64, 0, stream>>>(g, vx, vy, topA, nel);
        if (bbm) {
            sAdvX.alloc(size_t(DGs) * QS * opN);
            sAdvY.alloc(size_t(DGs) * QS * opN);
            siMass.alloc(size_t(DGs) * DGs * opN);
            topS = { sAdvX, sAdvY, siMass, opN, uniform ? 0 : 1 };
            setup_transport_kernel<DGs><<<blocksFor(nel, 64), 64, 0, stream>>>(g, vx, vy, topS, nel);
            helem_kernel<<<blocksFor(N), 128, 0, stream>>>(g, vx, vy, helem);
        }
        // the streamed per-element matrices (360 - 558 doubles per element) are built only for the kernel that streams
        // them; the factored-operator kernels need the 4 x 4 SSH matrices alone (ensureStreamedOps builds the rest on demand)
        const bool factored = !uniform && CG == 2 && DGA == 6 && !cfg.force_general && !std::getenv("NSDG_NO_FAST_PARAM")
            && (cfg.rheology == NSDG_MEVP || cfg.rheology == NSDG_BBM);
        odX.alloc(16 * opN);
        odY.alloc(16 * opN);
        streamedOpsBuilt = false;
        mop = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, odX, odY, opN, uniform ? 0 : 1 };
        // free drift runs no subcycle kernel at all: the SSH matrices alone (the rest only on demand, nsdg_get_internal)
        if (factored || cfg.rheology == NSDG_FREEDRIFT)
            setup_momentum_kernel<CG, DGA><<<blocksFor(nel, 64), 64, 0, stream>>>(g, vx, vy, mop, nel, false);
        else
            ensureStreamedOps();
        constexpr int CGGP = (CG == 1 ? 1 : 4);
        lumpedmass_kernel<CG, CGGP><<<blocksFor(size_t(g.cgnx) * g.cgny), 128, 0, stream>>>(g, g.cgnx, g.cgny, g.cgs, vx, vy, lmass);
        lumpedmass_kernel<1, 2><<<blocksFor(size_t(nx + 1) * (ny + 1)), 128, 0, stream>>>(g, nx + 1, ny + 1, cg1s, vx, vy, mass1);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (uniform && streamedOpsBuilt) { // the single operator set goes to __constant__ memory for the generic subcycle kernel
            MomentumOps& h = hostMops;
            h = MomentumOps {};
            auto pull = [&](double* dst, const DevBuf<double>& src, size_t n) {
                NSDG_CUDA_CHECK(cudaMemcpy(dst, src.p, n * 8, cudaMemcpyDeviceToHost));
            };
            pull(h.Gx, oGx, DGs * ND);
            pull(h.Gy, oGy, DGs * ND);
            pull(h.B, oB, DGs * Q);
            pull(h.Bd, oBd, DGA * Q);
            pull(h.D1, oD1, ND * DGs);
            pull(h.D2, oD2, ND * DGs);
            constOpsOwner(cfg.device) = nullptr; // uploaded again by ensureConstOps before the next kernel that reads it
        }

        // ---- strips and deferred-line buffers ----
        // element rows per warp strip: 12-24 on large grids (wave quantisation, below); on small ones as few as keep
        // every strip resident at once (a 128 x 128 grid has only 4 strips per element row)
        nsx = (nx + 31) / 32;
        {
            // small grids: the fewest rows per strip with which all strips are resident at once (one wave).  Warps per SM: 12
            // for the uniform mEVP kernel (168 registers), 6 for the spherical BBM kernel (shared memory), 8 for the others.
            // Measured per subcycle (strip + lines): 256^2 BBM 20.6 -> 18.3 us (R = 1 -> 2), 512^2 mEVP 47.8 -> 40.2 us (4 -> 5),
            // 512^2 BBM 62.7 -> 60.9 us (4 -> 7); the TOPAZ-like 128^2 grid keeps R = 1
            int smsR = 148;
            cudaDeviceGetAttribute(&smsR, cudaDevAttrMultiProcessorCount, cfg.device);
            const bool mevpFastR = uniform && cfg.rheology == NSDG_MEVP && CG == 2 && DGA == 6 && !std::getenv("NSDG_NO_FAST_UNIFORM");
            const bool mevp1R = uniform && cfg.rheology == NSDG_MEVP && CG == 1 && !cfg.force_general && !std::getenv("NSDG_NO_FAST_UNIFORM");
            const long warpsPerSM = mevp1R ? 16 : mevpFastR ? 12 : (cfg.rheology == NSDG_BBM && spherical ? 6 : 8);
            const long slotsR = long(smsR) * warpsPerSM;
            R = int(std::max<long>(1, std::min<long>(16, (long(nsx) * ny + slotsR - 1) / slotsR)));
        }
        if (R == 16) {
            // large grid: pick R in [12, 24] against wave quantisation.  The strip kernels run 4 warps per block and
            // `bps` blocks per SM (3 for the uniform mEVP kernel, 2 for the 246/255-register ones); with B blocks the
            // last of ceil(B / slots) rounds is partly empty.  Horizontal deferred lines cost ~ 1.07 / R of a strip pass.
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device);
            const bool mevpFast = uniform && cfg.rheology == NSDG_MEVP && CG == 2 && DGA == 6 && !std::getenv("NSDG_NO_FAST_UNIFORM");
            const double slots = double(sms) * (mevpFast ? 3 : 2);
            double best = -1.0;
            for (int r = 12; r <= 24; ++r) {
                const long blocks = (long(nsx) * ((ny + r - 1) / r) + 3) / 4;
                const double waves = blocks / slots;
                const double score = waves / std::ceil(waves) / (1.0 + 1.07 / r);
                if (score > best + 1e-9) {
                    best = score;
                    R = r;
                }
            }
        }
        if (const char* env = std::getenv("NSDG_STRIP_ROWS")) // tuning knob: element rows per warp strip
            R = std::max(1, std::atoi(env));
        nsx = (nx + 31) / 32;
        nsy = (ny + R - 1) / R;
        hbuf.alloc(size_t(nsy) * 2 * nx * NR * 2);
        vbuf.alloc(size_t(nsx) * 2 * ny * NR * 2);
        fastUniformMEVP = uniform && cfg.rheology == NSDG_MEVP && CG == 2 && DGA == 6;
        fastUniformBBM = uniform && cfg.rheology == NSDG_BBM && CG == 2 && DGA == 6;
        fastUniformMEVP1 = uniform && cfg.rheology == NSDG_MEVP && CG == 1 && !cfg.force_general;
        if (std::getenv("NSDG_NO_FAST_UNIFORM")) // testing knob: generic strip kernel on the uniform operator set
            fastUniformMEVP = fastUniformBBM = fastUniformMEVP1 = false;
        fastParamBBM = !uniform && cfg.rheology == NSDG_BBM && CG == 2 && DGA == 6 && !cfg.force_general
            && !std::getenv("NSDG_NO_FAST_PARAM");
        if (fastBBM()) {
            for (auto* f : { &ncCA, &ncRx, &ncRy, &ncIlm })
                f->alloc(ncg);
            gaussC.alloc(size_t(Q) * Npad);
            if constexpr (CG == 2 && DGA == 6) {
                prepareKernelsUBBM();
                prepareKernelsPBBM();
            }
            const double* fields[7] = { s11, s12, s22, damage, gaussA, gaussB, gaussC };
            const int comps[7] = { DGs, DGs, DGs, DGA, Q, Q, Q };
            for (int i = 0; i < 7; ++i)
                bbmTm[i] = planeTensorMap(fields[i], Npad, comps[i]);
        }
        fastParamMEVP = !uniform && cfg.rheology == NSDG_MEVP && CG == 2 && DGA == 6 && !cfg.force_general
            && !std::getenv("NSDG_NO_FAST_PARAM");
        if (fastMEVP()) {
            for (auto* f : { &ncCA, &ncRx, &ncRy, &ncIlm })
                f->alloc(ncg);
            if constexpr (CG == 2 && DGA == 6) {
                prepareKernelsUMEVP();
                prepareKernelsPMEVP();
            }
            const double* fields[4] = { s11, s12, s22, gaussA };
            const int comps[4] = { DGs, DGs, DGs, Q };
            for (int i = 0; i < 4; ++i)
                mevpTm[i] = planeTensorMap(fields[i], Npad, comps[i]);
        }
        if (fastUniformMEVP1) {
            for (auto* f : { &ncCA, &ncRx, &ncRy, &ncIlm })
                f->alloc(ncg);
            prepareKernelsUMEVP1();
            const double* fields[4] = { s11, s12, s22, gaussA };
            const int comps[4] = { DGs, DGs, DGs, Q };
            for (int i = 0; i < 4; ++i)
                mevpTm[i] = planeTensorMap(fields[i], Npad, comps[i]);
        }
        if (fastMEVP() || fastBBM() || fastUniformMEVP1)
            vcon.alloc(size_t(kVconPlanes) * nsx * g.cgny);
        vavg.alloc(fastBBM() ? size_t(2) * nsx * g.cgny : 0);
        if (fastParamMEVP || fastParamBBM) {
            geo.alloc(size_t(fastParamBBM ? geoPlanesBBM(spherical) : geoPlanes(spherical)) * Npad);
            if (spherical)
                paramgeom_kernel<true><<<blocksFor(N), 128, 0, stream>>>(g, vx, vy, geo);
            else
                paramgeom_kernel<false><<<blocksFor(N), 128, 0, stream>>>(g, vx, vy, geo);
            if (fastParamBBM) {
                if (spherical)
                    paramgeom_bbm_kernel<true><<<blocksFor(N), 128, 0, stream>>>(g, p, helem, geo);
                else
                    paramgeom_bbm_kernel<false><<<blocksFor(N), 128, 0, stream>>>(g, p, helem, geo);
                bbmTm[7] = planeTensorMap(geo, Npad, geoPlanesBBM(spherical));
            } else {
                mevpTm[4] = planeTensorMap(geo, Npad, geoPlanes(spherical));
            }
            NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        timing.uniform_path = uniform ? 1 : 0;
        haloActive = false;
        if (cfg.global_nx > 0) {
            arenaLayout.slotDoubles = alignUp(size_t(6) * std::max(g.cgnx, g.cgny) + 64, 64);
            arena.alloc(arenaLayout.totalBytes());
            haloError.alloc(1);
            haloState.alloc(1); // zeroed: epochs restart with a new arena
        }
        meshSet = true;
    }

    // ------------------------------------------------------------------------------------
    // fields in / out
    // ------------------------------------------------------------------------------------
    void requireMesh() const
    {
        if (!meshSet)
            throw std::runtime_error("nsdg: mesh not set (call nsdg_set_mesh first)");
    }

    //! host AoS (N x ncomp) -> device planes (nplanes), DGModelArray::ma2dg semantics (dynamics/src/include/DGModelArray.hpp:20-32;
    //! quirk Q4: a one-component source fills component 0 and zeroes the others); downloadPlanes = dg2ma (:34-48)
    void uploadPlanes(const double* host, int ncomp, int nplanes, double* planes)
    {
        if (ncomp != 1 && ncomp != nplanes)
            throw std::runtime_error("nsdg_set_field: ncomp must be 1 or the DG component count");
        const size_t N = g.N;
        if (ncomp == 1) {
            NSDG_CUDA_CHECK(cudaMemcpy2DAsync(planes, size_t(g.nxs) * 8, host, size_t(g.nx) * 8, size_t(g.nx) * 8, g.ny,
                cudaMemcpyHostToDevice, stream));
            if (nplanes > 1)
                NSDG_CUDA_CHECK(cudaMemsetAsync(planes + g.Npad, 0, size_t(nplanes - 1) * g.Npad * 8, stream));
        } else {
            NSDG_CUDA_CHECK(cudaMemcpyAsync(staging, host, N * ncomp * 8, cudaMemcpyHostToDevice, stream));
            aos2planes_kernel<<<blocksFor(N), 128, 0, stream>>>(g, ncomp, nplanes, staging, planes);
        }
    }
    void downloadPlanes(const double* planes, int ncomp, double* host)
    {
        const size_t N = g.N;
        if (ncomp == 1) {
            NSDG_CUDA_CHECK(cudaMemcpy2DAsync(host, size_t(g.nx) * 8, planes, size_t(g.nxs) * 8, size_t(g.nx) * 8, g.ny,
                cudaMemcpyDeviceToHost, stream));
        } else {
            planes2aos_kernel<<<blocksFor(N), 128, 0, stream>>>(g, ncomp, planes, staging);
            NSDG_CUDA_CHECK(cudaMemcpyAsync(host, staging, N * ncomp * 8, cudaMemcpyDeviceToHost, stream));
        }
    }
    //! ModelArray -> temp DG -> DG2CG (CGDynamicsKernel.cpp:56-86)
    void uploadViaDG2CG(const double* host, int ncomp, double* cgDest)
    {
        uploadPlanes(host, ncomp, ncomp == 1 ? 1 : DGA, scratchDG);
        const unsigned nb = blocksFor(size_t(g.cgnx) * g.cgny);
        if (ncomp == 1)
            dg2cg_kernel<CG, 1><<<nb, 128, 0, stream>>>(g, scratchDG, cgDest, -INFINITY, INFINITY);
        else
            dg2cg_kernel<CG, DGA><<<nb, 128, 0, stream>>>(g, scratchDG, cgDest, -INFINITY, INFINITY);
    }
    void setFieldAsync(int field, const double* host, int ncomp)
    {
        requireMesh();
        switch (field) {
        case NSDG_HICE:
            uploadPlanes(host, ncomp, DGA, hice);
            break;
        case NSDG_CICE:
            uploadPlanes(host, ncomp, DGA, cice);
            break;
        case NSDG_DAMAGE:
            if (cfg.rheology != NSDG_BBM)
                return; // the mEVP kernel just buckets unknown fields (DynamicsKernel.hpp:104-109): no effect on the path
            uploadPlanes(host, ncomp, DGA, damage);
            break;
        case NSDG_U:
            uploadViaDG2CG(host, ncomp, u);
            break;
        case NSDG_V:
            uploadViaDG2CG(host, ncomp, v);
            break;
        case NSDG_UWIND:
            uploadViaDG2CG(host, ncomp, uA);
            break;
        case NSDG_VWIND:
            uploadViaDG2CG(host, ncomp, vA);
            break;
        case NSDG_UOCEAN:
            uploadViaDG2CG(host, ncomp, uO);
            break;
        case NSDG_VOCEAN:
            uploadViaDG2CG(host, ncomp, vO);
            break;
        case NSDG_SSH:
            if (ncomp != 1)
                throw std::runtime_error("nsdg_set_field: ssh is a DG0 field (ncomp must be 1)");
            uploadPlanes(host, 1, 1, ssh);
            break;
        default:
            throw std::runtime_error("nsdg_set_field: field cannot be set");
        }
    }
    //! opt-in extension (nsdg_config::keep_dg_moments): the caller's array replaces the cell MEANS of an advected field; the
    //! higher moments stay and are limited again for the new mean (same limiters as DynamicsKernel::advectionAndLimits)
    void setMeanKeepMoments(int field, const double* host)
    {
        requireMesh();
        double* f = field == NSDG_HICE ? hice.p : (field == NSDG_CICE ? cice.p : damage.p);
        NSDG_CUDA_CHECK(cudaMemcpy2DAsync(f, size_t(g.nxs) * 8, host, size_t(g.nx) * 8, size_t(g.nx) * 8, g.ny, cudaMemcpyHostToDevice, stream));
        // hice, cice: limited again for the new mean (a no-op while the caller hands back the means it received: they left the
        // advection limited and the subcycles do not touch them).  The damage is NOT: the BBM subcycles rewrite all its
        // moments without limiting, and the reference limits it only after the next advection (BrittleCGDynamicsKernel.hpp:
        // 103-106) -- an extra limiter here would change a field the caller did not touch.
        if (field == NSDG_HICE)
            limit_kernel<DGA><<<blocksFor(g.N), 128, 0, stream>>>(g, f, 2, 0.0, 0.0);
        else if (field == NSDG_CICE)
            limit_kernel<DGA><<<blocksFor(g.N), 128, 0, stream>>>(g, f, 3, 1.0, 0.0);
    }
    void setField(int field, const double* host, int ncomp) override
    {
        setFieldAsync(field, host, ncomp);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    //! BenchmarkAtmosphere::update + BenchmarkOcean::setData evaluated on the device, then the DG0 -> CG path of setData
    void benchmarkForcing(double elapsed, double Lx, double Ly) override
    {
        requireMesh();
        const int gnx = cfg.global_nx > 0 ? cfg.global_nx : g.nx, gny = cfg.global_nx > 0 ? cfg.global_ny : g.ny;
        const int gi0 = cfg.global_nx > 0 ? cfg.box_x0 - (cfg.neighbour[NSDG_LEFT] >= 0 ? 1 : 0) : 0;
        const int gj0 = cfg.global_nx > 0 ? cfg.box_y0 - (cfg.neighbour[NSDG_BOTTOM] >= 0 ? 1 : 0) : 0;
        const double dx = Lx / gnx, dy = Ly / gny;
        const double timeFraction = elapsed / 86400.0, cycloneDuration = 5.;
        const double x0c = gnx * dx * 0.5 * (1 + timeFraction / cycloneDuration);
        const double y0c = gny * dy * 0.5 * (1 + timeFraction / cycloneDuration);
        const double alpha = 72. / 180. * M_PI;
        double* pl = scratchDG; // 4 of its DGA >= 3 planes... use tmp1/tmp2 as well to be safe for DGA = 3
        double* uw = pl;
        double* vw = pl + g.Npad;
        double* uo = tmp1;
        double* vo = tmp1.p + g.Npad;
        benchforcing_kernel<<<blocksFor(g.N), 128, 0, stream>>>(g, gi0, gj0, dx, dy, x0c, y0c, cos(alpha), sin(alpha), gnx * dx,
            gny * dy, uw, vw, uo, vo);
        const unsigned nb = blocksFor(size_t(g.cgnx) * g.cgny);
        dg2cg_kernel<CG, 1><<<nb, 128, 0, stream>>>(g, uw, uA, -INFINITY, INFINITY);
        dg2cg_kernel<CG, 1><<<nb, 128, 0, stream>>>(g, vw, vA, -INFINITY, INFINITY);
        dg2cg_kernel<CG, 1><<<nb, 128, 0, stream>>>(g, uo, uO, -INFINITY, INFINITY);
        dg2cg_kernel<CG, 1><<<nb, 128, 0, stream>>>(g, vo, vO, -INFINITY, INFINITY);
        NSDG_CUDA_CHECK(cudaMemsetAsync(ssh, 0, size_t(g.Npad) * 8, stream));
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    void cgToDG0(const double* cgSrc, double* host)
    {
        cg2dg_kernel<CG, DGA><<<blocksFor(g.N), 128, 0, stream>>>(g, vx, vy, cgSrc, topA, scratchDG);
        downloadPlanes(scratchDG, 1, host);
    }
    void getFieldAsync(int field, double* host, int ncomp)
    {
        requireMesh();
        const bool bbm = cfg.rheology == NSDG_BBM;
        if (ncomp != 1 && ncomp != DGA)
            throw std::runtime_error("nsdg_get_field: ncomp must be 1 or the DG component count");
        switch (field) {
        case NSDG_HICE:
            downloadPlanes(hice, ncomp, host);
            break;
        case NSDG_CICE:
            downloadPlanes(cice, ncomp, host);
            break;
        case NSDG_DAMAGE:
            if (!bbm)
                throw std::runtime_error("nsdg_get_field: damage exists for BBM only");
            downloadPlanes(damage, ncomp, host);
            break;
        case NSDG_U:
        case NSDG_V:
            if (ncomp != 1)
                throw std::runtime_error("nsdg_get_field: u, v are exported as DG0 cell means");
            cgToDG0(field == NSDG_U ? u : v, host);
            break;
        case NSDG_TAUX:
        case NSDG_TAUY: {
            if (ncomp != 1)
                throw std::runtime_error("nsdg_get_field: ice-ocean stress is exported as DG0 cell means");
            const unsigned nb = blocksFor(size_t(g.cgnx) * g.cgny);
            if (cfg.rheology == NSDG_FREEDRIFT)
                iostress_kernel<NSDG_BBM><<<nb, 128, 0, stream>>>(g, p, u, v, uO, vO, taux, tauy);
            else if (bbm)
                iostress_kernel<NSDG_BBM><<<nb, 128, 0, stream>>>(g, p, avgU, avgV, uO, vO, taux, tauy);
            else
                iostress_kernel<NSDG_MEVP><<<nb, 128, 0, stream>>>(g, p, u, v, uO, vO, taux, tauy);
            cgToDG0(field == NSDG_TAUX ? taux : tauy, host);
            break;
        }
        default:
            throw std::runtime_error("nsdg_get_field: field cannot be read");
        }
    }
    void getField(int field, double* host, int ncomp) override
    {
        getFieldAsync(field, host, ncomp);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    // ------------------------------------------------------------------------------------
    // halo exchange between partition boxes (nsdg_halo.cuh)
    // ------------------------------------------------------------------------------------
    bool hasNeighbour(int side) const { return cfg.global_nx > 0 && cfg.neighbour[side] >= 0; }

    void haloExport(unsigned char* handle) override
    {
        requireMesh();
        if (cfg.global_nx <= 0)
            throw std::runtime_error("nsdg_halo_export: the handle was created without a partition");
        cudaIpcMemHandle_t hnd;
        NSDG_CUDA_CHECK(cudaIpcGetMemHandle(&hnd, arena.p));
        static_assert(sizeof(hnd) == NSDG_IPC_HANDLE_BYTES, "IPC handle size");
        std::memcpy(handle, &hnd, sizeof(hnd));
    }
    void haloConnect(int side, const unsigned char* handle) override
    {
        requireMesh();
        if (side < 0 || side >= kHaloSides || !hasNeighbour(side))
            throw std::runtime_error("nsdg_halo_connect: no neighbour on that side");
        cudaIpcMemHandle_t hnd;
        std::memcpy(&hnd, handle, sizeof(hnd));
        if (peerArena[side]) { // reconnecting (the neighbour re-meshed): drop the mapping of its old arena
            haloActive = false;
            cudaIpcCloseMemHandle(peerArena[side]);
            peerArena[side] = nullptr;
        }
        void* ptr = nullptr;
        NSDG_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        peerArena[side] = static_cast<unsigned char*>(ptr);
    }
    void haloReady() override
    {
        for (int s = 0; s < kHaloSides; ++s)
            if (hasNeighbour(s) && !peerArena[s])
                throw std::runtime_error("nsdg_halo_ready: a neighbour side is not connected");
        // a (re)connected partition starts from epoch zero on every box: the launcher lets no box exchange before all have
        // passed this point (partition.connect_halos), so resetting my own flags and counters here cannot race a neighbour
        if (cfg.global_nx > 0) {
            NSDG_CUDA_CHECK(cudaMemsetAsync(haloState.p, 0, sizeof(HaloDevState), stream));
            NSDG_CUDA_CHECK(cudaMemsetAsync(arena.p + arenaLayout.flagsOffsetBytes(), 0, 256, stream));
            NSDG_CUDA_CHECK(cudaMemsetAsync(haloError.p, 0, sizeof(int), stream));
            NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        haloActive = cfg.global_nx > 0;
    }
    void closePeers()
    {
        for (auto& pa : peerArena)
            if (pa) {
                cudaIpcCloseMemHandle(pa);
                pa = nullptr;
            }
    }

    //! lines of a node field (CG grid) that go to / come from the neighbour across `side`.
    //! The right/top box owns the shared boundary line: a box receives CG lines from its left/bottom
    //! neighbour and CG+1 from its right/top one, and sends the complementary sets.
    HaloLineDesc nodeLines(int side, bool send, int nFields, int deg) const
    {
        // deg = CG for the velocity space, 1 for the CG1 sea-surface-height grid
        const int nxn = deg * g.nx + 1, nyn = deg * g.ny + 1, stride = (deg == CG) ? g.cgs : cg1s;
        HaloLineDesc d {};
        d.nFields = nFields;
        const bool vertical = (side == NSDG_LEFT || side == NSDG_RIGHT); // lines are node columns
        const int last = vertical ? nxn - 1 : nyn - 1;
        int first, count;
        if (side == NSDG_LEFT || side == NSDG_BOTTOM) {
            first = send ? deg : 0;
            count = send ? deg + 1 : deg;
        } else {
            first = send ? last - 2 * deg : last - deg;
            count = send ? deg : deg + 1;
        }
        d.nLines = count;
        d.lineLen = vertical ? nyn : nxn;
        d.stride = vertical ? stride : 1;
        for (int k = 0; k < count; ++k)
            d.firstLine[k] = vertical ? long(first + k) : long(first + k) * stride;
        return d;
    }
    //! the ring column/row of an element (DG plane) field: send the first/last OWNED line, receive the ring line
    HaloLineDesc elemLines(int side, bool send, int nFields) const
    {
        HaloLineDesc d {};
        d.nFields = nFields;
        const bool vertical = (side == NSDG_LEFT || side == NSDG_RIGHT);
        const int last = vertical ? g.nx - 1 : g.ny - 1;
        const int line = (side == NSDG_LEFT || side == NSDG_BOTTOM) ? (send ? 1 : 0) : (send ? last - 1 : last);
        d.nLines = 1;
        d.lineLen = vertical ? g.ny : g.nx;
        d.stride = vertical ? g.nxs : 1;
        d.firstLine[0] = vertical ? long(line) : long(line) * g.nxs;
        return d;
    }

    //! one full exchange (x phase, then y phase) of node fields (pitch == 0) or of the planes of a DG field
    void exchange(double* const* fields, int nFields, size_t pitch, bool nodes, int deg = CG)
    {
        if (!haloActive)
            return;
        const int phases[2] = { (1 << NSDG_LEFT) | (1 << NSDG_RIGHT), (1 << NSDG_BOTTOM) | (1 << NSDG_TOP) };
        for (int ph = 0; ph < 2; ++ph) {
            bool any = false;
            for (int s = 0; s < kHaloSides; ++s)
                any = any || ((phases[ph] & (1 << s)) && hasNeighbour(s));
            if (!any)
                continue;
            HaloExchangeArgs xa {};
            for (int f = 0; f < nFields && f < 8; ++f)
                xa.fields[f] = pitch ? nullptr : fields[f];
            xa.fields[0] = fields[0];
            xa.fieldPitch = pitch;
            xa.sideMask = phases[ph];
            xa.errorFlag = haloError;
            xa.state = haloState;
            xa.myArena = reinterpret_cast<double*>(arena.p);
            xa.layout = arenaLayout;
            for (int s = 0; s < kHaloSides; ++s) {
                if (!(phases[ph] & (1 << s)) || !hasNeighbour(s))
                    continue;
                xa.send[s] = nodes ? nodeLines(s, true, nFields, deg) : elemLines(s, true, nFields);
                xa.recv[s] = nodes ? nodeLines(s, false, nFields, deg) : elemLines(s, false, nFields);
                xa.peerArena[s] = reinterpret_cast<double*>(peerArena[s]);
            }
            // the exchange epoch is device state advanced by the kernel: the arguments are the same for every exchange of
            // these fields, so the launch can be captured in the subcycle graph
            halo_exchange_kernel<<<dim3(2 * kHaloBlocksPerSide, kHaloSides), 256, 0, stream>>>(xa);
            launches += 1;
        }
    }
    void exchangeNodes(double* a, double* b)
    {
        double* f[2] = { a, b };
        exchange(f, 2, 0, true);
    }
    //! a single field on the CG1 node grid (the sea-surface height before its gradient is taken)
    void exchangeNodesCG1(double* a)
    {
        double* f[1] = { a };
        exchange(f, 1, 0, true, 1);
    }
    void exchangePlanes(double* planes, int ncomp)
    {
        double* f[1] = { planes };
        exchange(f, ncomp, g.Npad, false);
    }
    //! Called after the final stream synchronisation of every entry point that exchanges halos: a box whose neighbour
    //! never delivered (the wall-clock timeout of halo_exchange_kernel) has NOT unpacked, so its ring holds stale data and the
    //! call must fail instead of returning success.  The flag is cleared once reported.
    void checkHaloError()
    {
        if (!haloActive)
            return;
        int e = 0;
        NSDG_CUDA_CHECK(cudaMemcpyAsync(&e, haloError.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (e) {
            NSDG_CUDA_CHECK(cudaMemsetAsync(haloError.p, 0, sizeof(int), stream));
            NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
            throw std::runtime_error("nsdg: halo exchange timed out waiting for a neighbour box; this box's ring data are stale");
        }
    }

    // ------------------------------------------------------------------------------------
    // advection
    // ------------------------------------------------------------------------------------
    template <int DG>
    void prepareAdvection(const double* cgU, const double* cgV, TransportOpPtrs op, double* vxd, double* vyd, double* nX, double* nY)
    {
        if (uniform && !std::getenv("NSDG_NO_UNIFORM_TRANSPORT"))
            cg2dg_pair_kernel<CG, DG, true><<<blocksFor(g.N), 128, 0, stream>>>(g, vx, vy, cgU, cgV, op, vxd, vyd);
        else
            cg2dg_pair_kernel<CG, DG, false><<<blocksFor(g.N), 128, 0, stream>>>(g, vx, vy, cgU, cgV, op, vxd, vyd);
        launches += 1;
        normalVelocity<DG>(vxd, vyd, nX, nY);
    }
    template <int DG> void launchStage(TransportStageArgs& a, int nf)
    {
        a.nf = nf;
        const dim3 grid(unsigned((g.nx + 127) / 128) * nf, g.ny);
        if (uniform && !std::getenv("NSDG_NO_UNIFORM_TRANSPORT")) {
            a.dxU = hvx[1] - hvx[0];
            a.dyU = hvy[g.nx + 1] - hvy[0];
            a.iAreaU = 1.0 / (a.dxU * a.dyU);
            transport_stage_kernel<DG, true><<<grid, 128, 0, stream>>>(a);
        } else
            transport_stage_kernel<DG, false><<<grid, 128, 0, stream>>>(a);
        launches += 1;
    }
    struct LimitSpec {
        int mode = 0; //!< bit 1: LimitMax(maxv), bit 2: LimitMin(minv)
        double maxv = 0, minv = 0;
    };
    /*
     * DGTransport::step (DGTransport.cpp:514-566) for up to three fields that share the transport object (operators,
     * velocities, edge normal velocities): one fused kernel per Runge-Kutta stage (nsdg_transport.cuh), the limiters of
     * DynamicsKernel::advectionAndLimits (DynamicsKernel.hpp:160-172) in the last one.  order = 1, 2, 3 (rk1, rk2, rk3);
     * the module path uses rk2 (DynamicsKernel.hpp:61).  tmp[s][f]: stage buffers (order - 1 ... at least `order` of them
     * for rk3, one for rk1 / rk2).  Partitioned boxes exchange the ring elements of every stage value.
     */
    template <int DG>
    void transportFields(double dt, TransportOpPtrs op, const double* vxd, const double* vyd, const double* nX, const double* nY,
        int order, int nf, double* const* phi, double* const (*tmp)[kTransportMaxFields], const LimitSpec* lim)
    {
        TransportStageArgs a {};
        a.g = g;
        a.dt = dt;
        a.landmask = d_landmask;
        a.dirmask = d_dirmask;
        a.velx = vxd;
        a.vely = vyd;
        a.nvX = nX;
        a.nvY = nY;
        a.pitchX = alignUp(size_t(g.nx) * (g.ny + 1), 32);
        a.pitchY = alignUp(size_t(g.nx + 1) * g.ny, 32);
        a.op = op;
        // parametric fast paths: the cell-term operators come from the geometry planes (DG6 and DG8 share the 3 x 3 Gauss points)
        a.geo = (fastParamMEVP || fastParamBBM) && !std::getenv("NSDG_NO_FACTORED_TRANSPORT") ? geo.p : nullptr;
        a.perNbr = perNbr.p;
        a.perEdge = perEdge.p;
        auto stage = [&](double* const* in, double* const* base, double* const* out, int epi, double c0, double c1, bool last) {
            for (int f = 0; f < nf; ++f) {
                a.in[f] = in[f];
                a.base[f] = base[f];
                a.out[f] = out[f];
                a.limitMode[f] = last ? lim[f].mode : 0;
                a.maxv[f] = lim[f].maxv;
                a.minv[f] = lim[f].minv;
            }
            a.epi = epi;
            a.c0 = c0;
            a.c1 = c1;
            launchStage<DG>(a, nf);
            for (int f = 0; f < nf; ++f)
                exchangePlanes(out[f], DG); // ring elements of the stage value come from their owners
        };
        if (order == 1) { // step_rk1: phi += k(phi); the stage cannot run in place (neighbours read phi)
            stage(phi, phi, tmp[0], 0, 0, 0, true);
            for (int f = 0; f < nf; ++f)
                NSDG_CUDA_CHECK(cudaMemcpyAsync(phi[f], tmp[0][f], size_t(DG) * g.Npad * 8, cudaMemcpyDeviceToDevice, stream));
        } else if (order == 2) { // step_rk2 (Heun): phi1 = phi + k1; phi = phi1 + (k2 - k1) / 2
            stage(phi, phi, tmp[0], 0, 0, 0, false);
            stage(tmp[0], phi, phi, 1, 0, 0, true);
        } else { // step_rk3: tmp1 = phi + k(phi); tmp2 = (tmp1 + k(tmp1)) / 4 + 3 phi / 4; phi = phi / 3 + 2 (tmp2 + k(tmp2)) / 3
            stage(phi, phi, tmp[0], 0, 0, 0, false);
            stage(tmp[0], phi, tmp[1], 2, 0.75, 0.25, false);
            stage(tmp[1], phi, phi, 2, 1.0 / 3.0, 2.0 / 3.0, true);
        }
    }
    //! DGTransport::reinitnormalvelocity (DGTransport.cpp:158-252) from the DG velocity currently in (vxd, vyd)
    template <int DG> void normalVelocity(const double* vxd, const double* vyd, double* nX, double* nY)
    {
        const size_t pX = alignUp(size_t(g.nx) * (g.ny + 1), 32), pY = alignUp(size_t(g.nx + 1) * g.ny, 32);
        const size_t nEdges = size_t(g.nx) * (g.ny + 1) + size_t(g.nx + 1) * g.ny;
        normalvel_kernel<DG><<<blocksFor(nEdges), 128, 0, stream>>>(g, vx, vy, d_landmask, d_dirmask, vxd, vyd, nX, pX, nY, pY);
        launches += 1;
    }

    /*
     * Replace the boundary lists of the mesh: ParametricMesh::dirichlet[4] and ParametricMesh::periodic
     * (ParametricMesh.hpp:76-79), which the .smesh 2.0 reader fills from a file (ParametricMesh.cpp:79-178) and the
     * reference's advection tests assign by hand (Advection_test.cpp:222-240, AdvectionPeriodicBC_test.cpp:228-249).
     * Periodic entries are {type (0: X-edge, bottom/top; 1: Y-edge, left/right), c1 = element left of / below the edge,
     * c2 = element right of / above it, edge = index of the edge whose normal velocity the flux uses}.
     */
    void setBoundaries(const long* const* dir, const size_t* ndir, const long* per, const size_t* segSizes, size_t nseg) override
    {
        requireMesh();
        if (cfg.global_nx > 0)
            throw std::runtime_error("nsdg_set_boundaries: not available on a partition box (its artificial edges are halo lines)");
        const long N = g.N, nx = g.nx, ny = g.ny;
        for (int edge = 0; edge < 4; ++edge) {
            if (!dir || !dir[edge])
                continue; // keep the list derived from the mask
            for (size_t i = 0; i < ndir[edge]; ++i) {
                const long el = dir[edge][i];
                if (el < 0 || el >= N)
                    throw std::runtime_error("nsdg_set_boundaries: Dirichlet element index out of range");
            }
            dirichlet[edge].assign(dir[edge], dir[edge] + ndir[edge]);
        }
        hdirmask.assign(size_t(N), 0);
        for (int edge = 0; edge < 4; ++edge)
            for (long el : dirichlet[edge])
                hdirmask[size_t(el)] |= uint8_t(1 << edge);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        NSDG_CUDA_CHECK(cudaMemcpy2D(d_dirmask, g.nxs, hdirmask.data(), nx, nx, ny, cudaMemcpyHostToDevice));
        legacySync();
        NSDG_CUDA_CHECK(cudaMemsetAsync(d_nodemask, 0, ncg, stream));
        nodemask_kernel<CG><<<blocksFor(size_t(N)), 128, 0, stream>>>(g, d_dirmask, d_nodemask);
        periodic.clear();
        periodicSegs.clear();
        perNbr.release();
        perEdge.release();
        d_periodic.release();
        d_seamNodes.release();
        seamX.release();
        seamY.release();
        size_t total = 0;
        for (size_t sgm = 0; sgm < nseg; ++sgm) {
            periodicSegs.emplace_back(total, segSizes[sgm]);
            total += segSizes[sgm];
        }
        if (total > 0) {
            if (!per)
                throw std::runtime_error("nsdg_set_boundaries: periodic list missing");
            std::vector<int> hn(size_t(4) * g.Npad, -1), he(size_t(4) * g.Npad, 0);
            auto plane = [&](long el) { return size_t(el / nx) * g.nxs + size_t(el % nx); };
            for (size_t i = 0; i < total; ++i) {
                const long type = per[4 * i], c1 = per[4 * i + 1], c2 = per[4 * i + 2], edge = per[4 * i + 3];
                const long nEdges = type == 0 ? nx * (ny + 1) : (nx + 1) * ny;
                if ((type != 0 && type != 1) || c1 < 0 || c1 >= N || c2 < 0 || c2 >= N || edge < 0 || edge >= nEdges)
                    throw std::runtime_error("nsdg_set_boundaries: bad periodic entry (DGTransport.cpp:475-478 aborts likewise)");
                // the flux is taken where the regular neighbour is missing: c1's top / right side and c2's bottom / left side
                // must lie on the edge of the domain
                const bool ok = type == 0 ? (c1 / nx == ny - 1 && c2 / nx == 0) : (c1 % nx == nx - 1 && c2 % nx == 0);
                if (!ok)
                    throw std::runtime_error("nsdg_set_boundaries: periodic edges must connect opposite edges of the domain");
                const int s1 = type == 0 ? NSDG_TOP : NSDG_RIGHT, s2 = type == 0 ? NSDG_BOTTOM : NSDG_LEFT;
                hn[size_t(s1) * g.Npad + plane(c1)] = int(plane(c2));
                he[size_t(s1) * g.Npad + plane(c1)] = int(edge);
                hn[size_t(s2) * g.Npad + plane(c2)] = int(plane(c1));
                he[size_t(s2) * g.Npad + plane(c2)] = int(edge);
                periodic.push_back(std::array<long, 4> { type, c1, c2, edge });
            }
            perNbr.alloc(hn.size());
            perEdge.alloc(he.size());
            d_periodic.alloc(4 * total);
            NSDG_CUDA_CHECK(cudaMemcpy(perNbr, hn.data(), hn.size() * sizeof(int), cudaMemcpyHostToDevice));
            NSDG_CUDA_CHECK(cudaMemcpy(perEdge, he.data(), he.size() * sizeof(int), cudaMemcpyHostToDevice));
            NSDG_CUDA_CHECK(cudaMemcpy(d_periodic, per, 4 * total * sizeof(long), cudaMemcpyHostToDevice));
            legacySync();
            // the momentum solve with periodic seams (CGDynamicsKernel.cpp:395-397) runs on the generic kernels: the nodes of
            // the seams, each once, flagged in the node mask; their stress divergence is averaged between lines kernel and update
            if (cfg.rheology != NSDG_FREEDRIFT) {
                if (haloActive)
                    throw std::runtime_error("nsdg_set_boundaries: periodic edges on a partitioned box are not supported");
                std::vector<long> nodes;
                for (const auto& e : periodic)
                    for (int j = 0; j <= CG; ++j) {
                        const long lb = e[2], rt = e[1];
                        const long n0lb = long(CG * (lb / nx)) * g.cgs + CG * (lb % nx), n0rt = long(CG * (rt / nx)) * g.cgs + CG * (rt % nx);
                        if (e[0] == 0) {
                            nodes.push_back(n0lb + j);
                            nodes.push_back(n0rt + long(CG) * g.cgs + j);
                        } else {
                            nodes.push_back(n0lb + long(j) * g.cgs);
                            nodes.push_back(n0rt + CG + long(j) * g.cgs);
                        }
                    }
                std::sort(nodes.begin(), nodes.end());
                nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
                d_seamNodes.alloc(nodes.size());
                NSDG_CUDA_CHECK(cudaMemcpy(d_seamNodes, nodes.data(), nodes.size() * sizeof(long), cudaMemcpyHostToDevice));
                legacySync();
                seamX.alloc(ncg);
                seamY.alloc(ncg);
                seam_flag_kernel<0><<<blocksFor(nodes.size()), 128, 0, stream>>>(d_seamNodes, long(nodes.size()), d_nodemask);
                fastUniformMEVP = fastUniformBBM = fastParamMEVP = fastParamBBM = fastUniformMEVP1 = false; // until the next nsdg_set_mesh
                ensureStreamedOps();
                constOpsOwner(cfg.device) = nullptr;
                if (graphExec) {
                    cudaGraphExecDestroy(graphExec);
                    graphExec = nullptr;
                }
            }
        }
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    //! VectorManipulations::CGAveragePeriodic (VectorManipulations.hpp:26-65) on a CG field: segments one after the other
    //! as in the reference; within a segment the node pairs j = 0 .. CG-1 of all entries at once, then the pairs j = CG
    //! (the pair an entry shares with the next one along the seam has then been averaged already: idempotent)
    void averagePeriodic(double* v)
    {
        for (const auto& sgm : periodicSegs) {
            if (sgm.second == 0)
                continue;
            const long* list = d_periodic.p + 4 * sgm.first;
            cg_average_periodic_kernel<CG><<<blocksFor(sgm.second * CG), 128, 0, stream>>>(g, list, long(sgm.second), 0, CG, v);
            cg_average_periodic_kernel<CG><<<blocksFor(sgm.second), 128, 0, stream>>>(g, list, long(sgm.second), CG, 1, v);
            launches += 2;
        }
    }
    /*
     * DGTransport::reinitnormalvelocity + DGTransport::step (+ LimitMax / LimitMin) on ONE named DG field, nsteps times,
     * with the DG velocity currently in the transport object (nsdg_set_internal "velx" / "vely" = DGTransport::GetVx /
     * GetVy, or the last prepareAdvection): the time loop of the reference's advection tests
     * (Advection_test.cpp:150-154, AdvectionPeriodicBC_test.cpp:186-192).
     */
    void advectField(int field, double dt, int order, int nsteps, int limitMode, double maxv, double minv) override
    {
        requireMesh();
        if (order < 1 || order > 3)
            throw std::runtime_error("nsdg_advect_field: the time stepping scheme must be 1 (rk1), 2 (rk2) or 3 (rk3) (DGTransport.cpp:553-565 aborts otherwise)");
        if (nsteps < 0)
            throw std::runtime_error("nsdg_advect_field: nsteps must not be negative");
        double* f = field == NSDG_HICE ? hice.p : field == NSDG_CICE ? cice.p : (field == NSDG_DAMAGE ? damage.p : nullptr);
        if (!f)
            throw std::runtime_error("nsdg_advect_field: the field must be one of the advected DG fields of this handle (hice, cice, damage)");
        launches = 0;
        double* phi[kTransportMaxFields] = { f, nullptr, nullptr };
        double* const tmp[2][kTransportMaxFields] = { { tmp1, nullptr, nullptr }, { tmp2, nullptr, nullptr } };
        const LimitSpec lim[kTransportMaxFields] = { { limitMode, maxv, minv }, {}, {} };
        for (int i = 0; i < nsteps; ++i) {
            normalVelocity<DGA>(velx, vely, nvX, nvY);
            transportFields<DGA>(dt, topA, velx, vely, nvX, nvY, order, 1, phi, tmp, lim);
        }
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        checkHaloError();
        timing.kernel_launches = launches;
    }

    //! DynamicsKernel::advectionAndLimits (DynamicsKernel.hpp:160-172): cice and hice (BBM: and the damage,
    //! BrittleCGDynamicsKernel.hpp:103-106) with the DGadvection transport object, in ONE pass per Runge-Kutta stage --
    //! the fields are independent of each other and share operators and velocities -- with their limiters:
    //! cice in [0, 1], hice >= 0, damage in [1e-12, 1]
    void advectTracers(double dt, bool withDamage)
    {
        double* phi[kTransportMaxFields] = { cice, hice, damage };
        double* const tmp[1][kTransportMaxFields] = { { tmp1, tmp2, tmp3 } };
        const LimitSpec lim[kTransportMaxFields] = { { 3, 1.0, 0.0 }, { 2, 0.0, 0.0 }, { 3, 1.0, 1e-12 } };
        transportFields<DGA>(dt, topA, velx, vely, nvX, nvY, 2, withDamage ? 3 : 2, phi, tmp, lim);
    }

    // ------------------------------------------------------------------------------------
    // prepareIteration (CGDynamicsKernel.cpp:260-276)
    // ------------------------------------------------------------------------------------
    void prepareIteration()
    {
        const unsigned nb = blocksFor(size_t(g.cgnx) * g.cgny);
        dg2cg_kernel<CG, DGA><<<nb, 128, 0, stream>>>(g, hice, cgH, 1.e-4, INFINITY);
        dg2cg_kernel<CG, DGA><<<nb, 128, 0, stream>>>(g, cice, cgA, 1.e-4, 1.0);
        if (!periodicSegs.empty()) {
            // CGDynamicsKernel.cpp:264-266 averages across the seam BEFORE the clamps of :272-275, which dg2cg_kernel fuses:
            // the clamps are monotone and the seam average of two clamped values is again inside the clamp range, but it is
            // not the clamp of the average -- redo both in the reference's order
            dg2cg_kernel<CG, DGA><<<nb, 128, 0, stream>>>(g, hice, cgH, -INFINITY, INFINITY);
            dg2cg_kernel<CG, DGA><<<nb, 128, 0, stream>>>(g, cice, cgA, -INFINITY, INFINITY);
            averagePeriodic(cgH);
            averagePeriodic(cgA);
            clamp_kernel<<<blocksFor(ncg, 256), 256, 0, stream>>>(ncg, cgH, 1.e-4, INFINITY);
            clamp_kernel<<<blocksFor(ncg, 256), 256, 0, stream>>>(ncg, cgA, 1.e-4, 1.0);
            launches += 4;
        }
        // ComputeGradientOfSeaSurfaceHeight
        GridDims g1 = g;
        g1.CG = 1;
        g1.cgnx = g.nx + 1;
        g1.cgny = g.ny + 1;
        g1.cgs = cg1s;
        const unsigned nb1 = blocksFor(size_t(g1.cgnx) * g1.cgny);
        dg2cg_kernel<1, 1><<<nb1, 128, 0, stream>>>(g1, ssh, cgSSH, -INFINITY, INFINITY);
        exchangeNodesCG1(cgSSH); // the gradient at the first owned CG1 line needs the complete outer line
        sshgrad_cg1_kernel<<<nb1, 128, 0, stream>>>(g, cg1s, cgSSH, mop, mass1, gu1, gv1);
        sshgrad_cg_kernel<CG><<<nb, 128, 0, stream>>>(g, cg1s, gu1, gv1, gradX, gradY);
        launches += 5;
    }

    // ------------------------------------------------------------------------------------
    // subcycles
    // ------------------------------------------------------------------------------------
    SubcycleArgs makeArgs(double deltaT) const
    {
        SubcycleArgs a {};
        a.g = g;
        a.R = R;
        a.nsx = nsx;
        a.nsy = nsy;
        a.s11 = s11;
        a.s12 = s12;
        a.s22 = s22;
        a.damage = damage;
        a.gaussA = gaussA;
        a.gaussB = gaussB;
        a.helem = helem;
        a.landmask = d_landmask;
        a.Gx = oGx;
        a.Gy = oGy;
        a.GM = oGM;
        a.B = oB;
        a.Bd = oBd;
        a.D1 = oD1;
        a.D2 = oD2;
        a.DM = oDM;
        a.u = u;
        a.v = v;
        a.avgU = avgU;
        a.avgV = avgV;
        a.u0 = u0;
        a.v0 = v0;
        a.cgH = cgH;
        a.cgA = cgA;
        a.uAtm = uA;
        a.vAtm = vA;
        a.uOcn = uO;
        a.vOcn = vO;
        a.gradX = gradX;
        a.gradY = gradY;
        a.lmass = lmass;
        a.nodemask = d_nodemask;
        a.seamX = seamX;
        a.seamY = seamY;
        a.hbuf = hbuf;
        a.vbuf = vbuf;
        a.deltaT = deltaT;
        a.nSteps = double(cfg.nsteps);
        a.p = p;
        return a;
    }

    UniformArgs makeUniformArgs(double deltaT) const
    {
        UniformArgs a {};
        a.g = g;
        a.R = R;
        a.nsx = nsx;
        a.nsy = nsy;
        a.s11 = s11;
        a.s12 = s12;
        a.s22 = s22;
        a.Pa = gaussA;
        a.landmask = d_landmask;
        a.u = u;
        a.v = v;
        a.cA = ncCA;
        a.rx = ncRx;
        a.ry = ncRy;
        a.uO = uO;
        a.vO = vO;
        a.ilm = ncIlm;
        a.geo = geo;
        a.vcon = vcon;
        a.nodemask = d_nodemask;
        a.hbuf = hbuf;
        a.vbuf = vbuf;
        a.dx = hvx[1] - hvx[0];
        a.dy = hvy[g.nx + 1] - hvy[0];
        a.keep = 1.0 - 1.0 / p.alpha;
        a.beta = p.beta;
        a.dtfc = deltaT * p.fc;
        a.DeltaMin2 = p.DeltaMin * p.DeltaMin;
        for (int i = 0; i < 5; ++i)
            a.tm[i] = mevpTm[i];
        return a;
    }
    UniformBBMArgs makeUniformBBMArgs(double deltaT) const
    {
        UniformBBMArgs a {};
        a.g = g;
        a.R = R;
        a.nsx = nsx;
        a.nsy = nsy;
        a.s11 = s11;
        a.s12 = s12;
        a.s22 = s22;
        a.damage = damage;
        a.gH = gaussA;
        a.gE = gaussB;
        a.gP = gaussC;
        a.landmask = d_landmask;
        a.u = u;
        a.v = v;
        a.avgU = avgU;
        a.avgV = avgV;
        a.cA = ncCA;
        a.ax = ncRx;
        a.ay = ncRy;
        a.uO = uO;
        a.vO = vO;
        a.ilm = ncIlm;
        a.nodemask = d_nodemask;
        a.hbuf = hbuf;
        a.vbuf = vbuf;
        a.dx = hvx[1] - hvx[0];
        a.dy = hvy[g.nx + 1] - hvy[0];
        a.deltaT = deltaT;
        a.dtfc = deltaT * p.fc;
        a.invNSteps = 1.0 / double(cfg.nsteps);
        a.cosA = p.cosOceanAngle;
        a.sinA = p.sinOceanAngle;
        a.young = p.young;
        a.nu0 = p.nu0;
        a.lambda0 = p.undamaged_time_relaxation_sigma;
        a.tan_phi = p.tan_phi;
        const double hel = std::sqrt(a.dx * a.dy); // smesh.h(i) = sqrt(area), ParametricMesh.hpp:299
        const double scale = std::sqrt(0.1 / hel);
        a.cohScale = p.C_lab * scale;
        a.comprScale = p.compr_strength * scale;
        a.invTdK = 1.0 / (hel * std::sqrt(2. * (1. + p.nu0) * p.rho_ice));
        a.dunitK = deltaT / (1. - p.nu0 * p.nu0);
        a.geo = geo;
        a.vcon = vcon;
        a.vavg = vavg;
        a.C_lab = p.C_lab;
        a.compr_strength = p.compr_strength;
        for (int i = 0; i < 8; ++i)
            a.tm[i] = bbmTm[i];
        return a;
    }
    void launchPairFastBBM(const UniformBBMArgs& ba, unsigned nbStrip, size_t nLine, bool stripOnly = false, bool linesOnly = false)
    {
        if constexpr (CG == 2 && DGA == 6) {
            const unsigned nStrips = unsigned(nsx) * nsy;
            if (!linesOnly && fastParamBBM)
                launchStripPBBM(ba, g.spherical != 0, nStrips, stream);
            else if (!linesOnly)
                launchStripUBBM(ba, nStrips, stream);
            if (!stripOnly)
                launchLinesUBBM(ba, nLine, stream);
        }
    }
    void launchStripFast(const UniformArgs& ua, unsigned nbStrip)
    {
        if constexpr (CG == 2 && DGA == 6) {
            const unsigned nStrips = unsigned(nsx) * nsy;
            if (fastParamMEVP)
                launchStripPMEVP(ua, g.spherical != 0, nStrips, stream);
            else
                launchStripUMEVP(ua, nStrips, stream);
        }
    }
    void launchLinesFast(const UniformArgs& ua, size_t nLine) { launchLinesUMEVP(ua, nLine, stream); }
    template <int RHEO> void launchStrip(const SubcycleArgs& a, unsigned nbStrip)
    {
        if (uniform)
            subcycle_strip<CG, DGA, RHEO, true, false><<<nbStrip, 128, 0, stream>>>(a);
        else if (g.spherical)
            subcycle_strip<CG, DGA, RHEO, false, true><<<nbStrip, 128, 0, stream>>>(a);
        else
            subcycle_strip<CG, DGA, RHEO, false, false><<<nbStrip, 128, 0, stream>>>(a);
    }
    template <int RHEO> void launchSubcycle(const SubcycleArgs& a)
    {
        const unsigned nwarps = unsigned(nsx) * nsy;
        const unsigned nbStrip = (nwarps + 3) / 4;
        const size_t nLine = size_t(nsy) * g.cgnx + size_t(nsx) * g.cgny;
        if (uniform)
            subcycle_strip<CG, DGA, RHEO, true, false><<<nbStrip, 128, 0, stream>>>(a);
        else if (g.spherical)
            subcycle_strip<CG, DGA, RHEO, false, true><<<nbStrip, 128, 0, stream>>>(a);
        else
            subcycle_strip<CG, DGA, RHEO, false, false><<<nbStrip, 128, 0, stream>>>(a);
        subcycle_lines<CG, RHEO><<<blocksFor(nLine), 128, 0, stream>>>(a);
        if (d_seamNodes.n > 0) { // CGAveragePeriodic(tx), (ty), then the seam nodes' momentum update
            averagePeriodic(seamX);
            averagePeriodic(seamY);
            seam_update_kernel<RHEO><<<blocksFor(d_seamNodes.n), 128, 0, stream>>>(a, d_seamNodes, long(d_seamNodes.n));
        }
    }

    /*
     * c_mops is ONE __constant__ symbol per device, but every handle with a uniform mesh has its own operator set
     * (other cell sizes, the other CG/DG build): the generic uniform kernel reads whatever the last upload left there.
     * Each handle keeps its set on the host and re-uploads it, stream-ordered, whenever another handle (or nobody) owns the
     * symbol (constOpsOwner: one table at namespace scope, shared by both template instantiations).  Handles are driven from one thread at a time (nsdg.h), so two streams never need different sets at once.
     */
    MomentumOps hostMops {};
    //! only the generic kernel on a uniform mesh reads c_mops (the fast kernels carry compile-time unit operators,
    //! free drift runs no subcycle kernel)
    bool readsConstOps() const { return uniform && !fastMEVP() && !fastBBM() && !fastUniformMEVP1 && cfg.rheology != NSDG_FREEDRIFT; }
    void ensureConstOps()
    {
        if (!readsConstOps() || constOpsOwner(cfg.device) == this)
            return;
        NSDG_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_mops, &hostMops, sizeof(hostMops), 0, cudaMemcpyHostToDevice, stream));
        constOpsOwner(cfg.device) = this;
    }

    void runSubcycles(int n, double deltaT)
    {
        const long launchesBefore = launches;
        ensureConstOps();
        const SubcycleArgs a = makeArgs(deltaT);
        const UniformArgs ua = makeUniformArgs(deltaT);
        const UniformBBMArgs ba = makeUniformBBMArgs(deltaT);
        const unsigned nbStripF = (unsigned(nsx) * nsy + 3) / 4;
        const size_t nLineF = size_t(nsy) * g.cgnx + size_t(nsx) * g.cgny;
        const dim3 vavgGrid(unsigned((g.cgny + 127) / 128), unsigned(nsx));
        auto body = [&]() {
            if (fastBBM())
                vavg_kernel<true><<<vavgGrid, 128, 0, stream>>>(g, nsx, R, avgU, avgV, vavg);
            for (int i = 0; i < n; ++i) {
                if (fastUniformMEVP1) {
                    launchStripUMEVP1(ua, unsigned(nsx) * nsy, stream);
                    launchLinesUMEVP1(ua, stream);
                } else if (fastMEVP()) {
                    launchStripFast(ua, nbStripF);
                    launchLinesFast(ua, nLineF);
                } else if (fastBBM()) {
                    launchPairFastBBM(ba, nbStripF, nLineF);
                } else if (cfg.rheology == NSDG_BBM)
                    launchSubcycle<NSDG_BBM>(a);
                else
                    launchSubcycle<NSDG_MEVP>(a);
                exchangeNodes(u, v); // no-op for a single domain
            }
            if (fastBBM())
                vavg_kernel<false><<<vavgGrid, 128, 0, stream>>>(g, nsx, R, avgU, avgV, vavg);
        };
        /*
         * Hiding the exchange behind strips that do not need it was built twice and is NOT in the library.  Round 1: frame
         * strips and their lines first, exchange beside the interior strips, a second lines pass -- 5 % slower (two extra
         * launches, a split lines pass, a frame kernel alone at poor occupancy).  Round 2: after the lines of subcycle k the
         * graph forks into {exchange(k), then the edge band of subcycle k+1 as a compact launch} and {interior strips of k+1},
         * one lines pass after the join -- measured A/B on the same boxes at 2048^2 per GPU: 68.2 against 67.4 ms per step on
         * 2 GPUs, 70.0 against 69.8 ms on 4.  With the one-kernel exchange inside the graph there is nothing left to hide: the
         * exchange costs a few microseconds per subcycle when the boxes run in step; what separates N GPUs from one is the
         * ring's 65th strip column and strip kernels that run ~3 % slower when every GPU of the box is busy.
         */
        // partitioned boxes too: the exchange epochs are device state, the launch arguments never change
        if (cfg.use_cuda_graph && n > 1) {
            if (!graphExec || graphN != n || graphDeltaT != deltaT) {
                if (graphExec) {
                    cudaGraphExecDestroy(graphExec);
                    graphExec = nullptr;
                }
                cudaGraph_t graph;
                NSDG_CUDA_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
                body();
                NSDG_CUDA_CHECK(cudaStreamEndCapture(stream, &graph));
                NSDG_CUDA_CHECK(cudaGraphInstantiate(&graphExec, graph, 0));
                cudaGraphDestroy(graph);
                graphN = n;
                graphDeltaT = deltaT;
            }
            NSDG_CUDA_CHECK(cudaGraphLaunch(graphExec, stream));
        } else {
            body();
        }
        NSDG_CUDA_CHECK(cudaGetLastError());
        // kernels per subcycle: strip + lines + one exchange kernel per active phase (a graph replay runs them without
        // passing through exchange(), which counts only at capture time)
        const int phasesActive = haloActive ? (hasNeighbour(NSDG_LEFT) || hasNeighbour(NSDG_RIGHT) ? 1 : 0) + (hasNeighbour(NSDG_BOTTOM) || hasNeighbour(NSDG_TOP) ? 1 : 0) : 0;
        launches = launchesBefore + long(n) * (2 + phasesActive) + (fastBBM() ? 2 : 0);
    }

    void subcycles(int n, float* ms) override
    {
        requireMesh();
        launches = 0;
        const double deltaT = graphDeltaT != 0 ? graphDeltaT : lastDeltaT;
        NSDG_CUDA_CHECK(cudaEventRecord(ev[0], stream));
        runSubcycles(n, deltaT);
        NSDG_CUDA_CHECK(cudaEventRecord(ev[1], stream));
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        checkHaloError();
        float t = 0;
        NSDG_CUDA_CHECK(cudaEventElapsedTime(&t, ev[0], ev[1]));
        if (ms)
            *ms = t;
        timing.subcycle_ms = t;
        timing.kernel_launches = launches;
    }

    double lastDeltaT = 1.0;

    //! per-kernel timing: events around every launch of the two subcycle kernels
    void timeKernels(int n, float* stripMs, float* linesMs) override
    {
        requireMesh();
        ensureConstOps();
        const SubcycleArgs a = makeArgs(lastDeltaT);
        const UniformArgs ua = makeUniformArgs(lastDeltaT);
        const unsigned nwarps = unsigned(nsx) * nsy, nbStrip = (nwarps + 3) / 4;
        const size_t nLine = size_t(nsy) * g.cgnx + size_t(nsx) * g.cgny;
        double ts = 0, tl = 0, th = 0;
        const dim3 vavgGrid(unsigned((g.cgny + 127) / 128), unsigned(nsx));
        if (fastBBM())
            vavg_kernel<true><<<vavgGrid, 128, 0, stream>>>(g, nsx, R, avgU, avgV, vavg);
        for (int i = 0; i < n; ++i) {
            NSDG_CUDA_CHECK(cudaEventRecord(ev[0], stream));
            if (fastUniformMEVP1)
                launchStripUMEVP1(ua, unsigned(nsx) * nsy, stream);
            else if (fastMEVP())
                launchStripFast(ua, nbStrip);
            else if (fastBBM())
                launchPairFastBBM(makeUniformBBMArgs(lastDeltaT), nbStrip, nLine, true, false);
            else if (cfg.rheology == NSDG_BBM)
                launchStrip<NSDG_BBM>(a, nbStrip);
            else
                launchStrip<NSDG_MEVP>(a, nbStrip);
            NSDG_CUDA_CHECK(cudaEventRecord(ev[1], stream));
            if (fastUniformMEVP1)
                launchLinesUMEVP1(ua, stream);
            else if (fastMEVP())
                launchLinesFast(ua, nLine);
            else if (fastBBM())
                launchPairFastBBM(makeUniformBBMArgs(lastDeltaT), nbStrip, nLine, false, true);
            else if (cfg.rheology == NSDG_BBM)
                subcycle_lines<CG, NSDG_BBM><<<blocksFor(nLine), 128, 0, stream>>>(a);
            else
                subcycle_lines<CG, NSDG_MEVP><<<blocksFor(nLine), 128, 0, stream>>>(a);
            NSDG_CUDA_CHECK(cudaEventRecord(ev[2], stream));
            exchangeNodes(u, v);
            NSDG_CUDA_CHECK(cudaEventRecord(ev[3], stream));
            NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
            float x = 0, y = 0, z = 0;
            NSDG_CUDA_CHECK(cudaEventElapsedTime(&x, ev[0], ev[1]));
            NSDG_CUDA_CHECK(cudaEventElapsedTime(&y, ev[1], ev[2]));
            NSDG_CUDA_CHECK(cudaEventElapsedTime(&z, ev[2], ev[3]));
            ts += x;
            tl += y;
            th += z;
        }
        if (fastBBM()) {
            vavg_kernel<false><<<vavgGrid, 128, 0, stream>>>(g, nsx, R, avgU, avgV, vavg);
            NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        checkHaloError();
        *stripMs = float(ts / n);
        *linesMs = float(tl / n);
        timing.halo_ms = float(th / n);
    }

    // ------------------------------------------------------------------------------------
    // one timestep
    // ------------------------------------------------------------------------------------
    void stepAsync(double dt)
    {
        requireMesh();
        launches = 0;
        const bool bbm = cfg.rheology == NSDG_BBM;
        const size_t cgBytes = ncg * 8;
        NSDG_CUDA_CHECK(cudaEventRecord(ev[0], stream));
        if (cfg.rheology == NSDG_FREEDRIFT) { // FreeDriftDynamicsKernel::update, FreeDriftDynamicsKernel.hpp:43-50
            freedrift_kernel<<<blocksFor(size_t(g.cgnx) * g.cgny), 128, 0, stream>>>(g, p, uO, vO, uA, vA, d_nodemask, u, v);
            launches += 1;
            NSDG_CUDA_CHECK(cudaEventRecord(ev[1], stream));
            exchangeNodes(u, v);
            prepareAdvection<DGA>(u, v, topA, velx, vely, nvX, nvY);
            advectTracers(dt, false);
            NSDG_CUDA_CHECK(cudaEventRecord(ev[2], stream));
            NSDG_CUDA_CHECK(cudaEventRecord(ev[3], stream));
            return;
        }
        // ---- advection + limiters (DynamicsKernel.hpp:160-172) ----
        exchangeNodes(u, v); // partitioned: non-owned node lines come from their owners (no-op otherwise)
        if (bbm)
            exchangeNodes(avgU, avgV);
        prepareAdvection<DGA>(bbm ? avgU : u, bbm ? avgV : v, topA, velx, vely, nvX, nvY);
        advectTracers(dt, bbm);
        if (bbm) { // BrittleCGDynamicsKernel.hpp:97-106: the three stresses with the DG8 transport object
            prepareAdvection<DGs>(avgU, avgV, topS, velxS, velyS, nvXS, nvYS);
            double* phiS[kTransportMaxFields] = { s11, s12, s22 };
            double* const tmpS[1][kTransportMaxFields] = { { tmp1S, tmp2S, tmp3S } };
            const LimitSpec none[kTransportMaxFields] = {};
            transportFields<DGs>(dt, topS, velxS, velyS, nvXS, nvYS, 2, 3, phiS, tmpS, none);
        }
        NSDG_CUDA_CHECK(cudaEventRecord(ev[1], stream));
        if (forcingPending) { // update(): the forcing fields were uploaded on the copy stream while the advection ran
            NSDG_CUDA_CHECK(cudaStreamWaitEvent(stream, evForcing, 0));
            forcingPending = false;
        }
        // ---- prepareIteration ----
        prepareIteration();
        double deltaT;
        if (!bbm) { // VPCGDynamicsKernel.hpp:70-74
            NSDG_CUDA_CHECK(cudaMemcpyAsync(u0, u, cgBytes, cudaMemcpyDeviceToDevice, stream));
            NSDG_CUDA_CHECK(cudaMemcpyAsync(v0, v, cgBytes, cudaMemcpyDeviceToDevice, stream));
            deltaT = dt;
            gaussconst_kernel<DGA, GS, NSDG_MEVP><<<blocksFor(g.N), 128, 0, stream>>>(
                g, p, hice, cice, gaussA, gaussB, (fastMEVP() || fastUniformMEVP1) ? 1.0 / p.alpha : 1.0);
            if (fastMEVP() || fastUniformMEVP1) {
                nodeconst_kernel<0><<<blocksFor(size_t(g.cgnx) * g.cgny), 128, 0, stream>>>(
                    g, p, deltaT, cgH, cgA, uA, vA, gradX, gradY, u0, v0, lmass, ncCA, ncRx, ncRy, ncIlm);
                launches += 1;
            }
            if (fastMEVP() || fastUniformMEVP1) {
                vcon_kernel<CG><<<blocksFor(size_t(nsx) * g.cgny), 128, 0, stream>>>(
                    g, nsx, ncCA, ncRx, ncRy, uO, vO, ncIlm, d_nodemask, vcon);
                launches += 1;
            }
        } else { // BrittleCGDynamicsKernel.hpp:110-114
            deltaT = dt / double(cfg.nsteps);
            NSDG_CUDA_CHECK(cudaMemsetAsync(avgU, 0, cgBytes, stream));
            NSDG_CUDA_CHECK(cudaMemsetAsync(avgV, 0, cgBytes, stream));
            if (fastBBM()) {
                gaussconst_bbm3_kernel<DGA, GS><<<blocksFor(g.N), 128, 0, stream>>>(g, p, hice, cice, gaussA, gaussB, gaussC);
                nodeconst_bbm_kernel<0><<<blocksFor(size_t(g.cgnx) * g.cgny), 128, 0, stream>>>(
                    g, p, deltaT, cgH, cgA, uA, vA, gradX, gradY, lmass, ncCA, ncRx, ncRy, ncIlm);
                vcon_kernel<CG><<<blocksFor(size_t(nsx) * g.cgny), 128, 0, stream>>>(
                    g, nsx, ncCA, ncRx, ncRy, uO, vO, ncIlm, d_nodemask, vcon);
                launches += 2;
            } else
                gaussconst_kernel<DGA, GS, NSDG_BBM><<<blocksFor(g.N), 128, 0, stream>>>(g, p, hice, cice, gaussA, gaussB, 1.0);
        }
        launches += 1;
        lastDeltaT = deltaT;
        NSDG_CUDA_CHECK(cudaEventRecord(ev[2], stream));
        // ---- the subcycle loop ----
        runSubcycles(cfg.nsteps, deltaT);
        NSDG_CUDA_CHECK(cudaEventRecord(ev[3], stream));
    }
    void finishTiming()
    {
        NSDG_CUDA_CHECK(cudaEventElapsedTime(&timing.advection_ms, ev[0], ev[1]));
        NSDG_CUDA_CHECK(cudaEventElapsedTime(&timing.prepare_ms, ev[1], ev[2]));
        NSDG_CUDA_CHECK(cudaEventElapsedTime(&timing.subcycle_ms, ev[2], ev[3]));
        NSDG_CUDA_CHECK(cudaEventElapsedTime(&timing.total_ms, ev[0], ev[3]));
        timing.kernel_launches = launches;
    }
    void step(double dt) override
    {
        stepAsync(dt);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        checkHaloError();
        finishTiming();
    }

    //! pin the model's buffers once so that the per-step copies run at PCIe speed and asynchronously
    void pin(const void* ptr, size_t bytes)
    {
        if (!ptr || !cfg.pin_host_buffers)
            return;
        for (auto& r : registered)
            if (r.first == ptr && r.second >= bytes)
                return;
        if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess)
            registered.emplace_back(ptr, bytes);
        else
            cudaGetLastError(); // already pinned by the caller or not pinnable: copies still work
    }
    //! a buffer the caller no longer passes (a ModelArray that was reallocated) must not stay page-locked until nsdg_destroy:
    //! registrations that are not among this call's pointers are dropped
    void unpinStale(const void* const* current, int n)
    {
        if (!cfg.pin_host_buffers)
            return;
        for (size_t i = 0; i < registered.size();) {
            bool live = false;
            for (int k = 0; k < n; ++k)
                live = live || current[k] == registered[i].first;
            if (live)
                ++i;
            else {
                cudaHostUnregister(const_cast<void*>(registered[i].first));
                cudaGetLastError(); // the memory may already have been freed by its owner
                registered.erase(registered.begin() + long(i));
            }
        }
    }

    //! MEVPDynamics::update / BBMDynamics::update in one call
    void update(const nsdg_update_io* io, double dt) override
    {
        requireMesh();
        const size_t bytes = size_t(g.N) * 8;
        const double* ins[] = { io->hice_in, io->cice_in, io->damage_in, io->uwind, io->vwind, io->uocean, io->vocean, io->ssh };
        const int inField[] = { NSDG_HICE, NSDG_CICE, NSDG_DAMAGE, NSDG_UWIND, NSDG_VWIND, NSDG_UOCEAN, NSDG_VOCEAN, NSDG_SSH };
        double* outs[] = { io->hice_out, io->cice_out, io->damage_out, io->u_out, io->v_out, io->taux_out, io->tauy_out };
        const int outField[] = { NSDG_HICE, NSDG_CICE, NSDG_DAMAGE, NSDG_U, NSDG_V, NSDG_TAUX, NSDG_TAUY };
        {
            const void* current[15];
            int nc = 0;
            for (const double* ptr : ins)
                current[nc++] = ptr;
            for (double* ptr : outs)
                current[nc++] = ptr;
            unpinStale(current, nc);
        }
        for (const double* ptr : ins)
            pin(ptr, bytes);
        for (double* ptr : outs)
            pin(ptr, bytes);
        /*
         * The module-level call with host buffers.  What the advection needs (hice, cice, damage) goes up on the compute
         * stream; the five forcing fields go up, and are interpolated to CG, on a second stream while the advection
         * runs (they are first read by prepareIteration); hice and cice are final after advection + limiters and come
         * down on the second stream while the subcycles run.  Only u, v and the ice-ocean stress wait for the last subcycle.
         */
        const bool overlap = cfg.rheology != NSDG_FREEDRIFT && !std::getenv("NSDG_NO_COPY_OVERLAP");
        if (overlap && !copyStream) {
            NSDG_CUDA_CHECK(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
            NSDG_CUDA_CHECK(cudaEventCreateWithFlags(&evForcing, cudaEventDisableTiming));
            NSDG_CUDA_CHECK(cudaEventCreateWithFlags(&evCopyDone, cudaEventDisableTiming));
        }
        for (int i = 0; i < 3; ++i)
            if (ins[i]) {
                if (cfg.keep_dg_moments && !(inField[i] == NSDG_DAMAGE && cfg.rheology != NSDG_BBM))
                    setMeanKeepMoments(inField[i], ins[i]);
                else
                    setFieldAsync(inField[i], ins[i], 1);
            }
        if (overlap) {
            // the previous work of the compute stream (this handle's earlier calls) must not be overtaken
            NSDG_CUDA_CHECK(cudaEventRecord(evCopyDone, stream));
            NSDG_CUDA_CHECK(cudaStreamWaitEvent(copyStream, evCopyDone, 0));
            std::swap(stream, copyStream);
        }
        for (int i = 3; i < 8; ++i)
            if (ins[i])
                setFieldAsync(inField[i], ins[i], 1);
        if (overlap) {
            NSDG_CUDA_CHECK(cudaEventRecord(evForcing, stream));
            std::swap(stream, copyStream);
            forcingPending = true;
        }
        stepAsync(dt);
        int firstLate = 0;
        if (overlap) { // hice, cice: final since ev[1] (end of advection + limiters)
            NSDG_CUDA_CHECK(cudaStreamWaitEvent(copyStream, ev[1], 0));
            std::swap(stream, copyStream);
            for (int i = 0; i < 2; ++i)
                if (outs[i])
                    getFieldAsync(outField[i], outs[i], 1);
            NSDG_CUDA_CHECK(cudaEventRecord(evCopyDone, stream));
            std::swap(stream, copyStream);
            firstLate = 2;
        }
        // u, v and the ice-ocean stress can only leave after the last subcycle: on a uniform DG2/CG2 mesh one kernel forms the
        // four cell-mean fields (the general path: iostress + a full DG projection per field), then the four copies follow
        bool lateDone = false;
        if constexpr (CG == 2 && DGA == 6) {
            if (uniform && cfg.rheology != NSDG_FREEDRIFT && io->u_out && io->v_out && io->taux_out && io->tauy_out
                && !std::getenv("NSDG_NO_FUSED_EXPORT")) {
                const bool bbmR = cfg.rheology == NSDG_BBM;
                if (bbmR)
                    export_dg0_uniform_kernel<CG, NSDG_BBM><<<blocksFor(g.N), 128, 0, stream>>>(g, p, u, v, avgU, avgV, uO, vO, scratchDG);
                else
                    export_dg0_uniform_kernel<CG, NSDG_MEVP><<<blocksFor(g.N), 128, 0, stream>>>(g, p, u, v, u, v, uO, vO, scratchDG);
                double* late[4] = { io->u_out, io->v_out, io->taux_out, io->tauy_out };
                for (int k = 0; k < 4; ++k)
                    downloadPlanes(scratchDG.p + size_t(k) * g.Npad, 1, late[k]);
                lateDone = true;
            }
        }
        for (int i = firstLate; i < 7; ++i) {
            if (lateDone && i >= 3)
                continue;
            if (outs[i] && !(outField[i] == NSDG_DAMAGE && cfg.rheology != NSDG_BBM))
                getFieldAsync(outField[i], outs[i], 1);
        }
        if (overlap)
            NSDG_CUDA_CHECK(cudaStreamWaitEvent(stream, evCopyDone, 0));
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        checkHaloError();
        finishTiming();
    }

    // ------------------------------------------------------------------------------------
    // test access in the reference layouts
    // ------------------------------------------------------------------------------------
    struct Internal {
        enum Kind { CGF, CG1F, DGF, OPF, EDGEX, EDGEY } kind;
        double* ptr;
        int comps; //!< DGF: planes; OPF: entries
    };
    bool lookup(const std::string& n, Internal& out)
    {
        const std::map<std::string, Internal> t = {
            { "cg_u", { Internal::CGF, u, 1 } }, { "cg_v", { Internal::CGF, v, 1 } }, { "cgH", { Internal::CGF, cgH, 1 } },
            { "cgA", { Internal::CGF, cgA, 1 } }, { "uGradSSH", { Internal::CGF, gradX, 1 } },
            { "vGradSSH", { Internal::CGF, gradY, 1 } }, { "uOcean", { Internal::CGF, uO, 1 } },
            { "vOcean", { Internal::CGF, vO, 1 } }, { "uAtmos", { Internal::CGF, uA, 1 } }, { "vAtmos", { Internal::CGF, vA, 1 } },
            { "avgU", { Internal::CGF, avgU, 1 } }, { "avgV", { Internal::CGF, avgV, 1 } }, { "u0", { Internal::CGF, u0, 1 } },
            { "v0", { Internal::CGF, v0, 1 } }, { "lumpedcgmass", { Internal::CGF, lmass, 1 } },
            { "lumpedcg1mass", { Internal::CG1F, mass1, 1 } }, { "hice", { Internal::DGF, hice, DGA } },
            { "cice", { Internal::DGF, cice, DGA } }, { "damage", { Internal::DGF, damage, DGA } },
            { "s11", { Internal::DGF, s11, DGs } }, { "s12", { Internal::DGF, s12, DGs } }, { "s22", { Internal::DGF, s22, DGs } },
            { "velx", { Internal::DGF, velx, DGA } }, { "vely", { Internal::DGF, vely, DGA } },
            { "ssh", { Internal::DGF, ssh, 1 } }, { "divS1", { Internal::OPF, oD1, ND * DGs } },
            { "divS2", { Internal::OPF, oD2, ND * DGs } }, { "divM", { Internal::OPF, oDM, ND * DGs } },
            { "iMgradX", { Internal::OPF, oGx, DGs * ND } }, { "iMgradY", { Internal::OPF, oGy, DGs * ND } },
            { "iMM", { Internal::OPF, oGM, DGs * ND } }, { "iMJwPSI", { Internal::OPF, oB, DGs * Q } },
            { "iMJwPSI_dam", { Internal::OPF, oBd, DGA * Q } }, { "dX_SSH", { Internal::OPF, odX, 16 } },
            { "dY_SSH", { Internal::OPF, odY, 16 } }, { "AdvX", { Internal::OPF, tAdvX, DGA * QA } },
            { "AdvY", { Internal::OPF, tAdvY, DGA * QA } }, { "iMass", { Internal::OPF, tiMass, DGA * DGA } },
            { "normalvel_X", { Internal::EDGEX, nvX, EDA } }, { "normalvel_Y", { Internal::EDGEY, nvY, EDA } },
        };
        auto it = t.find(n);
        if (it == t.end() || it->second.ptr == nullptr)
            return false;
        out = it->second;
        return true;
    }
    void healDamage(double dt, double td, const double* deltaCi) override
    {
        requireMesh();
        if (cfg.rheology != NSDG_BBM)
            throw std::runtime_error("nsdg_heal_damage: the handle has no damage field (BBM only)");
        if (!(td > 0.0))
            throw std::runtime_error("nsdg_heal_damage: td must be positive");
        const double* dci = nullptr;
        if (deltaCi) { // DG0 host field -> plane 0 of the scratch field
            uploadPlanes(deltaCi, 1, 1, scratchDG);
            dci = scratchDG;
        }
        healing_kernel<<<blocksFor(g.N), 128, 0, stream>>>(g, dt, td, dci, cice, damage);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    bool streamedOpsBuilt = false;
    //! allocate and fill the reference's per-element momentum matrices (generic kernel; nsdg_get_internal of an operator)
    void ensureStreamedOps()
    {
        if (streamedOpsBuilt)
            return;
        const bool spherical = g.spherical != 0;
        const size_t opN = uniform ? 1 : g.Npad;
        const int nel = uniform ? 1 : g.N;
        oGx.alloc(size_t(DGs) * ND * opN);
        oGy.alloc(size_t(DGs) * ND * opN);
        oB.alloc(size_t(DGs) * Q * opN);
        oBd.alloc(size_t(DGA) * Q * opN);
        oD1.alloc(size_t(ND) * DGs * opN);
        oD2.alloc(size_t(ND) * DGs * opN);
        oGM.alloc(spherical ? size_t(DGs) * ND * opN : 0);
        oDM.alloc(spherical ? size_t(ND) * DGs * opN : 0);
        mop = { oGx, oGy, oGM, oB, oBd, oD1, oD2, oDM, odX, odY, opN, uniform ? 0 : 1 };
        setup_momentum_kernel<CG, DGA><<<blocksFor(nel, 64), 64, 0, stream>>>(g, vx, vy, mop, nel, true);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        streamedOpsBuilt = true;
    }
    void dims(int* nx, int* ny) override
    {
        requireMesh();
        *nx = g.nx;
        *ny = g.ny;
    }
    void getInternal(const std::string& name, double* host, size_t cap, size_t* count) override
    {
        requireMesh();
        for (const char* opName : { "divS1", "divS2", "divM", "iMgradX", "iMgradY", "iMM", "iMJwPSI", "iMJwPSI_dam" })
            if (name == opName)
                ensureStreamedOps();
        Internal in;
        if (!lookup(name, in))
            throw std::runtime_error("nsdg_get_internal: unknown or unallocated array " + name);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        const size_t N = g.N;
        size_t n = 0;
        std::vector<double> tmp;
        auto pull = [&](size_t cnt) {
            tmp.resize(cnt);
            NSDG_CUDA_CHECK(cudaMemcpy(tmp.data(), in.ptr, cnt * 8, cudaMemcpyDeviceToHost));
        };
        if (in.kind == Internal::CGF || in.kind == Internal::CG1F) {
            const int nxn = in.kind == Internal::CGF ? g.cgnx : g.nx + 1, nyn = in.kind == Internal::CGF ? g.cgny : g.ny + 1;
            const int st = in.kind == Internal::CGF ? g.cgs : cg1s;
            n = size_t(nxn) * nyn;
            if (count)
                *count = n;
            if (!host || cap < n)
                return;
            pull(size_t(st) * nyn);
            for (int r = 0; r < nyn; ++r)
                std::copy(tmp.begin() + size_t(r) * st, tmp.begin() + size_t(r) * st + nxn, host + size_t(r) * nxn);
        } else if (in.kind == Internal::DGF) {
            n = N * in.comps;
            if (count)
                *count = n;
            if (!host || cap < n)
                return;
            pull(size_t(in.comps) * g.Npad);
            for (size_t d = 0; d < N; ++d)
                for (int c = 0; c < in.comps; ++c)
                    host[d * in.comps + c] = tmp[size_t(c) * g.Npad + (d / g.nx) * g.nxs + d % g.nx];
        } else if (in.kind == Internal::OPF) {
            n = N * in.comps;
            if (count)
                *count = n;
            if (!host || cap < n)
                return;
            const size_t opN = uniform ? 1 : g.Npad;
            pull(size_t(in.comps) * opN);
            for (size_t d = 0; d < N; ++d)
                for (int k = 0; k < in.comps; ++k)
                    host[d * in.comps + k] = tmp[size_t(k) * opN + (uniform ? 0 : (d / g.nx) * g.nxs + d % g.nx)];
        } else {
            const size_t ne = in.kind == Internal::EDGEX ? size_t(g.nx) * (g.ny + 1) : size_t(g.nx + 1) * g.ny;
            const size_t pitch = alignUp(ne, 32);
            n = ne * in.comps;
            if (count)
                *count = n;
            if (!host || cap < n)
                return;
            pull(pitch * in.comps);
            for (size_t e = 0; e < ne; ++e)
                for (int c = 0; c < in.comps; ++c)
                    host[e * in.comps + c] = tmp[size_t(c) * pitch + e];
        }
    }
    void setInternal(const std::string& name, const double* host, size_t count) override
    {
        requireMesh();
        Internal in;
        if (!lookup(name, in))
            throw std::runtime_error("nsdg_set_internal: unknown or unallocated array " + name);
        NSDG_CUDA_CHECK(cudaStreamSynchronize(stream));
        const size_t N = g.N;
        std::vector<double> tmp;
        if (in.kind == Internal::CGF) {
            if (count != size_t(g.cgnx) * g.cgny)
                throw std::runtime_error("nsdg_set_internal: wrong size for " + name);
            tmp.assign(ncg, 0.0);
            for (int r = 0; r < g.cgny; ++r)
                std::copy(host + size_t(r) * g.cgnx, host + size_t(r + 1) * g.cgnx, tmp.begin() + size_t(r) * g.cgs);
            NSDG_CUDA_CHECK(cudaMemcpy(in.ptr, tmp.data(), ncg * 8, cudaMemcpyHostToDevice));
            legacySync();
        } else if (in.kind == Internal::DGF) {
            if (count != N * in.comps)
                throw std::runtime_error("nsdg_set_internal: wrong size for " + name);
            tmp.assign(size_t(in.comps) * g.Npad, 0.0);
            for (size_t d = 0; d < N; ++d)
                for (int c = 0; c < in.comps; ++c)
                    tmp[size_t(c) * g.Npad + (d / g.nx) * g.nxs + d % g.nx] = host[d * in.comps + c];
            NSDG_CUDA_CHECK(cudaMemcpy(in.ptr, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice));
            legacySync();
        } else
            throw std::runtime_error("nsdg_set_internal: array is read-only: " + name);
    }
};

static HandleBase* makeHandle(const nsdg_config& c)
{
    if (c.rheology != NSDG_MEVP && c.rheology != NSDG_BBM && c.rheology != NSDG_FREEDRIFT)
        throw std::runtime_error("nsdg_create: unknown rheology");
    if (c.nsteps < 1)
        throw std::runtime_error("nsdg_create: nsteps must be >= 1");
    if (c.dgadv == 6 && c.cgdegree == 2)
        return new Handle<2, 6>(c);
    if (c.dgadv == 3 && c.cgdegree == 1)
        return new Handle<1, 3>(c);
    throw std::runtime_error("nsdg_create: supported (dgadv, cgdegree) builds are (6,2) and (3,1)");
}

} // namespace nsdg

using namespace nsdg;

#define NSDG_TRY try {
#define NSDG_CATCH                                                                                                   \
    }                                                                                                                \
    catch (const std::exception& ex)                                                                                 \
    {                                                                                                                \
        g_lastError = ex.what();                                                                                     \
        return 1;                                                                                                    \
    }                                                                                                                \
    return 0;

static HandleBase* H(nsdg_handle h)
{
    if (!h)
        throw std::runtime_error("nsdg: null handle");
    HandleBase* b = reinterpret_cast<HandleBase*>(h);
    // every entry point works on the handle's own device, whatever the caller (or another handle) made current
    NSDG_CUDA_CHECK(cudaSetDevice(b->cfg.device));
    return b;
}

extern "C" {

void nsdg_config_default(nsdg_config* cfg)
{
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->rheology = NSDG_MEVP;
    cfg->dgadv = 6;
    cfg->cgdegree = 2;
    cfg->nsteps = 100;
    cfg->device = -1;
    cfg->use_cuda_graph = 1;
    cfg->force_general = 0;
    cfg->pin_host_buffers = 0;
    cfg->alpha = 1500.0;
    cfg->beta = 1500.0;
    cfg->rank = 0;
    cfg->nranks = 1;
    for (int& n : cfg->neighbour)
        n = -1;
}

int nsdg_create(const nsdg_config* cfg, nsdg_handle* out)
{
    NSDG_TRY
    if (!cfg || !out)
        throw std::runtime_error("nsdg_create: null argument");
    *out = reinterpret_cast<nsdg_handle>(makeHandle(*cfg));
    NSDG_CATCH
}
int nsdg_destroy(nsdg_handle h)
{
    NSDG_TRY
    delete H(h);
    NSDG_CATCH
}
int nsdg_set_mesh(nsdg_handle h, int nx, int ny, const double* coords_xy, const double* mask, int spherical)
{
    NSDG_TRY
    if (!coords_xy || !mask)
        throw std::runtime_error("nsdg_set_mesh: coords and mask are required (IDynamics.hpp:107-120 throws likewise)");
    H(h)->setMesh(nx, ny, coords_xy, mask, spherical);
    NSDG_CATCH
}
int nsdg_set_field(nsdg_handle h, int field, const double* host, int ncomp)
{
    NSDG_TRY
    if (!host)
        throw std::runtime_error("nsdg_set_field: null data");
    H(h)->setField(field, host, ncomp);
    NSDG_CATCH
}
int nsdg_get_field(nsdg_handle h, int field, double* host, int ncomp)
{
    NSDG_TRY
    if (!host)
        throw std::runtime_error("nsdg_get_field: null data");
    H(h)->getField(field, host, ncomp);
    NSDG_CATCH
}
int nsdg_step(nsdg_handle h, double dt_seconds)
{
    NSDG_TRY
    H(h)->step(dt_seconds);
    NSDG_CATCH
}
int nsdg_set_benchmark_forcing(nsdg_handle h, double elapsed_seconds, double domain_x, double domain_y)
{
    NSDG_TRY
    if (!(domain_x > 0) || !(domain_y > 0))
        throw std::runtime_error("nsdg_set_benchmark_forcing: the domain extent must be positive");
    H(h)->benchmarkForcing(elapsed_seconds, domain_x, domain_y);
    NSDG_CATCH
}
int nsdg_update(nsdg_handle h, const nsdg_update_io* io, double dt_seconds)
{
    NSDG_TRY
    if (!io)
        throw std::runtime_error("nsdg_update: null io");
    H(h)->update(io, dt_seconds);
    NSDG_CATCH
}
int nsdg_subcycles(nsdg_handle h, int n, float* ms)
{
    NSDG_TRY
    H(h)->subcycles(n, ms);
    NSDG_CATCH
}
int nsdg_time_kernels(nsdg_handle h, int n, float* strip_ms, float* lines_ms)
{
    NSDG_TRY
    if (n < 1 || !strip_ms || !lines_ms)
        throw std::runtime_error("nsdg_time_kernels: bad arguments");
    H(h)->timeKernels(n, strip_ms, lines_ms);
    NSDG_CATCH
}
int nsdg_get_timing(nsdg_handle h, nsdg_timing* t)
{
    NSDG_TRY
    *t = H(h)->timing;
    NSDG_CATCH
}
int nsdg_get_landmask(nsdg_handle h, unsigned char* out)
{
    NSDG_TRY
    HandleBase* b = H(h);
    if (!b->meshSet)
        throw std::runtime_error("nsdg_get_landmask: mesh not set");
    std::copy(b->landmask.begin(), b->landmask.end(), out);
    NSDG_CATCH
}
int nsdg_get_dirichlet(nsdg_handle h, int edge, long* out, size_t capacity, size_t* count)
{
    NSDG_TRY
    HandleBase* b = H(h);
    if (!b->meshSet || edge < 0 || edge > 3)
        throw std::runtime_error("nsdg_get_dirichlet: mesh not set or bad edge");
    const auto& d = b->dirichlet[edge];
    if (count)
        *count = d.size();
    if (out && capacity >= d.size())
        std::copy(d.begin(), d.end(), out);
    NSDG_CATCH
}
int nsdg_get_internal(nsdg_handle h, const char* name, double* host, size_t capacity, size_t* count)
{
    NSDG_TRY
    H(h)->getInternal(name, host, capacity, count);
    NSDG_CATCH
}
int nsdg_set_internal(nsdg_handle h, const char* name, const double* host, size_t count)
{
    NSDG_TRY
    H(h)->setInternal(name, host, count);
    NSDG_CATCH
}
int nsdg_heal_damage(nsdg_handle h, double dt_seconds, double td_seconds, const double* delta_cice)
{
    NSDG_TRY
    H(h)->healDamage(dt_seconds, td_seconds, delta_cice);
    NSDG_CATCH
}
int nsdg_set_boundaries(nsdg_handle h, const long* const* dirichlet, const size_t* ndirichlet, const long* periodic,
    const size_t* periodic_segment_sizes, size_t nsegments)
{
    NSDG_TRY
    if ((dirichlet && !ndirichlet) || (nsegments > 0 && !periodic_segment_sizes))
        throw std::runtime_error("nsdg_set_boundaries: list sizes missing");
    H(h)->setBoundaries(dirichlet, ndirichlet, periodic, periodic_segment_sizes, nsegments);
    NSDG_CATCH
}
int nsdg_advect_field(nsdg_handle h, int field, double dt_seconds, int rk_order, int nsteps, int limit_mode, double maxv, double minv)
{
    NSDG_TRY
    H(h)->advectField(field, dt_seconds, rk_order, nsteps, limit_mode, maxv, minv);
    NSDG_CATCH
}
namespace {
constexpr double kStateMagic = 1314079815.0; // 'NSDG'
constexpr double kStateVersion = 2.0; // 2: the header carries the partition box origin
constexpr size_t kStateHeader = 10;
//! a double of a state buffer that must be a small non-negative integer (a corrupted buffer may hold anything, and
//! converting NaN / negative / huge doubles to an integer type is undefined behaviour)
bool stateInt(double v, double maxv, size_t* out)
{
    if (!(v >= 0.0) || !(v <= maxv) || v != std::floor(v))
        return false;
    *out = size_t(v);
    return true;
}
std::vector<std::string> stateFields(const nsdg_config& c)
{
    std::vector<std::string> f = { "hice", "cice" };
    if (c.rheology == NSDG_BBM)
        f.push_back("damage");
    for (const char* n : { "cg_u", "cg_v", "s11", "s12", "s22" })
        f.push_back(n);
    if (c.rheology == NSDG_BBM) {
        f.push_back("avgU");
        f.push_back("avgV");
    }
    return f;
}
}
int nsdg_get_state(nsdg_handle h, double* host, size_t capacity, size_t* count)
{
    NSDG_TRY
    HandleBase* hb = H(h);
    const auto fields = stateFields(hb->cfg);
    size_t total = kStateHeader;
    std::vector<size_t> len(fields.size());
    for (size_t i = 0; i < fields.size(); ++i) {
        hb->getInternal(fields[i], nullptr, 0, &len[i]);
        total += 1 + len[i];
    }
    if (count)
        *count = total;
    if (!host || capacity < total) {
        if (host)
            throw std::runtime_error("nsdg_get_state: buffer too small");
        return 0;
    }
    int nx = 0, ny = 0;
    hb->dims(&nx, &ny);
    const double hdr[kStateHeader] = { kStateMagic, kStateVersion, double(hb->cfg.rheology), double(hb->cfg.dgadv),
        double(hb->cfg.cgdegree), double(nx), double(ny), double(fields.size()), double(hb->cfg.global_nx > 0 ? hb->cfg.box_x0 : 0),
        double(hb->cfg.global_nx > 0 ? hb->cfg.box_y0 : 0) };
    std::copy(hdr, hdr + kStateHeader, host);
    size_t off = kStateHeader;
    for (size_t i = 0; i < fields.size(); ++i) {
        host[off++] = double(len[i]);
        size_t got = 0;
        hb->getInternal(fields[i], host + off, len[i], &got);
        off += len[i];
    }
    NSDG_CATCH
}
int nsdg_set_state(nsdg_handle h, const double* host, size_t count)
{
    NSDG_TRY
    HandleBase* hb = H(h);
    const auto fields = stateFields(hb->cfg);
    int nx = 0, ny = 0;
    hb->dims(&nx, &ny);
    if (!host || count < kStateHeader || host[0] != kStateMagic || host[1] != kStateVersion)
        throw std::runtime_error("nsdg_set_state: not a state buffer of this library version");
    size_t hv[8] = {};
    for (int i = 0; i < 8; ++i)
        if (!stateInt(host[2 + i], 1e9, &hv[i]))
            throw std::runtime_error("nsdg_set_state: corrupted header");
    const size_t bx = hb->cfg.global_nx > 0 ? size_t(hb->cfg.box_x0) : 0, by = hb->cfg.global_nx > 0 ? size_t(hb->cfg.box_y0) : 0;
    if (hv[0] != size_t(hb->cfg.rheology) || hv[1] != size_t(hb->cfg.dgadv) || hv[2] != size_t(hb->cfg.cgdegree) || hv[3] != size_t(nx)
        || hv[4] != size_t(ny) || hv[5] != fields.size() || hv[6] != bx || hv[7] != by)
        throw std::runtime_error("nsdg_set_state: state was written for a different configuration, mesh or partition box");
    size_t off = kStateHeader;
    for (const std::string& f : fields) {
        if (off >= count)
            throw std::runtime_error("nsdg_set_state: truncated buffer");
        size_t len = 0, expect = 0;
        const double lenField = host[off++];
        if (!stateInt(lenField, double(count - off), &len)) // finite, integral, and off + len <= count without wrapping
            throw std::runtime_error("nsdg_set_state: corrupted or truncated buffer (field " + f + ")");
        hb->getInternal(f, nullptr, 0, &expect);
        if (len != expect)
            throw std::runtime_error("nsdg_set_state: field " + f + " has the wrong length for this mesh");
        hb->setInternal(f, host + off, len);
        off += len;
    }
    NSDG_CATCH
}
int nsdg_halo_export(nsdg_handle h, unsigned char* ipc_handle)
{
    NSDG_TRY
    if (!ipc_handle)
        throw std::runtime_error("nsdg_halo_export: null argument");
    H(h)->haloExport(ipc_handle);
    NSDG_CATCH
}
int nsdg_halo_connect(nsdg_handle h, int side, const unsigned char* peer_ipc_handle)
{
    NSDG_TRY
    if (!peer_ipc_handle)
        throw std::runtime_error("nsdg_halo_connect: null argument");
    H(h)->haloConnect(side, peer_ipc_handle);
    NSDG_CATCH
}
int nsdg_halo_ready(nsdg_handle h)
{
    NSDG_TRY
    H(h)->haloReady();
    NSDG_CATCH
}
const char* nsdg_last_error(void) { return g_lastError.c_str(); }
const char* nsdg_version(void) { return "nsdg-cuda 0.1 (sm_100a)"; }
}
