/*
 * ref_driver.cpp -- TEST INFRASTRUCTURE. C entry points around the UNMODIFIED reference dynamics kernels
 * (MEVPDynamicsKernel / BBMDynamicsKernel / FreeDriftDynamicsKernel), compiled from the sources where they lie
 * under /root/reference by `make -C oracle ref` into oracle/_ref/libnsdg_ref_cg{1,2}.so.  Eigen (the reference's
 * un-vendored third-party dependency, absent from this image) is replaced by oracle/mini_eigen.
 *
 * The exported symbols are the nso_* subset of oracle/nsdg_oracle_capi.cpp, so oracle/__init__.py drives the
 * real reference and the restatement through the same Python class.  The call sequence is what
 * MEVPDynamics / BBMDynamics / FreeDriftDynamics do with their kernel member
 * (core/src/modules/DynamicsModule/MEVPDynamics.cpp:37-87, BBMDynamics.cpp:32-101, include/FreeDriftDynamics.hpp:30-58).
 * Protected kernel members are reached through derived classes (no reference file is modified or copied).
 */
#include "include/BBMDynamicsKernel.hpp"
#include "include/DynamicsParameters.hpp"
#include "include/FreeDriftDynamicsKernel.hpp"
#include "include/MEBParameters.hpp"
#include "include/MEVPDynamicsKernel.hpp"
#include "include/ModelArray.hpp"
#include "include/Time.hpp"
#include "include/VPParameters.hpp"
#include "include/gridNames.hpp"

#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <omp.h>
#include <stdexcept>
#include <string>

using namespace Nextsim;

namespace {
thread_local std::string lastError;

struct Raw {
    double* p;
    size_t n;
};
template <class M> Raw rawOf(M& m) { return { m.data(), static_cast<size_t>(m.rows() * m.cols()) }; }

struct IRef {
    int nx = 0, ny = 0;
    double lastUpdateSeconds = 0;
    virtual ~IRef() = default;
    virtual void setNSteps(size_t n) = 0;
    virtual void initialise(const ModelArray& coords, bool spherical, const ModelArray& mask) = 0;
    virtual void setData(const std::string& name, const ModelArray& data) = 0;
    virtual void update(const TimestepTime& tst) = 0;
    virtual ModelArray dg0(const std::string& name) = 0;
    virtual ModelArray dg(const std::string& name) = 0;
    virtual Raw raw(const std::string& name) = 0;
    virtual const ParametricMesh& mesh() = 0;
    virtual void sweep(const std::string& which) { throw std::runtime_error("sweep " + which + " not available"); }
};

// the storage of the per-element operator lists is a std::vector of fixed-size matrices: contiguous
template <class V> Raw rawOfList(V& v)
{
    if (v.empty())
        return { nullptr, 0 };
    return { v[0].data(), v.size() * static_cast<size_t>(v[0].rows() * v[0].cols()) };
}

// protected members of DGTransport, reached through a pointer-to-member named in a derived class
template <int DG> struct TransportPeek : public DGTransport<DG> {
    static ParametricTransportMap<DG>& map(DGTransport<DG>& t) { return t.*(&TransportPeek::parammap); }
    static DGVector<DG>& vx(DGTransport<DG>& t) { return t.*(&TransportPeek::velx); }
    static DGVector<DG>& vy(DGTransport<DG>& t) { return t.*(&TransportPeek::vely); }
    static auto& nvX(DGTransport<DG>& t) { return t.*(&TransportPeek::normalvel_X); }
    static auto& nvY(DGTransport<DG>& t) { return t.*(&TransportPeek::normalvel_Y); }
};

#define NSR_COMMON_RAW(K)                                                                                              \
    if (name == "hice") return rawOf(this->DK::hice);                                                                      \
    if (name == "cice") return rawOf(this->DK::cice);                                                                      \
    if (name == "ssh") return rawOf(this->DK::seaSurfaceHeight);                                                           \
    if (name == "e11") return rawOf(this->DK::e11);                                                                        \
    if (name == "e12") return rawOf(this->DK::e12);                                                                        \
    if (name == "e22") return rawOf(this->DK::e22);                                                                        \
    if (name == "s11") return rawOf(this->DK::s11);                                                                        \
    if (name == "s12") return rawOf(this->DK::s12);                                                                        \
    if (name == "s22") return rawOf(this->DK::s22);                                                                        \
    if (name == "cg_u") return rawOf(this->CG::u);                                                                         \
    if (name == "cg_v") return rawOf(this->CG::v);                                                                         \
    if (name == "cgA") return rawOf(this->CG::cgA);                                                                        \
    if (name == "cgH") return rawOf(this->CG::cgH);                                                                        \
    if (name == "uGradSSH") return rawOf(this->CG::uGradSeasurfaceHeight);                                                 \
    if (name == "vGradSSH") return rawOf(this->CG::vGradSeasurfaceHeight);                                                 \
    if (name == "dStressX") return rawOf(this->CG::dStressX);                                                              \
    if (name == "dStressY") return rawOf(this->CG::dStressY);                                                              \
    if (name == "uOcean") return rawOf(this->CG::uOcean);                                                                  \
    if (name == "vOcean") return rawOf(this->CG::vOcean);                                                                  \
    if (name == "uAtmos") return rawOf(this->CG::uAtmos);                                                                  \
    if (name == "vAtmos") return rawOf(this->CG::vAtmos);                                                                  \
    if (name == "lumpedcgmass") return rawOf(this->CG::pmap->lumpedcgmass);                                                \
    if (name == "lumpedcg1mass") return rawOf(this->CG::pmap->lumpedcg1mass);                                              \
    if (name == "divS1") return rawOfList(this->CG::pmap->divS1);                                                          \
    if (name == "divS2") return rawOfList(this->CG::pmap->divS2);                                                          \
    if (name == "divM") return rawOfList(this->CG::pmap->divM);                                                            \
    if (name == "iMgradX") return rawOfList(this->CG::pmap->iMgradX);                                                      \
    if (name == "iMgradY") return rawOfList(this->CG::pmap->iMgradY);                                                      \
    if (name == "iMM") return rawOfList(this->CG::pmap->iMM);                                                              \
    if (name == "iMJwPSI") return rawOfList(this->CG::pmap->iMJwPSI);                                                      \
    if (name == "iMJwPSI_dam") return rawOfList(this->CG::pmap->iMJwPSI_dam);                                              \
    if (name == "dX_SSH") return rawOfList(this->CG::pmap->dX_SSH);                                                        \
    if (name == "dY_SSH") return rawOfList(this->CG::pmap->dY_SSH);                                                        \
    if (name == "velx") return rawOf(TransportPeek<DG>::vx(*this->DK::dgtransport));                                 \
    if (name == "vely") return rawOf(TransportPeek<DG>::vy(*this->DK::dgtransport));                                 \
    if (name == "normalvel_X") return rawOf(TransportPeek<DG>::nvX(*this->DK::dgtransport));                           \
    if (name == "normalvel_Y") return rawOf(TransportPeek<DG>::nvY(*this->DK::dgtransport));                           \
    if (name == "AdvX") return rawOfList(TransportPeek<DG>::map(*this->DK::dgtransport).AdvectionCellTermX);           \
    if (name == "AdvY") return rawOfList(TransportPeek<DG>::map(*this->DK::dgtransport).AdvectionCellTermY);           \
    if (name == "iMass") return rawOfList(TransportPeek<DG>::map(*this->DK::dgtransport).InverseDGMassMatrix);

template <int DG> struct MEVPRef : public IRef, public MEVPDynamicsKernel<DG> {
    using K = MEVPDynamicsKernel<DG>;
    using CG = CGDynamicsKernel<DG>;
    using DK = DynamicsKernel<DG, DGstressComp>;
    MEVPRef(const DynamicsParameters& p)
        : K(p)
    {
    }
    void setNSteps(size_t n) override { this->DK::nSteps = n; }
    void initialise(const ModelArray& c, bool s, const ModelArray& m) override { K::initialise(c, s, m); }
    void setData(const std::string& n, const ModelArray& d) override { K::setData(n, d); }
    void update(const TimestepTime& t) override { K::update(t); }
    ModelArray dg0(const std::string& n) override { return K::getDG0Data(n); }
    ModelArray dg(const std::string& n) override { return K::getDGData(n); }
    const ParametricMesh& mesh() override { return *this->DK::smesh; }
    Raw raw(const std::string& name) override
    {
        NSR_COMMON_RAW(K)
        if (name == "u0") return rawOf(this->u0);
        if (name == "v0") return rawOf(this->v0);
        return { nullptr, 0 };
    }
    void sweep(const std::string& w) override
    {
        if (w == "strain")
            this->projectVelocityToStrain();
        else if (w == "divergence")
            this->stressDivergence();
        else if (w == "boundaries")
            this->applyBoundaries();
        else if (w == "prepare") {
            this->prepareIteration({ { hiceName, this->hice }, { ciceName, this->cice } });
        } else
            IRef::sweep(w);
    }
};

template <int DG> struct BBMRef : public IRef, public BBMDynamicsKernel<DG> {
    using K = BBMDynamicsKernel<DG>;
    using CG = CGDynamicsKernel<DG>;
    using DK = DynamicsKernel<DG, DGstressComp>;
    BBMRef(const DynamicsParameters& p)
        : K(p)
    {
    }
    void setNSteps(size_t n) override { this->DK::nSteps = n; }
    void initialise(const ModelArray& c, bool s, const ModelArray& m) override { K::initialise(c, s, m); }
    void setData(const std::string& n, const ModelArray& d) override { K::setData(n, d); }
    void update(const TimestepTime& t) override { K::update(t); }
    ModelArray dg0(const std::string& n) override { return K::getDG0Data(n); }
    ModelArray dg(const std::string& n) override { return K::getDGData(n); }
    const ParametricMesh& mesh() override { return *this->DK::smesh; }
    Raw raw(const std::string& name) override
    {
        NSR_COMMON_RAW(K)
        if (name == "damage") return rawOf(this->damage);
        if (name == "avgU") return rawOf(this->avgU);
        if (name == "avgV") return rawOf(this->avgV);
        return { nullptr, 0 };
    }
    void sweep(const std::string& w) override
    {
        if (w == "strain")
            this->projectVelocityToStrain();
        else if (w == "divergence")
            this->stressDivergence();
        else if (w == "boundaries")
            this->applyBoundaries();
        else
            IRef::sweep(w);
    }
};

template <int DG> struct FreeDriftRef : public IRef, public FreeDriftDynamicsKernel<DG> {
    using K = FreeDriftDynamicsKernel<DG>;
    using CG = CGDynamicsKernel<DG>;
    using DK = DynamicsKernel<DG, DGstressComp>;
    FreeDriftRef(const DynamicsParameters& p)
        : K(p)
    {
    }
    void setNSteps(size_t n) override { this->DK::nSteps = n; }
    void initialise(const ModelArray& c, bool s, const ModelArray& m) override { K::initialise(c, s, m); }
    void setData(const std::string& n, const ModelArray& d) override { K::setData(n, d); }
    void update(const TimestepTime& t) override { K::update(t); }
    ModelArray dg0(const std::string& n) override { return K::getDG0Data(n); }
    ModelArray dg(const std::string& n) override { return K::getDGData(n); }
    const ParametricMesh& mesh() override { return *this->DK::smesh; }
    Raw raw(const std::string& name) override
    {
        NSR_COMMON_RAW(K)
        return { nullptr, 0 };
    }
};

struct Handle {
    // parameter blocks must outlive the kernels (held by reference, VPCGDynamicsKernel.hpp:55)
    VPParameters vp;
    MEBParameters meb;
    DynamicsParameters dp;
    std::unique_ptr<IRef> k;
    int dgadv = 0;
};

template <int DG> IRef* makeKernel(Handle& h, int rheology)
{
    if (rheology == 0)
        return new MEVPRef<DG>(h.vp);
    if (rheology == 1)
        return new BBMRef<DG>(h.meb);
    return new FreeDriftRef<DG>(h.dp);
}

ModelArray makeField(const Handle& h, const double* data, int ncomp)
{
    if (ncomp == 1) {
        ModelArray ma(ModelArray::Type::H);
        ma.resize();
        ma.setData(data);
        return ma;
    }
    if (ncomp == h.dgadv) {
        ModelArray ma(ModelArray::Type::DG);
        ma.resize();
        ma.setData(data);
        return ma;
    }
    if (ncomp == DGstressComp) {
        ModelArray ma(ModelArray::Type::DGSTRESS);
        ma.resize();
        ma.setData(data);
        return ma;
    }
    throw std::runtime_error("unsupported component count");
}
} // namespace

#define NSR_TRY try {
#define NSR_CATCH                                                                                                      \
    }                                                                                                                  \
    catch (const std::exception& ex)                                                                                   \
    {                                                                                                                  \
        lastError = ex.what();                                                                                         \
        return -1;                                                                                                     \
    }

extern "C" {

const char* nso_last_error() { return lastError.c_str(); }
void nso_set_threads(int n) { omp_set_num_threads(n); }
int nso_max_threads() { return omp_get_max_threads(); }
int nso_is_reference() { return 1; }
int nso_cgdegree() { return CGdegree; }

void* nso_create(int rheology, int dgadv, int cg, int nsteps)
{
    if (cg != CGdegree || (dgadv != 1 && dgadv != 3 && dgadv != 6))
        return nullptr;
    Handle* h = new Handle;
    h->dgadv = dgadv;
    h->k.reset(dgadv == 6 ? makeKernel<6>(*h, rheology) : (dgadv == 3 ? makeKernel<3>(*h, rheology) : makeKernel<1>(*h, rheology)));
    if (nsteps > 0)
        h->k->setNSteps(static_cast<size_t>(nsteps));
    return h;
}
void nso_destroy(void* hv) { delete static_cast<Handle*>(hv); }

int nso_set_mesh(void* hv, int nx, int ny, const double* coords, const double* mask, int spherical)
{
    NSR_TRY Handle* h = static_cast<Handle*>(hv);
    // the array sizes are process-global in the reference (ModelArray::definedDimensions)
    ModelArray::setDimension(ModelArray::Dimension::X, nx);
    ModelArray::setDimension(ModelArray::Dimension::Y, ny);
    ModelArray::setDimension(ModelArray::Dimension::XVERTEX, nx + 1);
    ModelArray::setDimension(ModelArray::Dimension::YVERTEX, ny + 1);
    ModelArray::setDimension(ModelArray::Dimension::XCG, CGdegree * nx + 1);
    ModelArray::setDimension(ModelArray::Dimension::YCG, CGdegree * ny + 1);
    ModelArray::setNComponents(ModelArray::Type::DG, h->dgadv);
    ModelArray::setNComponents(ModelArray::Type::DGSTRESS, DGstressComp);
    ModelArray::setNComponents(ModelArray::Type::VERTEX, 2);
    h->k->nx = nx;
    h->k->ny = ny;
    ModelArray c(ModelArray::Type::VERTEX);
    c.resize();
    c.setData(coords);
    ModelArray m(ModelArray::Type::H);
    m.resize();
    m.setData(mask);
    h->k->initialise(c, spherical != 0, m);
    return 0;
    NSR_CATCH
}
int nso_set_field(void* hv, const char* name, const double* data, int ncomp)
{
    NSR_TRY Handle* h = static_cast<Handle*>(hv);
    h->k->setData(name, makeField(*h, data, ncomp));
    return 0;
    NSR_CATCH
}
int nso_update(void* hv, double dt)
{
    NSR_TRY Handle* h = static_cast<Handle*>(hv);
    TimestepTime tst = { TimePoint(), Duration(dt) };
    const auto t0 = std::chrono::steady_clock::now();
    h->k->update(tst);
    h->k->lastUpdateSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
    NSR_CATCH
}
double nso_last_update_seconds(void* hv) { return static_cast<Handle*>(hv)->k->lastUpdateSeconds; }
int nso_sweep(void* hv, const char* which)
{
    NSR_TRY static_cast<Handle*>(hv)->k->sweep(which);
    return 0;
    NSR_CATCH
}
int nso_get_dg0(void* hv, const char* name, double* out)
{
    NSR_TRY Handle* h = static_cast<Handle*>(hv);
    ModelArray ma = h->k->dg0(name);
    std::copy(ma.getData(), ma.getData() + size_t(h->k->nx) * h->k->ny, out);
    return 0;
    NSR_CATCH
}
int nso_get_dg(void* hv, const char* name, double* out)
{
    NSR_TRY Handle* h = static_cast<Handle*>(hv);
    ModelArray ma = h->k->dg(name);
    const size_t nc = ma.nComponents();
    std::copy(ma.getData(), ma.getData() + size_t(h->k->nx) * h->k->ny * nc, out);
    return static_cast<int>(nc);
    NSR_CATCH
}
long nso_raw_size(void* hv, const char* name)
{
    Raw r = static_cast<Handle*>(hv)->k->raw(name);
    return r.p ? static_cast<long>(r.n) : -1;
}
int nso_get_raw(void* hv, const char* name, double* out)
{
    Raw r = static_cast<Handle*>(hv)->k->raw(name);
    if (!r.p) {
        lastError = std::string("unknown array ") + name;
        return -1;
    }
    std::copy(r.p, r.p + r.n, out);
    return 0;
}
int nso_set_raw(void* hv, const char* name, const double* in)
{
    Raw r = static_cast<Handle*>(hv)->k->raw(name);
    if (!r.p) {
        lastError = std::string("unknown array ") + name;
        return -1;
    }
    std::copy(in, in + r.n, r.p);
    return 0;
}
#define MESH_OF(h) const_cast<ParametricMesh&>(static_cast<Handle*>(h)->k->mesh())
//! ParametricMesh::dirichlet[edge] / ::periodic assigned directly, as the reference's advection tests do
//! (Advection_test.cpp:222-240, AdvectionPeriodicBC_test.cpp:228-249).  dir[e] == nullptr keeps list e.
int nso_set_boundaries(void* hv, const long* const* dir, const size_t* ndir, const long* per, const size_t* segSizes, size_t nseg)
{
    auto& m = MESH_OF(hv);
    for (int e = 0; e < 4; ++e)
        if (dir && dir[e]) {
            m.dirichlet[e].clear();
            for (size_t i = 0; i < ndir[e]; ++i)
                m.dirichlet[e].push_back(static_cast<size_t>(dir[e][i]));
        }
    m.periodic.clear();
    m.periodic.resize(nseg);
    size_t k = 0;
    for (size_t s = 0; s < nseg; ++s)
        for (size_t i = 0; i < segSizes[s]; ++i, ++k)
            m.periodic[s].push_back({ size_t(per[4 * k]), size_t(per[4 * k + 1]), size_t(per[4 * k + 2]), size_t(per[4 * k + 3]) });
    return 0;
}
#undef MESH_OF
long nso_dirichlet_size(void* hv, int edge)
{
    return static_cast<long>(static_cast<Handle*>(hv)->k->mesh().dirichlet[edge].size());
}
void nso_get_dirichlet(void* hv, int edge, long* out)
{
    const auto& d = static_cast<Handle*>(hv)->k->mesh().dirichlet[edge];
    for (size_t i = 0; i < d.size(); ++i)
        out[i] = static_cast<long>(d[i]);
}
void nso_get_landmask(void* hv, unsigned char* out)
{
    const auto& l = static_cast<Handle*>(hv)->k->mesh().landmask;
    for (size_t i = 0; i < l.size(); ++i)
        out[i] = l[i] ? 1 : 0;
}
void nso_get_vertices(void* hv, double* out)
{
    const ParametricMesh& m = static_cast<Handle*>(hv)->k->mesh();
    for (size_t i = 0; i < m.nnodes; ++i) {
        out[2 * i] = m.vertices(i, 0);
        out[2 * i + 1] = m.vertices(i, 1);
    }
}
}
