/*
 * nsdg_oracle_capi.cpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * C entry points of the CPU oracle, loaded with ctypes by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs ONLY.  The product library
 * (libnsdg_cuda.so) never links or calls this.
 *
 * The call sequence mirrors what MEVPDynamics / BBMDynamics do with their kernel object
 * (core/src/modules/DynamicsModule/MEVPDynamics.cpp:37-87, BBMDynamics.cpp:32-101).
 */
#include "nsdg_dynamics.hpp"

#include <cstring>

using namespace nso;

namespace {
thread_local std::string lastError;

IDynOracle* make(int rheology, int dgadv, int cg)
{
    const Rheology r = rheology == 1 ? BBM : (rheology == 2 ? FREEDRIFT : MEVP);
    if (dgadv == 6 && cg == 2)
        return new DynOracle<6, 2>(r);
    if (dgadv == 3 && cg == 2)
        return new DynOracle<3, 2>(r);
    if (dgadv == 3 && cg == 1)
        return new DynOracle<3, 1>(r);
    if (dgadv == 6 && cg == 1)
        return new DynOracle<6, 1>(r);
    if (dgadv == 1 && cg == 1)
        return new DynOracle<1, 1>(r);
    if (dgadv == 1 && cg == 2)
        return new DynOracle<1, 2>(r);
    return nullptr;
}
}

#define NSO_TRY try {
#define NSO_CATCH                                                                                  \
    }                                                                                              \
    catch (const std::exception& ex)                                                               \
    {                                                                                              \
        lastError = ex.what();                                                                     \
        return -1;                                                                                 \
    }

extern "C" {

const char* nso_last_error() { return lastError.c_str(); }
void nso_set_threads(int n) { omp_set_num_threads(n); }
int nso_max_threads() { return omp_get_max_threads(); }

void* nso_create(int rheology, int dgadv, int cg, int nsteps)
{
    IDynOracle* d = make(rheology, dgadv, cg);
    if (d && nsteps > 0)
        d->nSteps = nsteps;
    return d;
}
void nso_destroy(void* h) { delete static_cast<IDynOracle*>(h); }

int nso_set_mesh(void* h, int nx, int ny, const double* coords, const double* mask, int spherical)
{
    NSO_TRY static_cast<IDynOracle*>(h)->initialise(nx, ny, coords, mask, spherical != 0);
    return 0;
    NSO_CATCH
}
int nso_set_field(void* h, const char* name, const double* data, int ncomp)
{
    NSO_TRY static_cast<IDynOracle*>(h)->setData(name, data, ncomp);
    return 0;
    NSO_CATCH
}
int nso_update(void* h, double dt)
{
    NSO_TRY static_cast<IDynOracle*>(h)->update(dt);
    return 0;
    NSO_CATCH
}
//! run only `n` bare subcycles on the current state (no advection / prepare); returns seconds
double nso_subcycles(void* h, int n)
{
    IDynOracle* d = static_cast<IDynOracle*>(h);
    const double t0 = omp_get_wtime();
    for (int i = 0; i < n; ++i)
        d->subcycle();
    return omp_get_wtime() - t0;
}
int nso_sweep(void* h, const char* which)
{
    NSO_TRY static_cast<IDynOracle*>(h)->sweep(which);
    return 0;
    NSO_CATCH
}
int nso_set_delta_t(void* h, double deltaT);
double nso_last_subcycle_seconds(void* h) { return static_cast<IDynOracle*>(h)->subcycleSeconds; }
int nso_get_dg0(void* h, const char* name, double* out)
{
    NSO_TRY static_cast<IDynOracle*>(h)->getDG0Data(name, out);
    return 0;
    NSO_CATCH
}
int nso_get_dg(void* h, const char* name, double* out)
{
    NSO_TRY return static_cast<IDynOracle*>(h)->getDGData(name, out);
    NSO_CATCH
}
//! size of an internal array (reference layouts), or -1
long nso_raw_size(void* h, const char* name)
{
    const Vec* v = static_cast<IDynOracle*>(h)->raw(name);
    return v ? static_cast<long>(v->size()) : -1;
}
int nso_get_raw(void* h, const char* name, double* out)
{
    const Vec* v = static_cast<IDynOracle*>(h)->raw(name);
    if (!v) {
        lastError = std::string("unknown array ") + name;
        return -1;
    }
    std::copy(v->begin(), v->end(), out);
    return 0;
}
int nso_set_raw(void* h, const char* name, const double* in)
{
    Vec* v = static_cast<IDynOracle*>(h)->rawMutable(name);
    if (!v) {
        lastError = std::string("unknown array ") + name;
        return -1;
    }
    std::copy(in, in + v->size(), v->begin());
    return 0;
}
int nso_set_param(void* h, const char* name, double value)
{
    Params& p = static_cast<IDynOracle*>(h)->params;
    const std::string n(name);
    if (n == "alpha")
        p.alpha = value;
    else if (n == "beta")
        p.beta = value;
    else if (n == "Pstar")
        p.Pstar = value;
    else if (n == "DeltaMin")
        p.DeltaMin = value;
    else if (n == "fc")
        p.fc = value;
    else if (n == "ocean_turning_angle")
        p.ocean_turning_angle = value;
    else {
        lastError = "unknown parameter " + n;
        return -1;
    }
    return 0;
}

int nso_set_delta_t(void* h, double deltaT)
{
    static_cast<IDynOracle*>(h)->setDeltaT(deltaT);
    return 0;
}

#define MESH_OF(h) static_cast<IDynOracle*>(h)->meshMutable()
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

// ---- mesh lists (bit-exact integer state) ----
long nso_dirichlet_size(void* h, int edge)
{
    return static_cast<long>(static_cast<IDynOracle*>(h)->mesh().dirichlet[edge].size());
}
void nso_get_dirichlet(void* h, int edge, long* out)
{
    const auto& d = static_cast<IDynOracle*>(h)->mesh().dirichlet[edge];
    for (size_t i = 0; i < d.size(); ++i)
        out[i] = static_cast<long>(d[i]);
}
void nso_get_landmask(void* h, unsigned char* out)
{
    const auto& l = static_cast<IDynOracle*>(h)->mesh().landmask;
    std::copy(l.begin(), l.end(), out);
}
void nso_get_vertices(void* h, double* out)
{
    const Mesh& m = static_cast<IDynOracle*>(h)->mesh();
    for (size_t i = 0; i < m.nnodes; ++i) {
        out[2 * i] = m.vx[i];
        out[2 * i + 1] = m.vy[i];
    }
}
}
