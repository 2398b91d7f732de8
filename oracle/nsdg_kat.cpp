/*
 * nsdg_kat.cpp -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Drivers that set up the reference's two advection known-answer tests on the oracle,
 * so that tests/ can pin the oracle against the L2 errors stored in the reference tests.
 *
 * Reference: dynamics/test/Advection_test.cpp:44-308          (rotating bump, box mesh)
 *            dynamics/test/AdvectionPeriodicBC_test.cpp:35-298 (ring mesh, periodic + limiter)
 */
#include "nsdg_transport.hpp"

using namespace nso;

namespace {

constexpr int dg2deg(int DG) { return DG == 1 ? 0 : (DG == 3 ? 1 : 2); }

// ---- Advection_test.cpp ----
const double Lx = 409600.0, Ly = 512000.0;

//! Advection_test.cpp:198-241
void boxMesh(Mesh& m, size_t Nx, size_t Ny, double distort)
{
    m.reset();
    m.spherical = false;
    m.nx = Nx;
    m.ny = Ny;
    m.nnodes = (Nx + 1) * (Ny + 1);
    m.nelements = Nx * Ny;
    m.vx.resize(m.nnodes);
    m.vy.resize(m.nnodes);
    size_t ii = 0;
    for (size_t iy = 0; iy <= Ny; ++iy)
        for (size_t ix = 0; ix <= Nx; ++ix, ++ii) {
            m.vx[ii] = Lx * ix / Nx + Lx * distort * sin(M_PI * ix / Nx * 3.0) * sin(M_PI * iy / Ny);
            m.vy[ii] = Ly * iy / Ny
                + Ly * distort * sin(M_PI * iy / Ny * 2.0) * sin(M_PI * ix / Nx * 2.0);
        }
    m.landmask.assign(m.nelements, 1);
    for (size_t i = 0; i < Nx; ++i)
        m.dirichlet[0].push_back(i);
    for (size_t i = 0; i < Ny; ++i)
        m.dirichlet[1].push_back(i * Nx + Nx - 1);
    for (size_t i = 0; i < Nx; ++i)
        m.dirichlet[2].push_back(Nx * (Ny - 1) + i);
    for (size_t i = 0; i < Ny; ++i)
        m.dirichlet[3].push_back(i * Nx);
}

//! Advection_test.cpp:71-84
double smoothBump(double x, double y)
{
    const double X = x / Lx, Y = y / Lx;
    const double r = (pow(X - 0.25, 2.0) + pow(Y - 0.5, 2.0)) / 0.025;
    return r < 1 ? exp(-1.0 / (1.0 - r)) : 0.0;
}

template <int DG> double katAdvection(int it, double distort, double* massloss)
{
    const size_t Nx = 12 * (1 << it), Ny = 13 * (1 << it);
    const size_t NT = 100 * (dg2deg(DG) + 1) * (dg2deg(DG) + 1) * (1 << it);
    Mesh m;
    boxMesh(m, Nx, Ny, distort);
    Transport<DG> tr(m);
    tr.scheme = (DG < 3) ? "rk2" : "rk3"; // Advection_test.cpp:131-135
    const double dt = Lx / NT;
    Vec phi;
    Function2DG<DG>(m, phi, smoothBump);
    const double mass0 = MeanValue<DG>(m, phi);
    Function2DG<DG>(m, tr.velx, [](double, double y) { return (y - 0.5 * Lx) * 2.0 * M_PI / Lx; });
    Function2DG<DG>(m, tr.vely, [](double x, double) { return (0.5 * Lx - x) * 2.0 * M_PI / Lx; });
    for (size_t iter = 1; iter <= NT; ++iter) {
        tr.reinitnormalvelocity();
        tr.step(dt, phi);
    }
    if (massloss)
        *massloss = (mass0 - MeanValue<DG>(m, phi)) / mass0;
    return sqrt(L2ErrorFunctionDG<DG>(m, phi, smoothBump)) / Lx;
}

// ---- AdvectionPeriodicBC_test.cpp ----
const double R0 = 100000.0, R1 = 250000.0;

//! AdvectionPeriodicBC_test.cpp:210-250
void ringMesh(Mesh& m, size_t Nx, size_t Ny)
{
    m.reset();
    m.spherical = false;
    m.nx = Nx;
    m.ny = Ny;
    m.nnodes = (Nx + 1) * (Ny + 1);
    m.nelements = Nx * Ny;
    m.vx.resize(m.nnodes);
    m.vy.resize(m.nnodes);
    size_t ii = 0;
    for (size_t iy = 0; iy <= Ny; ++iy)
        for (size_t ix = 0; ix <= Nx; ++ix, ++ii) {
            const double r = R0 + (R1 - R0) * iy / Ny;
            const double p = -2.0 * M_PI * ix / Nx;
            m.vx[ii] = r * cos(p);
            m.vy[ii] = r * sin(p);
        }
    m.landmask.assign(m.nelements, 1);
    for (size_t i = 0; i < Nx; ++i)
        m.dirichlet[0].push_back(i);
    for (size_t i = 0; i < Nx; ++i)
        m.dirichlet[2].push_back(Nx * (Ny - 1) + i);
    m.periodic.resize(1);
    m.periodic[0].resize(Ny);
    for (size_t i = 0; i < Ny; ++i)
        m.periodic[0][i] = { 1, (i + 1) * Nx - 1, i * Nx, i * (Nx + 1) + i };
}

//! AdvectionPeriodicBC_test.cpp:71-101
double packman(double x, double y)
{
    double r = pow((x + 175000.0) / 50000, 2.0) + pow(y / 50000.0, 2.0);
    if (r < 1.0)
        return exp(1.0) * exp(-1.0 / (1.0 - r));
    r = pow(x / 50000, 2.0) + pow((y - 175000.0) / 50000.0, 2.0);
    if (r < 1.0) {
        if (y > 175000)
            return 1.0;
        if (fabs(x) > 15000.0)
            return 1.0;
        return 0.0;
    }
    r = sqrt(pow((x - 175000.0) / 50000, 2.0) + pow(y / 50000.0, 2.0));
    if (r < 1.0)
        return 1.0 - r;
    r = sqrt(pow(x / 50000, 2.0) + pow((y + 175000.) / 50000.0, 2.0));
    if (r < 1.0)
        return 1.0;
    return 0.0;
}

template <int DG> double katPeriodic(int it, double* massloss)
{
    const size_t Nx = 32 * (1 << it), Ny = 4 * (1 << it);
    const size_t NT = 50 * (dg2deg(DG) + 1) * (dg2deg(DG) + 1) * (1 << it);
    Mesh m;
    ringMesh(m, Nx, Ny);
    Transport<DG> tr(m);
    tr.scheme = (DG < 3) ? "rk2" : "rk3";
    const double dt = R1 / NT;
    Vec phi;
    Function2DG<DG>(m, phi, packman);
    LimitMax<DG>(phi, 1.0);
    const double mass0 = MeanValue<DG>(m, phi);
    Function2DG<DG>(m, tr.velx, [](double, double y) { return y * 2.0 * M_PI / R1; });
    Function2DG<DG>(m, tr.vely, [](double x, double) { return -x * 2.0 * M_PI / R1; });
    for (size_t iter = 1; iter <= NT; ++iter) {
        tr.reinitnormalvelocity();
        tr.step(dt, phi);
        LimitMax<DG>(phi, 1.0);
        LimitMin<DG>(phi, 0.0);
    }
    if (massloss)
        *massloss = (mass0 - MeanValue<DG>(m, phi)) / mass0;
    return sqrt(L2ErrorFunctionDG<DG>(m, phi, packman)) / R0;
}

// ---- the same two tests with the TIME LOOP left to the caller (the CUDA library, through its C ABI): the oracle only
// builds the mesh, projects the initial data and the velocity (Function2DG) and measures the error at the end ----
template <int DG> int katStage(int kind, int it, double distort, double* coords, double* phi0, double* velx, double* vely, long* nt, double* dt)
{
    Mesh m;
    size_t NT;
    Vec phi, vx, vy;
    if (kind == 0) {
        boxMesh(m, 12 * (1 << it), 13 * (1 << it), distort);
        NT = 100 * (dg2deg(DG) + 1) * (dg2deg(DG) + 1) * (1 << it);
        *dt = Lx / NT;
        Function2DG<DG>(m, phi, smoothBump);
        Function2DG<DG>(m, vx, [](double, double y) { return (y - 0.5 * Lx) * 2.0 * M_PI / Lx; });
        Function2DG<DG>(m, vy, [](double x, double) { return (0.5 * Lx - x) * 2.0 * M_PI / Lx; });
    } else {
        ringMesh(m, 32 * (1 << it), 4 * (1 << it));
        NT = 50 * (dg2deg(DG) + 1) * (dg2deg(DG) + 1) * (1 << it);
        *dt = R1 / NT;
        Function2DG<DG>(m, phi, packman);
        LimitMax<DG>(phi, 1.0); // AdvectionPeriodicBC_test.cpp:169
        Function2DG<DG>(m, vx, [](double, double y) { return y * 2.0 * M_PI / R1; });
        Function2DG<DG>(m, vy, [](double x, double) { return -x * 2.0 * M_PI / R1; });
    }
    *nt = long(NT);
    for (size_t i = 0; i < m.nnodes; ++i) {
        coords[2 * i] = m.vx[i];
        coords[2 * i + 1] = m.vy[i];
    }
    std::copy(phi.begin(), phi.end(), phi0);
    std::copy(vx.begin(), vx.end(), velx);
    std::copy(vy.begin(), vy.end(), vely);
    return 0;
}
template <int DG> double katError(int kind, int it, double distort, const double* phiEnd)
{
    Mesh m;
    if (kind == 0)
        boxMesh(m, 12 * (1 << it), 13 * (1 << it), distort);
    else
        ringMesh(m, 32 * (1 << it), 4 * (1 << it));
    Vec phi(phiEnd, phiEnd + m.nelements * DG);
    return kind == 0 ? sqrt(L2ErrorFunctionDG<DG>(m, phi, smoothBump)) / Lx : sqrt(L2ErrorFunctionDG<DG>(m, phi, packman)) / R0;
}

} // namespace

extern "C" {

//! kind 0: Advection_test.cpp's box, kind 1: AdvectionPeriodicBC_test.cpp's ring.  Fills the vertex coordinates
//! ((nx+1)(ny+1) x 2), the projected initial field and velocity (N x DG each), the number of time steps and dt.
int nso_kat_stage(int kind, int DG, int it, double distort, double* coords, double* phi0, double* velx, double* vely, long* nt, double* dt)
{
    switch (DG) {
    case 3:
        return katStage<3>(kind, it, distort, coords, phi0, velx, vely, nt, dt);
    case 6:
        return katStage<6>(kind, it, distort, coords, phi0, velx, vely, nt, dt);
    }
    return -1;
}
//! `nsteps` steps of the staged problem with scheme rk<order> on the oracle's own transport object (no limiter): the
//! step-by-step checker of the CUDA Runge-Kutta stages
int nso_kat_steps(int kind, int DG, int it, double distort, int order, int nsteps, double* out)
{
    if (DG != 6 || kind != 0)
        return -1;
    Mesh m;
    boxMesh(m, 12 * (1 << it), 13 * (1 << it), distort);
    Transport<6> tr(m);
    tr.scheme = order == 1 ? "rk1" : (order == 2 ? "rk2" : "rk3");
    const double dt = Lx / (100 * 9 * (1 << it));
    Vec phi;
    Function2DG<6>(m, phi, smoothBump);
    Function2DG<6>(m, tr.velx, [](double, double y) { return (y - 0.5 * Lx) * 2.0 * M_PI / Lx; });
    Function2DG<6>(m, tr.vely, [](double x, double) { return (0.5 * Lx - x) * 2.0 * M_PI / Lx; });
    for (int i = 0; i < nsteps; ++i) {
        tr.reinitnormalvelocity();
        tr.step(dt, phi);
    }
    std::copy(phi.begin(), phi.end(), out);
    return 0;
}
//! the tests' error measure (sqrt(L2ErrorFunctionDG) / L) of a final field computed elsewhere
double nso_kat_error(int kind, int DG, int it, double distort, const double* phi)
{
    switch (DG) {
    case 3:
        return katError<3>(kind, it, distort, phi);
    case 6:
        return katError<6>(kind, it, distort, phi);
    }
    return -1;
}

//! L2 error of Advection_test.cpp's run<DG>(distort) at refinement `it`; -1 on bad DG.
double nso_kat_advection(int DG, int it, double distort, double* massloss)
{
    switch (DG) {
    case 1:
        return katAdvection<1>(it, distort, massloss);
    case 3:
        return katAdvection<3>(it, distort, massloss);
    case 6:
        return katAdvection<6>(it, distort, massloss);
    }
    return -1;
}
//! L2 error of AdvectionPeriodicBC_test.cpp's run<DG>() at refinement `it`.
double nso_kat_periodic(int DG, int it, double* massloss)
{
    switch (DG) {
    case 1:
        return katPeriodic<1>(it, massloss);
    case 3:
        return katPeriodic<3>(it, massloss);
    case 6:
        return katPeriodic<6>(it, massloss);
    }
    return -1;
}
}
