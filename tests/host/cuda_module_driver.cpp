/*
 * Drives Nextsim::CUDAMEVPDynamics / CUDABBMDynamics / CUDAFreeDriftDynamics (nextsimdg_b200/host/CUDADynamics.cpp) exactly as PrognosticData
 * does (core/src/PrognosticData.cpp:56-100): configure, setData(ms), fill the shared arrays, update(tst) -- against the
 * MOCK nextsim headers, linked with libnsdg_cuda.so.  Prints one line of results for tests/test_host_adapter.py.
 *   usage: cuda_module_driver mevp|bbm|freedrift n nupdates
 */
#include "include/CUDADynamics.hpp"
#include "include/gridNames.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>

using namespace Nextsim;

int main(int argc, char** argv)
{
    const std::string rheo = argc > 1 ? argv[1] : "mevp";
    const size_t n = argc > 2 ? std::atoi(argv[2]) : 16;
    const int nupd = argc > 3 ? std::atoi(argv[3]) : 2;
    const double L = 512000.0, dx = L / n, dt = 120.0;
    ModelArray::setDimension(ModelArray::Dimension::X, n);
    ModelArray::setDimension(ModelArray::Dimension::Y, n);
    ModelArray::setDimension(ModelArray::Dimension::XVERTEX, n + 1);
    ModelArray::setDimension(ModelArray::Dimension::YVERTEX, n + 1);

    // the benchmark box of run/make_init_benchmark.py (same as nextsimdg_b200/synthetic.py::benchmark_box)
    ModelState::DataMap ms;
    ModelArray coords(ModelArray::Type::VERTEX), mask(ModelArray::Type::H), hice(ModelArray::Type::H), cice(ModelArray::Type::H),
        zero(ModelArray::Type::H);
    coords.resize();
    mask.resize();
    hice.resize();
    cice.resize();
    zero.resize();
    for (size_t j = 0; j <= n; ++j)
        for (size_t i = 0; i <= n; ++i) {
            coords[2 * (i + (n + 1) * j)] = i * L / n;
            coords[2 * (i + (n + 1) * j) + 1] = j * L / n;
        }
    for (size_t j = 0; j < n; ++j)
        for (size_t i = 0; i < n; ++i) {
            const size_t e = i + n * j;
            mask[e] = (i == 0 || j == 0 || i == n - 1 || j == n - 1) ? 0.0 : 1.0;
            hice[e] = (0.3 + 0.005 * (std::sin(60e-6 * (j * dx)) + std::sin(30e-6 * (i * dx)))) * mask[e];
            cice[e] = mask[e];
        }
    ms[coordsName] = coords;
    ms[maskName] = mask;
    ms[hiceName] = hice;
    ms[ciceName] = cice;
    ms[uName] = zero;
    ms[vName] = zero;
    ms[xName] = zero;
    ms[yName] = zero;
    ModelComponent::oceanMaskPtr() = &mask;

    SharedArrays& sh = SharedArrays::get();
    for (HField* f : { &sh.hice, &sh.cice, &sh.hsnow, &sh.damage0, &sh.uwind, &sh.vwind, &sh.uocean, &sh.vocean, &sh.ssh })
        f->resize();
    sh.hice = hice;
    sh.cice = cice;
    sh.damage0 = mask;

    std::unique_ptr<IDynamics> dyn;
    if (rheo == "bbm")
        dyn.reset(new CUDABBMDynamics());
    else if (rheo == "freedrift")
        dyn.reset(new CUDAFreeDriftDynamics());
    else {
        auto* m = new CUDAMEVPDynamics();
        m->configure();
        dyn.reset(m);
    }
    dyn->setData(ms);
    for (int k = 0; k < nupd; ++k) {
        const double t = k * dt, x0 = (L / 2) * (1 + t / (5 * 86400.0)), al = 72. / 180. * M_PI;
        for (size_t j = 0; j < n; ++j)
            for (size_t i = 0; i < n; ++i) {
                const size_t e = i + n * j;
                const double x = i * dx, y = j * dx, xp = x - x0, yp = y - x0;
                const double s = 1e-5 * std::exp(-1e-5 * std::hypot(xp, yp));
                sh.uwind[e] = -s * 30.0 * (std::cos(al) * xp + std::sin(al) * yp);
                sh.vwind[e] = -s * 30.0 * (-std::sin(al) * xp + std::cos(al) * yp);
                sh.uocean[e] = 0.01 * (2 * y / L - 1);
                sh.vocean[e] = 0.01 * (1 - 2 * x / L);
            }
        TimestepTime tst;
        tst.start = t;
        tst.step.s = dt;
        dyn->update(tst);
    }
    double su = 0, sv = 0, sh_ = 0, st = 0;
    for (size_t e = 0; e < n * n; ++e) {
        su += std::fabs(dyn->getU()[e]);
        sv += std::fabs(dyn->getV()[e]);
        sh_ += sh.hice[e];
        if (dyn->getTauX().trueSize() == n * n) // FreeDriftDynamics never exports (or sizes) the ice-ocean stress
            st += std::fabs(dyn->getTauX()[e]);
    }
    std::printf("RESULT %s %.17e %.17e %.17e %.17e\n", dyn->getName().c_str(), su, sv, sh_, st);
    return 0;
}
