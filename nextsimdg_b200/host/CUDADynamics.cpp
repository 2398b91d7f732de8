/*!
 * @file CUDADynamics.cpp
 *
 * Implementation of Nextsim::CUDAMEVPDynamics / CUDABBMDynamics / CUDAFreeDriftDynamics above the C ABI.
 * Control flow follows core/src/modules/DynamicsModule/MEVPDynamics.cpp:37-101, BBMDynamics.cpp:32-132 and
 * include/FreeDriftDynamics.hpp:37-81; every kernel.* call of the reference is one nsdg_* call here.
 *
 * Goes to core/src/modules/DynamicsModule/ in the nextsimdg tree.
 */
#include "include/CUDADynamics.hpp"

#include "include/gridNames.hpp"

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace Nextsim {

// Degrees to radians as a hex float (same constant as MEVPDynamics.cpp:21)
static const double radians = 0x1.1df46a2529d39p-6;

static const std::vector<std::string> namedFields = { hiceName, ciceName, uName, vName };
static const std::map<std::string, int> fieldIds = { { hiceName, NSDG_HICE }, { ciceName, NSDG_CICE },
    { damageName, NSDG_DAMAGE }, { uName, NSDG_U }, { vName, NSDG_V } };

void CUDADynamicsBase::check(int status)
{
    if (status != 0)
        throw std::runtime_error(std::string("libnsdg_cuda: ") + nsdg_last_error());
}

CUDADynamicsBase::CUDADynamicsBase(int rheologyIn, bool usesDamageIn)
    : IDynamics(usesDamageIn)
    , handle(nullptr)
    , rheology(rheologyIn)
{
    getStore().registerArray(Protected::ICE_U, &uice, RO);
    getStore().registerArray(Protected::ICE_V, &vice, RO);

    nsdg_config cfg;
    nsdg_config_default(&cfg);
    cfg.rheology = rheology;
    cfg.dgadv = DGCOMP;
    cfg.cgdegree = CGDEGREE;
    cfg.pin_host_buffers = 1; // the shared ModelArrays live as long as the model
    check(nsdg_create(&cfg, &handle));
}

CUDADynamicsBase::~CUDADynamicsBase()
{
    if (handle)
        nsdg_destroy(handle);
}

void CUDADynamicsBase::setData(const ModelState::DataMap& ms)
{
    IDynamics::setData(ms);

    bool isSpherical = checkSpherical(ms);

    ModelArray coords = ms.at(coordsName);
    if (isSpherical) {
        coords *= radians;
    }
    const ModelArray& mask = ms.at(maskName);
    // kernel.initialise(coords, isSpherical, mask)
    check(nsdg_set_mesh(handle, static_cast<int>(ModelArray::size(ModelArray::Dimension::X)),
        static_cast<int>(ModelArray::size(ModelArray::Dimension::Y)), coords.getData(), mask.getData(),
        isSpherical ? 1 : 0));

    uice = ms.at(uName);
    vice = ms.at(vName);

    // Set the data in the kernel arrays.
    for (const auto& fieldName : namedFields) {
        const ModelArray& data = ms.at(fieldName);
        check(nsdg_set_field(
            handle, fieldIds.at(fieldName), data.getData(), static_cast<int>(data.nComponents())));
    }
}

void CUDADynamicsBase::update(const TimestepTime& tst)
{
    std::cout << tst.start << std::endl;

    const bool bbm = rheology == NSDG_BBM;
    if (bbm) {
        // Fill the updated damage array with the initial value (BBMDynamics.cpp:71)
        damage = damage0.data();
    }
    if (taux.trueSize() != hice.data().trueSize()) {
        taux.resize();
        tauy.resize();
    }

    // One call = setData(hice, cice, [damage], uwind, vwind, uocean, vocean, ssh) + kernel.update(tst)
    // + getDG0Data(hice, cice, [damage], u, v, uiostress, viostress)   (MEVPDynamics.cpp:63-86)
    nsdg_update_io io {};
    io.hice_in = hice.data().getData();
    io.cice_in = cice.data().getData();
    io.damage_in = bbm ? damage.getData() : nullptr;
    io.uwind = uwind.data().getData();
    io.vwind = vwind.data().getData();
    io.uocean = uocean.data().getData();
    io.vocean = vocean.data().getData();
    io.ssh = ssh.data().getData();
    io.hice_out = const_cast<double*>(hice.data().getData());
    io.cice_out = const_cast<double*>(cice.data().getData());
    io.damage_out = bbm ? const_cast<double*>(damage.getData()) : nullptr;
    io.u_out = const_cast<double*>(uice.getData());
    io.v_out = const_cast<double*>(vice.getData());
    io.taux_out = const_cast<double*>(taux.getData());
    io.tauy_out = const_cast<double*>(tauy.getData());
    check(nsdg_update(handle, &io, tst.step.seconds()));
}

ModelArray CUDADynamicsBase::getDG0Data(int field, ModelArray::Type type) const
{
    ModelArray data(type);
    data.resize();
    check(nsdg_get_field(handle, field, const_cast<double*>(data.getData()), 1));
    return data;
}

ModelArray CUDADynamicsBase::getDGData(int field) const
{
    DGField data(ModelArray::Type::DG);
    data.resize();
    check(nsdg_get_field(handle, field, const_cast<double*>(data.getData()), DGCOMP));
    return data;
}

// ---------------------------------------------------------------------------------------------

CUDAMEVPDynamics::CUDAMEVPDynamics()
    : CUDADynamicsBase(NSDG_MEVP, false)
{
}

void CUDAMEVPDynamics::configure()
{
    Module::Module<Nextsim::IDamageHealing>::setImplementation("Nextsim::NoHealing");
}

ModelState CUDAMEVPDynamics::getStateRecursive(const OutputSpec& os) const
{
    ModelState state(IDynamics::getStateRecursive(os));
    if (os.allComponents()) {
        state.merge({
            { hiceName, getDG0Data(NSDG_HICE, ModelArray::Type::H) },
            { ciceName, getDG0Data(NSDG_CICE, ModelArray::Type::H) },
        });
    }
    return state;
}

// ---------------------------------------------------------------------------------------------

static const std::map<std::string, std::pair<ModelArray::Type, double>> defaultFields = {
    { damageName, { ModelArray::Type::H, 1.0 } },
};

CUDABBMDynamics::CUDABBMDynamics()
    : CUDADynamicsBase(NSDG_BBM, true)
{
}

void CUDABBMDynamics::setData(const ModelState::DataMap& ms)
{
    CUDADynamicsBase::setData(ms);
    // Data that can have a default value (BBMDynamics.cpp:49-63)
    for (const auto& entry : defaultFields) {
        const std::string& fieldName = entry.first;
        if (ms.count(fieldName) > 0) {
            const ModelArray& data = ms.at(fieldName);
            check(nsdg_set_field(
                handle, fieldIds.at(fieldName), data.getData(), static_cast<int>(data.nComponents())));
        } else {
            ModelArray data(entry.second.first);
            data.resize();
            data = entry.second.second;
            const ModelArray masked = mask(data);
            check(nsdg_set_field(handle, fieldIds.at(fieldName), masked.getData(), 1));
        }
    }
}

ModelState CUDABBMDynamics::getState() const
{
    ModelState state(IDynamics::getState());
    state.merge({
        { hiceName, getDGData(NSDG_HICE) },
        { ciceName, getDGData(NSDG_CICE) },
        { damageName, getDGData(NSDG_DAMAGE) },
    });
    return state;
}

ModelState CUDABBMDynamics::getStateRecursive(const OutputSpec& os) const
{
    ModelState state(IDynamics::getStateRecursive(os));
    if (os.allComponents()) {
        state.merge({
            { hiceName, getDGData(NSDG_HICE) },
            { ciceName, getDGData(NSDG_CICE) },
            { damageName, getDGData(NSDG_DAMAGE) },
        });
    }
    return state;
}

// ---------------------------------------------------------------------------------------------

CUDAFreeDriftDynamics::CUDAFreeDriftDynamics()
    : CUDADynamicsBase(NSDG_FREEDRIFT, false)
{
}

void CUDAFreeDriftDynamics::setData(const ModelState::DataMap& ms)
{
    IDynamics::setData(ms);

    bool isSpherical = checkSpherical(ms);

    ModelArray coords = ms.at(coordsName);
    if (isSpherical) {
        coords *= radians;
    }
    // kernel.initialise(coords, false, mask): the reference passes isSpherical = false whatever the coordinates are
    // (FreeDriftDynamics.hpp:69), after scaling longitude / latitude to radians
    const ModelArray& mask = ms.at(maskName);
    check(nsdg_set_mesh(handle, static_cast<int>(ModelArray::size(ModelArray::Dimension::X)),
        static_cast<int>(ModelArray::size(ModelArray::Dimension::Y)), coords.getData(), mask.getData(), 0));

    // Set the data in the kernel arrays.
    for (const auto& fieldName : namedFields) {
        const ModelArray& data = ms.at(fieldName);
        check(nsdg_set_field(
            handle, fieldIds.at(fieldName), data.getData(), static_cast<int>(data.nComponents())));
    }
}

void CUDAFreeDriftDynamics::update(const TimestepTime& tst)
{
    std::cout << tst.start << std::endl;

    // only the updated ice thickness and concentration and the ocean velocity go in; hice, cice, u, v come out: no wind,
    // no sea-surface height, no ice-ocean stress (FreeDriftDynamics.hpp:39-58)
    nsdg_update_io io {};
    io.hice_in = hice.data().getData();
    io.cice_in = cice.data().getData();
    io.uocean = uocean.data().getData();
    io.vocean = vocean.data().getData();
    io.hice_out = const_cast<double*>(hice.data().getData());
    io.cice_out = const_cast<double*>(cice.data().getData());
    io.u_out = const_cast<double*>(uice.getData());
    io.v_out = const_cast<double*>(vice.getData());
    check(nsdg_update(handle, &io, tst.step.seconds()));
}

} /* namespace Nextsim */
