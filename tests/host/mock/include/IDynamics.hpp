/*! MOCK of core/src/modules/include/IDynamics.hpp:17-121: same members and virtuals; the ModelArrayRef members are plain
 *  references into a static store filled by the test driver. */
#ifndef MOCK_IDYNAMICS_HPP
#define MOCK_IDYNAMICS_HPP
#include "ModelComponent.hpp"
#include "gridNames.hpp"
#include <stdexcept>

namespace Nextsim {
struct SharedArrays { // stands in for the ModelArrayRef store (IDynamics.hpp:89-99)
    HField hice, cice, hsnow, damage0, uwind, vwind, uocean, vocean, ssh;
    static SharedArrays& get()
    {
        static SharedArrays s;
        return s;
    }
};
class IDynamics : public ModelComponent {
public:
    IDynamics(bool usesDamageIn = false)
        : hice(SharedArrays::get().hice)
        , cice(SharedArrays::get().cice)
        , hsnow(SharedArrays::get().hsnow)
        , damage0(SharedArrays::get().damage0)
        , uwind(SharedArrays::get().uwind)
        , vwind(SharedArrays::get().vwind)
        , uocean(SharedArrays::get().uocean)
        , vocean(SharedArrays::get().vocean)
        , ssh(SharedArrays::get().ssh)
        , m_usesDamage(usesDamageIn)
    {
    }
    virtual ModelState getState() const { return { { { uName, mask(uice) }, { vName, mask(vice) } }, {} }; }
    virtual ModelState getStateRecursive(const OutputSpec& os) const { return os ? IDynamics::getState() : ModelState(); }
    std::string getName() const override { return "IDynamics"; }
    virtual void setData(const ModelState::DataMap&)
    {
        uice.resize();
        vice.resize();
        damage.resize();
        if (!m_usesDamage)
            damage = 0.;
    }
    virtual void update(const TimestepTime& tst) = 0;
    virtual bool usesDamage() const { return m_usesDamage; }
    // test access
    const HField& getU() const { return uice; }
    const HField& getV() const { return vice; }
    const HField& getTauX() const { return taux; }

protected:
    HField uice, vice, damage, taux, tauy;
    HField &hice, &cice, &hsnow, &damage0, &uwind, &vwind, &uocean, &vocean, &ssh;
    bool m_usesDamage;
    static bool checkSpherical(const ModelState::DataMap& ms)
    {
        if (ms.count(longitudeName) > 0 && ms.count(latitudeName) > 0)
            return true;
        if (ms.count(xName) > 0 && ms.count(yName) > 0)
            return false;
        throw std::runtime_error("Input data must contain either Cartesian or spherical coordinates.");
    }
};
}
#endif
