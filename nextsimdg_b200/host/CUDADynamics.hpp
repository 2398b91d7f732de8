/*!
 * @file CUDADynamics.hpp
 *
 * Nextsim::CUDAMEVPDynamics / Nextsim::CUDABBMDynamics / Nextsim::CUDAFreeDriftDynamics -- drop-in IDynamics modules
 * whose kernel object lives on a B200 behind the C ABI of libnsdg_cuda.so (include/nsdg.h).
 *
 * They replace, method for method,
 *   Nextsim::MEVPDynamics       core/src/modules/DynamicsModule/MEVPDynamics.cpp:23-101
 *   Nextsim::BBMDynamics        core/src/modules/DynamicsModule/BBMDynamics.cpp:19-132
 *   Nextsim::FreeDriftDynamics  core/src/modules/DynamicsModule/include/FreeDriftDynamics.hpp:26-83
 * and are selected like them from the .cfg:   [Modules]  DynamicsModule = Nextsim::CUDAMEVPDynamics
 * (registration: INTEGRATION.md, integration/module.cfg.patch; the module builder includes include/<file_prefix>.hpp,
 * hence the one-line headers CUDAMEVPDynamics.hpp, CUDABBMDynamics.hpp, CUDAFreeDriftDynamics.hpp next to this file).  Host code only marshals ModelArray buffers; there is no CPU
 * fallback -- construction throws std::runtime_error if no CUDA device is usable.
 *
 * Goes to core/src/modules/DynamicsModule/include/ in the nextsimdg tree.
 */
#ifndef CUDADYNAMICS_HPP
#define CUDADYNAMICS_HPP

#include "include/IDamageHealing.hpp"
#include "include/IDynamics.hpp"
#include "include/ModelArray.hpp"
#include "include/ModelComponent.hpp"
#include "include/NextsimModule.hpp"

#include "nsdg.h"

#ifndef DGCOMP
#define DGCOMP 6
#endif
#ifndef CGDEGREE
#define CGDEGREE 2
#endif

namespace Nextsim {

//! What MEVPDynamics and BBMDynamics share, with the kernel member replaced by an nsdg_handle.
class CUDADynamicsBase : public IDynamics {
public:
    CUDADynamicsBase(int rheology, bool usesDamage);
    ~CUDADynamicsBase() override;

    void setData(const ModelState::DataMap& ms) override;
    void update(const TimestepTime& tst) override;

protected:
    //! kernel.getDG0Data(name) of the reference
    ModelArray getDG0Data(int field, ModelArray::Type type) const;
    //! kernel.getDGData(name) of the reference
    ModelArray getDGData(int field) const;
    //! throws std::runtime_error(nsdg_last_error()) on a non-zero status
    static void check(int status);

    nsdg_handle handle;
    int rheology;
};

class CUDAMEVPDynamics : public CUDADynamicsBase, public Configured<CUDAMEVPDynamics> {
public:
    CUDAMEVPDynamics();
    std::string getName() const override { return "CUDAMEVPDynamics"; }
    ModelState getStateRecursive(const OutputSpec& os) const override;
    void configure() override;
};

class CUDABBMDynamics : public CUDADynamicsBase {
public:
    CUDABBMDynamics();
    std::string getName() const override { return "CUDABBMDynamics"; }
    void setData(const ModelState::DataMap& ms) override;
    ModelState getState() const override;
    ModelState getStateRecursive(const OutputSpec& os) const override;
};

//! Drop-in for Nextsim::FreeDriftDynamics (FreeDriftDynamics.hpp:26-83): the ice follows the ocean, hice and cice are advected
class CUDAFreeDriftDynamics : public CUDADynamicsBase {
public:
    CUDAFreeDriftDynamics();
    std::string getName() const override { return "CUDAFreeDriftDynamics"; }
    void setData(const ModelState::DataMap& ms) override;
    void update(const TimestepTime& tst) override;
};

} /* namespace Nextsim */

#endif /* CUDADYNAMICS_HPP */
