/*! MOCK of core/src/include/ModelComponent.hpp, ModelState.hpp, OutputSpec.hpp, Configured.hpp (see ModelArray.hpp mock). */
#ifndef MOCK_MODELCOMPONENT_HPP
#define MOCK_MODELCOMPONENT_HPP
#include "ModelArray.hpp"
#include <iostream>
#include <map>
#include <string>

namespace Nextsim {
struct ModelState {
    typedef std::map<std::string, ModelArray> DataMap;
    DataMap data;
    ModelState() = default;
    ModelState(const DataMap& d, const std::map<std::string, std::string>& = {})
        : data(d)
    {
    }
    ModelState& merge(DataMap&& src)
    {
        for (auto& kv : src)
            data[kv.first] = kv.second;
        return *this;
    }
};
struct OutputSpec {
    bool all = true;
    bool allComponents() const { return all; }
    operator bool() const { return all; }
};
typedef int OutputLevel;
enum { RO, RW };
namespace Protected {
    enum { ICE_U, ICE_V, IO_STRESS_U, IO_STRESS_V };
}
namespace Shared {
    enum { DAMAGE };
}
struct Store {
    template <typename K> void registerArray(K, ModelArray*, int) { }
};
class ModelComponent {
public:
    virtual ~ModelComponent() = default;
    static Store& getStore()
    {
        static Store s;
        return s;
    }
    static ModelArray*& oceanMaskPtr()
    {
        static ModelArray* p = nullptr;
        return p;
    }
    //! ModelComponent::mask (ModelComponent.cpp:96-110): land gets MissingData::value (1.7e38)
    static ModelArray mask(const ModelArray& in)
    {
        ModelArray out = in;
        if (oceanMaskPtr())
            for (size_t i = 0; i < out.size(); ++i)
                out[i] = in[i] * (*oceanMaskPtr())[i] + 1.7e38 * (1 - (*oceanMaskPtr())[i]);
        return out;
    }
    virtual std::string getName() const = 0;
};
template <typename T> class Configured {
public:
    virtual ~Configured() = default;
    virtual void configure() = 0;
};
struct Duration {
    double s = 0;
    double seconds() const { return s; }
};
struct TimestepTime {
    double start = 0;
    Duration step;
};
}
#endif
