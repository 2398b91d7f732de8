/*! MOCK of core/src/include/NextsimModule.hpp / Module.hpp:74-223 (setImplementation only). */
#ifndef MOCK_NEXTSIMMODULE_HPP
#define MOCK_NEXTSIMMODULE_HPP
#include <string>
namespace Module {
template <typename I> struct Module {
    static std::string& chosen()
    {
        static std::string s;
        return s;
    }
    static void setImplementation(const std::string& name) { chosen() = name; }
};
}
#endif
