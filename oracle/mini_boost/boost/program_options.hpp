// mini_boost -- TEST INFRASTRUCTURE ONLY.  The declarations of boost::program_options that the reference's
// core/src/include/Configured.hpp and Configurator.hpp mention, so that nextsimdg_b200/host/CUDADynamics.{hpp,cpp}
// can be COMPILED (not linked, not run) against the reference's real headers in tests/test_host_adapter.py.
// Written from the public Boost API documentation; nothing here parses anything.
#ifndef NSDG_MINI_BOOST_PROGRAM_OPTIONS
#define NSDG_MINI_BOOST_PROGRAM_OPTIONS
#include <map>
#include <stdexcept>
#include <string>
namespace boost {
namespace program_options {
    struct value_semantic {
        virtual ~value_semantic() = default;
    };
    template <class T> struct typed_value : value_semantic {
        T def {};
        typed_value* default_value(const T& v)
        {
            def = v;
            return this;
        }
    };
    template <class T> typed_value<T>* value() { return new typed_value<T>(); }
    class options_description;
    struct options_description_easy_init {
        options_description* owner;
        options_description_easy_init& operator()(const char*, const value_semantic*, const char* = "") { return *this; }
        options_description_easy_init& operator()(const char*, const char* = "") { return *this; }
    };
    class options_description {
    public:
        options_description() = default;
        explicit options_description(const std::string&) { }
        options_description_easy_init add_options() { return { this }; }
        options_description& add(const options_description&) { return *this; }
    };
    struct variable_value {
        template <class T> const T& as() const { throw std::logic_error("mini_boost: no values"); }
        bool empty() const { return true; }
        bool defaulted() const { return true; }
    };
    struct variables_map : std::map<std::string, variable_value> {
        const variable_value& operator[](const std::string& k) const
        {
            static variable_value v;
            auto it = find(k);
            return it == end() ? v : it->second;
        }
        size_t count(const std::string& k) const { return std::map<std::string, variable_value>::count(k); }
    };
}
}
#endif
