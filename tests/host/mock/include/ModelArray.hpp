/*!
 * MOCK of core/src/include/ModelArray.hpp -- just enough surface (names, signatures, row-major N x ncomp
 * storage, ModelArray.hpp:92) to compile and run CUDADynamics.cpp without Eigen/Boost/netCDF.
 * Not part of the product; the real header replaces it in the nextsimdg tree.
 */
#ifndef MOCK_MODELARRAY_HPP
#define MOCK_MODELARRAY_HPP
#include <cstddef>
#include <map>
#include <string>
#include <vector>

namespace Nextsim {
class ModelArray {
public:
    enum class Type { H, U, V, Z, DG, DGSTRESS, CG, VERTEX };
    enum class Dimension { X, Y, Z, XVERTEX, YVERTEX };
    static std::map<Dimension, size_t>& dims()
    {
        static std::map<Dimension, size_t> d;
        return d;
    }
    static std::map<Type, size_t>& comps()
    {
        static std::map<Type, size_t> c = { { Type::H, 1 }, { Type::U, 1 }, { Type::V, 1 }, { Type::DG, 6 }, { Type::VERTEX, 2 } };
        return c;
    }
    static void setDimension(Dimension d, size_t n) { dims()[d] = n; }
    static size_t size(Dimension d) { return dims().at(d); }
    static size_t size(Type t)
    {
        return t == Type::VERTEX ? dims().at(Dimension::XVERTEX) * dims().at(Dimension::YVERTEX)
                                 : dims().at(Dimension::X) * dims().at(Dimension::Y);
    }
    ModelArray(Type t = Type::H)
        : type(t)
    {
    }
    Type getType() const { return type; }
    size_t nComponents() const { return comps().at(type); }
    size_t size() const { return size(type); }
    size_t trueSize() const { return m_data.size() / nComponents(); }
    void resize() { m_data.assign(size() * nComponents(), 0.0); }
    const double* getData() const { return m_data.data(); }
    double& operator[](size_t i) { return m_data[i]; }
    double operator[](size_t i) const { return m_data[i]; }
    ModelArray& operator=(double v)
    {
        for (auto& x : m_data)
            x = v;
        return *this;
    }
    ModelArray& operator*=(double v)
    {
        for (auto& x : m_data)
            x *= v;
        return *this;
    }
    const ModelArray& data() const { return *this; }
    ModelArray& data() { return *this; }

private:
    Type type;
    std::vector<double> m_data;
};
typedef ModelArray HField;
typedef ModelArray DGField;
}
#endif
