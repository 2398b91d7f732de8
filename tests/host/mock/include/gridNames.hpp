/*! MOCK of core/src/include/gridNames.hpp:17-46 (the names only). */
#ifndef MOCK_GRIDNAMES_HPP
#define MOCK_GRIDNAMES_HPP
#include <string>
namespace Nextsim {
static const std::string hiceName = "hice", ciceName = "cice", maskName = "mask", uName = "u", vName = "v", damageName = "damage",
                         coordsName = "coords", latitudeName = "latitude", longitudeName = "longitude", xName = "x", yName = "y";
}
#endif
