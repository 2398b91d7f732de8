/*! MOCK of the IDamageHealing interface header (only the type name is needed). */
#ifndef MOCK_IDAMAGEHEALING_HPP
#define MOCK_IDAMAGEHEALING_HPP
namespace Nextsim {
class IDamageHealing { };
}
#endif
