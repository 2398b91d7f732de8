/*!
 * @file CUDAMEVPDynamics.hpp
 *
 * The header the generated module table includes for the `[Nextsim::CUDAMEVPDynamics]` stanza of
 * core/src/modules/DynamicsModule/module.cfg (scripts/module_builder.py:49-50 writes
 * `#include "include/<file_prefix>.hpp"`); the class itself is declared in CUDADynamics.hpp.
 * Goes to core/src/modules/DynamicsModule/include/ in the nextsimdg tree.
 */
#ifndef CUDAMEVPDYNAMICS_HPP
#define CUDAMEVPDYNAMICS_HPP
#include "include/CUDADynamics.hpp"
#endif /* CUDAMEVPDYNAMICS_HPP */
