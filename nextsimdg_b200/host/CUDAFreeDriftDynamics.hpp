/*!
 * @file CUDAFreeDriftDynamics.hpp
 *
 * The header the generated module table includes for the `[Nextsim::CUDAFreeDriftDynamics]` stanza of
 * core/src/modules/DynamicsModule/module.cfg (scripts/module_builder.py:49-50 writes
 * `#include "include/<file_prefix>.hpp"`); the class itself is declared in CUDADynamics.hpp.
 * Goes to core/src/modules/DynamicsModule/include/ in the nextsimdg tree.
 */
#ifndef CUDAFREEDRIFTDYNAMICS_HPP
#define CUDAFREEDRIFTDYNAMICS_HPP
#include "include/CUDADynamics.hpp"
#endif /* CUDAFREEDRIFTDYNAMICS_HPP */
