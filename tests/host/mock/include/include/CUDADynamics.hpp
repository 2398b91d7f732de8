#include "../../../../../nextsimdg_b200/host/CUDADynamics.hpp"
