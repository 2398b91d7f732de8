#include "../../../CUDADynamics.hpp"
