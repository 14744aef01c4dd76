#ifndef HALAB200_GPU_PLU_HPP
#define HALAB200_GPU_PLU_HPP
// potrf/potrs/getrf/getrs (reference gpu/hala_gpu_plu.hpp, cuSOLVER-Dn): dense direct solves, out of scope.
// This link of the include chain carries the fused solver front doors instead (hala_gpu_solvers.hpp).
#include "hala_gpu_solvers.hpp"
#endif
