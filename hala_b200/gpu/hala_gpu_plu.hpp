#ifndef HALAB200_GPU_PLU_HPP
#define HALAB200_GPU_PLU_HPP
// potrf/potrs/getrf/getrs (reference gpu/hala_gpu_plu.hpp, cuSOLVER-Dn): dense direct solves, out of scope.
#include "hala_gpu_ilu.hpp"
#endif
