#ifndef HALAB200_CUDA_HPP
#define HALAB200_CUDA_HPP
// reference gpu/hala_cuda.hpp: user-facing include that switches the CUDA backend on.
#ifdef HALA_ENABLE_ROCM
    #error "Cannot include both cuda and rocm!"
#endif
#ifndef HALA_ENABLE_CUDA
#define HALA_ENABLE_CUDA
#endif
#ifndef HALA_ENABLE_GPU
#define HALA_ENABLE_GPU
#endif
#include "hala_gpu.hpp"
namespace hala{}
#endif
