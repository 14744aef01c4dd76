#ifndef HALAB200_GPU_BLAS3_HPP
#define HALAB200_GPU_BLAS3_HPP
// BLAS-3 (reference gpu/hala_gpu_blas3.hpp) is dense, compute-bound and never reached from CG/GMRES:
// nothing to provide on the B200 hot path; the header exists so that the include chain of the reference layout is preserved.
#include "hala_gpu_blas2.hpp"
#endif
