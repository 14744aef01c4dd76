#ifndef HALAB200_GPU_HPP
#define HALAB200_GPU_HPP
// hala_b200 replacement of the reference's gpu/ header directory (entry point, reference gpu/hala_gpu.hpp).
// Put THIS directory on the include path instead of the reference's gpu/ and define HALA_ENABLE_CUDA (+ HALA_ENABLE_GPU):
// wax/hala_lib_extensions.hpp:17-19 then pulls in this layer and every template above it (wax, hex/solvers, user code)
// compiles against the B200 backend with no edits.  See INTEGRATION.md.
#include "hala_gpu_overloads.hpp"
#endif
