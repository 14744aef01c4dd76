#ifndef HALAB200_CUDA_COMMON_HPP
#define HALAB200_CUDA_COMMON_HPP
// Foundation of the B200 header layer — takes the place of reference gpu/hala_cuda_common.hpp (:78-330):
// status -> exception, device count, op-character helpers, raw device allocation and copies, all through the C ABI of
// libhalab200 (include/halab200.h).  No cuBLAS / cuSPARSE / cuSOLVER anywhere in this layer.
#include "hala_input_checker.hpp"      // HALA common/ (stock checkout)
#include "halab200.h"

#if defined(__has_include)
  #if __has_include(<cuda_runtime_api.h>)
    #include <cuda_runtime_api.h>
    #define HALAB200_HAS_CUDART_HEADERS
  #endif
  #if __has_include(<cuComplex.h>)
    #include <cuComplex.h>
    #define HALAB200_HAS_CUCOMPLEX
  #endif
#endif
#ifndef HALAB200_HAS_CUDART_HEADERS
typedef struct CUstream_st *cudaStream_t;   // only the handle type is needed by gpu_engine::set_stream
#endif

#ifndef __HALA_CUDA_API_VERSION__
  #ifdef CUDART_VERSION
    #define __HALA_CUDA_API_VERSION__ CUDART_VERSION
  #else
    #define __HALA_CUDA_API_VERSION__ 12090
  #endif
#endif

namespace hala{

#ifdef HALAB200_HAS_CUCOMPLEX
//! ABI-compatible CUDA complex types are accepted wherever std::complex is (reference :64-70).
template<> struct is_fcomplex<cuComplex> : std::true_type{};
template<> struct is_dcomplex<cuDoubleComplex> : std::true_type{};
#endif

//! Non-zero libhalab200 status -> std::runtime_error carrying the library's message and the call site (reference check_cuda, :78-152).
inline void check_hb(int status, const char *function_name){
    if (status != HB_OK)
        throw std::runtime_error(std::string(function_name) + " failed with message: " + hb_last_error());
}

//! dtype code of the C ABI for a HALA scalar type (the reference's 4-way cuda_call_backend dispatch, :159-184).
template<typename T> constexpr int hb_type(){
    static_assert(is_float<T>::value || is_double<T>::value || is_fcomplex<T>::value || is_dcomplex<T>::value,
                  "the B200 backend works with float, double and their complex counterparts");
    return is_float<T>::value ? HB_F32 : (is_double<T>::value ? HB_F64 : (is_fcomplex<T>::value ? HB_C32 : HB_C64));
}

inline int gpu_device_count(){
    int count = 0;
    hb_device_count(&count);
    return count;
}

//! 'T' stays 'T' for complex data and 'C' degenerates to 'T' for real data (reference trans_to_cuda_sparse, :214-222).
template<typename scalar_type> inline char trans_to_hb(char trans){
    if (is_n(trans)) return 'N';
    return (is_complex<scalar_type>::value && is_c(trans)) ? 'C' : 'T';
}

template<typename T> T* gpu_allocate(int gpu_device, size_t num_elements){
    void *p = nullptr;
    check_hb(hb_dev_malloc(gpu_device, num_elements * sizeof(T), &p), "hala::gpu_allocate()");
    return reinterpret_cast<T*>(p);
}
template<typename T> void gpu_free(T *gpu_data){
    if (gpu_data != nullptr) check_hb(hb_dev_free(const_cast<void*>(reinterpret_cast<void const*>(gpu_data))), "hala::gpu_free()");
}
//! Host-synchronous copy in any of the three directions (reference gpu_copy_n, :322-330).
template<copy_direction dir, typename T> void gpu_copy_n(T const *source, size_t num_entries, T *destination){
    constexpr int kind = (dir == copy_direction::host2device) ? HB_H2D : ((dir == copy_direction::device2host) ? HB_D2H : HB_D2D);
    check_hb(hb_dev_memcpy(destination, source, num_entries * sizeof(T), kind), "hala::gpu_copy_n()");
}

}
#endif
