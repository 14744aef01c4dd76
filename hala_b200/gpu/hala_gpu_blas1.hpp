#ifndef HALAB200_GPU_BLAS1_HPP
#define HALAB200_GPU_BLAS1_HPP
// BLAS-1 on the B200 backend: vcopy, norm2, dot<conj>, axpy, scal through hb_copy/hb_nrm2/hb_dot/hb_axpy/hb_scal
// (reference gpu/hala_gpu_blas1.hpp:48-65, 102-121, 178-198, 204-222, 228-245: cublas{S,D,C,Z}{copy,nrm2,dot*,axpy,scal}).
// Same argument meaning: (N, x, incx[, y, incy]); scalars by value (host) or by pointer (host/device by pointer mode);
// results of dot / norm2 are returned by value on the host (host-synchronous), as the reference does.
#include "hala_gpu_blas0.hpp"

namespace hala{

template<class VectorLikeX, class VectorLikeY>
inline void vcopy(gpu_engine const &engine, int N, VectorLikeX const &x, int incx, VectorLikeY &&y, int incy){
    check_types(x, y);
    engine.check_gpu(x, y);
    check_set_size(assume_output, y, 1 + (N - 1) * incy);
    assert( valid::vcopy(N, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeX>;
    check_hb(hb_copy(engine, hb_type<scalar_type>(), N, get_data(x), incx, get_data(y), incy), "hala::vcopy(gpu_engine)");
}
template<class VectorLikeX, class VectorLikeY>
inline void vcopy(gpu_engine const &engine, VectorLikeX const &x, VectorLikeY &&y, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    vcopy(engine, N, x, incx, y, incy);
}

template<class VectorLike> auto norm2(gpu_engine const &engine, int N, VectorLike const &x, int incx){
    check_types(x);
    engine.check_gpu(x);
    assert( valid::norm2(N, x, incx) );
    using scalar_type    = get_scalar_type<VectorLike>;
    using precision_type = get_precision_type<VectorLike>;
    precision_type cpu_result = get_cast<precision_type>(0.0);
    check_hb(hb_nrm2(engine, hb_type<scalar_type>(), N, get_data(x), incx, &cpu_result), "hala::norm2(gpu_engine)");
    return cpu_result;
}

template<bool conjugate = true, class VectorLikeX, class VectorLikeY>
auto dot(gpu_engine const &engine, int N, VectorLikeX const &x, int incx, VectorLikeY const &y, int incy){
    check_types(x, y);
    engine.check_gpu(x, y);
    assert( valid::dot(N, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeX>;
    scalar_type cpu_result = get_cast<scalar_type>(0.0);
    check_hb(hb_dot(engine, hb_type<scalar_type>(), conjugate ? 1 : 0, N, get_data(x), incx, get_data(y), incy, &cpu_result),
             "hala::dot(gpu_engine)");
    return cpu_result;
}

template<typename FS, class VectorLikeX, class VectorLikeY>
void axpy(gpu_engine const &engine, int N, FS alpha, VectorLikeX const &x, int incx, VectorLikeY &&y, int incy){
    check_types(x, y);
    engine.check_gpu(x, y);
    assert( valid::axpy(N, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeX>;
    hb_scalar<scalar_type, FS> a(alpha);
    check_hb(hb_axpy(engine, hb_type<scalar_type>(), N, a.get(), get_data(x), incx, get_data(y), incy), "hala::axpy(gpu_engine)");
}

template<typename FS, class VectorLike>
void scal(gpu_engine const &engine, int N, FS alpha, VectorLike &&x, int incx){
    check_types(x);
    engine.check_gpu(x);
    assert( valid::scal(N, x, incx) );
    using scalar_type = get_scalar_type<VectorLike>;
    hb_scalar<scalar_type, FS> a(alpha);
    check_hb(hb_scal(engine, hb_type<scalar_type>(), N, a.get(), get_data(x), incx), "hala::scal(gpu_engine)");
}

// vswap (reference :83-100) -> hb_swap
template<class VectorLikeX, class VectorLikeY>
inline void vswap(gpu_engine const &engine, int N, VectorLikeX &&x, int incx, VectorLikeY &&y, int incy){
    check_types(x, y);
    engine.check_gpu(x, y);
    assert( valid::vswap(N, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeX>;
    check_hb(hb_swap(engine, hb_type<scalar_type>(), N, get_data(x), incx, get_data(y), incy), "hala::vswap(gpu_engine)");
}
template<class VectorLikeX> inline auto asum(gpu_engine const &engine, int N, VectorLikeX const &x, int incx){
    check_types(x);
    engine.check_gpu(x);
    assert( valid::norm2(N, x, incx) );     // same shape rules as norm2, as in the reference
    using scalar_type    = get_scalar_type<VectorLikeX>;
    using precision_type = get_precision_type<VectorLikeX>;
    precision_type cpu_result = get_cast<precision_type>(0.0);
    check_hb(hb_asum(engine, hb_type<scalar_type>(), N, get_data(x), incx, &cpu_result), "hala::asum(gpu_engine)");
    return cpu_result;
}
// iamax (reference :153-172): cublas' 1-based index minus one
template<class VectorLikeX> inline int iamax(gpu_engine const &engine, int N, VectorLikeX const &x, int incx){
    check_types(x);
    engine.check_gpu(x);
    assert( valid::norm2(N, x, incx) );     // inputs are the same
    using scalar_type = get_scalar_type<VectorLikeX>;
    int cpu_result = 0;
    check_hb(hb_iamax(engine, hb_type<scalar_type>(), N, get_data(x), incx, &cpu_result), "hala::iamax(gpu_engine)");
    return cpu_result - 1;
}
// rotg (reference :251-262): scalars are host or device pointers according to the engine's pointer mode
template<typename T>
void rotg(gpu_engine const &engine, T *SA, T *SB, typename define_standard_precision<T>::value_type *C, T *S){
    check_types(std::vector<T>());
    check_hb(hb_rotg(engine, hb_type<T>(), SA, SB, C, S), "hala::rotg(gpu_engine)");
}
// rot (reference :269-300): C is of the precision type; S is a scalar of the vector type, or real for complex vectors (csrot / zdrot)
template<typename FC, typename FS, class VectorLikeX, class VectorLikeY>
void rot(gpu_engine const &engine, int N, VectorLikeX &x, int incx, VectorLikeY &&y, int incy, FC C, FS S){
    check_types(x, y);
    engine.check_gpu(x, y);
    assert( valid::rot(N, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeX>;
    using precision_type = get_precision_type<VectorLikeX>;
    constexpr bool mixed = (is_fcomplex<scalar_type>::value || is_dcomplex<scalar_type>::value) &&
                           (std::is_same<FS, int>::value || is_float<FS>::value || is_double<FS>::value);
    hb_scalar<precision_type, FC> effc(C);
    hb_scalar<typename std::conditional<mixed, precision_type, scalar_type>::type, FS> effs(S);
    check_hb(hb_rot(engine, hb_type<scalar_type>(), N, get_data(x), incx, get_data(y), incy, effc.get(), effs.get(), mixed ? 1 : 0),
             "hala::rot(gpu_engine)");
}
// rotmg / rotm (reference :318-371): real types only; param follows the pointer mode (a host vector in the default mode)
template<typename T, class VectorLike>
void rotmg(gpu_engine const &engine, T &D1, T &D2, T &X, T const &Y, VectorLike &&param){
    check_types(param);
    static_assert(is_float<T>::value || is_double<T>::value, "Givens rotations work only with real numbers.");
    using standard_type = get_standard_type<VectorLike>;
    static_assert(is_compatible<T, standard_type>::value, "rotmg() requires that the types of all inputs (vector and scalars) match");
    assert( check_size(param, 5) );
    check_hb(hb_rotmg(engine, hb_type<T>(), &D1, &D2, &X, &Y, get_data(param)), "hala::rotmg(gpu_engine)");
}
template<class VectorLikeX, class VectorLikeY, class VectorLikeP>
void rotm(gpu_engine const &engine, int N, VectorLikeX &&x, int incx, VectorLikeY &&y, int incy, VectorLikeP const &param){
    check_types(x, y, param);
    engine.check_gpu(x, y);
    using scalar_type = get_scalar_type<VectorLikeX>;
    static_assert(is_float<scalar_type>::value || is_double<scalar_type>::value, "Givens rotations work only with real numbers.");
    assert( valid::rot(N, x, incx, y, incy) );
    assert( check_size(param, 5) );
    check_hb(hb_rotm(engine, hb_type<scalar_type>(), N, get_data(x), incx, get_data(y), incy, get_data(param)), "hala::rotm(gpu_engine)");
}

}
#endif
