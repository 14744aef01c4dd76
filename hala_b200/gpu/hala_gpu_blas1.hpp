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

// vswap / iamax / rot* (reference :83-100, 145-172, 269-373): Givens rotations run on the host in the solvers; the rest is row f4.
template<class VectorLikeX, class VectorLikeY>
inline void vswap(gpu_engine const&, int, VectorLikeX&&, int, VectorLikeY&&, int){ HALAB200_OUT_OF_SCOPE(VectorLikeX, "hala::vswap(gpu_engine)"); }
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
template<class VectorLikeX> inline int iamax(gpu_engine const&, int, VectorLikeX const&, int){
    HALAB200_OUT_OF_SCOPE(VectorLikeX, "hala::iamax(gpu_engine)");
    return 0;
}
template<typename FC, typename FS, class VectorLikeX, class VectorLikeY>
void rot(gpu_engine const&, int, VectorLikeX&, int, VectorLikeY&&, int, FC, FS){ HALAB200_OUT_OF_SCOPE(FC, "hala::rot(gpu_engine)"); }
template<class VectorLikeX, class VectorLikeY, class VectorLikeP>
void rotm(gpu_engine const&, int, VectorLikeX&, int, VectorLikeY&&, int, VectorLikeP const&){ HALAB200_OUT_OF_SCOPE(VectorLikeX, "hala::rotm(gpu_engine)"); }

}
#endif
