#ifndef HALAB200_GPU_OVERLOADS_HPP
#define HALAB200_GPU_OVERLOADS_HPP
// Default-argument overloads of the hot-path operations on gpu_engine (reference gpu/hala_gpu_overloads.hpp:69-128) and the
// engine's vcopy member.  N defaults to 1 + (size - 1) / incx (valid::default_size), lda to M (valid::default_ld).
#include "hala_gpu_plu.hpp"

namespace hala{

template<class vec>
auto gpu_engine::vcopy(vec const &x) const{
    auto y = new_vector(*this, x);
    const int xdevice = get_device(x);
    if ((xdevice > -1) && (xdevice != cgpu)){       // other device: through the host, as the reference does
        using standard_type = get_standard_type<vec>;
        std::vector<standard_type> cpuy(get_size(x));
        gpu_copy_n<copy_direction::device2host>(reinterpret_cast<standard_type const*>(get_data(x)), get_size(x), get_data(cpuy));
        y.load(cpuy);
    }else{
        hala::vcopy(*this, x, y);
    }
    return y;
}

template<class VectorLikeX> inline auto norm2(gpu_engine const &engine, VectorLikeX const &x, int incx = 1, int N = -1){
    valid::default_size(x, incx, N);
    return norm2(engine, N, x, incx);
}
template<class VectorLikeX> inline auto asum(gpu_engine const &engine, VectorLikeX const &x, int incx = 1, int N = -1){
    valid::default_size(x, incx, N);
    return asum(engine, N, x, incx);
}
template<bool conjugate = true, class VectorLikeX, class VectorLikeY>
auto dot(gpu_engine const &engine, VectorLikeX const &x, VectorLikeY const &y, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    return dot<conjugate>(engine, N, x, incx, y, incy);
}
template<class VectorLikeX, class VectorLikeY>
auto dotu(gpu_engine const &engine, VectorLikeX const &x, VectorLikeY const &y, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    return dot<false>(engine, N, x, incx, y, incy);
}
template<class VectorLikeX, class VectorLikeY>
auto dotu(gpu_engine const &engine, int N, VectorLikeX const &x, int incx, VectorLikeY const &y, int incy){
    return dot<false>(engine, N, x, incx, y, incy);
}
template<typename FP, class VectorLikeX, class VectorLikeY>
void axpy(gpu_engine const &engine, FP alpha, VectorLikeX const &x, VectorLikeY &&y, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    axpy(engine, N, alpha, x, incx, y, incy);
}
template<typename FP, class VectorLike>
void scal(gpu_engine const &engine, FP alpha, VectorLike &&x, int incx = 1, int N = -1){
    valid::default_size(x, incx, N);
    scal(engine, N, alpha, x, incx);
}
template<typename FPA, typename FPB, class VectorLikeA, class VectorLikeX, class VectorLikeY>
void gemv(gpu_engine const &engine, char trans, int M, int N,
          FPA alpha, VectorLikeA const &A, VectorLikeX const &x, FPB beta, VectorLikeY &&y, int lda = -1, int incx = 1, int incy = 1){
    valid::default_ld(M, lda);
    gemv(engine, trans, M, N, alpha, A, lda, x, incx, beta, y, incy);
}

// default-argument forms of the batch helpers (reference gpu/hala_gpu_overloads.hpp:31-46)
template<typename FPa, class VectorLikeA, typename FPb, class VectorLikeB, class VectorLikeC>
inline void geam(gpu_engine const &engine, char transa, char transb, int M, int N, FPa alpha, VectorLikeA const &A,
                 FPb beta, VectorLikeB const &B, VectorLikeC &&C, int lda = -1, int ldb = -1, int ldc = -1){
    valid::default_ld(is_n(transa), M, N, lda);
    valid::default_ld(is_n(transb), M, N, ldb);
    valid::default_ld(M, ldc);
    geam(engine, transa, transb, M, N, alpha, A, lda, beta, B, ldb, C, ldc);
}
template<class VectorLikeA, class VectorLikeB, class VectorLikeC>
inline void dgmm(gpu_engine const &engine, char side, int M, int N, VectorLikeA const &A,
                 VectorLikeB const &x, VectorLikeC &&C, int lda = -1, int incx = 1, int ldc = -1){
    valid::default_ld(M, lda);
    valid::default_ld(M, ldc);
    dgmm(engine, side, M, N, A, lda, x, incx, C, ldc);
}
template<class VectorLikeX, class VectorLikeY>
inline void vswap(gpu_engine const &engine, VectorLikeX &&x, VectorLikeY &&y, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    vswap(engine, N, x, incx, y, incy);
}
template<typename FC, typename FS, class VectorLikeX, class VectorLikeY>
void rot(gpu_engine const &engine, VectorLikeX &x, VectorLikeY &&y, FC C, FS S, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    rot(engine, N, x, incx, y, incy, C, S);
}
template<class VectorLikeX, class VectorLikeY, class VectorLikeP>
void rotm(gpu_engine const &engine, VectorLikeX &x, VectorLikeY &&y, VectorLikeP const &param, int incx = 1, int incy = 1, int N = -1){
    valid::default_size(x, incx, N);
    rotm(engine, N, x, incx, y, incy, param);
}
template<class VectorLikeX> inline int iamax(gpu_engine const &engine, VectorLikeX const &x, int incx = 1, int N = -1){
    valid::default_size(x, incx, N);
    return iamax(engine, N, x, incx);
}
// default-argument form of the one-shot SpMM (reference gpu/hala_gpu_overloads.hpp:390-397)
template<typename FSA, class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeB, typename FSB, class VectorLikeC>
inline void sparse_gemm(gpu_engine const &engine, char transa, char transb, int M, int N, int K,
                        FSA alpha, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals,
                        VectorLikeB const &B, FSB beta, VectorLikeC &&C, int ldb = -1, int ldc = -1){
    valid::default_ld(is_n(transb), K, N, ldb);
    valid::default_ld(M, ldc);
    sparse_gemm(engine, transa, transb, M, N, K, alpha, pntr, indx, vals, B, ldb, beta, C, ldc);
}
template<class VectorLikeA, class VectorLikeX>
inline void tbsv(gpu_engine const &engine, char uplo, char trans, char diag, int N, int k, const VectorLikeA &A, VectorLikeX &&x, int lda = -1, int incx = 1){
    valid::default_ld(k+1, lda);
    tbsv(engine, uplo, trans, diag, N, k, A, lda, x, incx);
}

}
#endif
