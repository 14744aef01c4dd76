#ifndef HALAB200_GPU_BLAS0_HPP
#define HALAB200_GPU_BLAS0_HPP
// geam / dgmm / tr2tp / tp2tr (reference gpu/hala_gpu_blas0.hpp) are dense helpers of the batch solvers — outside the hot path
// (SURVEY.md §8 rows f2/f4).  They are declared so that the wax templates that name them still parse; instantiating one is a
// compile-time error with a pointer to the scope table.
#include "hala_gpu_engine.hpp"

namespace hala{

template<typename T> struct hb_not_on_hot_path : std::false_type{};
#define HALAB200_OUT_OF_SCOPE(T, what) static_assert(hb_not_on_hot_path<T>::value, what " is not part of the B200 hot path yet (SURVEY.md §8 f-rows); use the reference gpu/ layer for it")

// geam / dgmm (reference gpu/hala_gpu_blas0.hpp:46-103, cublas?geam / cublas?dgmm): the dense helpers the batch solvers are built from
template<typename FPa, class VectorLikeA, typename FPb, class VectorLikeB, class VectorLikeC>
inline void geam(gpu_engine const &engine, char transa, char transb, int M, int N, FPa alpha, VectorLikeA const &A, int lda,
                 FPb beta, VectorLikeB const &B, int ldb, VectorLikeC &&C, int ldc){
    check_types(A, B, C);
    engine.check_gpu(A, B, C);
    assert( check_trans(transa) );
    assert( check_trans(transb) );
    assert( lda >= (is_n(transa) ? M : N) );
    assert( ldb >= (is_n(transb) ? M : N) );
    assert( ldc >= M );
    assert( check_size(A, lda, is_n(transa) ? N : M) );
    assert( check_size(B, ldb, is_n(transb) ? N : M) );
    check_set_size(assume_output, C, ldc, N);
    using scalar_type = get_scalar_type<VectorLikeA>;
    hb_scalar<scalar_type, FPa> a(alpha);
    hb_scalar<scalar_type, FPb> b(beta);
    check_hb(hb_geam(engine, hb_type<scalar_type>(), transa, transb, M, N, a.get(), get_data(A), lda, b.get(), get_data(B), ldb, get_data(C), ldc),
             "hala::geam(gpu_engine)");
}
template<class VectorLikeA, class VectorLikeB, class VectorLikeC>
inline void dgmm(gpu_engine const &engine, char side, int M, int N, VectorLikeA const &A, int lda, VectorLikeB const &x, int incx, VectorLikeC &&C, int ldc){
    check_types(A, x, C);
    engine.check_gpu(A, x, C);
    assert( check_side(side) );
    assert( lda >= M );
    assert( ldc >= M );
    assert( incx > 0 );
    assert( check_size(A, lda, N) );
    check_set_size(assume_output, C, ldc, N);
    using scalar_type = get_scalar_type<VectorLikeA>;
    check_hb(hb_dgmm(engine, hb_type<scalar_type>(), side, M, N, get_data(A), lda, get_data(x), incx, get_data(C), ldc), "hala::dgmm(gpu_engine)");
}
template<class VectorLikeA, class VectorLikeAP>
void tr2tp(gpu_engine const&, char, int, VectorLikeA const&, int, VectorLikeAP&&){ HALAB200_OUT_OF_SCOPE(VectorLikeA, "hala::tr2tp(gpu_engine)"); }
template<class VectorLikeAP, class VectorLikeA>
void tp2tr(gpu_engine const&, char, int, VectorLikeAP const&, VectorLikeA&&, int = -1){ HALAB200_OUT_OF_SCOPE(VectorLikeA, "hala::tp2tr(gpu_engine)"); }

}
#endif
