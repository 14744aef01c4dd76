#ifndef HALAB200_GPU_BLAS2_HPP
#define HALAB200_GPU_BLAS2_HPP
// gemv on the B200 backend (reference gpu/hala_gpu_blas2.hpp:39-62, cublas?gemv): the Gram-Schmidt pair of GMRES
// (hex/solvers/hala_solvers_gmres.hpp:67-77).  'T'/'C' on a tall-skinny column-major A runs as one fused multi-dot pass.
// The other BLAS-2 routines of the reference file are dense and not on the CG/GMRES path.
#include "hala_gpu_blas1.hpp"

namespace hala{

template<typename FPA, typename FPB, class VectorLikeA, class VectorLikeX, class VectorLikeY>
void gemv(gpu_engine const &engine, char trans, int M, int N, FPA alpha, VectorLikeA const &A, int lda,
          VectorLikeX const &x, int incx, FPB beta, VectorLikeY &&y, int incy){
    check_types(A, x, y);
    engine.check_gpu(A, x, y);
    pntr_check_set_size(beta, y, 1 + incy * ((is_n(trans) ? M : N) - 1), 1);
    assert( valid::gemv(trans, M, N, A, lda, x, incx, y, incy) );
    using scalar_type = get_scalar_type<VectorLikeA>;
    hb_scalar<scalar_type, FPA> a(alpha);
    hb_scalar<scalar_type, FPB> b(beta);
    const char op = is_n(trans) ? 'N' : (is_c(trans) ? 'C' : 'T');
    check_hb(hb_gemv(engine, hb_type<scalar_type>(), op, M, N, a.get(), get_data(A), lda, get_data(x), incx, b.get(), get_data(y), incy),
             "hala::gemv(gpu_engine)");
}

// tbsv (reference :368-388, cublas?tbsv): bandwidth 0 is the element-wise divide of hala::vdivide (wax/hala_blas_extensions.hpp:251-260)
template<class VectorLikeA, class VectorLikeX>
inline void tbsv(gpu_engine const &engine, char uplo, char trans, char diag, int N, int k, VectorLikeA const &A, int lda, VectorLikeX &&x, int incx){
    check_types(A, x);
    engine.check_gpu(A, x);
    assert( valid::tbmsv(uplo, trans, diag, N, k, A, lda, x, incx) );
    using scalar_type = get_scalar_type<VectorLikeA>;
    check_hb(hb_tbsv(engine, hb_type<scalar_type>(), uplo, trans, diag, N, k, get_data(A), lda, get_data(x), incx), "hala::tbsv(gpu_engine)");
}

}
#endif
