#ifndef HALAB200_GPU_SPARSE_GENERAL_HPP
#define HALAB200_GPU_SPARSE_GENERAL_HPP
// gpu_sparse_matrix: non-owning CSR view + the one-time analysis of libhalab200 (hb_csr), in place of the cusparseSpMatDescr_t
// wrapper of the reference (gpu/hala_cuda_sparse_general.hpp:191-375).  gemv -> hb_spmv; no per-call descriptors, no work
// buffer (gemv_buffer_size is 0; a buffer argument is accepted and ignored so reference call sites compile unchanged).
#include "hala_gpu_blas3.hpp"

namespace hala{

template<typename T>
struct gpu_sparse_matrix{
public:
    using value_type = std::remove_cv_t<T>;
    using engine_type = gpu_engine;

    template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
    gpu_sparse_matrix(gpu_engine const &engine, int num_rows, int num_cols, int num_nz,
                      VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals)
        : rengine(engine), rows(num_rows), cols(num_cols), nnz(num_nz), handle(nullptr){
        check_types(vals);
        check_types_int(pntr, indx);
        assert( check_size(pntr, rows+1) );
        assert( check_size(indx, nnz) );
        assert( check_size(vals, nnz) );
        engine.check_gpu(pntr, indx, vals);
        create(get_data(pntr), get_data(indx), get_standard_data(vals));
    }
    template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
    gpu_sparse_matrix(gpu_engine const &engine, int num_cols, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals)
        : rengine(engine), rows(get_size_int(pntr) - 1), cols(num_cols), nnz(get_size_int(indx)), handle(nullptr){
        check_types(vals);
        check_types_int(pntr, indx);
        assert( check_size(vals, nnz) );
        engine.check_gpu(pntr, indx, vals);
        create(get_data(pntr), get_data(indx), get_standard_data(vals));
    }
    ~gpu_sparse_matrix(){ if (handle) hb_csr_destroy(handle); }

    gpu_sparse_matrix(gpu_sparse_matrix const&) = delete;
    gpu_sparse_matrix& operator = (gpu_sparse_matrix const&) = delete;
    gpu_sparse_matrix(gpu_sparse_matrix &&other)
        : rengine(other.rengine), rows(other.rows), cols(other.cols), nnz(other.nnz), handle(std::exchange(other.handle, nullptr)){}
    gpu_sparse_matrix& operator = (gpu_sparse_matrix &&other){
        if (this != &other){
            if (handle) hb_csr_destroy(handle);
            rows = other.rows; cols = other.cols; nnz = other.nnz; handle = std::exchange(other.handle, nullptr);
        }
        return *this;
    }

    gpu_engine const& engine() const{ return rengine; }
    hb_csr* csr() const{ return handle; }
    //! extension: how op 'T'/'C' products use the cached CSR of the transpose (HB_TRANS_CHECKED default, HB_TRANS_FROZEN, HB_TRANS_SCATTER; halab200.h)
    void set_transpose_mode(int mode) const{ check_hb(hb_csr_set_transpose_mode(handle, mode), "hala::gpu_sparse_matrix::set_transpose_mode()"); }
    //! extension: tell a HB_TRANS_FROZEN matrix that the value array was rewritten
    void values_changed() const{ check_hb(hb_csr_values_changed(handle), "hala::gpu_sparse_matrix::values_changed()"); }

    template<typename FPa, class VectorLikeX, typename FPb, class VectorLikeY>
    size_t gemv_buffer_size(char trans, FPa, VectorLikeX const&, FPb beta, VectorLikeY &&y) const{
        pntr_check_set_size(beta, y, (is_n(trans)) ? rows : cols, 1);
        return 0;
    }
    template<typename FPa, class VectorLikeX, typename FPb, class VectorLikeY, class VectorLikeBuff>
    void gemv(char trans, FPa alpha, VectorLikeX const &x, FPb beta, VectorLikeY &&y, VectorLikeBuff &&) const{
        gemv(trans, alpha, x, beta, y);
    }
    template<typename FPa, class VectorLikeX, typename FPb, class VectorLikeY>
    void gemv(char trans, FPa alpha, VectorLikeX const &x, FPb beta, VectorLikeY &&y) const{
        check_types(x, y);
        static_assert(std::is_same<get_standard_type<VectorLikeX>, typename define_standard_type<value_type>::value_type>::value
                      || std::is_same<get_scalar_type<VectorLikeX>, value_type>::value, "vector type does not match the matrix");
        rengine.check_gpu(x, y);
        pntr_check_set_size(beta, y, (is_n(trans)) ? rows : cols, 1);
        hb_scalar<value_type, FPa> a(alpha);
        hb_scalar<value_type, FPb> b(beta);
        check_hb(hb_spmv(rengine, handle, trans_to_hb<value_type>(trans), a.get(), get_data(x), b.get(), get_data(y)), "hala::gpu_sparse_matrix::gemv()");
    }
    // sparse matrix - dense matrix product (reference :284-332, cusparseSpMM) -> hb_spmm; no work buffer
    template<typename FSA, class VectorLikeB, typename FSB, class VectorLikeC>
    size_t gemm_buffer_size(char transa, char transb, int b_rows, int b_cols, FSA, VectorLikeB const &B, int, FSB beta, VectorLikeC &C, int ldc) const{
        check_types(B, C);
        rengine.check_gpu(B, C);
        int N = (is_n(transb)) ? b_cols : b_rows;
        (void) transa;
        pntr_check_set_size(beta, C, ldc, N);
        return 0;
    }
    template<typename FSA, class VectorLikeB, typename FSB, class VectorLikeC, class VectorLikeT>
    void gemm(char transa, char transb, int b_rows, int b_cols, FSA alpha, VectorLikeB const &B, int ldb, FSB beta, VectorLikeC &C, int ldc, VectorLikeT &&) const{
        gemm(transa, transb, b_rows, b_cols, alpha, B, ldb, beta, C, ldc);
    }
    template<typename FSA, class VectorLikeB, typename FSB, class VectorLikeC>
    void gemm(char transa, char transb, int b_rows, int b_cols, FSA alpha, VectorLikeB const &B, int ldb, FSB beta, VectorLikeC &C, int ldc) const{
        check_types(B, C);
        rengine.check_gpu(B, C);
        int N = (is_n(transb)) ? b_cols : b_rows;
        pntr_check_set_size(beta, C, ldc, N);
        hb_scalar<value_type, FSA> a(alpha);
        hb_scalar<value_type, FSB> b(beta);
        check_hb(hb_spmm(rengine, handle, trans_to_hb<value_type>(transa), trans_to_hb<value_type>(transb), b_rows, b_cols, a.get(), get_data(B), ldb,
                         b.get(), get_data(C), ldc), "hala::gpu_sparse_matrix::gemm()");
    }

private:
    void create(int const *p, int const *i, void const *v){
        check_hb(hb_csr_create(rengine, hb_type<value_type>(), rows, cols, nnz, p, i, v, &handle), "hala::gpu_sparse_matrix()");
    }
    gpu_engine rengine;     // aliasing copy, as the reference holds (:369)
    int rows, cols, nnz;
    hb_csr *handle;
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_sparse_matrix(gpu_engine const &engine, int num_rows, int num_cols, int num_nz,
                        VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals){
    check_types(vals);
    check_types_int(pntr, indx);
    using scalar_type = get_scalar_type<VectorLikeV>;
    return gpu_sparse_matrix<scalar_type>(engine, num_rows, num_cols, num_nz, pntr, indx, vals);
}
template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_sparse_matrix(gpu_engine const &engine, int num_cols, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals){
    check_types(vals);
    check_types_int(pntr, indx);
    using scalar_type = get_scalar_type<VectorLikeV>;
    return gpu_sparse_matrix<scalar_type>(engine, num_cols, pntr, indx, vals);
}

//! One-shot SpMV (reference :407-419): a temporary view per call; nnz is read from the size of indx.
template<typename FPa, class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeX, typename FPb, class VectorLikeY>
void sparse_gemv(gpu_engine const &engine, char trans, int M, int N,
                 FPa alpha, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, VectorLikeX const &x,
                 FPb beta, VectorLikeY &y){
    check_types(vals, x, y);
    check_types_int(pntr, indx);
    engine.check_gpu(pntr, indx, vals, x, y);
    pntr_check_set_size(beta, y, (is_n(trans)) ? M : N, 1);
    assert( valid::sparse_gemv(trans, M, N, 0, pntr, indx, vals, x, y) );
    auto matrix = make_sparse_matrix(engine, M, N, get_size_int(indx), pntr, indx, vals);
    matrix.set_transpose_mode(HB_TRANS_SCATTER);      // a view that lives for one product: building the transposed copy cannot pay off
    matrix.gemv(trans, alpha, x, beta, y);
}

//! One-shot SpMM (reference :444-457): a temporary view per call.
template<typename FSA, class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeB, typename FSB, class VectorLikeC>
void sparse_gemm(gpu_engine const &engine, char transa, char transb, int M, int N, int K,
                 FSA alpha, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals,
                 VectorLikeB const &B, int ldb, FSB beta, VectorLikeC &C, int ldc){
    check_types(vals, B, C);
    check_types_int(pntr, indx);
    engine.check_gpu(pntr, indx, vals, B, C);
    pntr_check_set_size(beta, C, ldc, N);
    int nnz = get_size_int(indx);
    assert( valid::sparse_gemm(transa, transb, M, N, K, nnz, pntr, indx, vals, B, ldb, C, ldc) );
    auto matrix = make_sparse_matrix(engine, (is_n(transa)) ? M : K, (is_n(transa)) ? K : M, nnz, pntr, indx, vals);
    if (N < 8) matrix.set_transpose_mode(HB_TRANS_SCATTER);     // one-product view: the transposed copy pays off only over many columns
    matrix.gemm(transa, transb, (is_n(transb)) ? K : N, (is_n(transb)) ? N : K, alpha, B, ldb, beta, C, ldc);
}

}
#endif
