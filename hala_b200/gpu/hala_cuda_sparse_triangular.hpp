#ifndef HALAB200_GPU_SPARSE_TRIANGULAR_HPP
#define HALAB200_GPU_SPARSE_TRIANGULAR_HPP
// gpu_triangular_matrix: non-owning view of a CSR of which one triangle is used, with the dependency analysis of libhalab200
// (hb_tri) in place of the cusparseSpSV / cusparseSpSM descriptors of the reference (gpu/hala_cuda_sparse_triangular.hpp:38-454).
// trsv -> hb_sptrsv, trsm -> hb_sptrsm; no work buffers (the *_buffer_size calls return 0 and a buffer argument is accepted
// and ignored, as in the reference's CUDA >= 11.7 branch, :411-454).
#include "hala_cuda_sparse_general.hpp"

namespace hala{

template<typename T>
class gpu_triangular_matrix{
public:
    using value_type = std::remove_cv_t<T>;
    using engine_type = gpu_engine;

    template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
    gpu_triangular_matrix(gpu_engine const &cengine, char uplo, char diag,
                          VectorLikeP const &pn, VectorLikeI const &in, VectorLikeV const &va, char)
        : rengine(cengine), rpntr(get_data(pn)), rindx(get_data(in)), rvals(get_standard_data(va)),
          nrows(get_size_int(pn) - 1), nz(get_size_int(in)), handle(nullptr){
        check_types(va);
        check_types_int(pn, in);
        assert( get_size(in) == get_size(va) );
        assert( check_uplo(uplo) );
        assert( check_diag(diag) );
        cengine.check_gpu(pn, in, va);
        check_hb(hb_tri_create(rengine, hb_type<value_type>(), uplo, diag, nrows, nz, rpntr, rindx, rvals, &handle), "hala::gpu_triangular_matrix()");
    }
    ~gpu_triangular_matrix(){ if (handle) hb_tri_destroy(handle); }

    gpu_triangular_matrix(gpu_triangular_matrix const&) = delete;
    gpu_triangular_matrix& operator = (gpu_triangular_matrix const&) = delete;
    gpu_triangular_matrix(gpu_triangular_matrix &&other)
        : rengine(other.rengine), rpntr(other.rpntr), rindx(other.rindx), rvals(other.rvals), nrows(other.nrows), nz(other.nz),
          handle(std::exchange(other.handle, nullptr)){}
    gpu_triangular_matrix& operator = (gpu_triangular_matrix &&other){
        if (this != &other){
            if (handle) hb_tri_destroy(handle);
            rpntr = other.rpntr; rindx = other.rindx; rvals = other.rvals; nrows = other.nrows; nz = other.nz;
            handle = std::exchange(other.handle, nullptr);
        }
        return *this;
    }

    int const* pntr() const{ return rpntr; }
    int const* indx() const{ return rindx; }
    auto vals() const{ return rvals; }
    gpu_engine const& engine() const{ return rengine; }
    int rows() const{ return nrows; }
    int nnz() const{ return nz; }
    void* policy() const{ return nullptr; }

    template<typename FPA, class VectorLikeB, class VectorLikeX>
    size_t trsv_buffer_size(char, FPA, VectorLikeB const&, VectorLikeX &) const{ return 0; }
    template<typename FPA, class VectorLikeB, class VectorLikeX, class VectorLikeT>
    void trsv(char trans, FPA alpha, VectorLikeB const &b, VectorLikeX &&x, VectorLikeT &&) const{
        check_types(b, x);
        rengine.check_gpu(b, x);
        assert( valid::sparse_trsv(trans, *this, b) );
        check_set_size(assume_output, x, nrows);
        hb_scalar<value_type, FPA> a(alpha);
        check_hb(hb_sptrsv(rengine, handle, trans_to_hb<value_type>(trans), a.get(), get_data(b), 1, get_data(x), 1), "hala::gpu_triangular_matrix::trsv()");
    }
    template<typename FPA, class VectorLikeB, class VectorLikeX>
    void trsv(char trans, FPA alpha, VectorLikeB const &b, VectorLikeX &&x) const{ trsv(trans, alpha, b, x, 0); }

    template<typename FPA, class VectorLikeB>
    size_t trsm_buffer_size(char, char, int, FPA, VectorLikeB &&, int = -1) const{ return 0; }
    template<typename FPA, class VectorLikeB, class VectorLikeT>
    void trsm(char transa, char transb, int nrhs, FPA alpha, VectorLikeB &&B, int ldb, VectorLikeT &&) const{
        check_types(B);
        rengine.check_gpu(B);
        valid::default_ld(is_n(transb), nrows, nrhs, ldb);
        assert( valid::sparse_trsm(transa, transb, nrhs, *this, B, ldb) );
        hb_scalar<value_type, FPA> a(alpha);
        check_hb(hb_sptrsm(rengine, handle, trans_to_hb<value_type>(transa), is_n(transb) ? 'N' : 'T', nrhs, a.get(), get_data(B), ldb),
                 "hala::gpu_triangular_matrix::trsm()");
    }
    template<typename FPA, class VectorLikeB>
    void trsm(char transa, char transb, int nrhs, FPA alpha, VectorLikeB &&B, int ldb = -1) const{
        valid::default_ld(is_n(transb), nrows, nrhs, ldb);
        trsm(transa, transb, nrhs, alpha, B, ldb, 0);
    }

private:
    gpu_engine rengine;
    int const *rpntr, *rindx;
    value_type const *rvals;
    int nrows, nz;
    hb_tri *handle;
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_triangular_matrix(gpu_engine const &engine, char uplo, char diag, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, char policy = 'N'){
    check_types(vals);
    check_types_int(pntr, indx);
    using scalar_type = get_scalar_type<VectorLikeV>;
    return gpu_triangular_matrix<scalar_type>(engine, uplo, diag, pntr, indx, vals, policy);
}

}
#endif
