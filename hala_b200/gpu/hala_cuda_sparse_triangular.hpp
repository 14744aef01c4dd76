#ifndef HALAB200_GPU_SPARSE_TRIANGULAR_HPP
#define HALAB200_GPU_SPARSE_TRIANGULAR_HPP
// gpu_triangular_matrix (reference gpu/hala_cuda_sparse_triangular.hpp, cusparseSpSV/SpSM) — preconditioner machinery,
// SURVEY.md §8 row f1 ("next").  Declared so that wax/hala_lib_extensions.hpp:248-287 parses; construction is a compile-time error.
#include "hala_cuda_sparse_general.hpp"

namespace hala{

template<typename T>
struct gpu_triangular_matrix{
    using value_type = std::remove_cv_t<T>;
    using engine_type = gpu_engine;
    template<class... Args> gpu_triangular_matrix(gpu_engine const &e, Args&&...) : rengine(e){
        HALAB200_OUT_OF_SCOPE(T, "hala::gpu_triangular_matrix");
    }
    gpu_engine const& engine() const{ return rengine; }
    template<class... Args> size_t trsv_buffer_size(Args&&...) const{ return 0; }
    template<class... Args> void trsv(Args&&...) const{}
    template<class... Args> size_t trsm_buffer_size(Args&&...) const{ return 0; }
    template<class... Args> void trsm(Args&&...) const{}
private:
    gpu_engine rengine;
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_triangular_matrix(gpu_engine const &engine, char uplo, char diag, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, char policy = 'N'){
    using scalar_type = get_scalar_type<VectorLikeV>;
    return gpu_triangular_matrix<scalar_type>(engine, uplo, diag, pntr, indx, vals, policy);
}

}
#endif
