#ifndef HALAB200_GPU_ILU_HPP
#define HALAB200_GPU_ILU_HPP
// gpu_ilu (reference gpu/hala_gpu_ilu.hpp, cusparse?csrilu02 + two triangular solves) — SURVEY.md §8 row f1 ("next").
#include "hala_cuda_sparse_triangular.hpp"

namespace hala{

template<typename T>
class gpu_ilu{
public:
    using value_type = std::remove_cv_t<T>;
    using engine_type = gpu_engine;
    template<class... Args> gpu_ilu(gpu_engine const &e, Args&&...) : rengine(e){ HALAB200_OUT_OF_SCOPE(T, "hala::gpu_ilu"); }
    gpu_ilu(gpu_ilu const&) = delete;
    gpu_ilu(gpu_ilu &&) = default;
    gpu_engine const& engine() const{ return rengine; }
    template<class... Args> size_t buffer_size(Args&&...) const{ return 0; }
    template<class... Args> void apply(Args&&...) const{}
private:
    gpu_engine rengine;
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_ilu(gpu_engine const &engine, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, char policy){
    using scalar_type = get_scalar_type<VectorLikeV>;
    return gpu_ilu<scalar_type>(engine, pntr, indx, vals, policy);
}

}
#endif
