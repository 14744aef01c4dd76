#ifndef HALAB200_GPU_ILU_HPP
#define HALAB200_GPU_ILU_HPP
// gpu_ilu: ILU(0) preconditioner on the B200 backend (reference gpu/hala_gpu_ilu.hpp:45-199: cusparse?csrilu02 + two triangular
// matrices over the same factor array).  Factorisation -> hb_ilu0; apply = unit-lower solve, then upper solve (hb_sptrsv / hb_sptrsm).
#include "hala_cuda_sparse_triangular.hpp"

namespace hala{

template<typename T> struct gpu_ilu{
    using value_type = typename define_standard_type<T>::value_type;
    using engine_type = gpu_engine;

    template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
    gpu_ilu(gpu_engine const &cengine, VectorLikeP const &mpntr, VectorLikeI const &mindx, VectorLikeV const &mvals, char policy)
        : rengine(cengine), ilu(rengine.device()), num_rows(get_size_int(mpntr)-1), nnz(get_size_int(mindx)){
        check_types(mvals);
        check_types_int(mpntr, mindx);
        cengine.check_gpu(mpntr, mindx, mvals);
        ilu.resize(static_cast<size_t>(nnz));
        check_hb(hb_ilu0(rengine, hb_type<value_type>(), num_rows, nnz, get_data(mpntr), get_data(mindx), get_data(mvals), ilu.data()),
                 "hala::gpu_ilu()");
        upper = std::make_unique<gpu_triangular_matrix<value_type>>(rengine, 'U', 'N', mpntr, mindx, ilu, policy);
        lower = std::make_unique<gpu_triangular_matrix<value_type>>(rengine, 'L', 'U', mpntr, mindx, ilu, policy);
    }
    ~gpu_ilu() = default;
    gpu_ilu(gpu_ilu const &other) = delete;
    gpu_ilu& operator = (gpu_ilu const &other) = delete;
    gpu_ilu& operator = (gpu_ilu &&other) = default;
    gpu_ilu(gpu_ilu &&other) = default;

    template<class VectorLikeX, class VectorLikeR>
    size_t buffer_size(VectorLikeX const&, VectorLikeR &&, int) const{ return 0; }
    template<class VectorLikeX, class VectorLikeR>
    gpu_vector<value_type> get_temp_buffer(VectorLikeX const&, VectorLikeR &&, int) const{ return gpu_vector<value_type>(rengine.device()); }

    template<class VectorLikeX, class VectorLikeR, class VectorLikeT>
    void apply(VectorLikeX const &x, VectorLikeR &&r, int num_rhs, VectorLikeT &&) const{
        check_types(x, r);
        check_set_size(assume_output, r, num_rhs, num_rows);
        gpu_pntr<host_pntr> hold(engine());
        if (num_rhs == 1){
            auto tmp = new_vector(engine(), r);
            lower->trsv('N', 1.0, x, tmp);
            upper->trsv('N', 1.0, tmp, r);
        }else{
            vcopy(engine(), x, r);
            lower->trsm('N', 'N', num_rhs, 1.0, r, num_rows);
            upper->trsm('N', 'N', num_rhs, 1.0, r, num_rows);
        }
    }
    template<class VectorLikeX, class VectorLikeR>
    void apply(VectorLikeX const &x, VectorLikeR &&r, int num_rhs = 1) const{ apply(x, r, num_rhs, 0); }

    gpu_engine const& engine() const{ return rengine; }

private:
    gpu_engine rengine;
    gpu_vector<value_type> ilu;
    int num_rows, nnz;
    std::unique_ptr<gpu_triangular_matrix<value_type>> upper, lower;
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV>
auto make_ilu(gpu_engine const &cengine, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, char policy){
    using standard_type = get_standard_type<VectorLikeV>;
    return gpu_ilu<standard_type>(cengine, pntr, indx, vals, policy);
}

}
#endif
