#ifndef HALAB200_GPU_BLAS0_HPP
#define HALAB200_GPU_BLAS0_HPP
// geam / dgmm / tr2tp / tp2tr (reference gpu/hala_gpu_blas0.hpp) are dense helpers of the batch solvers — outside the hot path
// (SURVEY.md §8 rows f2/f4).  They are declared so that the wax templates that name them still parse; instantiating one is a
// compile-time error with a pointer to the scope table.
#include "hala_gpu_engine.hpp"

namespace hala{

template<typename T> struct hb_not_on_hot_path : std::false_type{};
#define HALAB200_OUT_OF_SCOPE(T, what) static_assert(hb_not_on_hot_path<T>::value, what " is not part of the B200 hot path yet (SURVEY.md §8 f-rows); use the reference gpu/ layer for it")

template<typename FPa, class VectorLikeA, typename FPb, class VectorLikeB, class VectorLikeC>
inline void geam(gpu_engine const&, char, char, int, int, FPa, VectorLikeA const&, int, FPb, VectorLikeB const&, int, VectorLikeC&&, int){
    HALAB200_OUT_OF_SCOPE(FPa, "hala::geam(gpu_engine)");
}
template<class VectorLikeA, class VectorLikeB, class VectorLikeC>
inline void dgmm(gpu_engine const&, char, int, int, VectorLikeA const&, int, VectorLikeB const&, int, VectorLikeC&&, int){
    HALAB200_OUT_OF_SCOPE(VectorLikeA, "hala::dgmm(gpu_engine)");
}
template<class VectorLikeA, class VectorLikeAP>
void tr2tp(gpu_engine const&, char, int, VectorLikeA const&, int, VectorLikeAP&&){ HALAB200_OUT_OF_SCOPE(VectorLikeA, "hala::tr2tp(gpu_engine)"); }
template<class VectorLikeAP, class VectorLikeA>
void tp2tr(gpu_engine const&, char, int, VectorLikeAP const&, VectorLikeA&&, int = -1){ HALAB200_OUT_OF_SCOPE(VectorLikeA, "hala::tp2tr(gpu_engine)"); }

}
#endif
