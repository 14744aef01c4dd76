#ifndef HALAB200_GPU_WRAP_ARRAY_HPP
#define HALAB200_GPU_WRAP_ARRAY_HPP
// gpu_wrapped_array: non-owning view of a device array (reference gpu/hala_gpu_wrap_array.hpp:57-131).
#include "hala_cuda_common.hpp"

namespace hala{

template<typename ScalarType>
class gpu_wrapped_array{
public:
    using value_type = std::remove_const_t<ScalarType>;

    gpu_wrapped_array(ScalarType *arr, size_t num_entries) : ptr(arr), count(num_entries){ check_gpu_type<value_type>(); }
    gpu_wrapped_array(gpu_wrapped_array const&) = delete;
    void operator =(gpu_wrapped_array const&) = delete;
    gpu_wrapped_array(gpu_wrapped_array &&other) : ptr(std::exchange(other.ptr, nullptr)), count(std::exchange(other.count, 0)){}
    void operator =(gpu_wrapped_array &&other){ ptr = std::exchange(other.ptr, nullptr); count = std::exchange(other.count, 0); }
    ~gpu_wrapped_array() = default;     // never frees

    size_t size() const{ return count; }
    ScalarType* data(){ return ptr; }
    ScalarType const* data() const{ return ptr; }

    template<class VectorLike> void load(VectorLike const &cpu_data){
        static_assert(std::is_same<value_type, typename define_type<VectorLike>::value_type>::value, "type mismatch in gpu_wrapped_array::load()");
        assert(count == get_size(cpu_data));
        gpu_copy_n<copy_direction::host2device>(get_data(cpu_data), count, ptr);
    }
    template<class VectorLike> void unload(VectorLike &cpu_data) const{
        static_assert(std::is_same<value_type, typename define_type<VectorLike>::value_type>::value, "type mismatch in gpu_wrapped_array::unload()");
        check_set_size(assume_output, cpu_data, count);
        gpu_copy_n<copy_direction::device2host>(static_cast<value_type const*>(ptr), count, get_data(cpu_data));
    }
    std::vector<value_type> unload() const{ std::vector<value_type> out(count); unload(out); return out; }
    std::valarray<value_type> unload_valarray() const{ std::valarray<value_type> out(count); unload(out); return out; }

private:
    ScalarType *ptr;
    size_t count;
};

template<typename ArrayType>
gpu_wrapped_array<ArrayType> wrap_gpu_array(ArrayType arr[], size_t num_entries){ return gpu_wrapped_array<ArrayType>(arr, num_entries); }

}
#endif
