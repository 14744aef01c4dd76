#ifndef HALAB200_GPU_VECTOR_HPP
#define HALAB200_GPU_VECTOR_HPP
// gpu_vector / binded_gpu_vector: owning device arrays (reference gpu/hala_gpu_vector.hpp:49-249).
// Same semantics: resize() discards contents, load/unload are host-synchronous, copy keeps the device id.
// One deliberate difference: copy-assignment on the same device copies `other` (the reference copies *this, :80-86).
#include "hala_gpu_wrap_array.hpp"

namespace hala{

template<typename T>
class gpu_vector{
public:
    using value_type = T;

    gpu_vector(int deviceid = 0) : gpu(deviceid), count(0), ptr(nullptr), owner(true){ check_gpu_type<T>(); }
    gpu_vector(size_t num_entries, int deviceid) : gpu(deviceid), count(0), ptr(nullptr), owner(true){ check_gpu_type<T>(); allocate(num_entries); }
    gpu_vector(gpu_vector<T> const &other) : gpu(other.gpu), count(0), ptr(nullptr), owner(true){
        allocate(other.count);
        gpu_copy_n<copy_direction::device2device>(static_cast<T const*>(other.ptr), count, ptr);
    }
    gpu_vector(gpu_vector<T> &&other)
        : gpu(other.gpu), count(std::exchange(other.count, 0)), ptr(std::exchange(other.ptr, nullptr)), owner(std::exchange(other.owner, true)){}
    ~gpu_vector(){ clear(); }

    //! extension (hala_gpu_solvers.hpp): a gpu_vector over device memory somebody else owns — the work arrays of the fused solvers as the
    //! caller's preconditioner sees them.  Same-size operations (vcopy into it, copy-assignment, ilu.apply) write through; resizing
    //! it to another size detaches it (the vector then owns fresh memory and the borrowed array is left alone).
    static gpu_vector<T> view(int deviceid, T *device_array, size_t num_entries){
        gpu_vector<T> v(deviceid);
        v.ptr = device_array; v.count = num_entries; v.owner = false;
        return v;
    }
    bool owns_memory() const{ return owner; }

    void clear(){ if (owner) gpu_free(ptr); ptr = nullptr; count = 0; owner = true; }

    void operator =(gpu_vector<T> const &other){
        if (this == &other) return;
        if (other.gpu == gpu){
            resize(other.count);
            gpu_copy_n<copy_direction::device2device>(static_cast<T const*>(other.ptr), count, ptr);
        }else{
            load(other.unload());       // different devices: through the host, as the reference does
        }
    }
    void operator =(gpu_vector<T> &&other){
        gpu_vector<T> tmp(std::move(other));
        std::swap(gpu, tmp.gpu); std::swap(count, tmp.count); std::swap(ptr, tmp.ptr); std::swap(owner, tmp.owner);
    }

    void resize(size_t new_size){
        if (new_size == count) return;
        if (owner) gpu_free(ptr);
        ptr = nullptr; owner = true;
        allocate(new_size);
    }
    template<class VectorLike> void load(VectorLike const &cpu_data){
        static_assert(std::is_same<T, typename define_type<VectorLike>::value_type>::value, "type mismatch in gpu_vector::load()");
        resize(get_size(cpu_data));
        gpu_copy_n<copy_direction::host2device>(get_data(cpu_data), count, ptr);
    }
    template<class VectorLike> void unload(VectorLike &cpu_data) const{
        static_assert(std::is_same<T, typename define_type<VectorLike>::value_type>::value, "type mismatch in gpu_vector::unload()");
        check_set_size(assume_output, cpu_data, count);
        gpu_copy_n<copy_direction::device2host>(static_cast<T const*>(ptr), count, get_data(cpu_data));
    }
    std::vector<T> unload() const{ std::vector<T> out(count); unload(out); return out; }
    std::valarray<T> unload_valarray() const{ std::valarray<T> out(count); unload(out); return out; }

    T* data(){ return ptr; }
    T const* data() const{ return ptr; }
    operator T *(){ return ptr; }
    operator T const *() const{ return ptr; }

    size_t size() const{ return count; }
    int device() const{ return gpu; }
    bool empty() const{ return count == 0; }

    //! One kernel (hb_fill) instead of the reference's log2(n) device-to-device copies (:147-156).
    void fill(T value){
        if (count == 0) return;
        check_hb(hb_dev_fill(gpu, fill_code(), count, &value, ptr), "hala::gpu_vector::fill()");
    }

protected:
    void allocate(size_t new_size){ count = new_size; ptr = (count > 0) ? gpu_allocate<T>(gpu, count) : nullptr; }
    static constexpr int fill_code(){
        return is_float<T>::value ? HB_F32 : (is_double<T>::value ? HB_F64 : (is_fcomplex<T>::value ? HB_C32 : (is_dcomplex<T>::value ? HB_C64 : -1)));
    }

private:
    int gpu;
    size_t count;
    T *ptr;
    bool owner;
};

//! extension: page-locked host storage for the containers that feed load() / unload() and mixed_engine calls.  The reference moves
//! std::vector data with synchronous cudaMemcpy from pageable memory (gpu/hala_gpu_vector.hpp:224-249, wax/hala_lib_extensions.hpp:
//! 126-241: every mixed_engine operation loads its operands and unloads its result); from a hala::pinned_vector the same calls run at
//! the full PCIe rate (B200, 8 MiB .. 1 GiB: ~55 GB/s against ~17 GB/s pageable) with no staging copy by the driver.
template<typename T> struct pinned_allocator{
    using value_type = T;
    pinned_allocator() = default;
    template<class U> pinned_allocator(pinned_allocator<U> const&){}
    T* allocate(size_t n){
        void *p = nullptr;
        check_hb(hb_host_alloc(n * sizeof(T), &p), "hala::pinned_allocator::allocate()");
        return static_cast<T*>(p);
    }
    void deallocate(T *p, size_t){ hb_host_free(p); }
    template<class U> bool operator ==(pinned_allocator<U> const&) const{ return true; }
    template<class U> bool operator !=(pinned_allocator<U> const&) const{ return false; }
};
template<typename T> using pinned_vector = std::vector<T, pinned_allocator<T>>;

template<class VectorLike>
inline auto make_gpu_vector(VectorLike const &cpu_data, int gpuid = 0){
    gpu_vector<typename define_type<VectorLike>::value_type> out(gpuid);
    out.load(cpu_data);
    return out;
}
template<typename T> inline auto make_gpu_vector(size_t num_entries, int gpuid = 0){ return gpu_vector<T>(num_entries, gpuid); }

template<typename T> struct deviceid_extractor<gpu_vector<T>>{ static int device(gpu_vector<T> const &x){ return x.device(); } };
template<typename T> struct deviceid_extractor<const gpu_vector<T>>{ static int device(gpu_vector<T> const &x){ return x.device(); } };

//! Loads a host container on construction and writes it back on destruction (reference :224-249).
template<typename T, class VectorLike> struct binded_gpu_vector{
    using value_type = T;
    binded_gpu_vector(int gpuid, VectorLike &v) : host(v), dev(gpuid){
        static_assert(!std::is_const<VectorLike>::value, "Cannot bind to a const vector!");
        dev.load(v);
    }
    ~binded_gpu_vector(){ dev.unload(host); }
    T* data(){ return dev.data(); }
    T const* data() const{ return dev.data(); }
    void resize(size_t new_size){ dev.resize(new_size); }
    size_t size() const{ return dev.size(); }
    int device() const{ return dev.device(); }
private:
    VectorLike &host;
    gpu_vector<T> dev;
};

}
#endif
