#ifndef HALAB200_GPU_ENGINE_HPP
#define HALAB200_GPU_ENGINE_HPP
// gpu_engine: one libhalab200 context (device id, stream, pointer mode, reduction scratch, solver arena) in place of the
// cuBLAS + cuSPARSE + cuSOLVER handle triple of the reference (gpu/hala_gpu_engine.hpp:47-307).  Same contract: a copy is a
// NON-owning alias of the same context (:80), one engine per host thread, not thread-safe.
#include "hala_gpu_vector.hpp"

namespace hala{

struct device_pntr{};
struct host_pntr{};

struct gpu_engine{
private:
    int cgpu;
    hb_ctx *ctx;
    bool owner;

    static hb_ctx* make_context(int deviceid){
        hb_ctx *c = nullptr;
        check_hb(hb_ctx_create(deviceid, &c), "hala::gpu_engine()");
        return c;
    }

public:
    gpu_engine(int deviceid = 0) : cgpu(deviceid), ctx(make_context(deviceid)), owner(true){}
    gpu_engine(cudaStream_t streamid, int deviceid) : gpu_engine(deviceid){ set_stream(streamid); }
    //! Alias of an existing context without owning it (the reference's ctor from external handles).
    gpu_engine(int deviceid, hb_ctx *extern_context) : cgpu(deviceid), ctx(extern_context), owner(false){}
    gpu_engine(gpu_engine const &other) : cgpu(other.cgpu), ctx(other.ctx), owner(false){}
    gpu_engine(gpu_engine &&other) : cgpu(other.cgpu), ctx(std::exchange(other.ctx, nullptr)), owner(std::exchange(other.owner, false)){}
    gpu_engine& operator =(gpu_engine &&other){
        if (this != &other){
            release();
            cgpu = other.cgpu; ctx = std::exchange(other.ctx, nullptr); owner = std::exchange(other.owner, false);
        }
        return *this;
    }
    gpu_engine& operator =(gpu_engine const&){ return *this; }      // as the reference: assigning an alias is a no-op
    ~gpu_engine(){ release(); }

    //! The context every hala::<op>(gpu_engine const&, ...) hands to the C ABI.
    hb_ctx* context() const{ return ctx; }
    operator hb_ctx* () const{ return ctx; }

    int device() const{ return cgpu; }
    void synchronize() const{ check_hb(hb_ctx_sync(ctx), "hala::gpu_engine::synchronize()"); }
    void set_stream(cudaStream_t streamid) const{ check_hb(hb_ctx_set_stream(ctx, (void*) streamid), "hala::gpu_engine::set_stream()"); }
    void set_active_device() const{ void *p = nullptr; hb_dev_malloc(cgpu, 0, &p); }   // cudaSetDevice side effect only

    //! Scalars given by POINTER are device pointers while this mode is on (reference :105-117, gpu_pntr below).
    void set_blas_device_pntr() const{ check_hb(hb_ctx_set_pointer_mode(ctx, HB_POINTER_DEVICE), "set_blas_device_pntr()"); }
    void reset_blas_device_pntr() const{ check_hb(hb_ctx_set_pointer_mode(ctx, HB_POINTER_HOST), "reset_blas_device_pntr()"); }
    void set_cusparse_device_pntr() const{ set_blas_device_pntr(); }     // one pointer mode for the whole context
    void reset_cusparse_device_pntr() const{ reset_blas_device_pntr(); }
    int get_blas_pointer_mode() const{
        int mode = HB_POINTER_HOST;
        check_hb(hb_ctx_get_pointer_mode(ctx, &mode), "get_blas_pointer_mode()");
        return mode;
    }

    template<class VectorLike> auto load(VectorLike const &cpu_data) const{ return make_gpu_vector(cpu_data, cgpu); }
    template<class VectorLike> auto load(VectorLike const &cpu_data, size_t size) const{
        assert( check_size(cpu_data, size) );
        using T = typename define_type<VectorLike>::value_type;
        gpu_vector<T> out(size, cgpu);
        gpu_copy_n<copy_direction::host2device>(static_cast<T const*>(get_data(cpu_data)), size, out.data());
        return out;
    }
    template<class VectorLike> auto unload(VectorLike const &gpu_data) const{
        using standard_type = get_standard_type<VectorLike>;
        cpu_engine e;
        auto result = new_vector(e, std::vector<standard_type>());
        force_size(get_size(gpu_data), result);
        gpu_copy_n<copy_direction::device2host>(reinterpret_cast<standard_type const*>(get_data(gpu_data)), get_size(gpu_data), get_data(result));
        return result;
    }
    template<typename T> auto vector(size_t num_entries, T value) const{
        gpu_vector<T> x(num_entries, cgpu);
        x.fill(value);
        return x;
    }
    template<class vec> auto vcopy(vec const &x) const;                 // defined in hala_gpu_overloads.hpp
    template<typename T> auto wrap_array(T *array, size_t num_entries) const{ return wrap_gpu_array(array, num_entries); }
    template<typename T> using dfvector = typename hala::gpu_vector<T>;

    template<class one_gpu_vec> void check_gpu(one_gpu_vec const &x) const{
        int vector_device = get_device(x);
        if (vector_device > -1) assert(cgpu == vector_device);
        (void) vector_device;
    }
    template<class first_gpu_vec, class second_gpu_vec, class... other_gpu_vecs>
    void check_gpu(first_gpu_vec const &x, second_gpu_vec const &y, other_gpu_vecs const &... other) const{
        check_gpu(x);
        check_gpu(y, other...);
    }

private:
    void release(){ if (owner && ctx) hb_ctx_destroy(ctx); ctx = nullptr; owner = false; }
};

template<typename T> struct vector_constructor<gpu_vector<T>>{
    static auto make_one(gpu_engine const &engine){ return hala::gpu_vector<T>(engine.device()); }
};

template<class VectorLike>
auto new_vector(gpu_engine const &engine, VectorLike const &){
    return vector_constructor<get_vdefault<gpu_engine, VectorLike>>::make_one(engine);
}

//! One kernel instead of the reference's doubling copies (:335-352).
template<class VectorLike>
void set_zero(gpu_engine const &engine, size_t num_entries, VectorLike &&x){
    if (num_entries == 0) return;
    check_set_size(assume_output, x, num_entries);
    using standard_type = get_standard_type<VectorLike>;
    check_hb(hb_memset_zero(engine, get_standard_data(x), num_entries * sizeof(standard_type)), "hala::set_zero()");
}

template<class VectorLike>
auto gpu_bind_vector(gpu_engine const &e, VectorLike &x){
    using standard_type = get_standard_type<VectorLike>;
    return binded_gpu_vector<standard_type, VectorLike>(e.device(), x);
}

template<typename ArrayType>
auto wrap_array(gpu_engine const &engine, ArrayType arr[], size_t num_entries){ return engine.wrap_array(arr, num_entries); }

//! RAII switch of the scalar pointer mode (reference gpu_pntr, :410-446).
template<typename pntr_mode>
struct gpu_pntr{
    explicit gpu_pntr(gpu_engine const &engine) : eng(engine), original_mode(engine.get_blas_pointer_mode()){
        static_assert(std::is_same<pntr_mode, device_pntr>::value || std::is_same<pntr_mode, host_pntr>::value,
                      "gpu_pntr can be used only with device_pntr and host_pntr pntr_mode types");
        if (std::is_same<pntr_mode, device_pntr>::value) eng.set_blas_device_pntr(); else eng.reset_blas_device_pntr();
    }
    ~gpu_pntr(){ hb_ctx_set_pointer_mode(eng, original_mode); }
private:
    gpu_engine const &eng;
    int original_mode;
};

//! Turns "value or pointer" scalars into the `const void*` the C ABI takes; the holder keeps a by-value scalar alive.
template<typename scalar_type, typename FS> struct hb_scalar{
    explicit hb_scalar(FS v) : value(get_cast<scalar_type>(v)){}
    const void* get() const{ return &value; }
    scalar_type value;
};
template<typename scalar_type, typename P> struct hb_scalar<scalar_type, P*>{
    explicit hb_scalar(P *v) : pntr(v){}
    const void* get() const{ return pntr; }
    P *pntr;
};

}
#endif
