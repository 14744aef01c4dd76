// hb_context.cu — context, error plumbing and device memory of libhalab200.
// Replaces: gpu_engine handle management (reference gpu/hala_gpu_engine.hpp:60-163), gpu_allocate/gpu_free/gpu_copy_n
// (gpu/hala_cuda_common.hpp:253-330), gpu_vector::fill (gpu/hala_gpu_vector.hpp:147-156), set_zero (gpu_engine.hpp:335-352).
// Device memory comes from the driver's stream-ordered pool (cudaMallocAsync / cudaFreeAsync with the release threshold lifted), so
// the load -> operation -> unload pattern of mixed_engine and binded_gpu_vector (wax/hala_lib_extensions.hpp:126-241,
// gpu/hala_gpu_vector.hpp:224-249) and every new_vector / resize stop paying a cudaMalloc plus a device-synchronising cudaFree per
// call: a freed block is handed out again without touching the OS.  HB_MEM_POOL=0 restores cudaMalloc/cudaFree.
#include "hb_common.cuh"
#include <algorithm>
#include <cstdlib>

static thread_local std::string g_last_error;

void hb_set_error(const std::string &msg){ g_last_error = msg; }
int hb_cuda_fail(cudaError_t e, const char *what){
    g_last_error = std::string(what) + " failed with message: " + cudaGetErrorString(e);
    return HB_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ pooled device memory
static constexpr int HB_MAX_DEVICES = 64;
static bool g_pool_ready[HB_MAX_DEVICES];
static bool g_custom_stream[HB_MAX_DEVICES];     // a context of this device runs on a caller-supplied stream
static bool pool_enabled(){
    static int v = -1;
    if (v < 0){ const char *e = getenv("HB_MEM_POOL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
static void pool_trim(int device){
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
}
// the current device is `device`
static cudaError_t pool_malloc(int device, void **ptr, size_t bytes, cudaStream_t stream){
    if (!pool_enabled() || device < 0 || device >= HB_MAX_DEVICES) return cudaMalloc(ptr, bytes);
    if (!g_pool_ready[device]){
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess){
            unsigned long long keep = ~0ull;                // never give freed blocks back to the OS on a synchronisation
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        g_pool_ready[device] = true;
    }
    cudaError_t e = cudaMallocAsync(ptr, bytes, stream);
    if (e != cudaSuccess){                                   // pool exhausted or unsupported: give the cache back and ask the plain allocator
        cudaGetLastError();
        cudaDeviceSynchronize();
        pool_trim(device);
        return cudaMalloc(ptr, bytes);
    }
    // an allocation is ordered on `stream`; contexts on other non-blocking streams may touch it at once, so make it visible to all
    if (g_custom_stream[device]) e = cudaStreamSynchronize(stream);
    return e;
}
static cudaError_t pool_free(void *ptr, cudaStream_t stream){
    if (!pool_enabled()) return cudaFree(ptr);
    cudaPointerAttributes attr;
    int cur = 0;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice || cudaGetDevice(&cur) != cudaSuccess){
        cudaGetLastError();
        return cudaFree(ptr);
    }
    const int dev = attr.device;
    if (dev != cur){ cudaSetDevice(dev); stream = nullptr; }
    // cudaFree synchronises the device, which is what protects a block still in use on another stream; keep that when a caller
    // brought its own stream, otherwise the legacy default stream orders the free after all work of the blocking streams
    if (dev >= 0 && dev < HB_MAX_DEVICES && g_custom_stream[dev]) cudaDeviceSynchronize();
    cudaError_t e = cudaFreeAsync(ptr, stream);
    if (e != cudaSuccess){ cudaGetLastError(); e = cudaFree(ptr); }
    if (dev != cur) cudaSetDevice(cur);
    return e;
}

template<typename T> __global__ void fill_kernel(size_t n, T value, T *x){
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) x[i] = value;
}

int hb_ctx_workspace(hb_ctx *ctx, size_t bytes, void **ptr){
    if (bytes > ctx->work_bytes){
        if (ctx->work){ HB_CUDA(cudaStreamSynchronize(ctx->stream)); HB_CUDA(cudaFree(ctx->work)); ctx->work = nullptr; ctx->work_bytes = 0; }
        cudaError_t e = cudaMalloc(&ctx->work, bytes);
        if (e != cudaSuccess){ hb_cuda_fail(e, "solver workspace cudaMalloc"); return HB_ERR_ALLOC; }
        ctx->work_bytes = bytes;
    }
    *ptr = ctx->work;
    return HB_OK;
}

// ---- per-kernel event timing of the solver loops (hb_ctx_profile)
void hb_prof_begin(hb_ctx *ctx){ if (ctx->prof && ctx->prof->on) ctx->prof->marked = 0; }
void hb_prof_mark(hb_ctx *ctx, long long it, int k){
    hb_prof *p = ctx->prof;
    if (!p || !p->on || it >= hb_prof::MAXIT || k > hb_prof::SLOTS) return;
    cudaEvent_t &e = p->ev[it * (hb_prof::SLOTS + 1) + k];
    if (!e && cudaEventCreate(&e) != cudaSuccess){ cudaGetLastError(); e = nullptr; return; }
    cudaEventRecord(e, ctx->stream);
    if (it + 1 > p->marked) p->marked = (int) (it + 1);
}
void hb_prof_collect(hb_ctx *ctx, long long valid_iterations, int slots){
    hb_prof *p = ctx->prof;
    if (!p || !p->on) return;
    const long long n = valid_iterations < p->marked ? valid_iterations : p->marked;
    for (long long it = 0; it < n; it++)
        for (int s = 0; s < slots && s < hb_prof::SLOTS; s++){
            cudaEvent_t a = p->ev[it * (hb_prof::SLOTS + 1) + s], b = p->ev[it * (hb_prof::SLOTS + 1) + s + 1];
            float ms = 0;
            if (a && b && cudaEventElapsedTime(&ms, a, b) == cudaSuccess){ p->ms[s] += ms; p->n[s]++; }
            else cudaGetLastError();
        }
    p->marked = 0;
}

extern "C" {

const char* hb_version(void){ return "halab200 0.1 (sm_100a)"; }
const char* hb_last_error(void){ return g_last_error.c_str(); }

int hb_device_count(int *count){
    HB_ARG(count, "count is null");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess){ *count = 0; cudaGetLastError(); }
    return HB_OK;
}

int hb_ctx_create(int device, hb_ctx **out){
    HB_ARG(out, "ctx output is null");
    int n = 0;
    HB_CUDA(cudaGetDeviceCount(&n));
    HB_ARG(device >= 0 && device < n, "device id out of range");
    HB_CUDA(cudaSetDevice(device));
    hb_ctx *ctx = new hb_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    HB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    HB_CUDA(cudaMalloc(&ctx->partials, HB_PARTIAL_BYTES));
    HB_CUDA(cudaMalloc((void**) &ctx->tickets, HB_NUM_TICKETS * sizeof(unsigned int)));
    HB_CUDA(cudaMemset(ctx->tickets, 0, HB_NUM_TICKETS * sizeof(unsigned int)));
    HB_CUDA(cudaMalloc(&ctx->dscalars, HB_SCALAR_BYTES));
    HB_CUDA(cudaMemset(ctx->dscalars, 0, HB_SCALAR_BYTES));
    HB_CUDA(cudaHostAlloc(&ctx->hscalars, HB_SCALAR_BYTES, cudaHostAllocMapped));
    memset(ctx->hscalars, 0, HB_SCALAR_BYTES);
    HB_CUDA(cudaHostGetDevicePointer(&ctx->hscalars_dev, ctx->hscalars, 0));
    HB_CUDA(cudaEventCreate(&ctx->timer[0]));
    HB_CUDA(cudaEventCreate(&ctx->timer[1]));
    *out = ctx;
    return HB_OK;
}

int hb_ctx_destroy(hb_ctx *ctx){
    if (!ctx) return HB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->timer[0]) cudaEventDestroy(ctx->timer[0]);
    if (ctx->timer[1]) cudaEventDestroy(ctx->timer[1]);
    if (ctx->work) cudaFree(ctx->work);
    if (ctx->prof){ for (cudaEvent_t e : ctx->prof->ev) if (e) cudaEventDestroy(e); delete ctx->prof; }
    cudaFree(ctx->partials); cudaFree(ctx->tickets); cudaFree(ctx->dscalars); cudaFreeHost(ctx->hscalars);
    delete ctx;
    return HB_OK;
}

int hb_ctx_profile(hb_ctx *ctx, int enable){
    HB_ARG(ctx, "ctx is null");
    if (enable){
        if (!ctx->prof) ctx->prof = new hb_prof();
        for (int s = 0; s < hb_prof::SLOTS; s++){ ctx->prof->ms[s] = 0; ctx->prof->n[s] = 0; }
        ctx->prof->marked = 0;
    }
    if (ctx->prof) ctx->prof->on = enable ? 1 : 0;
    return HB_OK;
}
int hb_ctx_profile_read(const hb_ctx *ctx, int slot, double *ms_total, long long *launches){
    HB_ARG(ctx && ms_total && launches, "null");
    HB_ARG(slot >= 0 && slot < hb_prof::SLOTS, "slot");
    *ms_total = ctx->prof ? ctx->prof->ms[slot] : 0.0;
    *launches = ctx->prof ? ctx->prof->n[slot] : 0;
    return HB_OK;
}

int hb_ctx_trim(hb_ctx *ctx){
    HB_ARG(ctx, "ctx is null");
    if (ctx->work){ HB_CUDA(cudaFree(ctx->work)); ctx->work = nullptr; ctx->work_bytes = 0; }
    if (pool_enabled()){ HB_CUDA(cudaDeviceSynchronize()); pool_trim(ctx->device); }     // and the blocks the memory pool keeps
    return HB_OK;
}
int hb_ctx_device(const hb_ctx *ctx, int *device){ HB_ARG(ctx && device, "null"); *device = ctx->device; return HB_OK; }
int hb_ctx_set_stream(hb_ctx *ctx, void *s){
    HB_ARG(ctx, "ctx is null");
    ctx->stream = (cudaStream_t) s;
    if (s && ctx->device >= 0 && ctx->device < HB_MAX_DEVICES) g_custom_stream[ctx->device] = true;
    return HB_OK;
}
int hb_ctx_get_stream(const hb_ctx *ctx, void **s){ HB_ARG(ctx && s, "null"); *s = (void*) ctx->stream; return HB_OK; }
int hb_ctx_sync(hb_ctx *ctx){
    HB_ARG(ctx, "ctx is null");
    HB_CUDA(cudaSetDevice(ctx->device));
    HB_CUDA(cudaDeviceSynchronize());
    return HB_OK;
}
int hb_ctx_set_pointer_mode(hb_ctx *ctx, int mode){
    HB_ARG(ctx, "ctx is null");
    HB_ARG(mode == HB_POINTER_HOST || mode == HB_POINTER_DEVICE, "pointer mode");
    ctx->pointer_mode = mode;
    return HB_OK;
}
int hb_ctx_get_pointer_mode(const hb_ctx *ctx, int *mode){ HB_ARG(ctx && mode, "null"); *mode = ctx->pointer_mode; return HB_OK; }
int hb_ctx_launch_count(const hb_ctx *ctx, long long *count){ HB_ARG(ctx && count, "null"); *count = ctx->launches; return HB_OK; }

int hb_timer_start(hb_ctx *ctx){
    HB_ARG(ctx, "ctx is null");
    HB_CUDA(cudaEventRecord(ctx->timer[0], ctx->stream));
    return HB_OK;
}
int hb_timer_stop(hb_ctx *ctx, float *ms){
    HB_ARG(ctx && ms, "null");
    HB_CUDA(cudaEventRecord(ctx->timer[1], ctx->stream));
    HB_CUDA(cudaEventSynchronize(ctx->timer[1]));
    HB_CUDA(cudaEventElapsedTime(ms, ctx->timer[0], ctx->timer[1]));
    return HB_OK;
}

int hb_malloc(hb_ctx *ctx, size_t bytes, void **ptr){
    HB_ARG(ctx && ptr, "null");
    HB_CUDA(cudaSetDevice(ctx->device));      // as gpu_allocate does (gpu/hala_cuda_common.hpp:255)
    *ptr = nullptr;
    if (bytes == 0) return HB_OK;
    cudaError_t e = pool_malloc(ctx->device, ptr, bytes, ctx->stream);
    if (e != cudaSuccess){ hb_cuda_fail(e, "hb_malloc"); return HB_ERR_ALLOC; }
    return HB_OK;
}
int hb_free(hb_ctx *ctx, void *ptr){
    if (ptr) HB_CUDA(pool_free(ptr, ctx ? ctx->stream : nullptr));
    return HB_OK;
}
static cudaMemcpyKind to_kind(int kind){
    return kind == HB_H2D ? cudaMemcpyHostToDevice : kind == HB_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
}
int hb_memcpy(hb_ctx *ctx, void *dst, const void *src, size_t bytes, int kind){
    HB_ARG(ctx, "ctx is null");
    if (bytes == 0) return HB_OK;
    HB_CUDA(cudaMemcpyAsync(dst, src, bytes, to_kind(kind), ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    return HB_OK;
}
int hb_memcpy_async(hb_ctx *ctx, void *dst, const void *src, size_t bytes, int kind){
    HB_ARG(ctx, "ctx is null");
    if (bytes == 0) return HB_OK;
    HB_CUDA(cudaMemcpyAsync(dst, src, bytes, to_kind(kind), ctx->stream));
    return HB_OK;
}
int hb_memset_zero(hb_ctx *ctx, void *ptr, size_t bytes){
    HB_ARG(ctx, "ctx is null");
    if (bytes == 0) return HB_OK;
    HB_CUDA(cudaMemsetAsync(ptr, 0, bytes, ctx->stream));
    return HB_OK;
}
int hb_fill(hb_ctx *ctx, int dtype, size_t n, const void *host_value, void *x){
    HB_ARG(ctx && host_value, "null");
    if (n == 0) return HB_OK;
    HB_ARG(x, "x is null");
    int grid = hb_grid_for(ctx, n, 1024, 8);
    if (dtype == -1){
        fill_kernel<int><<<grid, 256, 0, ctx->stream>>>(n, *(const int*) host_value, (int*) x);
    }else{
        HB_DISPATCH(dtype, (fill_kernel<T><<<grid, 256, 0, ctx->stream>>>(n, *(const T*) host_value, (T*) x)));
    }
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_dev_malloc(int device, size_t bytes, void **ptr){
    HB_ARG(ptr, "null");
    HB_CUDA(cudaSetDevice(device));
    *ptr = nullptr;
    if (bytes == 0) return HB_OK;
    cudaError_t e = pool_malloc(device, ptr, bytes, nullptr);
    if (e != cudaSuccess){ hb_cuda_fail(e, "hb_dev_malloc"); return HB_ERR_ALLOC; }
    return HB_OK;
}
int hb_dev_free(void *ptr){ if (ptr) HB_CUDA(pool_free(ptr, nullptr)); return HB_OK; }
int hb_dev_memcpy(void *dst, const void *src, size_t bytes, int kind){
    if (bytes == 0) return HB_OK;
    HB_CUDA(cudaMemcpy(dst, src, bytes, to_kind(kind)));
    return HB_OK;
}
int hb_dev_fill(int device, int dtype, size_t n, const void *host_value, void *x){
    HB_ARG(host_value, "null");
    if (n == 0) return HB_OK;
    HB_ARG(x, "x is null");
    HB_CUDA(cudaSetDevice(device));
    const int grid = (int) std::min<size_t>((n + 1023) / 1024, 148 * 8);
    if (dtype == -1){
        fill_kernel<int><<<grid, 256>>>(n, *(const int*) host_value, (int*) x);
    }else{
        HB_DISPATCH(dtype, (fill_kernel<T><<<grid, 256>>>(n, *(const T*) host_value, (T*) x)));
    }
    HB_CUDA(cudaPeekAtLastError());
    HB_CUDA(cudaDeviceSynchronize());
    return HB_OK;
}
int hb_host_alloc(size_t bytes, void **ptr){
    HB_ARG(ptr, "null");
    HB_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return HB_OK;
}
int hb_host_free(void *ptr){ if (ptr) HB_CUDA(cudaFreeHost(ptr)); return HB_OK; }

}
