// hb_blas1.cu — BLAS-1 on sm_100a: copy, axpy, scal, dot (conj / unconj), nrm2 for float, double, complex<float/double>.
// Replaces the cublas{S,D,C,Z}{copy,axpy,scal,dot,dotc,dotu,nrm2} call sites of the reference
// (gpu/hala_gpu_blas1.hpp:48-65, 204-222, 228-245, 178-198, 102-121).
//
// All five are HBM-bound streams: unit-stride arrays that are 16-byte aligned go through 128-bit loads/stores with four
// independent vectors in flight per thread; anything else (strides, odd alignment) takes the element-wise path.
// Reductions: warp shuffle -> shared memory -> one partial per block -> the last block to arrive (one atomic ticket per
// block) adds the partials in fixed order, so results are bit-reproducible run to run.
#include "hb_common.cuh"

static constexpr int B1_THREADS = 256;
static constexpr int B1_UNROLL  = 4;

// ---------------------------------------------------------------- element-wise (strided / unaligned) kernels
template<typename T> __global__ void copy_strided(int n, const T *x, long long incx, T *y, long long incy){
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        y[i * incy] = x[i * incx];
}
template<typename T> __global__ void axpy_strided(int n, scalar_arg<T> alpha, const T *x, long long incx, T *y, long long incy){
    const T a = get_scalar(alpha);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        y[i * incy] = hfma(a, x[i * incx], y[i * incy]);
}
template<typename T> __global__ void scal_strided(int n, scalar_arg<T> alpha, T *x, long long incx){
    const T a = get_scalar(alpha);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        x[i * incx] = hmul(a, x[i * incx]);
}

// ---------------------------------------------------------------- 128-bit streaming kernels (unit stride, aligned)
template<typename T> __global__ void __launch_bounds__(B1_THREADS) copy_vec(size_t nvec, const vec16<T> *x, vec16<T> *y){
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    for (; i + (B1_UNROLL - 1) * stride < nvec; i += B1_UNROLL * stride){
        vec16<T> a[B1_UNROLL];
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++) a[u] = x[i + u * stride];
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++) y[i + u * stride] = a[u];
    }
    for (; i < nvec; i += stride) y[i] = x[i];
}
template<typename T> __global__ void __launch_bounds__(B1_THREADS) axpy_vec(size_t nvec, scalar_arg<T> alpha, const vec16<T> *x, vec16<T> *y){
    const T a = get_scalar(alpha);
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    for (; i + (B1_UNROLL - 1) * stride < nvec; i += B1_UNROLL * stride){
        vec16<T> vx[B1_UNROLL], vy[B1_UNROLL];
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++){ vx[u] = x[i + u * stride]; vy[u] = y[i + u * stride]; }
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++) vy[u].v[k] = hfma(a, vx[u].v[k], vy[u].v[k]);
            y[i + u * stride] = vy[u];
        }
    }
    for (; i < nvec; i += stride){
        vec16<T> vx = x[i], vy = y[i];
        #pragma unroll
        for (int k = 0; k < vec16<T>::N; k++) vy.v[k] = hfma(a, vx.v[k], vy.v[k]);
        y[i] = vy;
    }
}
template<typename T> __global__ void __launch_bounds__(B1_THREADS) scal_vec(size_t nvec, scalar_arg<T> alpha, vec16<T> *x){
    const T a = get_scalar(alpha);
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    for (; i + (B1_UNROLL - 1) * stride < nvec; i += B1_UNROLL * stride){
        vec16<T> vx[B1_UNROLL];
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++) vx[u] = x[i + u * stride];
        #pragma unroll
        for (int u = 0; u < B1_UNROLL; u++){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++) vx[u].v[k] = hmul(a, vx[u].v[k]);
            x[i + u * stride] = vx[u];
        }
    }
    for (; i < nvec; i += stride){
        vec16<T> vx = x[i];
        #pragma unroll
        for (int k = 0; k < vec16<T>::N; k++) vx.v[k] = hmul(a, vx.v[k]);
        x[i] = vx;
    }
}

// ---------------------------------------------------------------- reductions
// MODE 0: dot (unconjugated), 1: dot (conj(x) * y), 2: sum |x|^2 accumulated in double (nrm2; y unused),
// 3: sum |re| + |im| accumulated in double (asum; y unused)
template<int MODE, typename T> struct red_type { using type = T; };
template<typename T> struct red_type<2, T> { using type = double; };
template<typename T> struct red_type<3, T> { using type = double; };
__device__ __forceinline__ double habs1(float a){ return fabs((double) a); }
__device__ __forceinline__ double habs1(double a){ return fabs(a); }
template<typename R> __device__ __forceinline__ double habs1(cplx<R> a){ return fabs((double) a.re) + fabs((double) a.im); }

template<int MODE, typename T>
__device__ __forceinline__ typename red_type<MODE, T>::type red_term(T a, T b, typename red_type<MODE, T>::type acc){
    if constexpr (MODE == 2) return acc + (double) habs2(a);
    else if constexpr (MODE == 3) return acc + habs1(a);
    else if constexpr (MODE == 1) return hfma(hconj(a), b, acc);
    else return hfma(a, b, acc);
}

template<int MODE, typename T>
__global__ void __launch_bounds__(B1_THREADS) reduce_kernel(int n, size_t nvec, const T *x, long long incx, const T *y, long long incy,
                                                            void *partials_v, unsigned int *ticket, void *out_v){
    using A = typename red_type<MODE, T>::type;
    __shared__ A red[32];
    A acc = zero_of<A>();
    if (nvec > 0){      // unit stride, aligned: 128-bit loads; the scalar tail (n - nvec*N) is handled below
        const vec16<T> *vx = reinterpret_cast<const vec16<T>*>(x), *vy = reinterpret_cast<const vec16<T>*>(y);
        const size_t stride = (size_t) gridDim.x * blockDim.x;
        size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
        for (; i + (B1_UNROLL - 1) * stride < nvec; i += B1_UNROLL * stride){
            vec16<T> a[B1_UNROLL], b[B1_UNROLL];
            #pragma unroll
            for (int u = 0; u < B1_UNROLL; u++){ a[u] = vx[i + u * stride]; if (MODE < 2) b[u] = vy[i + u * stride]; else b[u] = a[u]; }
            #pragma unroll
            for (int u = 0; u < B1_UNROLL; u++){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++) acc = red_term<MODE, T>(a[u].v[k], b[u].v[k], acc);
            }
        }
        for (; i < nvec; i += stride){
            vec16<T> a = vx[i], b = (MODE < 2) ? vy[i] : a;
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++) acc = red_term<MODE, T>(a.v[k], b.v[k], acc);
        }
        const long long done = (long long) nvec * vec16<T>::N;
        for (long long j = done + blockIdx.x * (long long) blockDim.x + threadIdx.x; j < n; j += (long long) gridDim.x * blockDim.x)
            acc = red_term<MODE, T>(x[j], (MODE < 2) ? y[j] : x[j], acc);
    }else{
        for (long long j = blockIdx.x * (long long) blockDim.x + threadIdx.x; j < n; j += (long long) gridDim.x * blockDim.x){
            T a = x[j * incx];
            acc = red_term<MODE, T>(a, (MODE < 2) ? y[j * incy] : a, acc);
        }
    }
    A *partials = reinterpret_cast<A*>(partials_v);
    A bsum = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = bsum;
    if (last_block_arrives(ticket)){
        A total = sum_partials<A>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0){
            if constexpr (MODE == 2) *reinterpret_cast<real_t<T>*>(out_v) = (real_t<T>) sqrt(total);
            else if constexpr (MODE == 3) *reinterpret_cast<real_t<T>*>(out_v) = (real_t<T>) total;
            else *reinterpret_cast<T*>(out_v) = total;
        }
    }
}

template<int MODE, typename T>
static int launch_reduce(hb_ctx *ctx, int n, const T *x, int incx, const T *y, int incy, void *result){
    const size_t out_bytes = (MODE >= 2) ? sizeof(real_t<T>) : sizeof(T);
    void *out = (ctx->pointer_mode == HB_POINTER_HOST) ? ctx->hscalars_dev : result;
    if (n <= 0){
        if (ctx->pointer_mode == HB_POINTER_HOST) memset(result, 0, out_bytes);
        else HB_CUDA(cudaMemsetAsync(result, 0, out_bytes, ctx->stream));
        return HB_OK;
    }
    const bool vec = (incx == 1) && (MODE >= 2 || incy == 1) && aligned16(x) && (MODE >= 2 || aligned16(y));
    const size_t nvec = vec ? (size_t) n / vec16<T>::N : 0;
    // when the vector path leaves nothing (n < N) fall back to the element-wise loop
    const size_t use_nvec = (nvec > 0) ? nvec : 0;
    int grid = hb_grid_for(ctx, (size_t) n, B1_THREADS * B1_UNROLL * vec16<T>::N, 4);
    reduce_kernel<MODE, T><<<grid, B1_THREADS, 0, ctx->stream>>>(n, use_nvec, x, incx, y, incy, ctx->partials, ctx->tickets + 0, out);
    HB_LAUNCH_CHECK(ctx);
    if (ctx->pointer_mode == HB_POINTER_HOST){
        HB_CUDA(cudaStreamSynchronize(ctx->stream));
        memcpy(result, ctx->hscalars, out_bytes);
    }
    return HB_OK;
}

extern "C" {

int hb_copy(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *y, int incy){
    HB_ARG(ctx, "ctx is null");
    if (n <= 0) return HB_OK;
    HB_ARG(x && y, "null vector");
    HB_ARG(incx != 0 && incy != 0, "zero increment");
    x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype));
    HB_DISPATCH(dtype, {
        if (incx == 1 && incy == 1 && aligned16(x) && aligned16(y) && (size_t) n >= vec16<T>::N){
            size_t nvec = (size_t) n / vec16<T>::N;
            int grid = hb_grid_for(ctx, nvec, B1_THREADS * B1_UNROLL, 8);
            copy_vec<T><<<grid, B1_THREADS, 0, ctx->stream>>>(nvec, (const vec16<T>*) x, (vec16<T>*) y);
            HB_LAUNCH_CHECK(ctx);
            size_t done = nvec * vec16<T>::N;
            if (done < (size_t) n){
                copy_strided<T><<<1, 32, 0, ctx->stream>>>(n - (int) done, (const T*) x + done, 1, (T*) y + done, 1);
                HB_LAUNCH_CHECK(ctx);
            }
        }else{
            int grid = hb_grid_for(ctx, (size_t) n, B1_THREADS, 8);
            copy_strided<T><<<grid, B1_THREADS, 0, ctx->stream>>>(n, (const T*) x, incx, (T*) y, incy);
            HB_LAUNCH_CHECK(ctx);
        }
    });
    return HB_OK;
}

int hb_axpy(hb_ctx *ctx, int dtype, int n, const void *alpha, const void *x, int incx, void *y, int incy){
    HB_ARG(ctx && alpha, "null");
    if (n <= 0) return HB_OK;
    HB_ARG(x && y, "null vector");
    HB_ARG(incx != 0 && incy != 0, "zero increment");
    x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype));
    HB_DISPATCH(dtype, {
        scalar_arg<T> a = make_scalar<T>(ctx, alpha);
        if (incx == 1 && incy == 1 && aligned16(x) && aligned16(y) && (size_t) n >= vec16<T>::N){
            size_t nvec = (size_t) n / vec16<T>::N;
            int grid = hb_grid_for(ctx, nvec, B1_THREADS * B1_UNROLL, 8);
            axpy_vec<T><<<grid, B1_THREADS, 0, ctx->stream>>>(nvec, a, (const vec16<T>*) x, (vec16<T>*) y);
            HB_LAUNCH_CHECK(ctx);
            size_t done = nvec * vec16<T>::N;
            if (done < (size_t) n){
                axpy_strided<T><<<1, 32, 0, ctx->stream>>>(n - (int) done, a, (const T*) x + done, 1, (T*) y + done, 1);
                HB_LAUNCH_CHECK(ctx);
            }
        }else{
            int grid = hb_grid_for(ctx, (size_t) n, B1_THREADS, 8);
            axpy_strided<T><<<grid, B1_THREADS, 0, ctx->stream>>>(n, a, (const T*) x, incx, (T*) y, incy);
            HB_LAUNCH_CHECK(ctx);
        }
    });
    return HB_OK;
}

int hb_scal(hb_ctx *ctx, int dtype, int n, const void *alpha, void *x, int incx){
    HB_ARG(ctx && alpha, "null");
    if (n <= 0 || incx <= 0) return HB_OK;                 // netlib ?scal: nothing to do for a non-positive increment
    HB_ARG(x, "null vector");
    HB_DISPATCH(dtype, {
        scalar_arg<T> a = make_scalar<T>(ctx, alpha);
        if (incx == 1 && aligned16(x) && (size_t) n >= vec16<T>::N){
            size_t nvec = (size_t) n / vec16<T>::N;
            int grid = hb_grid_for(ctx, nvec, B1_THREADS * B1_UNROLL, 8);
            scal_vec<T><<<grid, B1_THREADS, 0, ctx->stream>>>(nvec, a, (vec16<T>*) x);
            HB_LAUNCH_CHECK(ctx);
            size_t done = nvec * vec16<T>::N;
            if (done < (size_t) n){
                scal_strided<T><<<1, 32, 0, ctx->stream>>>(n - (int) done, a, (T*) x + done, 1);
                HB_LAUNCH_CHECK(ctx);
            }
        }else{
            int grid = hb_grid_for(ctx, (size_t) n, B1_THREADS, 8);
            scal_strided<T><<<grid, B1_THREADS, 0, ctx->stream>>>(n, a, (T*) x, incx);
            HB_LAUNCH_CHECK(ctx);
        }
    });
    return HB_OK;
}

int hb_dot(hb_ctx *ctx, int dtype, int conj, int n, const void *x, int incx, const void *y, int incy, void *result){
    HB_ARG(ctx && result, "null");
    HB_ARG(n <= 0 || (x && y), "null vector");
    HB_ARG(n <= 0 || (incx != 0 && incy != 0), "zero increment");
    if (n > 0){ x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype)); }
    HB_DISPATCH(dtype, {
        if (conj && is_cplx<T>::value) return launch_reduce<1, T>(ctx, n, (const T*) x, incx, (const T*) y, incy, result);
        return launch_reduce<0, T>(ctx, n, (const T*) x, incx, (const T*) y, incy, result);
    });
    return HB_OK;
}

int hb_asum(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *result){
    HB_ARG(ctx && result, "null");
    HB_ARG(n <= 0 || x, "null vector");
    if (incx <= 0) n = 0;                                   // netlib ?asum: zero for a non-positive increment
    HB_DISPATCH(dtype, { return launch_reduce<3, T>(ctx, n, (const T*) x, incx, (const T*) x, incx, result); });
    return HB_OK;
}

int hb_nrm2(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *result){
    HB_ARG(ctx && result, "null");
    HB_ARG(n <= 0 || x, "null vector");
    if (incx <= 0) n = 0;                                   // netlib ?nrm2: zero for a non-positive increment
    HB_DISPATCH(dtype, { return launch_reduce<2, T>(ctx, n, (const T*) x, incx, (const T*) x, incx, result); });
    return HB_OK;
}

}
