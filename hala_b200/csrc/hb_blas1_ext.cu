// hb_blas1_ext.cu — the rest of BLAS-1 on the engine, so that gpu_engine needs no cuBLAS handle at all:
// swap, iamax, rot (real and mixed complex/real), rotm, rotg, rotmg.
// Replaces cublas{S,D,C,Z}swap (reference gpu/hala_gpu_blas1.hpp:83-100), cublasI{s,d,c,z}amax (:153-172), cublas?rotg (:251-262),
// cublas{S,D,C,Z}rot / cublasCsrot / cublasZdrot (:269-300), cublas{S,D}rotmg (:318-341), cublas{S,D}rotm (:349-371).
// Vectors stream once (128-bit packets when unit-stride and aligned); the scalar routines run on the host in host pointer mode
// and in a one-thread kernel in device pointer mode, from the same __host__ __device__ code (netlib reference algorithms).
#include "hb_common.cuh"
#include <cmath>

static constexpr int X_THREADS = 256;

// ---------------------------------------------------------------- swap
template<typename T, bool VEC> __global__ void __launch_bounds__(X_THREADS) swap_kernel(int n, T *x, long long incx, T *y, long long incy){
    if (VEC){
        constexpr int U = 4;
        vec16<T> vx[U], vy[U];
        vec16<T> *x4 = reinterpret_cast<vec16<T>*>(x), *y4 = reinterpret_cast<vec16<T>*>(y);
        stream_sweep<T, true, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vy[u] = y4[i]; },
            [&](int u, size_t i){ x4[i] = vy[u]; y4[i] = vx[u]; },
            [&](size_t j){ T t = x[j]; x[j] = y[j]; y[j] = t; });
    }else{
        for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
            T t = x[i * incx]; x[i * incx] = y[i * incy]; y[i * incy] = t;
        }
    }
}

// ---------------------------------------------------------------- iamax: first index of the largest |re| + |im| (BLAS definition)
__host__ __device__ __forceinline__ double cabs1(float a){ return fabs((double) a); }
__host__ __device__ __forceinline__ double cabs1(double a){ return fabs(a); }
template<typename R> __host__ __device__ __forceinline__ double cabs1(cplx<R> a){ return fabs((double) a.re) + fabs((double) a.im); }

struct amax_pair { double v; long long i; };
__device__ __forceinline__ amax_pair amax_better(amax_pair a, amax_pair b){     // larger value wins, ties go to the smaller index; NaN never wins
    if (b.i >= 0 && (a.i < 0 || b.v > a.v || (b.v == a.v && b.i < a.i))) return b;
    return a;
}
__device__ __forceinline__ amax_pair amax_shfl_down(amax_pair p, int d){
    amax_pair q; q.v = __shfl_down_sync(0xffffffffu, p.v, d); q.i = __shfl_down_sync(0xffffffffu, p.i, d); return q;
}
__device__ __forceinline__ amax_pair amax_block(amax_pair p, amax_pair *red){
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    for (int d = 16; d > 0; d >>= 1) p = amax_better(p, amax_shfl_down(p, d));
    if (lane == 0) red[warp] = p;
    __syncthreads();
    amax_pair r; r.v = -1.0; r.i = -1;
    if (warp == 0){
        if (lane < nwarps) r = red[lane];
        for (int d = 16; d > 0; d >>= 1) r = amax_better(r, amax_shfl_down(r, d));
    }
    __syncthreads();
    return r;
}
template<typename T> __global__ void __launch_bounds__(X_THREADS) iamax_kernel(int n, const T *x, long long incx, amax_pair *partials, unsigned int *ticket, int *out){
    __shared__ amax_pair red[32];
    amax_pair best; best.v = -1.0; best.i = -1;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
        amax_pair c; c.v = cabs1(x[i * incx]); c.i = i;
        best = amax_better(best, c);
    }
    best = amax_block(best, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = best;
    if (last_block_arrives(ticket)){
        amax_pair b2; b2.v = -1.0; b2.i = -1;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += blockDim.x){
            amax_pair c; c.v = __ldcg(&partials[i].v); c.i = __ldcg(&partials[i].i);
            b2 = amax_better(b2, c);
        }
        b2 = amax_block(b2, red);
        if (threadIdx.x == 0) *out = (int) (b2.i + 1);          // 1-based like cublasI?amax; 0 for n == 0
    }
}

// ---------------------------------------------------------------- rot:  x' = c x + s y ;  y' = c y - conj(s) x
template<typename T> struct rot_args { real_t<T> c; T s; const real_t<T> *c_dev; const void *s_dev; int s_is_real; };
template<typename T> __device__ __forceinline__ void rot_fetch(const rot_args<T> &a, T &c, T &s){
    c = from_real<T>(a.c_dev ? *a.c_dev : a.c);
    if (a.s_dev) s = a.s_is_real ? from_real<T>(*reinterpret_cast<const real_t<T>*>(a.s_dev)) : *reinterpret_cast<const T*>(a.s_dev);
    else s = a.s;
}
template<typename T, bool VEC> __global__ void __launch_bounds__(X_THREADS) rot_kernel(int n, T *x, long long incx, T *y, long long incy, rot_args<T> a){
    T c, s;
    rot_fetch(a, c, s);
    const T ncs = hneg(hconj(s));
    auto one = [&](T &xv, T &yv){ T tx = hfma(s, yv, hmul(c, xv)); yv = hfma(ncs, xv, hmul(c, yv)); xv = tx; };
    if (VEC){
        constexpr int U = 2;
        vec16<T> vx[U], vy[U];
        vec16<T> *x4 = reinterpret_cast<vec16<T>*>(x), *y4 = reinterpret_cast<vec16<T>*>(y);
        stream_sweep<T, true, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vy[u] = y4[i]; },
            [&](int u, size_t i){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++) one(vx[u].v[k], vy[u].v[k]);
                x4[i] = vx[u]; y4[i] = vy[u];
            },
            [&](size_t j){ T xv = x[j], yv = y[j]; one(xv, yv); x[j] = xv; y[j] = yv; });
    }else{
        for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
            T xv = x[i * incx], yv = y[i * incy]; one(xv, yv); x[i * incx] = xv; y[i * incy] = yv;
        }
    }
}

// ---------------------------------------------------------------- rotm (real): [x; y] <- H [x; y], H coded in param[0] as in BLAS
template<typename R> struct rotm_args { R p[5]; const R *dev; };
template<typename R> __global__ void __launch_bounds__(X_THREADS) rotm_kernel(int n, R *x, long long incx, R *y, long long incy, rotm_args<R> a){
    R flag, h11, h21, h12, h22;
    if (a.dev){ flag = a.dev[0]; h11 = a.dev[1]; h21 = a.dev[2]; h12 = a.dev[3]; h22 = a.dev[4]; }
    else { flag = a.p[0]; h11 = a.p[1]; h21 = a.p[2]; h12 = a.p[3]; h22 = a.p[4]; }
    if (flag == R(-2)) return;                                  // identity
    if (flag == R(0)){ h11 = R(1); h22 = R(1); }
    else if (flag == R(1)){ h12 = R(1); h21 = R(-1); }
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
        const R xv = x[i * incx], yv = y[i * incy];
        x[i * incx] = xv * h11 + yv * h12;
        y[i * incy] = xv * h21 + yv * h22;
    }
}

// ---------------------------------------------------------------- rotg / rotmg (netlib reference BLAS 3.8 algorithms)
template<typename R> __host__ __device__ void rotg_real(R &a, R &b, R &c, R &s){
    R roe = b;
    const R absa = fabs(a), absb = fabs(b);
    if (absa > absb) roe = a;
    const R scale = absa + absb;
    if (scale == R(0)){ c = R(1); s = R(0); a = R(0); b = R(0); return; }
    R r = scale * sqrt((a / scale) * (a / scale) + (b / scale) * (b / scale));
    if (roe < R(0)) r = -r;
    c = a / r; s = b / r;
    R z = R(1);
    if (absa > absb) z = s;
    if (absb >= absa && c != R(0)) z = R(1) / c;
    a = r; b = z;
}
template<typename R> __host__ __device__ void rotg_cplx(cplx<R> &a, const cplx<R> &b, R &c, cplx<R> &s){
    const R absa = sqrt(habs2(a));
    if (absa == R(0)){ c = R(0); s = {R(1), R(0)}; a = b; return; }
    const R scale = absa + sqrt(habs2(b));
    const cplx<R> as = {a.re / scale, a.im / scale}, bs = {b.re / scale, b.im / scale};
    const R norm = scale * sqrt(habs2(as) + habs2(bs));
    const cplx<R> alpha = {a.re / absa, a.im / absa};
    c = absa / norm;
    const cplx<R> t = hmul(alpha, hconj(b));
    s = {t.re / norm, t.im / norm};
    a = {alpha.re * norm, alpha.im * norm};
}
template<typename R> __host__ __device__ void rotmg_real(R &d1, R &d2, R &x1, const R y1, R *param){
    const R gam = R(4096), gamsq = R(16777216), rgamsq = R(5.9604645e-8);
    R flag, h11 = 0, h12 = 0, h21 = 0, h22 = 0;
    if (d1 < R(0)){
        flag = R(-1); d1 = d2 = x1 = R(0);
    }else{
        const R p2 = d2 * y1;
        if (p2 == R(0)){ param[0] = R(-2); return; }
        const R p1 = d1 * x1, q2 = p2 * y1, q1 = p1 * x1;
        if (fabs(q1) > fabs(q2)){
            h21 = -y1 / x1; h12 = p2 / p1;
            const R u = R(1) - h12 * h21;
            if (u > R(0)){ flag = R(0); d1 /= u; d2 /= u; x1 *= u; }
            else { flag = R(-1); h11 = h12 = h21 = h22 = R(0); d1 = d2 = x1 = R(0); }
        }else{
            if (q2 < R(0)){ flag = R(-1); h11 = h12 = h21 = h22 = R(0); d1 = d2 = x1 = R(0); }
            else{
                flag = R(1); h11 = p1 / p2; h22 = x1 / y1;
                const R u = R(1) + h11 * h22, t = d2 / u;
                d2 = d1 / u; d1 = t; x1 = y1 * u;
            }
        }
        if (d1 != R(0)){
            while (d1 <= rgamsq || d1 >= gamsq){
                if (flag == R(0)){ h11 = R(1); h22 = R(1); flag = R(-1); }
                else if (flag == R(1)){ h21 = R(-1); h12 = R(1); flag = R(-1); }
                if (d1 <= rgamsq){ d1 *= gam * gam; x1 /= gam; h11 /= gam; h12 /= gam; }
                else { d1 /= gam * gam; x1 *= gam; h11 *= gam; h12 *= gam; }
            }
        }
        if (d2 != R(0)){
            while (fabs(d2) <= rgamsq || fabs(d2) >= gamsq){
                if (flag == R(0)){ h11 = R(1); h22 = R(1); flag = R(-1); }
                else if (flag == R(1)){ h21 = R(-1); h12 = R(1); flag = R(-1); }
                if (fabs(d2) <= rgamsq){ d2 *= gam * gam; h21 /= gam; h22 /= gam; }
                else { d2 /= gam * gam; h21 *= gam; h22 *= gam; }
            }
        }
    }
    if (flag < R(0)){ param[1] = h11; param[2] = h21; param[3] = h12; param[4] = h22; }
    else if (flag == R(0)){ param[2] = h21; param[3] = h12; }
    else { param[1] = h11; param[4] = h22; }
    param[0] = flag;
}
template<typename R> __global__ void rotg_real_kernel(R *a, R *b, R *c, R *s){ rotg_real(*a, *b, *c, *s); }
template<typename R> __global__ void rotg_cplx_kernel(cplx<R> *a, const cplx<R> *b, R *c, cplx<R> *s){ rotg_cplx(*a, *b, *c, *s); }
template<typename R> __global__ void rotmg_kernel(R *d1, R *d2, R *x1, const R *y1, R *param){ rotmg_real(*d1, *d2, *x1, *y1, param); }

static int x_grid(const hb_ctx *ctx, long long n, int per_block){ return hb_grid_for(ctx, (size_t) (n > 0 ? n : 1), per_block, 4); }

extern "C" {

int hb_swap(hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy){
    HB_ARG(ctx, "ctx is null");
    if (n <= 0) return HB_OK;
    HB_ARG(x && y, "null vector");
    HB_ARG(incx != 0 && incy != 0, "zero increment");
    x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype));
    const bool vec = incx == 1 && incy == 1 && aligned16(x) && aligned16(y);
    HB_DISPATCH(dtype, {
        if (vec) swap_kernel<T, true><<<x_grid(ctx, n, X_THREADS * 8), X_THREADS, 0, ctx->stream>>>(n, (T*) x, 1, (T*) y, 1);
        else     swap_kernel<T, false><<<x_grid(ctx, n, X_THREADS * 4), X_THREADS, 0, ctx->stream>>>(n, (T*) x, incx, (T*) y, incy);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_iamax(hb_ctx *ctx, int dtype, int n, const void *x, int incx, int *result){
    HB_ARG(ctx && result, "null");
    int *out = (ctx->pointer_mode == HB_POINTER_HOST) ? reinterpret_cast<int*>(ctx->hscalars_dev) : result;
    if (n <= 0 || incx <= 0){
        if (ctx->pointer_mode == HB_POINTER_HOST) *result = 0;
        else HB_CUDA(cudaMemsetAsync(result, 0, sizeof(int), ctx->stream));
        return HB_OK;
    }
    HB_ARG(x, "null vector");
    HB_DISPATCH(dtype, (iamax_kernel<T><<<x_grid(ctx, n, X_THREADS * 8), X_THREADS, 0, ctx->stream>>>(n, (const T*) x, incx, (amax_pair*) ctx->partials,
                                                                                                         ctx->tickets + 7, out)));
    HB_LAUNCH_CHECK(ctx);
    if (ctx->pointer_mode == HB_POINTER_HOST){
        HB_CUDA(cudaStreamSynchronize(ctx->stream));
        *result = *reinterpret_cast<volatile int*>(ctx->hscalars);
    }
    return HB_OK;
}

int hb_rot(hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy, const void *c, const void *s, int s_is_real){
    HB_ARG(ctx && c && s, "null");
    if (n <= 0) return HB_OK;
    HB_ARG(x && y, "null vector");
    HB_ARG(incx != 0 && incy != 0, "zero increment");
    x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype));
    const bool vec = incx == 1 && incy == 1 && aligned16(x) && aligned16(y);
    HB_DISPATCH(dtype, {
        rot_args<T> a;
        a.c = real_t<T>(0); a.s = zero_of<T>(); a.c_dev = nullptr; a.s_dev = nullptr; a.s_is_real = s_is_real;
        if (ctx->pointer_mode == HB_POINTER_HOST){
            a.c = *reinterpret_cast<const real_t<T>*>(c);
            a.s = s_is_real ? from_real<T>(*reinterpret_cast<const real_t<T>*>(s)) : *reinterpret_cast<const T*>(s);
        }else{ a.c_dev = reinterpret_cast<const real_t<T>*>(c); a.s_dev = s; }
        if (vec) rot_kernel<T, true><<<x_grid(ctx, n, X_THREADS * 4), X_THREADS, 0, ctx->stream>>>(n, (T*) x, 1, (T*) y, 1, a);
        else     rot_kernel<T, false><<<x_grid(ctx, n, X_THREADS * 4), X_THREADS, 0, ctx->stream>>>(n, (T*) x, incx, (T*) y, incy, a);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_rotm(hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy, const void *param){
    HB_ARG(ctx && param, "null");
    HB_ARG(dtype == HB_F32 || dtype == HB_F64, "rotm is defined for real types only");
    if (n <= 0) return HB_OK;
    HB_ARG(x && y, "null vector");
    HB_ARG(incx != 0 && incy != 0, "zero increment");
    x = hb_blas_base(x, n, incx, hb_dtype_size(dtype)); y = hb_blas_base(y, n, incy, hb_dtype_size(dtype));
    if (dtype == HB_F32){
        rotm_args<float> a; a.dev = nullptr;
        if (ctx->pointer_mode == HB_POINTER_HOST) memcpy(a.p, param, sizeof(a.p)); else a.dev = (const float*) param;
        rotm_kernel<float><<<x_grid(ctx, n, X_THREADS * 4), X_THREADS, 0, ctx->stream>>>(n, (float*) x, incx, (float*) y, incy, a);
    }else{
        rotm_args<double> a; a.dev = nullptr;
        if (ctx->pointer_mode == HB_POINTER_HOST) memcpy(a.p, param, sizeof(a.p)); else a.dev = (const double*) param;
        rotm_kernel<double><<<x_grid(ctx, n, X_THREADS * 4), X_THREADS, 0, ctx->stream>>>(n, (double*) x, incx, (double*) y, incy, a);
    }
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_rotg(hb_ctx *ctx, int dtype, void *a, void *b, void *c, void *s){
    HB_ARG(ctx && a && b && c && s, "null");
    const bool host = ctx->pointer_mode == HB_POINTER_HOST;
    switch (dtype){
        case HB_F32: if (host) rotg_real(*(float*) a, *(float*) b, *(float*) c, *(float*) s);
                     else rotg_real_kernel<float><<<1, 1, 0, ctx->stream>>>((float*) a, (float*) b, (float*) c, (float*) s); break;
        case HB_F64: if (host) rotg_real(*(double*) a, *(double*) b, *(double*) c, *(double*) s);
                     else rotg_real_kernel<double><<<1, 1, 0, ctx->stream>>>((double*) a, (double*) b, (double*) c, (double*) s); break;
        case HB_C32: if (host) rotg_cplx(*(cplx<float>*) a, *(const cplx<float>*) b, *(float*) c, *(cplx<float>*) s);
                     else rotg_cplx_kernel<float><<<1, 1, 0, ctx->stream>>>((cplx<float>*) a, (const cplx<float>*) b, (float*) c, (cplx<float>*) s); break;
        case HB_C64: if (host) rotg_cplx(*(cplx<double>*) a, *(const cplx<double>*) b, *(double*) c, *(cplx<double>*) s);
                     else rotg_cplx_kernel<double><<<1, 1, 0, ctx->stream>>>((cplx<double>*) a, (const cplx<double>*) b, (double*) c, (cplx<double>*) s); break;
        default: hb_set_error("unknown dtype"); return HB_ERR_ARG;
    }
    if (!host) HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_rotmg(hb_ctx *ctx, int dtype, void *d1, void *d2, void *x1, const void *y1, void *param){
    HB_ARG(ctx && d1 && d2 && x1 && y1 && param, "null");
    HB_ARG(dtype == HB_F32 || dtype == HB_F64, "rotmg is defined for real types only");
    const bool host = ctx->pointer_mode == HB_POINTER_HOST;
    if (dtype == HB_F32){
        if (host) rotmg_real(*(float*) d1, *(float*) d2, *(float*) x1, *(const float*) y1, (float*) param);
        else rotmg_kernel<float><<<1, 1, 0, ctx->stream>>>((float*) d1, (float*) d2, (float*) x1, (const float*) y1, (float*) param);
    }else{
        if (host) rotmg_real(*(double*) d1, *(double*) d2, *(double*) x1, *(const double*) y1, (double*) param);
        else rotmg_kernel<double><<<1, 1, 0, ctx->stream>>>((double*) d1, (double*) d2, (double*) x1, (const double*) y1, (double*) param);
    }
    if (!host) HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

}
