// hb_sptrsv.cu — sparse triangular solves and ILU(0) on sm_100a.
// Replaces cusparseSpSV / cusparseSpSM (reference gpu/hala_cuda_sparse_triangular.hpp:38-454) and cusparse?csrilu02
// (gpu/hala_gpu_ilu.hpp:45-99).  Semantics are the ones the reference's GPU path relies on: the matrix is a non-owning view of
// a GENERAL CSR of which only the `uplo` triangle is used (gpu_ilu hands the full ILU array to both triangular matrices,
// gpu_ilu.hpp:88-89); diag 'U' ignores a stored diagonal; values are read at solve time (they may change between calls).
//
// Scheme (no cuSPARSE, no per-level launches, no grid barrier):
//   analysis, once per (matrix, direction): level[i] = 1 + max level of the rows it depends on, computed by ONE sync-free
//     pass in natural order (a row waits on the level flags of its dependencies), then a counting sort of the rows by level.
//   solve op 'N': persistent CTAs claim chunks of the level-ordered row list through one atomic ticket, one thread per row.
//     A row's dependencies sit in earlier levels, i.e. earlier in the list, so they are finished or held by a resident CTA:
//     the row polls their per-row epoch flags (acquire) and never blocks — every lane of a warp makes non-blocking progress
//     in a common loop, so dependencies inside a warp cannot deadlock it.  Sums run left to right over the row, as
//     sparse_trsv_array does (sparse/hala_sparse_utils.hpp:283-311).
//   solve op 'T'/'C': the column-oriented form of sparse_trsv_array (:312-335) in natural order: row i waits until all
//     contributions to it have arrived (a per-row countdown), then scatters x_i * op(a_ij) with atomics.  First cut: the
//     order of the atomic additions is not fixed (results agree to round-off) and parallelism is limited to the rows in
//     flight; the reference reaches this path only from its small tests.
//   ILU(0): same level-ordered sync-free sweep; row i divides by the pivots of the rows it depends on and merges their upper
//     parts into its own (sorted rows, diagonal present) — the operations of factorize_ilu_array (:228-253) in the same order
//     per entry.
#include "hb_common.cuh"

static constexpr int TS_THREADS = 256;

struct hb_tri {
    hb_ctx *ctx = nullptr;
    int dtype = HB_F64, rows = 0, nnz = 0;
    bool lower = true, unit = false;
    const int *pntr = nullptr, *indx = nullptr;
    const void *vals = nullptr;
    // analysis of op 'N' (lazy): rows sorted by dependency level
    int *order = nullptr;
    int nlevels = 0;
    // run-time state
    void *xp = nullptr;             // op 'N': finished x_i in flag-in-data form (8-byte words {32-bit part, epoch tag}), 2 sizeof(T) per row
    unsigned epoch = 0;
    int *cnt0 = nullptr, *cnt = nullptr;   // op 'T'/'C': contributions each row waits for (analysis / working copy)
    void *acc = nullptr;            // op 'T'/'C': accumulated contributions
    unsigned int *ticket = nullptr; // chunk dispenser
    int grid = 0;
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p){
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v){
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void atomic_add_T(float *p, float v){ atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_T(double *p, double v){ atomicAdd(p, v); }
template<typename R> __device__ __forceinline__ void atomic_add_T(cplx<R> *p, cplx<R> v){ atomicAdd(&p->re, v.re); atomicAdd(&p->im, v.im); }

// chunk dispenser: every thread of the block gets the same chunk index (block-wide barrier inside)
__device__ __forceinline__ long long next_chunk(unsigned int *ticket){
    __shared__ unsigned int chunk_s;
    __syncthreads();
    if (threadIdx.x == 0) chunk_s = atomicAdd(ticket, 1u);
    __syncthreads();
    return (long long) chunk_s;
}

// ------------------------------------------------------------------------------------------------ analysis
// level[i] (>= 0) of every row for op 'N': natural order for the lower triangle, reversed for the upper one
__global__ void __launch_bounds__(TS_THREADS) tri_level_kernel(int n, int lower, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                               int *level, unsigned int *ticket, int *maxlevel){
    for (;;){
        const long long q = next_chunk(ticket) * TS_THREADS + threadIdx.x;
        if (q - threadIdx.x >= n) break;
        bool finished = q >= n;
        const int i = finished ? 0 : (lower ? (int) q : n - 1 - (int) q);
        int j = finished ? 0 : pntr[i], lv = 0;
        const int re = finished ? 0 : pntr[i + 1];
        while (!__all_sync(0xffffffffu, finished)){
            if (!finished){
                while (j < re){
                    const int c = indx[j];
                    if (lower ? c < i : c > i){
                        const int l = ld_acquire_gpu(level + c);
                        if (l < 0) break;
                        lv = max(lv, l + 1);
                    }
                    j++;
                }
                if (j == re){
                    st_release_gpu(level + i, lv);
                    atomicMax(maxlevel, lv);
                    finished = true;
                }
            }
        }
    }
}
__global__ void tri_hist_kernel(int n, const int *level, int *hist){
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(hist + level[i], 1);
}
// exclusive scan of hist[0..m) by one block (m = number of levels: thousands for stencils, n in the worst case)
__global__ void __launch_bounds__(1024) tri_scan_kernel(int m, int *hist){
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024){
        const int idx = base + threadIdx.x;
        const int v = idx < m ? hist[idx] : 0;
        int incl = v;
        for (int d = 1; d < 32; d <<= 1){ int t = __shfl_up_sync(0xffffffffu, incl, d); if ((threadIdx.x & 31) >= d) incl += t; }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32){
            int w = warp_tot[threadIdx.x], wi = w;
            for (int d = 1; d < 32; d <<= 1){ int t = __shfl_up_sync(0xffffffffu, wi, d); if (threadIdx.x >= d) wi += t; }
            warp_tot[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int excl = carry_s + warp_tot[threadIdx.x >> 5] + incl - v;
        if (idx < m) hist[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}
__global__ void tri_fill_kernel(int n, const int *level, int *cursor, int *order){
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) order[atomicAdd(cursor + level[i], 1)] = i;
}
// op 'T'/'C': number of off-diagonal entries of the used triangle in every COLUMN
__global__ void tri_colcount_kernel(int n, int lower, const int * __restrict__ pntr, const int * __restrict__ indx, int *cnt){
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        for (int j = pntr[i]; j < pntr[i + 1]; j++){
            const int c = indx[j];
            if (lower ? c < i : c > i) atomicAdd(cnt + c, 1);
        }
}

// ------------------------------------------------------------------------------------------------ solves
// x = alpha * T^-1 b, level-ordered, one thread per row.
// The critical path is one row per level, so a row must cost as few dependent memory round trips as possible:
//   * its entries are loaded TS_SEG at a time into registers (all loads of a segment in one round trip — and, because the
//     resident CTAs run a few levels ahead of the solve front, usually long before the dependencies are ready);
//   * a finished x_i is handed over in "flag-in-data" form: every 32-bit part of the value travels in an 8-byte word together
//     with the solve's epoch tag (8-byte stores are atomic), so a consumer gets readiness AND value from one relaxed load per
//     dependency, with no fence on either side (the protocol NCCL calls LL).  Measured against per-row flags + fences
//     (x store, fence, release flag / acquire flag, fence, x load): 12.0 ms -> see DESIGN.md for the 256^3 numbers;
//   * all pending dependencies of a segment are polled together; the sum still runs left to right.
static constexpr int TS_SEG = 8;
template<typename T> struct ll_parts { static constexpr int N = sizeof(T) / 4; };
template<typename T> __device__ __forceinline__ void ll_store(uint2 *slot, T v, unsigned tag){
    constexpr int NP = ll_parts<T>::N;
    unsigned w[NP];
    memcpy(w, &v, sizeof(T));
    if (NP == 1){
        asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" :: "l"(slot), "r"(w[0]), "r"(tag) : "memory");
    }else{
        #pragma unroll
        for (int p = 0; p < NP; p += 2)
            asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(slot + p), "r"(w[p]), "r"(tag), "r"(w[p + 1 < NP ? p + 1 : p]), "r"(tag) : "memory");
    }
}
template<typename T> __device__ __forceinline__ bool ll_load(const uint2 *slot, unsigned tag, T &v){
    constexpr int NP = ll_parts<T>::N;
    unsigned w[NP];
    bool ok = true;
    if (NP == 1){
        unsigned a, t;
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(t) : "l"(slot) : "memory");
        w[0] = a; ok = (t == tag);
    }else{
        #pragma unroll
        for (int p = 0; p < NP; p += 2){
            unsigned a0, t0, a1, t1;
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(t0), "=r"(a1), "=r"(t1) : "l"(slot + p) : "memory");
            w[p] = a0; w[p + 1 < NP ? p + 1 : p] = a1;
            ok = ok && (t0 == tag) && (t1 == tag);
        }
    }
    memcpy(&v, w, sizeof(T));
    return ok;
}
template<typename T>
__global__ void __launch_bounds__(TS_THREADS) tri_solve_n_kernel(int n, int lower, int unit, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                                 const T * __restrict__ vals, const int * __restrict__ order, scalar_arg<T> alpha_s,
                                                                 const T *b, long long incb, T *x, long long incx, uint2 *xp, unsigned epoch, unsigned int *ticket){
    constexpr int NP = ll_parts<T>::N;
    const T alpha = get_scalar(alpha_s);
    for (;;){
        const long long q = next_chunk(ticket) * TS_THREADS + threadIdx.x;
        if (q - threadIdx.x >= n) break;
        bool finished = q >= n;
        const int i = finished ? 0 : order[q];
        int j = finished ? 0 : pntr[i];
        const int re = finished ? 0 : pntr[i + 1];
        T s = zero_of<T>(), d = one_of<T>();
        const T ab = finished ? zero_of<T>() : hmul(alpha, b[(long long) i * incb]);
        int c[TS_SEG]; T v[TS_SEG], xv[TS_SEG];
        unsigned deps = 0, pend = 0;
        bool have = false;
        while (!__all_sync(0xffffffffu, finished)){
            if (!finished){
                if (!have && j < re){                            // next segment of the row into registers
                    #pragma unroll
                    for (int u = 0; u < TS_SEG; u++){
                        c[u] = -1;
                        if (j + u < re){ c[u] = __ldg(indx + j + u); v[u] = vals[j + u]; }
                    }
                    deps = 0;
                    #pragma unroll
                    for (int u = 0; u < TS_SEG; u++){
                        if (c[u] >= 0){
                            if (lower ? c[u] < i : c[u] > i) deps |= 1u << u;
                            else if (c[u] == i && !unit) d = v[u];
                        }
                    }
                    pend = deps;
                    have = true;
                }
                if (have){
                    #pragma unroll
                    for (int u = 0; u < TS_SEG; u++)
                        if ((pend >> u) & 1u){ if (ll_load<T>(xp + (size_t) c[u] * NP, epoch, xv[u])) pend &= ~(1u << u); }
                    if (pend == 0){
                        #pragma unroll
                        for (int u = 0; u < TS_SEG; u++) if ((deps >> u) & 1u) s = hfma(v[u], xv[u], s);
                        j += TS_SEG;
                        have = false;
                    }
                }
                if (!have && j >= re){
                    T r = hsub(ab, s);
                    if (!unit) r = hdiv(r, d);
                    ll_store<T>(xp + (size_t) i * NP, r, epoch);
                    x[(long long) i * incx] = r;
                    finished = true;
                }
            }
        }
    }
}
// x = alpha * op(T)^-1 b for op = T / C: natural order of the column-oriented sweep, atomic scatter of the contributions
template<typename T, bool CONJ>
__global__ void __launch_bounds__(TS_THREADS) tri_solve_t_kernel(int n, int lower, int unit, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                                 const T * __restrict__ vals, scalar_arg<T> alpha_s, const T *b, long long incb,
                                                                 T *x, long long incx, int *cnt, T *acc, unsigned int *ticket){
    const T alpha = get_scalar(alpha_s);
    for (;;){
        const long long q = next_chunk(ticket) * TS_THREADS + threadIdx.x;
        if (q - threadIdx.x >= n) break;
        bool finished = q >= n;
        // L^T is upper: rows are final from the last to the first; U^T is lower: first to last
        const int i = finished ? 0 : (lower ? n - 1 - (int) q : (int) q);
        while (!__all_sync(0xffffffffu, finished)){
            if (!finished && ld_acquire_gpu(cnt + i) == 0){
                T d = one_of<T>();
                const int rs = pntr[i], re = pntr[i + 1];
                if (!unit) for (int j = rs; j < re; j++) if (indx[j] == i) d = vals[j];
                T v = hsub(hmul(alpha, b[(long long) i * incb]), ld_cg_T(acc + i));
                if (!unit) v = hdiv(v, CONJ ? hconj(d) : d);
                x[(long long) i * incx] = v;
                for (int j = rs; j < re; j++){
                    const int c = indx[j];
                    if (lower ? c < i : c > i){
                        const T a = vals[j];
                        atomic_add_T(acc + c, hmul(v, CONJ ? hconj(a) : a));
                        __threadfence();
                        atomicSub(cnt + c, 1);
                    }
                }
                finished = true;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ ILU(0)
// diag[i] = position of the diagonal of row i; *bad = 1 + first row whose columns are not strictly increasing or that has no diagonal
__global__ void ilu_diag_kernel(int n, const int * __restrict__ pntr, const int * __restrict__ indx, int *diag, int *bad){
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x){
        int dpos = -1, ok = 1;
        for (int j = pntr[i]; j < pntr[i + 1]; j++){
            if (indx[j] == i) dpos = j;
            if (j > pntr[i] && indx[j] <= indx[j - 1]) ok = 0;
        }
        diag[i] = dpos;
        if (dpos < 0 || !ok) atomicMax(bad, i + 1);
    }
}
template<typename T>
__global__ void __launch_bounds__(TS_THREADS) ilu0_kernel(int n, const int * __restrict__ pntr, const int * __restrict__ indx, const int * __restrict__ diag,
                                                          const int * __restrict__ order, T *ilu, int *done, int epoch, unsigned int *ticket){
    for (;;){
        const long long q = next_chunk(ticket) * TS_THREADS + threadIdx.x;
        if (q - threadIdx.x >= n) break;
        bool finished = q >= n;
        const int i = finished ? 0 : order[q];
        int j = finished ? 0 : pntr[i];
        const int re = finished ? 0 : pntr[i + 1], dg = finished ? 0 : diag[i];
        while (!__all_sync(0xffffffffu, finished)){
            if (!finished){
                while (j < dg){                                 // the strictly-lower entries, ascending column = ascending pivot
                    const int k = indx[j];
                    if (ld_acquire_gpu(done + k) != epoch) break;
                    const int kd = diag[k], ke = pntr[k + 1];
                    const T l = hdiv(ilu[j], ld_cg_T(ilu + kd));
                    ilu[j] = l;
                    int ik = kd + 1, jk = j + 1;
                    while (ik < ke && jk < re){                 // merge row k's upper part into the rest of row i
                        const int ci = indx[ik], cj = indx[jk];
                        if (ci == cj){ ilu[jk] = hsub(ilu[jk], hmul(l, ld_cg_T(ilu + ik))); ik++; jk++; }
                        else if (ci < cj) ik++;
                        else jk++;
                    }
                    j++;
                }
                if (j >= dg){
                    __threadfence();
                    st_release_gpu(done + i, epoch);
                    finished = true;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
namespace {
int tri_grid(hb_ctx *ctx){ return ctx->num_sms * 4; }

// levels + level-ordered row list of op 'N' for the given structure
int analyse_levels(hb_ctx *ctx, int n, bool lower, const int *pntr, const int *indx, unsigned int *ticket, int **order_out, int *nlevels_out){
    int *level = nullptr, *hist = nullptr, *order = nullptr, *maxl = nullptr;
    struct temporaries { int **a, **b, **c, **d; bool keep_d = false;     // freed on every exit; `order` survives a successful one
        ~temporaries(){ cudaFree(*a); cudaFree(*b); cudaFree(*c); if (!keep_d) cudaFree(*d); } } tmp{&level, &hist, &maxl, &order};
    HB_CUDA(cudaMalloc((void**) &level, sizeof(int) * (size_t) (n + 1)));
    HB_CUDA(cudaMalloc((void**) &maxl, sizeof(int)));
    HB_CUDA(cudaMemsetAsync(level, 0xff, sizeof(int) * (size_t) (n + 1), ctx->stream));
    HB_CUDA(cudaMemsetAsync(maxl, 0, sizeof(int), ctx->stream));
    HB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ctx->stream));
    tri_level_kernel<<<tri_grid(ctx), TS_THREADS, 0, ctx->stream>>>(n, lower ? 1 : 0, pntr, indx, level, ticket, maxl);
    HB_LAUNCH_CHECK(ctx);
    int nlev = 0;
    HB_CUDA(cudaMemcpyAsync(&nlev, maxl, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    nlev += 1;
    HB_CUDA(cudaMalloc((void**) &hist, sizeof(int) * (size_t) (nlev + 1)));
    HB_CUDA(cudaMalloc((void**) &order, sizeof(int) * (size_t) (n > 0 ? n : 1)));
    HB_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * (size_t) (nlev + 1), ctx->stream));
    const int g = hb_grid_for(ctx, (size_t) n, 256, 8);
    tri_hist_kernel<<<g, 256, 0, ctx->stream>>>(n, level, hist);
    HB_LAUNCH_CHECK(ctx);
    tri_scan_kernel<<<1, 1024, 0, ctx->stream>>>(nlev, hist);
    HB_LAUNCH_CHECK(ctx);
    tri_fill_kernel<<<g, 256, 0, ctx->stream>>>(n, level, hist, order);
    HB_LAUNCH_CHECK(ctx);
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    tmp.keep_d = true;
    *order_out = order; *nlevels_out = nlev;
    return HB_OK;
}
int ensure_n_analysis(hb_tri *t){
    if (t->order || t->rows == 0) return HB_OK;
    return analyse_levels(t->ctx, t->rows, t->lower, t->pntr, t->indx, t->ticket, &t->order, &t->nlevels);
}
int ensure_t_analysis(hb_tri *t){
    if (t->cnt0 || t->rows == 0) return HB_OK;
    hb_ctx *ctx = t->ctx;
    const size_t n = (size_t) t->rows;
    HB_CUDA(cudaMalloc((void**) &t->cnt0, sizeof(int) * n));
    HB_CUDA(cudaMalloc((void**) &t->cnt, sizeof(int) * n));
    HB_CUDA(cudaMalloc(&t->acc, hb_dtype_size(t->dtype) * n));
    HB_CUDA(cudaMemsetAsync(t->cnt0, 0, sizeof(int) * n, ctx->stream));
    tri_colcount_kernel<<<hb_grid_for(ctx, n, 256, 8), 256, 0, ctx->stream>>>(t->rows, t->lower ? 1 : 0, t->pntr, t->indx, t->cnt0);
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

template<typename T>
int solve_typed(hb_ctx *ctx, hb_tri *t, char trans, scalar_arg<T> alpha, const T *b, long long incb, T *x, long long incx){
    const int n = t->rows;
    int rc;
    if (hb_is_n(trans)){
        if ((rc = ensure_n_analysis(t)) != HB_OK) return rc;
        HB_CUDA(cudaMemsetAsync(t->ticket, 0, sizeof(unsigned int), ctx->stream));      // after the analysis: it uses the same dispenser
        if (++t->epoch == 0){                                                           // tag wrap-around: start from a clean slate
            HB_CUDA(cudaMemsetAsync(t->xp, 0, 2 * sizeof(T) * (size_t) n, ctx->stream));
            t->epoch = 1;
        }
        tri_solve_n_kernel<T><<<t->grid, TS_THREADS, 0, ctx->stream>>>(n, t->lower ? 1 : 0, t->unit ? 1 : 0, t->pntr, t->indx, (const T*) t->vals, t->order,
                                                                       alpha, b, incb, x, incx, (uint2*) t->xp, t->epoch, t->ticket);
    }else{
        if ((rc = ensure_t_analysis(t)) != HB_OK) return rc;
        HB_CUDA(cudaMemsetAsync(t->ticket, 0, sizeof(unsigned int), ctx->stream));
        HB_CUDA(cudaMemcpyAsync(t->cnt, t->cnt0, sizeof(int) * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
        HB_CUDA(cudaMemsetAsync(t->acc, 0, sizeof(T) * (size_t) n, ctx->stream));
        if (hb_is_c(trans) && is_cplx<T>::value)
            tri_solve_t_kernel<T, true><<<t->grid, TS_THREADS, 0, ctx->stream>>>(n, t->lower ? 1 : 0, t->unit ? 1 : 0, t->pntr, t->indx, (const T*) t->vals,
                                                                                 alpha, b, incb, x, incx, t->cnt, (T*) t->acc, t->ticket);
        else
            tri_solve_t_kernel<T, false><<<t->grid, TS_THREADS, 0, ctx->stream>>>(n, t->lower ? 1 : 0, t->unit ? 1 : 0, t->pntr, t->indx, (const T*) t->vals,
                                                                                  alpha, b, incb, x, incx, t->cnt, (T*) t->acc, t->ticket);
    }
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
}

extern "C" {

int hb_tri_create(hb_ctx *ctx, int dtype, char uplo, char diag, int rows, int nnz, const int *pntr, const int *indx, const void *vals, hb_tri **out){
    HB_ARG(ctx && out, "null");
    HB_ARG(dtype >= HB_F32 && dtype <= HB_C64, "dtype");
    HB_ARG(uplo == 'L' || uplo == 'l' || uplo == 'U' || uplo == 'u', "uplo must be L or U");
    HB_ARG(diag == 'N' || diag == 'n' || diag == 'U' || diag == 'u', "diag must be N or U");
    HB_ARG(rows >= 0 && nnz >= 0, "negative size");
    HB_ARG(rows == 0 || (pntr && (nnz == 0 || (indx && vals))), "null array");
    hb_tri *t = new hb_tri();
    t->ctx = ctx; t->dtype = dtype; t->rows = rows; t->nnz = nnz;
    t->lower = (uplo == 'L' || uplo == 'l'); t->unit = (diag == 'U' || diag == 'u');
    t->pntr = pntr; t->indx = indx; t->vals = vals;
    t->grid = tri_grid(ctx);
    const size_t xp_bytes = 2 * hb_dtype_size(dtype) * (size_t) (rows > 0 ? rows : 1);
    cudaError_t e = cudaMalloc((void**) &t->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&t->xp, xp_bytes);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->xp, 0, xp_bytes, ctx->stream);
    if (e != cudaSuccess){ hb_tri_destroy(t); return hb_cuda_fail(e, "hb_tri_create"); }
    *out = t;
    return HB_OK;
}

int hb_tri_destroy(hb_tri *t){
    if (!t) return HB_OK;
    cudaFree(t->order); cudaFree(t->xp); cudaFree(t->cnt0); cudaFree(t->cnt); cudaFree(t->acc); cudaFree(t->ticket);
    cudaGetLastError();
    delete t;
    return HB_OK;
}

int hb_tri_info(const hb_tri *t, int *rows, int *nnz, int *nlevels){
    HB_ARG(t, "null");
    if (rows) *rows = t->rows;
    if (nnz) *nnz = t->nnz;
    if (nlevels) *nlevels = t->nlevels;
    return HB_OK;
}

int hb_sptrsv(hb_ctx *ctx, hb_tri *t, char trans, const void *alpha, const void *b, int incb, void *x, int incx){
    HB_ARG(ctx && t && alpha, "null");
    HB_ARG(incb != 0 && incx != 0, "zero stride");
    if (t->rows == 0) return HB_OK;
    HB_ARG(b && x, "null vector");
    HB_DISPATCH(t->dtype, { return solve_typed<T>(ctx, t, trans, make_scalar<T>(ctx, alpha), (const T*) b, incb, (T*) x, incx); });
    return HB_OK;
}

// in place on B: transb 'N' -> B is rows x nrhs (column-major, ldb >= rows); otherwise B is nrhs x rows (ldb >= nrhs) and every ROW
// is a right-hand side (the conjugation cancels on both sides: reference sparse/hala_sparse_structs.hpp:294-301)
int hb_sptrsm(hb_ctx *ctx, hb_tri *t, char transa, char transb, int nrhs, const void *alpha, void *B, int ldb){
    HB_ARG(ctx && t && alpha, "null");
    HB_ARG(nrhs >= 0, "negative nrhs");
    if (t->rows == 0 || nrhs == 0) return HB_OK;
    HB_ARG(B, "null matrix");
    const bool bn = hb_is_n(transb);
    HB_ARG(ldb >= (bn ? t->rows : nrhs), "ldb too small");
    const size_t es = hb_dtype_size(t->dtype);
    for (int k = 0; k < nrhs; k++){
        char *col = (char*) B + es * (bn ? (size_t) k * (size_t) ldb : (size_t) k);
        const int inc = bn ? 1 : ldb;
        int rc = hb_sptrsv(ctx, t, transa, alpha, col, inc, col, inc);
        if (rc != HB_OK) return rc;
    }
    return HB_OK;
}

// ILU(0) of a CSR with sorted rows and a full diagonal; `ilu` receives the factors in the pattern of the matrix (may alias vals)
int hb_ilu0(hb_ctx *ctx, int dtype, int rows, int nnz, const int *pntr, const int *indx, const void *vals, void *ilu){
    HB_ARG(ctx, "null");
    HB_ARG(dtype >= HB_F32 && dtype <= HB_C64, "dtype");
    HB_ARG(rows >= 0 && nnz >= 0, "negative size");
    if (rows == 0) return HB_OK;
    HB_ARG(pntr && indx && vals && ilu, "null array");
    const size_t es = hb_dtype_size(dtype);
    if (ilu != vals) HB_CUDA(cudaMemcpyAsync(ilu, vals, es * (size_t) nnz, cudaMemcpyDeviceToDevice, ctx->stream));
    int *diag = nullptr, *flags = nullptr, *order = nullptr;
    unsigned int *ticket = nullptr;
    struct temporaries { int **a, **b, **c; unsigned int **d; ~temporaries(){ cudaFree(*a); cudaFree(*b); cudaFree(*c); cudaFree(*d); } } tmp{&diag, &flags, &order, &ticket};
    HB_CUDA(cudaMalloc((void**) &diag, sizeof(int) * (size_t) rows));
    HB_CUDA(cudaMalloc((void**) &flags, sizeof(int) * ((size_t) rows + 2)));      // [0] bad-row marker, [1..] done flags
    HB_CUDA(cudaMalloc((void**) &ticket, sizeof(unsigned int)));
    int rc = HB_OK, bad = 0, nlev = 0;
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t) rows + 2), ctx->stream);
    if (e == cudaSuccess){
        ilu_diag_kernel<<<hb_grid_for(ctx, (size_t) rows, 256, 8), 256, 0, ctx->stream>>>(rows, pntr, indx, diag, flags);
        ctx->launches++;
        e = cudaMemcpyAsync(&bad, flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = hb_cuda_fail(e, "hb_ilu0 setup");
    else if (bad){
        hb_set_error("hb_ilu0: row " + std::to_string(bad - 1) + " has no diagonal entry or unsorted columns");
        rc = HB_ERR_ARG;
    }
    if (rc == HB_OK) rc = analyse_levels(ctx, rows, true, pntr, indx, ticket, &order, &nlev);
    if (rc == HB_OK){
        e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), ctx->stream);
        if (e == cudaSuccess){
            HB_DISPATCH(dtype, (ilu0_kernel<T><<<tri_grid(ctx), TS_THREADS, 0, ctx->stream>>>(rows, pntr, indx, diag, order, (T*) ilu, flags + 1, 1, ticket)));
            ctx->launches++;
            e = cudaStreamSynchronize(ctx->stream);
        }
        if (e != cudaSuccess) rc = hb_cuda_fail(e, "hb_ilu0");
    }
    return rc;
}

}
