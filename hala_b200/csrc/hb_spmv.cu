// hb_spmv.cu — CSR sparse matrix-vector product on sm_100a for float, double, complex<float/double>.
// Replaces cusparseCreateCsr / cusparseSpMV_bufferSize / cusparseSpMV behind gpu_sparse_matrix
// (reference gpu/hala_cuda_sparse_general.hpp:72-90, 245-277) and restates, per row, sparse_gemv_array
// (sparse/hala_sparse_utils.hpp:103-118).
//
// op 'N', two kernel families (chosen per matrix by a one-time analysis, hb_csr_create):
//   * staged tiles ("thread per row", with warp-cooperative handling of long row segments):
//       a CTA owns ROWS consecutive rows; their CSR slice [pntr[r0], pntr[r0+ROWS)) is contiguous in memory and is
//       streamed into shared memory in chunks of CH non-zeros with 128-bit, perfectly coalesced, evict-first loads
//       of col_idx and values; then thread t walks row r0+t inside shared memory, left to right (the reference's
//       summation order), gathering x through L1/L2.  Adjacent lanes own adjacent rows, so for banded/stencil
//       matrices the gather of step k is x[c_k + lane]: one or two 128-byte lines per warp instruction instead of
//       one line per lane.  Row segments longer than LONGSEG inside a chunk are handed to whole warps
//       (shuffle reduction) so that heavy-tailed row-length distributions do not serialise on one lane.
//   * row-vector ("sub-warp / warp per row"): TPR lanes per row, strided walk, shuffle reduction; used for matrices
//       whose mean row is long, and as the unaligned fallback.
// Both fuse an optional <x, y> dot (conjugated) into the same pass for CG (hb_spmv_dot).
// op 'T' / 'C': the op 'N' kernels on a cached CSR of A^T (hb_transpose.cu); in scatter mode y = beta y, then
//   atomic scatter y[col] += alpha x[row] op(val).
#include "hb_common.cuh"
#include "hb_spmv_pipe.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

// ------------------------------------------------------------------------------------------------ analysis
__global__ void csr_analyse_kernel(int rows, const int *pntr, int *stats){
    int mx = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
        mx = max(mx, pntr[i + 1] - pntr[i]);
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(&stats[0], mx);
}

// ------------------------------------------------------------------------------------------------ staging helpers
template<typename T> struct stage_cfg;      // CH = non-zeros per chunk
template<> struct stage_cfg<float>        { static constexpr int CH = 4096; };
template<> struct stage_cfg<double>       { static constexpr int CH = 4096; };
template<> struct stage_cfg<cplx<float>>  { static constexpr int CH = 4096; };
template<> struct stage_cfg<cplx<double>> { static constexpr int CH = 2048; };

static constexpr int LONGSEG = 128;         // row segment (within one chunk) handed to a whole warp
static constexpr int MAXLONG = 4096 / LONGSEG;

__device__ __forceinline__ int4 ldcs_int4(const int *p){ return __ldcs(reinterpret_cast<const int4*>(p)); }

// copy 4 consecutive values global -> shared with 128-bit transactions (both sides 16-byte aligned)
template<typename T> __device__ __forceinline__ void stage4(const T *g, T *s);
template<> __device__ __forceinline__ void stage4<float>(const float *g, float *s){
    *reinterpret_cast<float4*>(s) = __ldcs(reinterpret_cast<const float4*>(g));
}
template<> __device__ __forceinline__ void stage4<double>(const double *g, double *s){
    double2 a = __ldcs(reinterpret_cast<const double2*>(g)), b = __ldcs(reinterpret_cast<const double2*>(g) + 1);
    reinterpret_cast<double2*>(s)[0] = a; reinterpret_cast<double2*>(s)[1] = b;
}
template<> __device__ __forceinline__ void stage4<cplx<float>>(const cplx<float> *g, cplx<float> *s){
    float4 a = __ldcs(reinterpret_cast<const float4*>(g)), b = __ldcs(reinterpret_cast<const float4*>(g) + 1);
    reinterpret_cast<float4*>(s)[0] = a; reinterpret_cast<float4*>(s)[1] = b;
}
template<> __device__ __forceinline__ void stage4<cplx<double>>(const cplx<double> *g, cplx<double> *s){
    const double2 *gp = reinterpret_cast<const double2*>(g);
    double2 a = __ldcs(gp), b = __ldcs(gp + 1), c = __ldcs(gp + 2), d = __ldcs(gp + 3);
    double2 *sp = reinterpret_cast<double2*>(s);
    sp[0] = a; sp[1] = b; sp[2] = c; sp[3] = d;
}

// ------------------------------------------------------------------------------------------------ staged tiles, op N
// DOT: also accumulate conj(x[row]) * (A x)[row] and publish the grid-wide sum to *dot_out (alpha = 1, beta = 0 semantics).
template<typename T, int ROWS, bool VEC, bool DOT>
__global__ void __launch_bounds__(ROWS) spmv_tiles_kernel(int rows, int nnz, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                          const T * __restrict__ vals, const T * __restrict__ x, T *y,
                                                          scalar_arg<T> alpha_s, scalar_arg<T> beta_s,
                                                          void *partials_v, unsigned int *ticket, T *dot_out, const int *skip_flag){
    constexpr int CH = stage_cfg<T>::CH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T   *sv = reinterpret_cast<T*>(smem_raw);
    int *sc = reinterpret_cast<int*>(smem_raw + sizeof(T) * CH);
    __shared__ int   long_row[MAXLONG], long_lo[MAXLONG], long_hi[MAXLONG];
    __shared__ int   long_count;
    __shared__ T     red[32];
    T *long_sum = reinterpret_cast<T*>(smem_raw + (sizeof(T) + sizeof(int)) * CH);   // MAXLONG entries

    if (skip_flag && *skip_flag) return;

    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = (rows + ROWS - 1) / ROWS;
    T dot_acc = zero_of<T>();

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x){
        const int r0 = tile * ROWS, row = r0 + tid;
        const int rend = min(r0 + ROWS, rows);
        const int nz0 = __ldg(pntr + r0), nz1 = __ldg(pntr + rend);
        int rs = nz1, re = nz1;
        if (row < rows){ rs = __ldg(pntr + row); re = __ldg(pntr + row + 1); }
        T sum = zero_of<T>();
        const int abase = VEC ? (nz0 & ~3) : nz0;
        for (int cb = abase; cb < nz1; cb += CH){
            const int cend = min(cb + CH, nz1);
            if (tid == 0) long_count = 0;
            // ---- stage chunk [cb, cend) : coalesced, 128-bit when aligned
            if (VEC){
                const int nq = (cend - cb + 3) >> 2;
                for (int q = tid; q < nq; q += ROWS){
                    const int e = cb + 4 * q;
                    if (e + 3 < nnz){
                        *reinterpret_cast<int4*>(sc + 4 * q) = ldcs_int4(indx + e);
                        stage4<T>(vals + e, sv + 4 * q);
                    }else{
                        for (int k = 0; k < 4; k++) if (e + k < nnz){ sc[4 * q + k] = indx[e + k]; sv[4 * q + k] = vals[e + k]; }
                    }
                }
            }else{
                for (int i = tid; i < cend - cb; i += ROWS){ sc[i] = __ldcs(indx + cb + i); sv[i] = ld_stream(vals + cb + i); }
            }
            __syncthreads();
            // ---- consume: thread per row, left to right
            const int lo = max(rs, cb) - cb, hi = min(re, cend) - cb;
            if (hi - lo >= LONGSEG){
                int slot = atomicAdd(&long_count, 1);
                long_row[slot] = tid; long_lo[slot] = lo; long_hi[slot] = hi;
            }else{
                #pragma unroll 4
                for (int j = lo; j < hi; j++) sum = hfma(sv[j], ld_ro(x + sc[j]), sum);
            }
            __syncthreads();
            const int nlong = long_count;
            if (nlong > 0){         // block-uniform
                for (int s = warp; s < nlong; s += ROWS / 32){
                    T part = zero_of<T>();
                    for (int j = long_lo[s] + lane; j < long_hi[s]; j += 32) part = hfma(sv[j], ld_ro(x + sc[j]), part);
                    part = warp_sum(part);
                    if (lane == 0) long_sum[s] = part;
                }
                __syncthreads();
                for (int s = 0; s < nlong; s++) if (long_row[s] == tid) sum = hadd(sum, long_sum[s]);
                __syncthreads();
            }
        }
        if (row < rows){
            if (DOT){
                y[row] = sum;
                dot_acc = hfma(hconj(ld_ro(x + row)), sum, dot_acc);
            }else{
                T out = hmul(alpha, sum);
                if (use_beta) out = hfma(beta, y[row], out);
                y[row] = out;
            }
        }
    }
    if (DOT){
        T *partials = reinterpret_cast<T*>(partials_v);
        T b = block_sum(dot_acc, red);
        if (tid == 0) partials[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            T total = sum_partials<T>(partials, gridDim.x, 1, red);
            if (tid == 0) *dot_out = total;
        }
    }
}

// ------------------------------------------------------------------------------------------------ row-vector, op N
static constexpr int RV_UNR = 4;
template<typename T, int TPR, bool DOT>
__global__ void __launch_bounds__(256) spmv_rowvec_kernel(int rows, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                          const T * __restrict__ vals, const T * __restrict__ x, T *y,
                                                          scalar_arg<T> alpha_s, scalar_arg<T> beta_s,
                                                          void *partials_v, unsigned int *ticket, T *dot_out, const int *skip_flag){
    __shared__ T red[32];
    if (skip_flag && *skip_flag) return;
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    const int sub = threadIdx.x % TPR;
    const long long group = (blockIdx.x * (long long) blockDim.x + threadIdx.x) / TPR;
    const long long ngroups = (long long) gridDim.x * blockDim.x / TPR;
    T dot_acc = zero_of<T>();
    // every lane of a warp runs the same number of outer iterations (rows padded up) so the shuffles stay converged
    const long long rows_pad = ((long long) rows + (32 / TPR) - 1) / (32 / TPR) * (32 / TPR);
    for (long long row = group; row < rows_pad; row += ngroups){
        T sum = zero_of<T>();
        if (row < rows){
            const int rs = __ldg(pntr + row), re = __ldg(pntr + row + 1);
            // RV_UNR entries per lane per step, all col_idx/value loads issued before the first gather, all gathers before
            // the first FMA: a row of up to RV_UNR*TPR entries costs ONE memory round trip instead of one per entry.
            for (int base = rs + sub; base < re; base += RV_UNR * TPR){
                int c[RV_UNR]; T v[RV_UNR], xv[RV_UNR];
                #pragma unroll
                for (int u = 0; u < RV_UNR; u++){
                    const int j = base + u * TPR;
                    c[u] = (j < re) ? __ldcs(indx + j) : -1;
                    v[u] = (j < re) ? ld_stream(vals + j) : zero_of<T>();
                }
                #pragma unroll
                for (int u = 0; u < RV_UNR; u++) xv[u] = (c[u] >= 0) ? ld_ro(x + c[u]) : zero_of<T>();
                #pragma unroll
                for (int u = 0; u < RV_UNR; u++) sum = hfma(v[u], xv[u], sum);
            }
        }
        #pragma unroll
        for (int d = TPR / 2; d > 0; d >>= 1) sum = hadd(sum, shfl_down(sum, d));
        if (sub == 0 && row < rows){
            if (DOT){
                y[row] = sum;
                dot_acc = hfma(hconj(ld_ro(x + row)), sum, dot_acc);
            }else{
                T out = hmul(alpha, sum);
                if (use_beta) out = hfma(beta, y[row], out);
                y[row] = out;
            }
        }
    }
    if (DOT){
        T *partials = reinterpret_cast<T*>(partials_v);
        T b = block_sum(dot_acc, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            T total = sum_partials<T>(partials, gridDim.x, 1, red);
            if (threadIdx.x == 0) *dot_out = total;
        }
    }
}

// ------------------------------------------------------------------------------------------------ op T / C (scatter)
template<typename T> __global__ void scale_or_zero_kernel(int n, scalar_arg<T> beta_s, T *y){
    const T beta = get_scalar(beta_s);
    const bool z = hiszero(beta);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        y[i] = z ? zero_of<T>() : hmul(beta, y[i]);
}
__device__ __forceinline__ void atomic_add(float *p, float v){ atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double *p, double v){ atomicAdd(p, v); }
template<typename R> __device__ __forceinline__ void atomic_add(cplx<R> *p, cplx<R> v){
    atomicAdd(&p->re, v.re); atomicAdd(&p->im, v.im);
}
template<typename T, int TPR, bool CONJ>
__global__ void __launch_bounds__(256) spmv_trans_kernel(int rows, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                         const T * __restrict__ vals, const T * __restrict__ x, T *y, scalar_arg<T> alpha_s){
    const T alpha = get_scalar(alpha_s);
    const int sub = threadIdx.x % TPR;
    const long long group = (blockIdx.x * (long long) blockDim.x + threadIdx.x) / TPR;
    const long long ngroups = (long long) gridDim.x * blockDim.x / TPR;
    for (long long row = group; row < rows; row += ngroups){
        const int rs = __ldg(pntr + row), re = __ldg(pntr + row + 1);
        const T ax = hmul(alpha, ld_ro(x + row));
        for (int j = rs + sub; j < re; j += TPR){
            T v = ld_stream(vals + j);
            atomic_add(y + __ldcs(indx + j), hmul(ax, CONJ ? hconj(v) : v));
        }
    }
}

// ------------------------------------------------------------------------------------------------ host side
static inline bool aligned16p(const void *p){ return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template<typename T, int ROWS, bool DOT>
static int launch_tiles(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    constexpr int CH = stage_cfg<T>::CH;
    const size_t smem = (sizeof(T) + sizeof(int)) * CH + sizeof(T) * MAXLONG;
    const int ntiles = (A->rows + ROWS - 1) / ROWS;
    const int per_sm = (int) std::min<size_t>(2048 / ROWS, (220 * 1024) / (smem + 1024));
    int grid = std::min(ntiles, ctx->num_sms * per_sm);
    if (grid < 1) grid = 1;
    if (A->vec_aligned){
        static per_device_flag attr_set;
        auto k = spmv_tiles_kernel<T, ROWS, true, DOT>;
        if (attr_set.first_time(ctx->device)) HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k<<<grid, ROWS, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, A->indx, (const T*) A->vals, x, y, alpha, beta,
                                             ctx->partials, ctx->tickets + 1, dot_out, skip);
    }else{
        static per_device_flag attr_set;
        auto k = spmv_tiles_kernel<T, ROWS, false, DOT>;
        if (attr_set.first_time(ctx->device)) HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        k<<<grid, ROWS, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, A->indx, (const T*) A->vals, x, y, alpha, beta,
                                             ctx->partials, ctx->tickets + 1, dot_out, skip);
    }
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

template<typename T, int TPR, bool DOT>
static int launch_rowvec(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    const long long threads = (long long) A->rows * TPR;
    int grid = (int) std::min<long long>((threads + 255) / 256, (long long) ctx->num_sms * 8);
    if (grid < 1) grid = 1;
    spmv_rowvec_kernel<T, TPR, DOT><<<grid, 256, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, (const T*) A->vals, x, y, alpha, beta,
                                                                   ctx->partials, ctx->tickets + 1, dot_out, skip);
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

// ---- streaming pipeline launcher. Two (THREADS, STAGES) configurations are instantiated (HB_PIPE_CFG picks one for probing).
template<int CFG> struct pipe_cfg;
template<> struct pipe_cfg<0> { static constexpr int THREADS = 128, STAGES = 3; };
template<> struct pipe_cfg<1> { static constexpr int THREADS = 256, STAGES = 3; };
static size_t pipe_smem_bytes(int threads, int tpr, int stages, size_t es, int slot_div = 1){
    const size_t cap = (size_t) threads * (es == 16 ? 4 : 8) / slot_div, rows = threads / tpr;
    return stages * ((cap + 4) * (es + sizeof(int)) + (rows + 4) * sizeof(int));
}
// lanes per row: the smallest power of two that keeps a typical row within ~7 entries per lane (<= two batches of four)
static constexpr int HB_HEAVY_TPR_SHIFT = 0;     // notches added to lanes-per-row for heavy-tailed matrices (HB_PIPE_TPR_SHIFT overrides)
static int pipe_tpr(double mean, size_t es){
    const double slots = es == 16 ? 4.0 : 8.0;
    return mean > 7.5 * slots ? 16 : mean > 3.75 * slots ? 8 : mean > 1.9 * slots ? 4 : mean > 0.94 * slots ? 2 : 1;
}
template<typename T, int CFG, int TPR, bool DOT>
static int launch_pipe_cfg(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    using C = pipe_cfg<CFG>;
    const size_t smem = pipe_smem_bytes(C::THREADS, TPR, C::STAGES, sizeof(T));
    auto k = spmv_pipe_kernel<T, C::THREADS, TPR, C::STAGES, DOT>;
    k<<<A->pipe_grid[CFG], C::THREADS, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, A->indx, (const T*) A->vals, x, y, alpha, beta,
                                                            A->pipe_contiguous ? A->cta_rows[CFG] : nullptr, ctx->partials, ctx->tickets + 1, dot_out, skip,
                                                            (TPR <= 2 && A->rows == A->cols) ? 2 : 0, A->cols,
                                                            DOT ? (const peer_view*) ctx->peer_hook : nullptr, ctx->peer_epoch, (size_t) 0, (size_t) 0, (size_t) 0,
                                                            (DOT && ctx->peer_hook && !A->pipe_contiguous) ? ctx->peer_trot : 0,
                                                            (DOT && ctx->peer_hook && !A->pipe_contiguous) ? ctx->peer_twait : 0, vsplit_view());
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
template<typename T, int CFG, bool DOT>
static int launch_pipe(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    switch (A->tpr){
        case 16: return launch_pipe_cfg<T, CFG, 16, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        case 8:  return launch_pipe_cfg<T, CFG, 8, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        case 4:  return launch_pipe_cfg<T, CFG, 4, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        case 2:  return launch_pipe_cfg<T, CFG, 2, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        default: return launch_pipe_cfg<T, CFG, 1, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    }
}

// ---- interleaved multi right-hand-side product (hb_spmm fast path): Ct[row][kb] = sum_j a_ij Bt[col_j][kb], kb < NBP
template<typename T, int CFG, int TPR, int NBP>
static int launch_pipe_mm_cfg(hb_ctx *ctx, const hb_csr *A, const T *Bt, size_t ldbt, T *Ct, size_t ldct){
    using C = pipe_cfg<CFG>;
    const size_t smem = pipe_smem_bytes(C::THREADS, TPR, C::STAGES, sizeof(T));
    auto k = spmv_pipe_kernel<T, C::THREADS, TPR, C::STAGES, false, NBP>;
    static per_device_flag configured;
    if (configured.first_time(ctx->device)){
        HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    scalar_arg<T> one; one.value = one_of<T>(); one.dev = nullptr;
    scalar_arg<T> zero; zero.value = zero_of<T>(); zero.dev = nullptr;
    k<<<A->pipe_grid[CFG], C::THREADS, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, A->indx, (const T*) A->vals, Bt, Ct, one, zero,
                                                            nullptr, ctx->partials, ctx->tickets + 1, nullptr, nullptr, 0, A->cols, nullptr, 0ull, ldbt, ldct, (size_t) 0, 0, 0, vsplit_view());
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
template<typename T, int CFG, int NBP>
static int launch_pipe_mm(hb_ctx *ctx, const hb_csr *A, const T *Bt, size_t ldbt, T *Ct, size_t ldct){
    switch (A->tpr){
        case 16: return launch_pipe_mm_cfg<T, CFG, 16, NBP>(ctx, A, Bt, ldbt, Ct, ldct);
        case 8:  return launch_pipe_mm_cfg<T, CFG, 8, NBP>(ctx, A, Bt, ldbt, Ct, ldct);
        case 4:  return launch_pipe_mm_cfg<T, CFG, 4, NBP>(ctx, A, Bt, ldbt, Ct, ldct);
        case 2:  return launch_pipe_mm_cfg<T, CFG, 2, NBP>(ctx, A, Bt, ldbt, Ct, ldct);
        default: return launch_pipe_mm_cfg<T, CFG, 1, NBP>(ctx, A, Bt, ldbt, Ct, ldct);
    }
}
// ---- multi right-hand-side product, lane-per-column form (spmv_pipe_kernel<..., NBP, LPC = true>): B and C used where they lie
static int pipe_mm_tpr(const hb_csr *A, int nbp){
    const int t = A->tpr;
    return t > nbp ? nbp : t;
}
template<typename T, int CFG, int TPR, int NBP>
static int launch_pipe_lpc_cfg(hb_ctx *ctx, const hb_csr *A, int nb, const T *B, size_t sxr, size_t sxc, scalar_arg<T> alpha, scalar_arg<T> beta, T *Cm, size_t ldc){
    using C = pipe_cfg<CFG>;
    const size_t smem = pipe_smem_bytes(C::THREADS, TPR, C::STAGES, sizeof(T));
    auto k = spmv_pipe_kernel<T, C::THREADS, TPR, C::STAGES, false, NBP, true>;
    static int occ_by_device[256];                       // 0: not asked yet on that device
    int &occ = occ_by_device[(unsigned) ctx->device & 255u];
    if (occ <= 0){
        HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int n = 0;
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, C::THREADS, smem));
        occ = n < 1 ? 1 : n;
    }
    const int tile_rows = C::THREADS / TPR;
    const long long ntiles = ((long long) A->rows + tile_rows - 1) / tile_rows;
    const int grid = (int) std::min<long long>(ntiles, (long long) ctx->num_sms * occ);
    k<<<grid, C::THREADS, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, A->indx, (const T*) A->vals, B, Cm, alpha, beta,
                                               nullptr, ctx->partials, ctx->tickets + 1, nullptr, nullptr, nb, A->cols, nullptr, 0ull, sxr, ldc, sxc, 0, 0, vsplit_view());
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
template<typename T, int CFG, int NBP>
static int launch_pipe_lpc(hb_ctx *ctx, const hb_csr *A, int nb, const T *B, size_t sxr, size_t sxc, scalar_arg<T> alpha, scalar_arg<T> beta, T *Cm, size_t ldc){
    switch (pipe_mm_tpr(A, NBP)){
        case 4:  return launch_pipe_lpc_cfg<T, CFG, 4, NBP>(ctx, A, nb, B, sxr, sxc, alpha, beta, Cm, ldc);
        case 2:  return launch_pipe_lpc_cfg<T, CFG, 2, NBP>(ctx, A, nb, B, sxr, sxc, alpha, beta, Cm, ldc);
        default: return launch_pipe_lpc_cfg<T, CFG, 1, NBP>(ctx, A, nb, B, sxr, sxc, alpha, beta, Cm, ldc);
    }
}
int hb_spmv_variant(const hb_csr *A);
// can the lane-per-column streaming kernel take this matrix with blocks of nbp right-hand sides?  (every tile fits its stage)
bool hb_spmm_lpc_ok(const hb_csr *A, int nbp){
    if (hb_spmv_variant(A) != 3) return false;
    const size_t es = hb_dtype_size(A->dtype);
    const int threads = A->pipe_cfg == 0 ? pipe_cfg<0>::THREADS : pipe_cfg<1>::THREADS;
    const int rows_per_tile = threads / pipe_mm_tpr(A, nbp);
    const long long cap = (long long) threads * (es == 16 ? 4 : 8);
    return (long long) rows_per_tile * A->max_row_nnz + 4 <= cap;
}
// C[:, 0..nb) = alpha A op(B)[:, 0..nb) + beta C, nb <= nbp (= 4); operand (c, kb) at B[c * sxr + kb * sxc]; alpha / beta host or device
// pointers according to the pointer mode
int hb_spmm_lpc(hb_ctx *ctx, const hb_csr *A, int nbp, int nb, const void *B, size_t sxr, size_t sxc, const void *alpha, const void *beta, void *Cm, size_t ldc){
    HB_DISPATCH(A->dtype, {
        scalar_arg<T> a = make_scalar<T>(ctx, alpha), b = make_scalar<T>(ctx, beta);
        HB_ARG(nbp == 4, "lane-per-column blocks are 4 wide (8-wide measured 1.5x slower per column on B200)");
        if (A->pipe_cfg == 0) return launch_pipe_lpc<T, 0, 4>(ctx, A, nb, (const T*) B, sxr, sxc, a, b, (T*) Cm, ldc);
        return launch_pipe_lpc<T, 1, 4>(ctx, A, nb, (const T*) B, sxr, sxc, a, b, (T*) Cm, ldc);
    });
    return HB_OK;
}
// can the interleaved streaming kernel take this matrix?  (pipeline available, every tile fits its stage, no warp / CTA rows)
bool hb_spmm_interleaved_ok(const hb_csr *A){
    if (hb_spmv_variant(A) != 3) return false;
    const size_t es = hb_dtype_size(A->dtype);
    const int threads = A->pipe_cfg == 0 ? pipe_cfg<0>::THREADS : pipe_cfg<1>::THREADS;
    const int rows_per_tile = threads / A->tpr;
    const long long cap = (long long) threads * (es == 16 ? 4 : 8);
    return A->max_row_nnz < PIPE_WARPROW && (long long) rows_per_tile * A->max_row_nnz + 4 <= cap;
}
// nbp = 4 or 8 interleaved right-hand sides (a multiple of the 128-bit packet); Bt: cols x nbp (ldbt), Ct: rows x nbp (ldct)
int hb_spmm_interleaved(hb_ctx *ctx, const hb_csr *A, int nbp, const void *Bt, size_t ldbt, void *Ct, size_t ldct){
    HB_DISPATCH(A->dtype, {
        if (A->pipe_cfg == 0) return nbp == 8 ? launch_pipe_mm<T, 0, 8>(ctx, A, (const T*) Bt, ldbt, (T*) Ct, ldct) : launch_pipe_mm<T, 0, 4>(ctx, A, (const T*) Bt, ldbt, (T*) Ct, ldct);
        return nbp == 8 ? launch_pipe_mm<T, 1, 8>(ctx, A, (const T*) Bt, ldbt, (T*) Ct, ldct) : launch_pipe_mm<T, 1, 4>(ctx, A, (const T*) Bt, ldbt, (T*) Ct, ldct);
    });
    return HB_OK;
}

// Shared memory and L1 share one 256 KB array per SM, and every gather of x that misses L1 needs an L1 line to land in: a kernel that
// takes all the shared memory it can leaves ~30 KB of L1, i.e. a couple of hundred misses in flight per SM, and the x gather then runs
// at a fraction of what the L1TEX pipe can do (power-law matrix: 548 us with the largest carve-out, of which 383 us are the gathers).
// So the carve-out is set to what `want` resident CTAs need and no more; want <= 0: as many CTAs as fit (largest carve-out).
// Returns the number of CTAs per SM to size the grid with.
template<typename K>
static int pipe_configure(K kernel, int threads, size_t smem, int want){
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess){ cudaGetLastError(); return 0; }
    int carve = cudaSharedmemCarveoutMaxShared;
    if (want > 0){
        const double need = (double) want * (double) (smem + 4096) / (228.0 * 1024.0) * 100.0;     // + static shared memory and the per-CTA reserve
        carve = need >= 100.0 ? 100 : (int) (need + 0.999);
    }
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess){ cudaGetLastError(); return 0; }
    return (want > 0 && n > want) ? want : n;
}
static int env_int(const char *name, int dflt){ const char *e = getenv(name); return e ? atoi(e) : dflt; }

// ---- heavy-tailed row lengths: virtual-row / tile-table form of the streaming kernel (spmv_pipe_kernel<..., VS = true>)
static constexpr int VS_THREADS = pipe_cfg<0>::THREADS, VS_STAGES = 2;
static constexpr int VS_CTAS = 5;          // resident CTAs per SM the carve-out is sized for (HB_VS_CTAS overrides): the rest of the array is L1 for the gathers
static void vsplit_free(hb_vsplit *v){
    if (!v) return;
    for (void *p : {(void*) v->vpntr, (void*) v->vmap, (void*) v->trow, (void*) v->tnz, (void*) v->srow, (void*) v->spart, v->part, (void*) v->cta_tiles})
        if (p) cudaFree(p);
    delete v;
}
template<typename T, int TPR, bool DOT>
static int vsplit_occupancy(){
    const size_t smem = pipe_smem_bytes(VS_THREADS, TPR, VS_STAGES, sizeof(T), VS_SLOT_DIV);
    auto k = spmv_pipe_kernel<T, VS_THREADS, TPR, VS_STAGES, DOT, 0, false, true>;
    return pipe_configure(k, VS_THREADS, smem, env_int("HB_VS_CTAS", VS_CTAS));
}
static int vsplit_occupancy_any(int dtype, int tpr){
    HB_DISPATCH(dtype, {
        int a = 0, b = 0;
        switch (tpr){
            case 4:  a = vsplit_occupancy<T, 4, false>(); b = vsplit_occupancy<T, 4, true>(); break;
            case 2:  a = vsplit_occupancy<T, 2, false>(); b = vsplit_occupancy<T, 2, true>(); break;
            default: a = vsplit_occupancy<T, 1, false>(); b = vsplit_occupancy<T, 1, true>(); break;
        }
        return a < b ? a : b;
    });
    return 0;
}
// measurement probe (results are WRONG with it on): bit 0 skips the row sums, bit 1 replaces the gathers of x by a constant
static int vs_probe_bits(){
    const char *e = getenv("HB_VS_PROBE"), *w = getenv("HB_VS_WARPROW");      // HB_VS_WARPROW: threshold of the warp-summed rows (probe)
    return (e ? atoi(e) & 3 : 0) | (w ? (atoi(w) & 0xffff) << 8 : 0);
}
template<typename T, int TPR, bool DOT>
static int launch_pipe_vs_tpr(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    const hb_vsplit *v = A->vs;
    const size_t smem = pipe_smem_bytes(VS_THREADS, TPR, VS_STAGES, sizeof(T), VS_SLOT_DIV);
    auto k = spmv_pipe_kernel<T, VS_THREADS, TPR, VS_STAGES, DOT, 0, false, true>;
    vsplit_view view;
    view.trow = v->trow; view.tnz = v->tnz; view.vmap = v->vmap; view.part = v->part; view.ntiles = v->ntiles;
    k<<<v->grid, VS_THREADS, smem, ctx->stream>>>(v->nvrows, A->nnz, v->vpntr, A->indx, (const T*) A->vals, x, y, alpha, beta,
                                                  v->contiguous ? v->cta_tiles : nullptr, ctx->partials, ctx->tickets + 1, dot_out, skip, vs_probe_bits(), A->cols,
                                                  nullptr, 0ull, (size_t) 0, (size_t) 0, (size_t) 0, 0, 0, view);
    HB_LAUNCH_CHECK(ctx);
    if (v->nsplit > 0){
        const int grid = std::min((v->nsplit + 7) / 8, ctx->num_sms * 4);
        vsplit_combine_kernel<T, DOT><<<grid, 256, 0, ctx->stream>>>(v->nsplit, v->srow, v->spart, (const T*) v->part, alpha, beta, y, x,
                                                                    ctx->partials, ctx->tickets + 7, dot_out, skip);
        HB_LAUNCH_CHECK(ctx);
    }
    return HB_OK;
}
template<typename T, bool DOT>
static int launch_pipe_vs(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    switch (A->vs->tpr){
        case 4:  return launch_pipe_vs_tpr<T, 4, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        case 2:  return launch_pipe_vs_tpr<T, 2, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        default: return launch_pipe_vs_tpr<T, 1, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    }
}
// One-time analysis on the host (heavy-tailed matrices only: one 4(rows+1)-byte read-back, two linear host loops, ~35 MB of tables for the
// 2^22-row power-law matrix): segments of at most SEG = (CAP - 8) / 4 entries, so that ANY four consecutive virtual rows fit a ring stage
// and tiles can start at multiples of four virtual rows (16-byte aligned slices of vpntr for the bulk copies).
static int vsplit_build(hb_ctx *ctx, hb_csr *A){
    hb_range nvtx_range("hb_csr_create: virtual-row tables");
    const size_t es = hb_dtype_size(A->dtype);
    const int cap = VS_THREADS * (es == 16 ? 4 : 8) / VS_SLOT_DIV;
    // lanes per row of the summation phase only (the gathers are dealt by non-zero): half of what the general kernel would take, so
    // that tiles are filled by non-zeros rather than cut short by the row limit; HB_VS_TPR overrides (probe)
    int tpr = A->tpr / 2 < 1 ? 1 : (A->tpr / 2 > 4 ? 4 : A->tpr / 2);
    { const char *te = getenv("HB_VS_TPR"); if (te){ const int q = atoi(te); if (q == 1 || q == 2 || q == 4) tpr = q; } }
    const int occ = vsplit_occupancy_any(A->dtype, tpr);
    if (occ < 1) return HB_OK;                                  // cannot run here: the matrix keeps the general kernel
    const int tile_rows = VS_THREADS / tpr, seg = ((cap - 8) / 4) & ~3;
    const int rows = A->rows;
    std::vector<int> hp((size_t) rows + 1);
    HB_CUDA(cudaMemcpyAsync(hp.data(), A->pntr, sizeof(int) * hp.size(), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> vpntr, vmap, srow, spart(1, 0);
    vpntr.reserve((size_t) rows + rows / 8 + 8); vmap.reserve((size_t) rows + rows / 8 + 8);
    vpntr.push_back(hp[0]);
    int nparts = 0;
    for (int r = 0; r < rows; r++){
        const int b = hp[(size_t) r], e = hp[(size_t) r + 1], len = e - b;
        if (len <= seg){ vpntr.push_back(e); vmap.push_back(r); continue; }
        const int nseg = (len + seg - 1) / seg;
        const int sl = (((len + nseg - 1) / nseg) + 3) & ~3;   // equal segments, a multiple of 4 entries each (<= seg since seg % 4 == 0)
        for (int q = 0; q < nseg; q++){
            const long long end = (long long) b + (long long) (q + 1) * sl;
            vpntr.push_back((int) (end < e ? end : e));
            vmap.push_back(~nparts);
            nparts++;
        }
        srow.push_back(r); spart.push_back(nparts);
    }
    const int nv = (int) vmap.size();
    std::vector<int> trow, tnz;
    for (int t0 = 0; t0 < nv; ){
        const int a0 = vpntr[(size_t) t0] & ~3;
        int r = t0;
        while (r < nv && r - t0 < tile_rows){
            const int r4 = std::min(r + 4, nv);
            if (vpntr[(size_t) r4] - a0 > cap) break;
            r = r4;
        }
        if (r == t0){ hb_set_error("internal: a group of four virtual rows exceeds a ring stage"); return HB_ERR_ARG; }
        int has_long = 0;                                       // bit 0 of the table entry: the tile holds a row the warp path takes
        for (int q = t0; q < r && !has_long; q++) has_long = (vpntr[(size_t) q + 1] - vpntr[(size_t) q]) >= PIPE_WARPROW;
        trow.push_back(t0 | has_long);
        t0 = r;
    }
    trow.push_back(nv % 4 == 0 ? nv : ((nv + 3) & ~3));         // end marker, rounded up so that masking the flag bits keeps it >= nv
    const int nt = (int) trow.size() - 1;
    tnz.resize(trow.size());
    for (size_t t = 0; t < trow.size(); t++) tnz[t] = vpntr[(size_t) std::min(trow[t] & ~3, nv)];
    int G = ctx->num_sms * occ;
    if (G > nt) G = nt;
    std::vector<int> cta((size_t) G + 1);
    for (int g = 0; g <= G; g++){
        const long long target = (long long) hp[0] + ((long long) A->nnz * g) / G;
        cta[(size_t) g] = g == G ? nt : (int) (std::lower_bound(tnz.begin(), tnz.begin() + nt, (int) target) - tnz.begin());
    }
    hb_vsplit *v = new hb_vsplit();
    {   // tiles hold about the same number of non-zeros: dealt round-robin (HB_PIPE_MAP=c: contiguous equal-nnz pieces)
        const char *m = getenv("HB_PIPE_MAP");
        v->contiguous = (m && m[0] == 'c') ? 1 : 0;
    }
    v->nvrows = nv; v->ntiles = nt; v->nsplit = (int) srow.size(); v->nparts = nparts; v->seg = seg; v->tpr = tpr; v->grid = G;
    auto up = [&](int **dst, const std::vector<int> &src)->cudaError_t{
        cudaError_t e = cudaMalloc((void**) dst, sizeof(int) * std::max<size_t>(src.size(), 4));
        if (e == cudaSuccess && !src.empty()) e = cudaMemcpyAsync(*dst, src.data(), sizeof(int) * src.size(), cudaMemcpyHostToDevice, ctx->stream);
        return e;
    };
    cudaError_t e = up(&v->vpntr, vpntr);
    if (e == cudaSuccess) e = up(&v->vmap, vmap);
    if (e == cudaSuccess) e = up(&v->trow, trow);
    if (e == cudaSuccess) e = up(&v->tnz, tnz);
    if (e == cudaSuccess) e = up(&v->srow, srow);
    if (e == cudaSuccess) e = up(&v->spart, spart);
    if (e == cudaSuccess) e = up(&v->cta_tiles, cta);
    if (e == cudaSuccess) e = cudaMalloc(&v->part, es * (size_t) std::max(nparts, 1));
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);       // the host vectors go out of scope
    if (e != cudaSuccess){ vsplit_free(v); return hb_cuda_fail(e, "virtual-row tables"); }
    A->vs = v;
    return HB_OK;
}

// resident CTAs per SM of the instantiation that will run (asked of the driver, not guessed): the persistent grid and its
// partition table are sized from it at hb_csr_create time
template<typename T, int CFG, int TPR, bool DOT>
static int pipe_occupancy_one(){
    using C = pipe_cfg<CFG>;
    const size_t smem = pipe_smem_bytes(C::THREADS, TPR, C::STAGES, sizeof(T));
    auto k = spmv_pipe_kernel<T, C::THREADS, TPR, C::STAGES, DOT>;
    return pipe_configure(k, C::THREADS, smem, env_int("HB_PIPE_CTAS", 0));
}
template<typename T, int CFG>
static int pipe_occupancy(int tpr){
    int a = 0, b = 0;
    switch (tpr){
        case 16: a = pipe_occupancy_one<T, CFG, 16, false>(); b = pipe_occupancy_one<T, CFG, 16, true>(); break;
        case 8:  a = pipe_occupancy_one<T, CFG, 8, false>();  b = pipe_occupancy_one<T, CFG, 8, true>();  break;
        case 4:  a = pipe_occupancy_one<T, CFG, 4, false>();  b = pipe_occupancy_one<T, CFG, 4, true>();  break;
        case 2:  a = pipe_occupancy_one<T, CFG, 2, false>();  b = pipe_occupancy_one<T, CFG, 2, true>();  break;
        default: a = pipe_occupancy_one<T, CFG, 1, false>();  b = pipe_occupancy_one<T, CFG, 1, true>();  break;
    }
    return a < b ? a : b;
}
static int pipe_occupancy_any(int cfg, int tpr, int dtype){
    HB_DISPATCH(dtype, { return cfg == 0 ? pipe_occupancy<T, 0>(tpr) : pipe_occupancy<T, 1>(tpr); });
    return 0;
}

// variant: 0 auto, 1 row-vector, 2 staged tiles, 3 streaming pipeline (bulk-async ring)
int hb_spmv_variant(const hb_csr *A){
    int variant = A->variant;
    const bool pipe_ok = A->vec_aligned && A->cta_rows[A->pipe_cfg] != nullptr && A->nnz > 0;
    if (variant == 0) variant = (A->mean_row_nnz > 112.0) ? 1 : (pipe_ok ? 3 : 2);
    if (variant == 3 && !pipe_ok) variant = 2;
    return variant;
}
template<typename T, bool DOT>
int hb_spmv_n_typed(hb_ctx *ctx, const hb_csr *A, const T *x, T *y, scalar_arg<T> alpha, scalar_arg<T> beta, T *dot_out, const int *skip){
    const int variant = hb_spmv_variant(A);
    const double mean = A->mean_row_nnz;
    if (variant == 3){
        if (A->vs && !(DOT && ctx->peer_hook)) return launch_pipe_vs<T, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        if (A->pipe_cfg == 0) return launch_pipe<T, 0, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        return launch_pipe<T, 1, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    }
    if (variant == 2){
        if (mean >= 16.0) return launch_tiles<T, 128, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
        return launch_tiles<T, 256, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    }
    // TPR * RV_UNR slots per step: pick the smallest TPR whose single step covers a typical row
    if (mean > 64.0)      return launch_rowvec<T, 32, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    else if (mean > 16.0) return launch_rowvec<T, 8, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    else if (mean > 8.0)  return launch_rowvec<T, 4, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
    return launch_rowvec<T, 2, DOT>(ctx, A, x, y, alpha, beta, dot_out, skip);
}

// used by hb_solvers.cu (fused CG): y = A x, *dot_dev = <x,y>, skipped entirely when *skip != 0
int hb_spmv_dot_internal(hb_ctx *ctx, const hb_csr *A, const void *x, void *y, void *dot_dev, const int *skip){
    HB_DISPATCH(A->dtype, {
        scalar_arg<T> one; one.value = one_of<T>(); one.dev = nullptr;
        scalar_arg<T> zero; zero.value = zero_of<T>(); zero.dev = nullptr;
        return hb_spmv_n_typed<T, true>(ctx, A, (const T*) x, (T*) y, one, zero, (T*) dot_dev, skip);
    });
    return HB_OK;
}
int hb_spmv_internal(hb_ctx *ctx, const hb_csr *A, const void *x, void *y, const int *skip){
    HB_DISPATCH(A->dtype, {
        scalar_arg<T> one; one.value = one_of<T>(); one.dev = nullptr;
        scalar_arg<T> zero; zero.value = zero_of<T>(); zero.dev = nullptr;
        return hb_spmv_n_typed<T, false>(ctx, A, (const T*) x, (T*) y, one, zero, nullptr, skip);
    });
    return HB_OK;
}

// ---- row-partitioned runs: which tiles of the streaming kernel reference ghost columns (columns >= rows of the local matrix)
__global__ void csr_tile_ghost_kernel(int rows, const int * __restrict__ pntr, const int * __restrict__ indx, int tile_rows, unsigned char *flags){
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x){
        const int rs = pntr[r], re = pntr[r + 1];
        bool g = false;
        for (int j = rs; j < re; j++) g = g || (indx[j] >= rows);
        if (g) flags[r / tile_rows] = 1;                    // same value from every writer
    }
}
// trot: tiles to rotate the round-robin sweep by (= length of the leading run of ghost-touching tiles, so that the sweep starts
// at the first interior tile and the leading boundary block comes last); twait: first position of the rotated order whose tile
// touches a ghost column.  1-D row blocks of a stencil: [lower face | interior | upper face] -> interior, upper face, lower face.
int hb_csr_halo_order(hb_ctx *ctx, const hb_csr *A, int *trot, int *twait){
    if (!A->halo_state){
        A->halo_trot = 0; A->halo_twait = 0;
        if (hb_spmv_variant(A) == 3 && !A->pipe_contiguous && A->cols > A->rows && A->rows > 0){
            const int threads = A->pipe_cfg == 0 ? pipe_cfg<0>::THREADS : pipe_cfg<1>::THREADS;
            const int tile_rows = threads / A->tpr;
            const int ntiles = (A->rows + tile_rows - 1) / tile_rows;
            unsigned char *flags = nullptr;
            HB_CUDA(cudaMalloc((void**) &flags, (size_t) ntiles));
            std::vector<unsigned char> h((size_t) ntiles);
            cudaError_t e = cudaMemsetAsync(flags, 0, (size_t) ntiles, ctx->stream);
            if (e == cudaSuccess){
                csr_tile_ghost_kernel<<<hb_grid_for(ctx, (size_t) A->rows, 256, 8), 256, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, tile_rows, flags);
                ctx->launches++;
                e = cudaMemcpyAsync(h.data(), flags, (size_t) ntiles, cudaMemcpyDeviceToHost, ctx->stream);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            cudaFree(flags);
            if (e != cudaSuccess) return hb_cuda_fail(e, "halo order analysis");
            int rot = 0;
            while (rot < ntiles && h[(size_t) rot]) rot++;
            if (rot < ntiles){                              // at least one interior tile
                int w = 0;
                while (w < ntiles && !h[(size_t) ((w + rot) % ntiles)]) w++;
                A->halo_trot = rot; A->halo_twait = w;      // w == ntiles: nothing touches a ghost, nobody waits
            }
        }
        A->halo_state = 1;
    }
    *trot = A->halo_trot; *twait = A->halo_twait;
    return HB_OK;
}

extern "C" {

int hb_csr_create(hb_ctx *ctx, int dtype, int rows, int cols, int nnz, const int *pntr, const int *indx, const void *vals, hb_csr **out){
    hb_range nvtx_range("hb_csr_create (analysis)");
    HB_ARG(ctx && out, "null");
    hb_activate(ctx);
    HB_ARG(dtype >= HB_F32 && dtype <= HB_C64, "dtype");
    HB_ARG(rows >= 0 && cols >= 0 && nnz >= 0, "negative dimension");
    HB_ARG(rows == 0 || pntr, "pntr is null");
    HB_ARG(nnz == 0 || (indx && vals), "indx/vals null");
    hb_csr *A = new hb_csr();
    struct csr_guard { hb_csr *a; ~csr_guard(){ if (a) hb_csr_destroy(a); } } guard{A};       // every early return below frees the object and its tables
    A->ctx = ctx; A->dtype = dtype; A->rows = rows; A->cols = cols; A->nnz = nnz;
    A->pntr = pntr; A->indx = indx; A->vals = vals;
    A->vec_aligned = aligned16p(indx) && aligned16p(vals);
    A->tc = hb_tcache_new();
    A->mean_row_nnz = rows > 0 ? (double) nnz / rows : 0.0;
    // one-time analysis: longest row (one tiny kernel, one 4-byte read-back) — decides the tile shape and the tile-to-CTA map below
    A->stats_dev = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->dscalars) + 2048);
    HB_CUDA(cudaMemsetAsync(A->stats_dev, 0, sizeof(int), ctx->stream));
    if (rows > 0){
        int grid = hb_grid_for(ctx, (size_t) rows, 256, 8);
        csr_analyse_kernel<<<grid, 256, 0, ctx->stream>>>(rows, pntr, A->stats_dev);
        HB_LAUNCH_CHECK(ctx);
    }
    HB_CUDA(cudaMemcpyAsync(&A->max_row_nnz, A->stats_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    const bool heavy_tail = (double) A->max_row_nnz > 16.0 * (A->mean_row_nnz + 1.0);
    // lanes per row from the mean row length.  Heavy-tailed row lengths: a tile whose non-zeros do not fit a ring stage is walked
    // straight from global memory, so tiles are made of fewer rows (more lanes per row) to keep almost all of them staged
    // (power-law matrix of configs[4]: see DESIGN.md §4)
    A->tpr = pipe_tpr(A->mean_row_nnz, hb_dtype_size(dtype));
    {
        const char *sh = getenv("HB_PIPE_TPR_SHIFT");
        int shift = !heavy_tail ? 0 : (sh ? atoi(sh) : HB_HEAVY_TPR_SHIFT);
        while (shift-- > 0 && A->tpr < 16) A->tpr *= 2;
    }
    // equal-nnz row partition tables for the streaming kernel (one per pipeline configuration)
    const char *cfg_env = getenv("HB_PIPE_CFG");
    A->pipe_cfg = cfg_env ? ((cfg_env[0] == '1') ? 1 : 0) : (A->tpr >= 4 ? 1 : 0);
    A->vec_aligned = A->vec_aligned && aligned16p(pntr);
    {   // tile-to-CTA map of the streaming kernel: round-robin sweep by default, contiguous equal-nnz pieces on request
        const char *m = getenv("HB_PIPE_MAP");
        A->pipe_contiguous = m ? (m[0] == 'c' ? 1 : 0) : -1;       // -1: decided below from the row-length statistics
    }
    if (rows > 0 && nnz > 0 && A->vec_aligned){
        for (int c = 0; c < 2; c++){
            const int threads = c == 0 ? pipe_cfg<0>::THREADS : pipe_cfg<1>::THREADS;
            const int tile_rows = threads / A->tpr;
            const int ntiles = (rows + tile_rows - 1) / tile_rows;
            const int occ = pipe_occupancy_any(c, A->tpr, dtype);
            if (occ < 1) continue;                          // this configuration cannot run here: cta_rows stays null
            int G = ctx->num_sms * occ;
            if (G > ntiles) G = ntiles;
            A->pipe_grid[c] = G;
            HB_CUDA(cudaMalloc((void**) &A->cta_rows[c], sizeof(int) * (size_t) (G + 1)));
            csr_partition_kernel<<<(G + 1 + 127) / 128, 128, 0, ctx->stream>>>(rows, nnz, pntr, tile_rows, G, A->cta_rows[c]);
            HB_LAUNCH_CHECK(ctx);
        }
    }
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    // heavy-tailed row lengths: equal-nnz contiguous pieces balance better than the sweep (power-law matrix: 623 vs 674 us);
    // regular matrices: the sweep keeps the gather window of x in L2/L1 (27-point: 117 vs 151 us, 512^3 7-point: -13 % DRAM traffic)
    if (A->pipe_contiguous < 0) A->pipe_contiguous = heavy_tail ? 1 : 0;
    // heavy-tailed row lengths: long rows are cut into segments and the tiles come from a table, so that every tile is staged
    // (HB_VSPLIT=0 keeps the general kernel with its global-memory paths: A/B probe)
    {
        const char *vse = getenv("HB_VSPLIT");
        if (heavy_tail && hb_spmv_variant(A) == 3 && !(vse && vse[0] == '0')){
            int rc = vsplit_build(ctx, A);
            if (rc != HB_OK) return rc;
        }
    }
    guard.a = nullptr;
    *out = A;
    return HB_OK;
}

int hb_csr_destroy(hb_csr *csr){
    if (!csr) return HB_OK;
    for (int c = 0; c < 2; c++) if (csr->cta_rows[c]) cudaFree(csr->cta_rows[c]);
    hb_tcache_delete(csr->tc);
    vsplit_free(csr->vs);
    delete csr;
    return HB_OK;
}

int hb_csr_info(const hb_csr *A, int *dtype, int *rows, int *cols, int *nnz, int *max_row_nnz){
    HB_ARG(A, "csr is null");
    if (dtype) *dtype = A->dtype;
    if (rows) *rows = A->rows;
    if (cols) *cols = A->cols;
    if (nnz) *nnz = A->nnz;
    if (max_row_nnz) *max_row_nnz = A->max_row_nnz;
    return HB_OK;
}

int hb_csr_set_variant(hb_csr *A, int variant){ HB_ARG(A && variant >= 0 && variant <= 3, "variant"); A->variant = variant; return HB_OK; }

int hb_spmv_buffer_size(const hb_csr *A, char trans, size_t *bytes){ (void) trans; HB_ARG(A && bytes, "null"); *bytes = 0; return HB_OK; }

int hb_spmv(hb_ctx *ctx, const hb_csr *A, char trans, const void *alpha, const void *x, const void *beta, void *y){
    HB_ARG(ctx && A && alpha && beta, "null");
    hb_activate(ctx);
    const int ny = hb_is_n(trans) ? A->rows : A->cols;
    if (ny == 0) return HB_OK;
    HB_ARG(y, "y is null");
    HB_ARG(x || (hb_is_n(trans) ? A->cols : A->rows) == 0, "x is null");
    // op 'T' / 'C': the op 'N' kernel on the cached CSR of A^T (hb_transpose.cu) unless the object is in scatter mode
    const hb_csr *At = nullptr;
    if (!hb_is_n(trans)){ int rc = hb_csr_transposed(ctx, A, trans, &At); if (rc != HB_OK) return rc; }
    HB_DISPATCH(A->dtype, {
        scalar_arg<T> a = make_scalar<T>(ctx, alpha), b = make_scalar<T>(ctx, beta);
        if (hb_is_n(trans)) return hb_spmv_n_typed<T, false>(ctx, A, (const T*) x, (T*) y, a, b, nullptr, nullptr);
        if (At) return hb_spmv_n_typed<T, false>(ctx, At, (const T*) x, (T*) y, a, b, nullptr, nullptr);
        int grid = hb_grid_for(ctx, (size_t) ny, 256, 8);
        scale_or_zero_kernel<T><<<grid, 256, 0, ctx->stream>>>(ny, b, (T*) y);
        HB_LAUNCH_CHECK(ctx);
        if (A->rows > 0 && A->nnz > 0){
            const bool cj = hb_is_c(trans) && is_cplx<T>::value;
            const long long threads = (long long) A->rows * 4;
            int g2 = (int) std::min<long long>((threads + 255) / 256, (long long) ctx->num_sms * 8);
            if (cj) spmv_trans_kernel<T, 4, true><<<g2, 256, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, (const T*) A->vals, (const T*) x, (T*) y, a);
            else    spmv_trans_kernel<T, 4, false><<<g2, 256, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, (const T*) A->vals, (const T*) x, (T*) y, a);
            HB_LAUNCH_CHECK(ctx);
        }
    });
    return HB_OK;
}

int hb_spmv_dot(hb_ctx *ctx, const hb_csr *A, const void *x, void *y, void *dot_dev){
    HB_ARG(ctx && A && dot_dev, "null");
    hb_activate(ctx);
    HB_ARG(A->cols >= A->rows, "hb_spmv_dot needs cols >= rows (x[i] pairs with y[i]; extra columns are ghost entries)");
    if (A->rows == 0){ HB_CUDA(cudaMemsetAsync(dot_dev, 0, hb_dtype_size(A->dtype), ctx->stream)); return HB_OK; }
    return hb_spmv_dot_internal(ctx, A, x, y, dot_dev, nullptr);
}

}
