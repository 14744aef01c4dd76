// hb_batch.cu — what the reference's batch (multi right-hand-side) CG needs from the GPU layer:
//   hb_spmm : C = alpha op(A) op(B) + beta C, A sparse CSR, B / C dense column-major — gpu_sparse_matrix::gemm -> cusparseSpMM
//             (reference gpu/hala_cuda_sparse_general.hpp:284-332)
//   hb_geam : C = alpha op(A) + beta op(B)            — cublas?geam (gpu/hala_gpu_blas0.hpp:46-72)
//   hb_dgmm : C = diag(x) A  or  A diag(x)            — cublas?dgmm (gpu/hala_gpu_blas0.hpp:79-103)
//   hb_tbsv : banded triangular solve in place        — cublas?tbsv (gpu/hala_gpu_blas2.hpp), which wax/hala_blas_extensions.hpp:251-260
//             uses with bandwidth 0 as an element-wise divide
// The wax templates batch_dot / batch_axpy / batch_scal / batch_max_norm2 / vdivide (wax/hala_blas_extensions.hpp:203-275,353-365)
// are built from exactly these plus gemv('T'), axpy, vcopy, iamax and norm2, so solve_batch_cg runs unchanged on top.
#include "hb_common.cuh"
#include <cstdlib>

static constexpr int BT_THREADS = 256;

bool hb_spmm_interleaved_ok(const hb_csr *A);
bool hb_spmm_lpc_ok(const hb_csr *A, int nbp);
int  hb_spmm_lpc(hb_ctx *ctx, const hb_csr *A, int nbp, int nb, const void *B, size_t sxr, size_t sxc, const void *alpha, const void *beta, void *Cm, size_t ldc);
int  hb_spmm_interleaved(hb_ctx *ctx, const hb_csr *A, int nbp, const void *Bt, size_t ldbt, void *Ct, size_t ldct);

// ------------------------------------------------------------------------------------------------ SpMM, op(A) = A
// TPR lanes per row, NB right-hand sides per pass: every lane walks its share of the row once and feeds NB accumulators, so the
// matrix is read ceil(N / NB) times instead of N times.  BT != 0: op(B)[k][n] = B[n + k ldb] (rows of B are contiguous in n).
template<typename T, int TPR, int NB, bool BT, bool BCONJ>
__global__ void __launch_bounds__(BT_THREADS) spmm_n_kernel(int rows, int ncolsC, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                            const T * __restrict__ vals, scalar_arg<T> alpha_s, const T * __restrict__ B, long long ldb,
                                                            scalar_arg<T> beta_s, T *C, long long ldc){
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    const int sub = threadIdx.x % TPR;
    const long long grp0 = (blockIdx.x * (long long) BT_THREADS + threadIdx.x) / TPR, ngrp = (long long) gridDim.x * BT_THREADS / TPR;
    for (int n0 = 0; n0 < ncolsC; n0 += NB){
        for (long long base = 0; base < rows; base += ngrp){      // every lane keeps looping so the shuffles below stay converged
            const long long i = base + grp0;
            const bool live = i < rows;
            T acc[NB];
            #pragma unroll
            for (int kb = 0; kb < NB; kb++) acc[kb] = zero_of<T>();
            if (live){
                for (int j = pntr[i] + sub; j < pntr[i + 1]; j += TPR){
                    const T v = ld_stream(vals + j);
                    const long long c = __ldcs(indx + j);
                    #pragma unroll
                    for (int kb = 0; kb < NB; kb++){
                        if (n0 + kb < ncolsC){
                            T bv = BT ? ld_ro(B + (n0 + kb) + c * ldb) : ld_ro(B + c + (long long) (n0 + kb) * ldb);
                            acc[kb] = hfma(v, BCONJ ? hconj(bv) : bv, acc[kb]);
                        }
                    }
                }
            }
            #pragma unroll
            for (int kb = 0; kb < NB; kb++){
                #pragma unroll
                for (int d = TPR / 2; d > 0; d >>= 1) acc[kb] = hadd(acc[kb], shfl_down(acc[kb], d));
            }
            if (live && sub == 0){
                #pragma unroll
                for (int kb = 0; kb < NB; kb++){
                    if (n0 + kb < ncolsC){
                        T *cp = C + i + (long long) (n0 + kb) * ldc;
                        T out = hmul(alpha, acc[kb]);
                        if (use_beta) out = hfma(beta, *cp, out);
                        *cp = out;
                    }
                }
            }
        }
    }
}

// contiguous (optionally conjugated) copy of a strided vector: the row of B that op(B) = B^T / B^H turns into a column
template<typename T, bool CONJ> __global__ void gather_row_kernel(int n, const T *src, long long inc, T *dst){
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
        const T v = src[i * inc];
        dst[i] = CONJ ? hconj(v) : v;
    }
}

// Bt[c * NBP + kb] = op(B)[c][n0 + kb] for kb < nb (zero padding up to NBP): one thread per row c, coalesced reads per column, one
// contiguous NBP-element store per thread.  sb_row / sb_col = strides of B between consecutive c / consecutive right-hand sides.
template<typename T, int NBP>
__global__ void __launch_bounds__(256) interleave_kernel(long long K, int nb, const T * __restrict__ B, long long sb_row, long long sb_col, T *Bt){
    constexpr int NV = vec16<T>::N;
    for (long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < K; c += (long long) gridDim.x * blockDim.x){
        T v[NBP];
        #pragma unroll
        for (int kb = 0; kb < NBP; kb++) v[kb] = (kb < nb) ? ld_stream(B + c * sb_row + kb * sb_col) : zero_of<T>();
        vec16<T> *dst = reinterpret_cast<vec16<T>*>(Bt + c * NBP);
        #pragma unroll
        for (int q = 0; q < NBP / NV; q++){
            vec16<T> o;
            #pragma unroll
            for (int e = 0; e < NV; e++) o.v[e] = v[q * NV + e];
            dst[q] = o;
        }
    }
}
// C[i + kb ldc] = alpha Ct[i * NBP + kb] + beta C[i + kb ldc] for kb < nb; C is not read when beta == 0
template<typename T, int NBP>
__global__ void __launch_bounds__(256) deinterleave_kernel(long long M, int nb, const T * __restrict__ Ct, scalar_arg<T> alpha_s, scalar_arg<T> beta_s,
                                                          T *C, long long ldc){
    constexpr int NV = vec16<T>::N;
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < M; i += (long long) gridDim.x * blockDim.x){
        const vec16<T> *src = reinterpret_cast<const vec16<T>*>(Ct + i * NBP);
        T v[NBP];
        #pragma unroll
        for (int q = 0; q < NBP / NV; q++){
            const vec16<T> o = src[q];
            #pragma unroll
            for (int e = 0; e < NV; e++) v[q * NV + e] = o.v[e];
        }
        #pragma unroll
        for (int kb = 0; kb < NBP; kb++){
            if (kb < nb){
                T *cp = C + i + kb * ldc;
                T out = hmul(alpha, v[kb]);
                if (use_beta) out = hfma(beta, *cp, out);
                *cp = out;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ geam / dgmm / tbsv
template<typename T> __device__ __forceinline__ T op_elem(const T *A, long long lda, int i, int j, int mode){   // mode 0 N, 1 T, 2 C
    if (mode == 0) return A[i + (long long) j * lda];
    const T v = A[j + (long long) i * lda];
    return mode == 2 ? hconj(v) : v;
}
// 32 x 32 tiles through shared memory so that transposed operands are still read along their contiguous dimension
template<typename T>
__global__ void __launch_bounds__(256) geam_kernel(int M, int N, scalar_arg<T> alpha_s, const T *A, long long lda, int ma,
                                                   scalar_arg<T> beta_s, const T *B, long long ldb, int mb, T *C, long long ldc){
    __shared__ T ta[32][33], tb[32][33];
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_a = !hiszero(alpha), use_b = !hiszero(beta);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8 threads
    const int tiles_m = (M + 31) / 32, tiles_n = (N + 31) / 32;
    for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x){
        const int i0 = (t % tiles_m) * 32, j0 = (t / tiles_m) * 32;
        // operand tiles: element (i0 + r, j0 + c) of op(X) lands in tile[c][r]
        for (int k = ty; k < 32; k += 8){
            if (use_a){
                if (ma == 0){ const int i = i0 + tx, j = j0 + k; if (i < M && j < N) ta[k][tx] = A[i + (long long) j * lda]; }
                else { const int j = j0 + tx, i = i0 + k; if (i < M && j < N){ const T v = A[j + (long long) i * lda]; ta[tx][k] = ma == 2 ? hconj(v) : v; } }
            }
            if (use_b){
                if (mb == 0){ const int i = i0 + tx, j = j0 + k; if (i < M && j < N) tb[k][tx] = B[i + (long long) j * ldb]; }
                else { const int j = j0 + tx, i = i0 + k; if (i < M && j < N){ const T v = B[j + (long long) i * ldb]; tb[tx][k] = mb == 2 ? hconj(v) : v; } }
            }
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8){
            const int i = i0 + tx, j = j0 + k;
            if (i < M && j < N){
                T out = zero_of<T>();
                if (use_a) out = hmul(alpha, ta[k][tx]);
                if (use_b) out = hfma(beta, tb[k][tx], out);
                C[i + (long long) j * ldc] = out;
            }
        }
        __syncthreads();
    }
}
template<typename T> __global__ void __launch_bounds__(256) dgmm_kernel(int M, int N, int left, const T *A, long long lda,
                                                                        const T *x, long long incx, T *C, long long ldc){
    const long long total = (long long) M * N;
    for (long long e = blockIdx.x * (long long) blockDim.x + threadIdx.x; e < total; e += (long long) gridDim.x * blockDim.x){
        const long long i = e % M, j = e / M;
        C[i + j * ldc] = hmul(A[i + j * lda], x[(left ? i : j) * incx]);
    }
}
// bandwidth 0: x_i = x_i / op(a_ii)
template<typename T> __global__ void tbsv_diag_kernel(int n, int conj, const T * __restrict__ A, long long lda, T *x, long long incx){
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x){
        const T d = A[i * lda];
        x[i * incx] = hdiv(x[i * incx], conj ? hconj(d) : d);
    }
}
// general bandwidth: the netlib ?tbsv recurrences by one thread (a dependency chain of length n; the batch solvers never get here)
template<typename T> __global__ void tbsv_serial_kernel(int upper, int mode, int unit, int n, int k, const T * __restrict__ A, long long lda, T *x, long long incx){
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    auto a = [&](int r, int c){ T v = A[r + (long long) c * lda]; return mode == 2 ? hconj(v) : v; };
    if (mode == 0){
        if (upper){
            for (int j = n - 1; j >= 0; j--){
                T xj = x[j * incx];
                if (!unit){ xj = hdiv(xj, a(k, j)); x[j * incx] = xj; }
                for (int i = j - 1; i >= max(0, j - k); i--) x[i * incx] = hsub(x[i * incx], hmul(xj, a(k + i - j, j)));
            }
        }else{
            for (int j = 0; j < n; j++){
                T xj = x[j * incx];
                if (!unit){ xj = hdiv(xj, a(0, j)); x[j * incx] = xj; }
                for (int i = j + 1; i <= min(n - 1, j + k); i++) x[i * incx] = hsub(x[i * incx], hmul(xj, a(i - j, j)));
            }
        }
    }else{
        if (upper){
            for (int j = 0; j < n; j++){
                T t = x[j * incx];
                for (int i = max(0, j - k); i < j; i++) t = hsub(t, hmul(a(k + i - j, j), x[i * incx]));
                if (!unit) t = hdiv(t, a(k, j));
                x[j * incx] = t;
            }
        }else{
            for (int j = n - 1; j >= 0; j--){
                T t = x[j * incx];
                for (int i = min(n - 1, j + k); i > j; i--) t = hsub(t, hmul(a(i - j, j), x[i * incx]));
                if (!unit) t = hdiv(t, a(0, j));
                x[j * incx] = t;
            }
        }
    }
}

static int trans_mode(char t){ return hb_is_n(t) ? 0 : (hb_is_c(t) ? 2 : 1); }

template<typename T, bool BT, bool BCONJ>
static void launch_spmm_n(hb_ctx *ctx, const hb_csr *A, int ncolsC, scalar_arg<T> a, const T *B, long long ldb, scalar_arg<T> b, T *C, long long ldc){
    const double mean = A->mean_row_nnz;
    const int grid = hb_grid_for(ctx, (size_t) A->rows, BT_THREADS / (mean > 24 ? 8 : (mean > 10 ? 4 : 1)), 4);
    #define HB_SPMM(TPR) spmm_n_kernel<T, TPR, 4, BT, BCONJ><<<grid, BT_THREADS, 0, ctx->stream>>>(A->rows, ncolsC, A->pntr, A->indx, (const T*) A->vals, a, B, ldb, b, C, ldc)
    if (mean > 24) HB_SPMM(8); else if (mean > 10) HB_SPMM(4); else HB_SPMM(1);
    #undef HB_SPMM
}

extern "C" {

int hb_spmm(hb_ctx *ctx, const hb_csr *A, char transa, char transb, int b_rows, int b_cols, const void *alpha, const void *B, int ldb,
            const void *beta, void *C, int ldc){
    HB_ARG(ctx && A && alpha && beta, "null");
    hb_activate(ctx);
    HB_ARG(b_rows >= 0 && b_cols >= 0, "negative size");
    const bool an = hb_is_n(transa), bn = hb_is_n(transb);
    const int M = an ? A->rows : A->cols, K = an ? A->cols : A->rows, N = bn ? b_cols : b_rows;
    HB_ARG((bn ? b_rows : b_cols) == K, "inner dimensions of op(A) and op(B) differ");
    HB_ARG(ldb >= (b_rows > 1 ? b_rows : 1) && ldc >= (M > 1 ? M : 1), "leading dimension too small");
    if (M == 0 || N == 0) return HB_OK;
    HB_ARG(B && C, "null matrix");
    const size_t es = hb_dtype_size(A->dtype);
    if (!an){       // op(A) = A^T / A^H: the same product with the cached CSR of A^T (hb_transpose.cu) when the object keeps one
        const hb_csr *At = nullptr;
        int rc = hb_csr_transposed(ctx, A, transa, &At);
        if (rc != HB_OK) return rc;
        if (At) return hb_spmm(ctx, At, 'N', transb, b_rows, b_cols, alpha, B, ldb, beta, C, ldc);
    }
    // Fast path: right-hand sides in blocks of up to 8, interleaved (row-major) so that the operands of one non-zero are one
    // contiguous gather, through the streaming SpMV kernel in its multi right-hand-side mode: the matrix is read once per block.
    //   Bt = block of op(B), interleaved (interleave_kernel) -> Ct = A Bt (spmv_pipe_kernel<..., NBP>) -> C = alpha Ct + beta C (deinterleave_kernel)
    const bool bconj = hb_is_c(transb) && (A->dtype == HB_C32 || A->dtype == HB_C64);
    // Alternative form of the streaming kernel, "lane per column" (HB_SPMM_PATH=lpc): B and C used where they lie, NBP adjacent lanes share a
    // row, no interleaving passes, no workspace, rows summed left to right.  Measured on B200 it loses to the interleaved form on long rows
    // (27-point 128^3, 4 columns: 340 us against 272 us) and ties on short ones (7-point 256^3: 828 us against 838 us): both are bound by the
    // L1TEX data stage at about one wavefront per gathered sector (ncu: l1tex data-pipe 87 %, DRAM 28-44 %), see DESIGN.md §4c.
    static const char *path_env = getenv("HB_SPMM_PATH");
    if (an && A->nnz > 0 && path_env && path_env[0] == 'l' && !(bconj && !bn) && hb_spmm_lpc_ok(A, 4)){
        int rc = HB_OK;
        for (int n0 = 0; n0 < N && rc == HB_OK; n0 += 4){
            const int nb = (N - n0 < 4) ? (N - n0) : 4;
            // op(B)[c][n]: transb N -> B[c + n ldb]; otherwise B[n + c ldb]
            const char *Bblk = (const char*) B + es * (bn ? (size_t) n0 * (size_t) ldb : (size_t) n0);
            rc = hb_spmm_lpc(ctx, A, 4, nb, Bblk, bn ? (size_t) 1 : (size_t) ldb, bn ? (size_t) ldb : (size_t) 1, alpha, beta,
                             (char*) C + es * (size_t) n0 * (size_t) ldc, (size_t) ldc);
        }
        return rc;
    }
    const bool want_il = !path_env || path_env[0] != 'g';
    if (an && A->nnz > 0 && want_il && hb_spmm_interleaved_ok(A) && !(bconj && !bn)){
        const int saved_mode = ctx->pointer_mode;
        int rc = HB_OK;
        void *arena = nullptr;
        if ((rc = hb_ctx_workspace(ctx, es * 8 * ((size_t) K + (size_t) M) + 512, &arena)) != HB_OK) return rc;
        char *Bt = (char*) arena, *Ct = Bt + ((es * 8 * (size_t) K + 255) / 256) * 256;
        // blocks of 4: measured on B200 (27-point 128^3) one 8-wide pass costs 610 us, two 4-wide passes 2 x 273 us — the 8-wide
        // instantiation keeps only two entries per lane in flight under the kernel's 80-register budget
        for (int n0 = 0; n0 < N && rc == HB_OK; n0 += 4){
            const int nb = (N - n0 < 4) ? (N - n0) : 4, nbp = 4;
            const int gk = hb_grid_for(ctx, (size_t) K, 256, 8), gm = hb_grid_for(ctx, (size_t) M, 256, 8);
            HB_DISPATCH(A->dtype, {
                // op(B)[c][n]: transb N -> B[c + n ldb]; otherwise B[n + c ldb]
                const T *Bblk = bn ? (const T*) B + (size_t) n0 * (size_t) ldb : (const T*) B + n0;
                const long long sb_row = bn ? 1 : ldb, sb_col = bn ? ldb : 1;
                if (nbp == 4) interleave_kernel<T, 4><<<gk, 256, 0, ctx->stream>>>(K, nb, Bblk, sb_row, sb_col, (T*) Bt);
                else          interleave_kernel<T, 8><<<gk, 256, 0, ctx->stream>>>(K, nb, Bblk, sb_row, sb_col, (T*) Bt);
                ctx->launches++;
                rc = hb_spmm_interleaved(ctx, A, nbp, Bt, (size_t) nbp, Ct, (size_t) nbp);
                if (rc == HB_OK){
                    scalar_arg<T> a = make_scalar<T>(ctx, alpha), b = make_scalar<T>(ctx, beta);
                    T *Cblk = (T*) C + (size_t) n0 * (size_t) ldc;
                    if (nbp == 4) deinterleave_kernel<T, 4><<<gm, 256, 0, ctx->stream>>>(M, nb, (const T*) Ct, a, b, Cblk, ldc);
                    else          deinterleave_kernel<T, 8><<<gm, 256, 0, ctx->stream>>>(M, nb, (const T*) Ct, a, b, Cblk, ldc);
                    ctx->launches++;
                }
            });
        }
        ctx->pointer_mode = saved_mode;
        return rc;
    }
    if (an && A->nnz > 0){
        HB_DISPATCH(A->dtype, {
            scalar_arg<T> a = make_scalar<T>(ctx, alpha), b = make_scalar<T>(ctx, beta);
            const bool cj = hb_is_c(transb) && is_cplx<T>::value;
            if (bn) launch_spmm_n<T, false, false>(ctx, A, N, a, (const T*) B, ldb, b, (T*) C, ldc);
            else if (cj) launch_spmm_n<T, true, true>(ctx, A, N, a, (const T*) B, ldb, b, (T*) C, ldc);
            else launch_spmm_n<T, true, false>(ctx, A, N, a, (const T*) B, ldb, b, (T*) C, ldc);
        });
        HB_LAUNCH_CHECK(ctx);
        return HB_OK;
    }
    // op(A) = A^T / A^H (or an empty matrix): column by column through the SpMV path, as sparse_gemm_array does (hala_sparse_utils.hpp:121-160)
    void *tmp = nullptr;
    if (!bn) HB_CUDA(cudaMalloc(&tmp, es * (size_t) (K > 0 ? K : 1)));
    int rc = HB_OK;
    for (int n = 0; n < N && rc == HB_OK; n++){
        const void *xcol = (const char*) B + es * (size_t) n * (size_t) ldb;
        if (!bn){
            const int g = hb_grid_for(ctx, (size_t) K, 256, 4);
            HB_DISPATCH(A->dtype, {
                if (hb_is_c(transb) && is_cplx<T>::value) gather_row_kernel<T, true><<<g, 256, 0, ctx->stream>>>(K, (const T*) B + n, ldb, (T*) tmp);
                else gather_row_kernel<T, false><<<g, 256, 0, ctx->stream>>>(K, (const T*) B + n, ldb, (T*) tmp);
            });
            ctx->launches++;
            xcol = tmp;
        }
        rc = hb_spmv(ctx, A, transa, alpha, xcol, beta, (char*) C + es * (size_t) n * (size_t) ldc);
    }
    if (tmp){ cudaStreamSynchronize(ctx->stream); cudaFree(tmp); }
    return rc;
}

int hb_geam(hb_ctx *ctx, int dtype, char transa, char transb, int M, int N, const void *alpha, const void *A, int lda,
            const void *beta, const void *B, int ldb, void *C, int ldc){
    HB_ARG(ctx && alpha && beta, "null");
    HB_ARG(M >= 0 && N >= 0, "negative size");
    if (M == 0 || N == 0) return HB_OK;
    HB_ARG(A && B && C, "null matrix");
    HB_ARG(lda >= (hb_is_n(transa) ? M : N) && ldb >= (hb_is_n(transb) ? M : N) && ldc >= M, "leading dimension too small");
    const long long tiles = (long long) ((M + 31) / 32) * ((N + 31) / 32);
    const int grid = (int) (tiles < (long long) ctx->num_sms * 8 ? tiles : (long long) ctx->num_sms * 8);
    HB_DISPATCH(dtype, (geam_kernel<T><<<grid, 256, 0, ctx->stream>>>(M, N, make_scalar<T>(ctx, alpha), (const T*) A, lda, trans_mode(transa),
                                                                      make_scalar<T>(ctx, beta), (const T*) B, ldb, trans_mode(transb), (T*) C, ldc)));
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_dgmm(hb_ctx *ctx, int dtype, char side, int M, int N, const void *A, int lda, const void *x, int incx, void *C, int ldc){
    HB_ARG(ctx, "null");
    HB_ARG(side == 'L' || side == 'l' || side == 'R' || side == 'r', "side must be L or R");
    HB_ARG(M >= 0 && N >= 0, "negative size");
    if (M == 0 || N == 0) return HB_OK;
    HB_ARG(A && x && C, "null array");
    HB_ARG(lda >= M && ldc >= M && incx != 0, "leading dimension / stride");
    const int grid = hb_grid_for(ctx, (size_t) M * (size_t) N, 256 * 4, 8);
    HB_DISPATCH(dtype, (dgmm_kernel<T><<<grid, 256, 0, ctx->stream>>>(M, N, (side == 'L' || side == 'l') ? 1 : 0, (const T*) A, lda, (const T*) x, incx, (T*) C, ldc)));
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_tbsv(hb_ctx *ctx, int dtype, char uplo, char trans, char diag, int n, int k, const void *A, int lda, void *x, int incx){
    HB_ARG(ctx, "null");
    HB_ARG(n >= 0 && k >= 0 && lda >= k + 1 && incx != 0, "bad dimensions");
    if (n == 0) return HB_OK;
    HB_ARG(A && x, "null array");
    const int unit = (diag == 'U' || diag == 'u'), upper = (uplo == 'U' || uplo == 'u');
    if (k == 0){
        if (unit) return HB_OK;
        const int grid = hb_grid_for(ctx, (size_t) n, 256, 8);
        HB_DISPATCH(dtype, (tbsv_diag_kernel<T><<<grid, 256, 0, ctx->stream>>>(n, (hb_is_c(trans) && is_cplx<T>::value) ? 1 : 0, (const T*) A, lda, (T*) x, incx)));
    }else{
        HB_DISPATCH(dtype, (tbsv_serial_kernel<T><<<1, 32, 0, ctx->stream>>>(upper, trans_mode(trans), unit, n, k, (const T*) A, lda, (T*) x, incx)));
    }
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

}
