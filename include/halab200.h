/* halab200.h — C ABI of libhalab200.so, the B200-native (sm_100a) backend for LIBHALA/hala's sparse
 * iterative-solve hot path.  This is the drop-in boundary: the replacement `gpu/` header layer
 * (the .hpp files under hala_b200/gpu/, source-compatible with the reference's L4 headers) calls ONLY these entry points;
 * anything that can bind a C function (ctypes, cgo, JNI, ...) can call them too.  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns an int status (HB_OK == 0); hb_last_error() gives the message of the last
 *     failure on the calling thread.  The C++ header layer turns a non-zero status into std::runtime_error,
 *     as the reference does for CUDA/cuBLAS/cuSPARSE statuses (gpu/hala_cuda_common.hpp:78-152).
 *   - dtype: HB_F32/HB_F64/HB_C32/HB_C64 == float/double/complex<float>/complex<double>
 *     (the reference's cuda_call_backend 4-way dispatch, gpu/hala_cuda_common.hpp:159-184).
 *   - all array arguments are DEVICE pointers on the context's device unless the name says host.
 *   - scalars (alpha, beta, results) are passed by pointer to a value of `dtype` (results of nrm2: its real
 *     type).  In HB_POINTER_HOST mode (default) they are host pointers and scalar-returning calls are
 *     host-synchronous, exactly like the reference (gpu/hala_gpu_blas1.hpp:186-197); in HB_POINTER_DEVICE
 *     mode they are device pointers and nothing synchronises (reference: gpu_pntr<device_pntr>,
 *     gpu/hala_gpu_engine.hpp:410-446).
 *   - a context is NOT thread-safe (one engine per host thread per device — the reference's contract).
 *   - kernels are launched on the context's stream (default: the legacy default stream, as the reference).
 *   - 32-bit CSR indices, 0-based (reference: CUSPARSE_INDEX_32I / BASE_ZERO, gpu/hala_cuda_sparse_general.hpp:85-87);
 *     vector lengths are int like BLAS, but every internal offset is 64-bit (rows*restart > 2^31 is legal).
 */
#ifndef HALAB200_H
#define HALAB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_ctx hb_ctx;   /* replaces the cuBLAS+cuSPARSE handle pair held by gpu_engine (gpu/hala_gpu_engine.hpp:47-163) */
typedef struct hb_csr hb_csr;
typedef struct hb_tri hb_tri;   /* replaces cusparseSpMatDescr_t + the cached SpSV/SpSM analyses of gpu_triangular_matrix (gpu/hala_cuda_sparse_triangular.hpp:38-110) */   /* replaces cusparseSpMatDescr_t + cached buffer sizes of gpu_sparse_matrix (gpu/hala_cuda_sparse_general.hpp:191-375) */

enum { HB_F32 = 0, HB_F64 = 1, HB_C32 = 2, HB_C64 = 3 };
enum { HB_OK = 0, HB_ERR_CUDA = 1, HB_ERR_ARG = 2, HB_ERR_ALLOC = 3, HB_ERR_UNSUPPORTED = 4, HB_ERR_NCCL = 5, HB_ERR_NOT_CONVERGED = 6, HB_ERR_CALLBACK = 7 };
enum { HB_POINTER_HOST = 0, HB_POINTER_DEVICE = 1 };
enum { HB_H2D = 0, HB_D2H = 1, HB_D2D = 2 };

/* ---- library / context: gpu_engine (gpu/hala_gpu_engine.hpp:60-117), gpu_device_count (gpu/hala_cuda_common.hpp:195) ---- */
const char* hb_version(void);
const char* hb_last_error(void);
int hb_device_count(int *count);
int hb_ctx_create(int device, hb_ctx **ctx);
int hb_ctx_destroy(hb_ctx *ctx);
int hb_ctx_device(const hb_ctx *ctx, int *device);
int hb_ctx_set_stream(hb_ctx *ctx, void *cuda_stream);          /* gpu_engine::set_stream  (:86) */
int hb_ctx_get_stream(const hb_ctx *ctx, void **cuda_stream);
int hb_ctx_sync(hb_ctx *ctx);                                   /* gpu_engine::synchronize (:83) */
int hb_ctx_set_pointer_mode(hb_ctx *ctx, int mode);             /* set/reset_blas_device_pntr (:105-111) */
int hb_ctx_get_pointer_mode(const hb_ctx *ctx, int *mode);      /* get_blas_pointer_mode (:112-117) */
int hb_ctx_launch_count(const hb_ctx *ctx, long long *count);   /* kernels launched through this context so far */
/* per-kernel timing of the solver loops: while enabled, hb_cg / hb_pcg / hb_dist_cg bracket every launch of the first 512 iterations with
 * CUDA events on the context's stream; read = total milliseconds and number of launches of kernel `slot` of the iteration (CG: 0 SpMV
 * fused with <p,Ap>, 1 residual update + norm, 2 solution + direction update [hb_pcg: 2 <r,z>, 3 direction]) since it was enabled.
 * Costs four event records per iteration; off by default. */
int hb_ctx_profile(hb_ctx *ctx, int enable);
int hb_ctx_profile_read(const hb_ctx *ctx, int slot, double *ms_total, long long *launches);
int hb_ctx_trim(hb_ctx *ctx);                                   /* frees the cached solver workspace (kept across hb_cg / hb_gmres calls) */

/* ---- device timers (CUDA events on the context stream) — measurement plumbing for bench.py; the reference only has the
 *      wall-clock hala::chronometer (common/hala_core.hpp:155-162) ---- */
int hb_timer_start(hb_ctx *ctx);                                /* records the start event on the context stream */
int hb_timer_stop(hb_ctx *ctx, float *milliseconds);            /* records stop, synchronises on it, returns elapsed */

/* ---- memory: gpu_allocate / gpu_free / gpu_copy_n (gpu/hala_cuda_common.hpp:253-330), gpu_vector::fill (gpu/hala_gpu_vector.hpp:147),
 *      set_zero (gpu/hala_gpu_engine.hpp:335-352) ---- */
int hb_malloc(hb_ctx *ctx, size_t bytes, void **ptr);
int hb_free(hb_ctx *ctx, void *ptr);
int hb_memcpy(hb_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);        /* host-synchronous */
int hb_memcpy_async(hb_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);  /* on the context stream */
int hb_memset_zero(hb_ctx *ctx, void *ptr, size_t bytes);
int hb_fill(hb_ctx *ctx, int dtype, size_t n, const void *host_value, void *x);        /* dtype may also be -1: int32 */
/* context-free variants on an explicit device id — what gpu_vector needs: it knows a device id, not an engine (gpu/hala_gpu_vector.hpp:49-171) */
int hb_dev_malloc(int device, size_t bytes, void **ptr);
int hb_dev_free(void *ptr);
int hb_dev_memcpy(void *dst, const void *src, size_t bytes, int kind);                 /* host-synchronous, default stream */
int hb_dev_fill(int device, int dtype, size_t n, const void *host_value, void *x);     /* gpu_vector::fill; dtype -1: int32; host-synchronous */
int hb_host_alloc(size_t bytes, void **ptr);                                           /* pinned host memory */
int hb_host_free(void *ptr);

/* ---- CSR matrix view: gpu_sparse_matrix ctor / make_sparse_matrix (gpu/hala_cuda_sparse_general.hpp:215-227,382-401) ----
 * Non-owning of pntr/indx/vals (as the reference, :186-190,370-371); owns its one-time analysis tables. */
int hb_csr_create(hb_ctx *ctx, int dtype, int rows, int cols, int nnz,
                  const int *pntr, const int *indx, const void *vals, hb_csr **csr);
int hb_csr_destroy(hb_csr *csr);
int hb_csr_info(const hb_csr *csr, int *dtype, int *rows, int *cols, int *nnz, int *max_row_nnz);
/* workspace bytes a caller must supply to hb_spmv: always 0 here (reference: cusparseSpMV_bufferSize, :245-262,342-351) */
int hb_spmv_buffer_size(const hb_csr *csr, char trans, size_t *bytes);
/* y = alpha op(A) x + beta y; y is NOT read when beta == 0 (host mode) — gpu_sparse_matrix::gemv (:264-277) -> cusparseSpMV */
int hb_spmv(hb_ctx *ctx, const hb_csr *csr, char trans, const void *alpha, const void *x, const void *beta, void *y);
/* fused: y = A x and *dot_dev = <x, y> (conjugated for complex) in one pass; result stays on the device.
 * Replaces cusparseSpMV + cublas?dot of the CG loop (hex/solvers/hala_solvers_cg.hpp:129-133). */
int hb_spmv_dot(hb_ctx *ctx, const hb_csr *csr, const void *x, void *y, void *dot_dev);
/* selects the SpMV kernel variant for op 'N': 0 = auto, 1 = row-vector (sub-warp per row), 2 = staged tiles (LDG),
 * 3 = staged tiles (TMA bulk copy pipeline).  For benchmarking; auto is what the header layer uses. */
int hb_csr_set_variant(hb_csr *csr, int variant);
/* op 'T' / 'C' products (cusparseSpMV with a transposed operation, :264-277; pinned by tests/sparse_tests.hpp:184-190) run the op 'N'
 * kernel on a CSR of A^T that the object builds on the device at its first such product and keeps (nnz * (8 + sizeof value) + 4 cols
 * bytes).  The object is a non-owning view whose values the caller may change between products (:186-190), hence the mode:
 *   HB_TRANS_CHECKED (default) a 64-bit fingerprint pass over the value array per product, values re-gathered when it changed
 *   HB_TRANS_FROZEN            no check: the caller reports changes with hb_csr_values_changed()
 *   HB_TRANS_SCATTER           no cached copy: atomic scatter in the caller's row order (also the fallback when the copy does not fit)
 * The structure arrays (pntr, indx) are taken as fixed for the life of the object.  HB_TRANS_MODE=scatter|checked|frozen sets the
 * default of new objects. */
enum { HB_TRANS_SCATTER = 0, HB_TRANS_CHECKED = 1, HB_TRANS_FROZEN = 2 };
int hb_csr_set_transpose_mode(hb_csr *csr, int mode);
int hb_csr_values_changed(hb_csr *csr);
int hb_csr_transpose_info(const hb_csr *csr, int *mode, int *built, size_t *bytes);

/* ---- sparse triangular solves and ILU(0): gpu_triangular_matrix (gpu/hala_cuda_sparse_triangular.hpp:38-454 ->
 *      cusparseSpSV / cusparseSpSM) and gpu_ilu (gpu/hala_gpu_ilu.hpp:45-199 -> cusparse?csrilu02 + two triangular solves) ----
 * hb_tri_create : non-owning view of a CSR of which only the `uplo` ('L'/'U') triangle is used — the CSR may hold the whole
 *   matrix, as gpu_ilu passes it (gpu_ilu.hpp:88-89); diag 'U' = unit diagonal (a stored diagonal is ignored), 'N' = stored.
 *   The dependency analysis is done on first use per direction and cached; values are read at solve time.
 * hb_sptrsv : x = alpha * op(T)^-1 b, op = N / T / C (gpu_triangular_matrix::trsv -> cusparseSpSV_solve); b, x may alias (same stride)
 * hb_sptrsm : the same in place on nrhs right-hand sides: transb 'N' -> B is rows x nrhs column-major; 'T'/'C' -> B is nrhs x rows and
 *   every row is a right-hand side (gpu_triangular_matrix::trsm -> cusparseSpSM_solve)
 * hb_ilu0   : ILU(0) factors of a CSR with sorted rows and a full diagonal, in the pattern of the matrix (unit-lower L below the
 *   diagonal, U on and above it); `ilu` may alias `vals`; HB_ERR_ARG when a row lacks its diagonal (cusparse?csrilu02) */
int hb_tri_create(hb_ctx *ctx, int dtype, char uplo, char diag, int rows, int nnz,
                  const int *pntr, const int *indx, const void *vals, hb_tri **tri);
int hb_tri_destroy(hb_tri *tri);
int hb_tri_info(const hb_tri *tri, int *rows, int *nnz, int *nlevels);     /* nlevels: dependency levels of op 'N' (0 before its first solve) */
int hb_sptrsv(hb_ctx *ctx, hb_tri *tri, char trans, const void *alpha, const void *b, int incb, void *x, int incx);
int hb_sptrsm(hb_ctx *ctx, hb_tri *tri, char transa, char transb, int nrhs, const void *alpha, void *B, int ldb);
int hb_ilu0(hb_ctx *ctx, int dtype, int rows, int nnz, const int *pntr, const int *indx, const void *vals, void *ilu);

/* ---- multi right-hand-side pieces of the batch solvers ----
 * hb_spmm : C = alpha op(A) op(B) + beta C; C is M x N column-major (ldc), op(A) is M x K, B is stored b_rows x b_cols (ldb) and
 *   op(B) is K x N — gpu_sparse_matrix::gemm -> cusparseSpMM (gpu/hala_cuda_sparse_general.hpp:284-332).  C is not read when beta == 0.
 * hb_geam : C = alpha op(A) + beta op(B), M x N    — cublas?geam (gpu/hala_gpu_blas0.hpp:46-72)
 * hb_dgmm : C = diag(x) A ('L') or A diag(x) ('R') — cublas?dgmm (gpu/hala_gpu_blas0.hpp:79-103)
 * hb_tbsv : x <- op(A)^-1 x, A triangular banded with k off-diagonals in BLAS band storage — cublas?tbsv (gpu/hala_gpu_blas2.hpp);
 *   k == 0 is the element-wise divide behind hala::vdivide (wax/hala_blas_extensions.hpp:251-260) */
int hb_spmm(hb_ctx *ctx, const hb_csr *csr, char transa, char transb, int b_rows, int b_cols, const void *alpha,
            const void *B, int ldb, const void *beta, void *C, int ldc);
int hb_geam(hb_ctx *ctx, int dtype, char transa, char transb, int M, int N, const void *alpha, const void *A, int lda,
            const void *beta, const void *B, int ldb, void *C, int ldc);
int hb_dgmm(hb_ctx *ctx, int dtype, char side, int M, int N, const void *A, int lda, const void *x, int incx, void *C, int ldc);
int hb_tbsv(hb_ctx *ctx, int dtype, char uplo, char trans, char diag, int n, int k, const void *A, int lda, void *x, int incx);

/* ---- BLAS-1: gpu/hala_gpu_blas1.hpp  vcopy :48-65, norm2 :102-121, dot<conj> :178-198, axpy :204-222, scal :228-245 ---- */
int hb_copy(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *y, int incy);
int hb_axpy(hb_ctx *ctx, int dtype, int n, const void *alpha, const void *x, int incx, void *y, int incy);
int hb_scal(hb_ctx *ctx, int dtype, int n, const void *alpha, void *x, int incx);
int hb_dot (hb_ctx *ctx, int dtype, int conj, int n, const void *x, int incx, const void *y, int incy, void *result);
int hb_nrm2(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *result);
int hb_asum(hb_ctx *ctx, int dtype, int n, const void *x, int incx, void *result);   /* sum |re|+|im| (cublas?asum, gpu_blas1.hpp:124-143) */
/* the rest of BLAS-1, so that gpu_engine carries no cuBLAS handle:
 *   hb_swap  cublas?swap  (gpu_blas1.hpp:83-100)
 *   hb_iamax cublasI?amax (:153-172): *result = 1-BASED index of the first entry with the largest |re| + |im| (0 when n <= 0);
 *            the header layer subtracts 1 as the reference does
 *   hb_rot   cublas?rot / cublasCsrot / cublasZdrot (:269-300): x' = c x + s y, y' = c y - conj(s) x; c is always of the real
 *            type; s is of `dtype`, or of the real type when s_is_real != 0
 *   hb_rotm  cublas{S,D}rotm (:349-371), hb_rotg cublas?rotg (:251-262), hb_rotmg cublas{S,D}rotmg (:318-341): netlib semantics;
 *            scalar / param pointers are host or device pointers according to the pointer mode */
int hb_swap (hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy);
int hb_iamax(hb_ctx *ctx, int dtype, int n, const void *x, int incx, int *result);
int hb_rot  (hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy, const void *c, const void *s, int s_is_real);
int hb_rotm (hb_ctx *ctx, int dtype, int n, void *x, int incx, void *y, int incy, const void *param);
int hb_rotg (hb_ctx *ctx, int dtype, void *a, void *b, void *c, void *s);
int hb_rotmg(hb_ctx *ctx, int dtype, void *d1, void *d2, void *x1, const void *y1, void *param);

/* ---- BLAS-2 gemv, the Gram-Schmidt pair of GMRES: gpu/hala_gpu_blas2.hpp:39-62 (cublas?gemv), column-major A ---- */
int hb_gemv(hb_ctx *ctx, int dtype, char trans, int M, int N, const void *alpha, const void *A, int lda,
            const void *x, int incx, const void *beta, void *y, int incy);

/* ---- fused Krylov building blocks (device-resident scalars; no host synchronisation) ----
 * hb_multi_dot : h[c] = sum_i op(W[i,c]) r[i], c < k, one pass over r and the k basis columns (conj != 0 -> op = conj)
 * hb_multi_axpy_nrm2 : r -= W h (h on device), *nrm2sq_dev = sum |r_i|^2 of the UPDATED r, same pass
 *   together they replace gemv('T')+gemv('N')+nrm2 of krylov_project (hex/solvers/hala_solvers_gmres.hpp:67-72,192)
 * hb_axpy2_nrm2 : x += a p ; r -= a q ; *rr_dev = <r,r>   (CG update, hala_solvers_cg.hpp:135-138) with a read from device
 * hb_xpby : p = r + b p (b on device)                     (hala_solvers_cg.hpp:147-148)                              */
int hb_multi_dot(hb_ctx *ctx, int dtype, int conj, int rows, int k, const void *W, size_t ldw, const void *r, void *h_dev);
int hb_multi_axpy_nrm2(hb_ctx *ctx, int dtype, int rows, int k, const void *W, size_t ldw, const void *h_dev,
                       void *r, void *nrm2sq_dev);
int hb_axpy2_nrm2(hb_ctx *ctx, int dtype, int n, const void *a_dev, const void *p, const void *q, void *x, void *r, void *rr_dev);
int hb_xpby(hb_ctx *ctx, int dtype, int n, const void *r, const void *b_dev, void *p);

/* ---- solvers with the iteration on the device ----
 * hb_cg : unpreconditioned CG, recurrence, iteration counter and stop test of solve_cg_core
 *   (hex/solvers/hala_solvers_cg.hpp:92-156; stop: iterations == max_iter || ||r||_2 < tol, :225), three fused kernels per
 *   iteration, scalars never leave the device; the host only polls a completion flag.
 *   x: initial guess in, solution out. *iters = operator applications (reference return value), *res = final ||r||_2.
 * hb_gmres : restarted GMRES(m), identity preconditioner, classical Gram-Schmidt as fused multi-dot/multi-axpy,
 *   Givens/Hessenberg on the host exactly as hala_solvers_gmres.hpp:127-230 (quirks included); cproj != 0 uses the
 *   conjugated projection ('C'), which is what complex data needs (reference defect at :48,:69). */
int hb_cg(hb_ctx *ctx, const hb_csr *csr, const void *b, void *x, double tol, int max_iter, int *iters, double *res);
int hb_gmres(hb_ctx *ctx, const hb_csr *csr, const void *b, void *x, double tol, int max_outer, int restart, int cproj,
             int *iters, double *res);


/* ---- the same solvers with the caller's preconditioner (hala::preconditioner, hex/solvers/hala_solvers_core.hpp:75-118) ----
 * precon(user, in, out) must enqueue out = P^-1 in on the context's stream (device pointers, n scalars of the matrix type, never
 * aliased) and return 0; any other value aborts the solve with HB_ERR_CALLBACK.  It is called once per operator application; after
 * the device-side stop test fires it may still be called for a few iterations that are skipped (its output is then ignored).
 * hb_pcg   : solve_cg_core with z = P^-1 r (hala_solvers_cg.hpp:116,140-150): four launches of this library per iteration plus the
 *            preconditioner's own, no host synchronisation that stalls the stream.  precon == NULL is hb_cg.
 * hb_pgmres: solve_gmres with r = P^-1 (A w) (hala_solvers_gmres.hpp:176,186).  precon == NULL is hb_gmres. */
typedef int (*hb_precon_fn)(void *user, const void *in_dev, void *out_dev);
int hb_pcg(hb_ctx *ctx, const hb_csr *csr, const void *b, void *x, double tol, int max_iter, hb_precon_fn precon, void *user,
           int *iters, double *res);
int hb_pgmres(hb_ctx *ctx, const hb_csr *csr, const void *b, void *x, double tol, int max_outer, int restart, int cproj,
              hb_precon_fn precon, void *user, int *iters, double *res);

#ifdef __cplusplus
}
#endif
#endif
