/* TEST INFRASTRUCTURE — CPU oracle for the hala_b200 hot path. NOT part of the product.
 *
 * A plain-C, single-threaded restatement of what LIBHALA/hala's cpu_engine path computes for
 *   CSR SpMV (sparse/hala_sparse_utils.hpp:103-118), BLAS-1 (blas/hala_blas_1.hpp), the Gram-Schmidt
 *   gemv pair, CG (hex/solvers/hala_solvers_cg.hpp:92-156,181-227) and GMRES (hala_solvers_gmres.hpp:127-230),
 *   and the sparse triangular solve and ILU(0) (sparse/hala_sparse_utils.hpp:228-335).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load liboracle.so, and only
 * as the checker. The product (libhalab200.so) never links or calls it.
 *
 * Third-party algorithm note: for real types the reference forwards BLAS-1/2 to an external Fortran BLAS
 * ("an implementation of BLAS", CMakeLists.txt:124-138 — un-vendored, no pinned version; OpenBLAS 0.3.15
 * is what this image offers). Those routines are restated here from the published netlib reference BLAS
 * 3.8 algorithms (?copy ?axpy ?scal ?dot ?nrm2 ?gemv ?rot ?rotg ?tpsv); summation order therefore differs
 * from OpenBLAS's SIMD kernels in the last bits, which the parity tolerances (1e-13 / 1e-5) absorb.
 *
 * Parity pinned: tests/test_oracle.py checks this file against (1) the golden vectors harvested from
 * the reference's own tests (tests/golden/ref_tests.json: tests/sparse_tests.hpp:166-191,
 * tests/solvers_tests.hpp, cmake/post_install_test.sh:24-38) and (2) outputs of the unmodified reference
 * (oracle/_ref/libhala_ref.so) recorded in tests/golden/*.npz by tests/golden/make_golden.py, and live
 * against oracle/_ref when it is present.
 */
#include <stdlib.h>
#include <math.h>
#include <complex.h>
#include <string.h>

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

/* ---- float ---- */
#define T float
#define R float
#define NAME(x) CAT(x, _s)
#define CPLX 0
#define CONJ(z) (z)
#define ABS(z) fabsf(z)
#define SQRT(z) sqrtf(z)
#include "hb_oracle_impl.h"
#undef T
#undef R
#undef NAME
#undef CPLX
#undef CONJ
#undef ABS
#undef SQRT

/* ---- double ---- */
#define T double
#define R double
#define NAME(x) CAT(x, _d)
#define CPLX 0
#define CONJ(z) (z)
#define ABS(z) fabs(z)
#define SQRT(z) sqrt(z)
#include "hb_oracle_impl.h"
#undef T
#undef R
#undef NAME
#undef CPLX
#undef CONJ
#undef ABS
#undef SQRT

/* ---- complex<float> ---- */
#define T float _Complex
#define R float
#define NAME(x) CAT(x, _c)
#define CPLX 1
#define CONJ(z) conjf(z)
#define ABS(z) cabsf(z)
#define REAL(z) crealf(z)
#define IMAG(z) cimagf(z)
#define SQRT(z) sqrtf(z)
#include "hb_oracle_impl.h"
#undef T
#undef R
#undef NAME
#undef CPLX
#undef CONJ
#undef ABS
#undef REAL
#undef IMAG
#undef SQRT

/* ---- complex<double> ---- */
#define T double _Complex
#define R double
#define NAME(x) CAT(x, _z)
#define CPLX 1
#define CONJ(z) conj(z)
#define ABS(z) cabs(z)
#define REAL(z) creal(z)
#define IMAG(z) cimag(z)
#define SQRT(z) sqrt(z)
#include "hb_oracle_impl.h"
#undef T
#undef R
#undef NAME
#undef CPLX
#undef CONJ
#undef ABS
#undef REAL
#undef IMAG
#undef SQRT

/* ---------------- exported C ABI (same shapes as oracle/ref_driver.cpp, prefix orc_) ----------------
 * dtype: 0 = float, 1 = double, 2 = complex<float>, 3 = complex<double>; scalars by pointer to that dtype. */

#define SW(dtype, S, D, C, Z) switch(dtype){ case 0: S; break; case 1: D; break; case 2: C; break; case 3: Z; break; default: return 2; }

int orc_spmv(int dtype, char trans, int M, int N, const void *alpha, int nnz, const int *pntr, const int *indx,
             const void *vals, const void *x, const void *beta, void *y){
    (void) nnz;
    SW(dtype,
       spmv_s(trans, M, N, *(const float*) alpha, pntr, indx, vals, x, *(const float*) beta, y),
       spmv_d(trans, M, N, *(const double*) alpha, pntr, indx, vals, x, *(const double*) beta, y),
       spmv_c(trans, M, N, *(const float _Complex*) alpha, pntr, indx, vals, x, *(const float _Complex*) beta, y),
       spmv_z(trans, M, N, *(const double _Complex*) alpha, pntr, indx, vals, x, *(const double _Complex*) beta, y))
    return 0;
}

int orc_cg(int dtype, int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
           double tol, int max_iter, int *iters){
    (void) nnz;
    SW(dtype,
       *iters = cg_s(nrows, pntr, indx, vals, b, x, (float) tol, max_iter),
       *iters = cg_d(nrows, pntr, indx, vals, b, x, tol, max_iter),
       *iters = cg_c(nrows, pntr, indx, vals, b, x, (float) tol, max_iter),
       *iters = cg_z(nrows, pntr, indx, vals, b, x, tol, max_iter))
    return 0;
}

int orc_gmres(int dtype, int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
              double tol, int max_outer, int restart, int cproj, int *iters){
    (void) nnz;
    SW(dtype,
       *iters = gmres_s(nrows, pntr, indx, vals, b, x, (float) tol, max_outer, restart, cproj),
       *iters = gmres_d(nrows, pntr, indx, vals, b, x, tol, max_outer, restart, cproj),
       *iters = gmres_c(nrows, pntr, indx, vals, b, x, (float) tol, max_outer, restart, cproj),
       *iters = gmres_z(nrows, pntr, indx, vals, b, x, tol, max_outer, restart, cproj))
    return 0;
}

/* op: 0 copy, 1 axpy, 2 scal (on y), 3 dot (conj), 4 dotu, 5 nrm2 (of x) */
int orc_blas1(int dtype, int op, int n, const void *alpha, const void *x, int incx, void *y, int incy, void *result){
#define B1(sfx, TT, RR) \
    switch(op){ \
        case 0: copy_##sfx(n, x, incx, y, incy); break; \
        case 1: axpy_##sfx(n, *(const TT*) alpha, x, incx, y, incy); break; \
        case 2: scal_##sfx(n, *(const TT*) alpha, y, incy); break; \
        case 3: *(TT*) result = dot_##sfx(1, n, x, incx, y, incy); break; \
        case 4: *(TT*) result = dot_##sfx(0, n, x, incx, y, incy); break; \
        case 5: *(RR*) result = nrm2_##sfx(n, x, incx); break; \
        default: return 1; }
    SW(dtype, B1(s, float, float), B1(d, double, double), B1(c, float _Complex, float), B1(z, double _Complex, double))
    return 0;
}

int orc_gemv(int dtype, char trans, int M, int N, const void *alpha, const void *A, int lda, const void *x,
             const void *beta, void *y){
    SW(dtype,
       gemv_s(trans, M, N, *(const float*) alpha, A, lda, x, *(const float*) beta, y),
       gemv_d(trans, M, N, *(const double*) alpha, A, lda, x, *(const double*) beta, y),
       gemv_c(trans, M, N, *(const float _Complex*) alpha, A, lda, x, *(const float _Complex*) beta, y),
       gemv_z(trans, M, N, *(const double _Complex*) alpha, A, lda, x, *(const double _Complex*) beta, y))
    return 0;
}

/* general != 0: any CSR, only the uplo triangle is used (cuSPARSE fill-mode semantics); general == 0: the reference's layout
 * (one triangle, diagonal last / first) */
int orc_trsv(int dtype, int general, char uplo, char diag, char trans, int n, const void *alpha, const int *pntr, const int *indx,
             const void *vals, const void *b, void *x){
#define TR(sfx, TT) if (general) trsv_general_##sfx(uplo, diag, trans, n, *(const TT*) alpha, pntr, indx, vals, b, x); \
                    else trsv_##sfx(uplo, diag, trans, n, *(const TT*) alpha, pntr, indx, vals, b, x)
    SW(dtype, TR(s, float), TR(d, double), TR(c, float _Complex), TR(z, double _Complex))
    return 0;
}
/* ilu = ILU(0) factors in the pattern of the matrix; diag = position of each row's diagonal. Returns 3 when a diagonal is missing. */
int orc_ilu_factor(int dtype, int n, const int *pntr, const int *indx, const void *vals, int *diag, void *ilu){
    int rc = 0;
    SW(dtype, rc = ilu_factor_s(n, pntr, indx, vals, diag, ilu), rc = ilu_factor_d(n, pntr, indx, vals, diag, ilu),
              rc = ilu_factor_c(n, pntr, indx, vals, diag, ilu), rc = ilu_factor_z(n, pntr, indx, vals, diag, ilu))
    return rc ? 3 : 0;
}
int orc_ilu_apply(int dtype, int n, const int *pntr, const int *indx, const int *diag, const void *ilu, void *x){
    SW(dtype, ilu_apply_s(n, pntr, indx, diag, ilu, x), ilu_apply_d(n, pntr, indx, diag, ilu, x),
              ilu_apply_c(n, pntr, indx, diag, ilu, x), ilu_apply_z(n, pntr, indx, diag, ilu, x))
    return 0;
}

int orc_spmm(int dtype, char transa, char transb, int M, int N, int K, const void *alpha, int nnz, const int *pntr, const int *indx,
             const void *vals, const void *B, int ldb, const void *beta, void *C, int ldc){
    (void) nnz;
    SW(dtype,
       spmm_s(transa, transb, M, N, K, *(const float*) alpha, pntr, indx, vals, B, ldb, *(const float*) beta, C, ldc),
       spmm_d(transa, transb, M, N, K, *(const double*) alpha, pntr, indx, vals, B, ldb, *(const double*) beta, C, ldc),
       spmm_c(transa, transb, M, N, K, *(const float _Complex*) alpha, pntr, indx, vals, B, ldb, *(const float _Complex*) beta, C, ldc),
       spmm_z(transa, transb, M, N, K, *(const double _Complex*) alpha, pntr, indx, vals, B, ldb, *(const double _Complex*) beta, C, ldc))
    return 0;
}

const char* orc_version(void){ return "hala_b200 CPU oracle (restatement of LIBHALA/hala 1.1.0 cpu_engine path)"; }

/* Host generator of the bench matrix for the reference arm of bench.py (the numpy generator hala_b200/matgen._stencil needs minutes
 * and ~40 GB at 512^3): rows [row_lo, row_hi) of the n^3 7-point stencil with lexicographic index (k*n + j)*n + i, Dirichlet
 * truncation, lower neighbours `lower`, diagonal `diag`, upper neighbours `upper` (Laplacian: -1, 6, -1; convection-diffusion:
 * -1-d, 6, -1+d).  Bit-identical to matgen.lap3d7 / convdiff7 (tests/test_oracle.py).  pntr has row_hi - row_lo + 1 entries and
 * starts at 0; returns the number of non-zeros written (call with indx == NULL to size the arrays). */
long long orc_gen_stencil7(int n, long long row_lo, long long row_hi, double lower, double diag, double upper, int *pntr, int *indx, double *vals){
    const long long n2 = (long long) n * n;
    long long nz = 0;
    if (pntr) pntr[0] = 0;
    for (long long row = row_lo; row < row_hi; row++){
        const int i = (int) (row % n), j = (int) ((row / n) % n), k = (int) (row / n2);
        const long long col[7] = {row - n2, row - n, row - 1, row, row + 1, row + n, row + n2};
        const int ok[7] = {k > 0, j > 0, i > 0, 1, i < n - 1, j < n - 1, k < n - 1};
        const double v[7] = {lower, lower, lower, diag, upper, upper, upper};
        for (int s = 0; s < 7; s++) if (ok[s]){
            if (indx){ indx[nz] = (int) col[s]; vals[nz] = v[s]; }
            nz++;
        }
        if (pntr) pntr[row - row_lo + 1] = (int) nz;
    }
    return nz;
}
