// TEST INFRASTRUCTURE — not part of the product path.
//
// Thin extern "C" driver around the UNMODIFIED reference (LIBHALA/hala v1.1.0) cpu_engine path.
// It is compiled by oracle/Makefile from the reference headers where they lie under
// /root/reference (nothing is copied into this repository); the output goes to oracle/_ref/.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it.
//
// What it exposes (dtype: 0=float 1=double 2=complex<float> 3=complex<double>):
//   ref_spmv   -> hala::sparse_gemv(cpu_engine, ...)      sparse/hala_sparse_blas.hpp:232-236 -> sparse_gemv_array (hala_sparse_utils.hpp:103-118)
//   ref_copy/axpy/scal/dot/dotu/nrm2                      blas/hala_blas_1.hpp:53-61,209-217,316-335,277-290,104-121
//   ref_gemv                                              blas/hala_blas_2.hpp (gemv) — the Gram-Schmidt pair of hex/solvers/hala_solvers_gmres.hpp:47-50
//   ref_cg     -> hala::solve_cg(cpu_engine, ...)         hex/solvers/hala_solvers_cg.hpp:232-246 -> :181-227 -> solve_cg_core :92-156
//   ref_gmres  -> hala::solve_gmres(cpu_engine, ...)      hex/solvers/hala_solvers_gmres.hpp:127-230
//   ref_trsv   -> hala::sparse_trsv(cpu_triangular_matrix) sparse/hala_sparse_structs.hpp:257-269 -> sparse_trsv_array (hala_sparse_utils.hpp:283-335)
//   ref_spmm   -> hala::sparse_gemm(cpu_engine)           sparse/hala_sparse_utils.hpp:120-160
//   ref_batch_cg -> hala::solve_batch_cg(cpu_engine)      hex/solvers/hala_solvers_cg_batch.hpp:68-151
//   ref_ilu    -> factorize_ilu + make_ilu(cpu_engine).apply sparse/hala_sparse_ilu.hpp -> hala_sparse_utils.hpp:228-274
// The preconditioner is the identity lambda SURVEY.md §8(d) prescribes: hala::vcopy(engine, in, out).
//
// The same source is compiled twice: once against the stock headers (libhala_ref.so) and once with
// -DHALA_REF_CPATCH against a build-time shadow of hex/solvers in which krylov_project uses 'C'
// instead of 'T' (libhala_ref_cpatch.so) — the documented reference defect for complex GMRES
// (hala_solvers_gmres.hpp:48). Exported names carry the prefix ref_ / refc_.

#include "hala.hpp"
#include "hala_solvers.hpp"

#include <complex>
#include <chrono>
#include <cstring>

#ifdef HALA_REF_CPATCH
#define RNAME(x) refc_##x
#else
#define RNAME(x) ref_##x
#endif

namespace {

// minimal non-owning container: HALA needs value_type, size(), data() (common/hala_vector_defines.hpp:276-491)
template<typename T> struct view{
    using value_type = std::remove_const_t<T>;
    view(T *p, size_t n) : ptr(p), num(n){}
    size_t size() const{ return num; }
    T* data(){ return ptr; }
    T const* data() const{ return ptr; }
    T& operator[](size_t i){ return ptr[i]; }
    T const& operator[](size_t i) const{ return ptr[i]; }
    T *ptr; size_t num;
};

template<typename T> T rd(const void *p){ return *reinterpret_cast<T const*>(p); }

template<typename T>
int spmv(char trans, int M, int N, const void *alpha, int nnz, const int *pntr, const int *indx, const void *vals,
         const void *x, const void *beta, void *y){
    hala::cpu_engine e;
    view<const int> vp(pntr, (size_t) M + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz);
    bool n = (trans == 'N' || trans == 'n');
    view<const T> vx((T const*) x, (size_t) (n ? N : M));
    view<T> vy((T*) y, (size_t) (n ? M : N));
    hala::sparse_gemv(e, trans, M, N, rd<T>(alpha), vp, vi, vv, vx, rd<T>(beta), vy);
    return 0;
}

template<typename T>
int cg(int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
       double tol, int max_iter, int *iters){
    hala::cpu_engine e;
    using P = typename hala::define_standard_precision<T>::value_type;
    view<const int> vp(pntr, (size_t) nrows + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vb((T const*) b, (size_t) nrows);
    view<T> vx((T*) x, (size_t) nrows);
    hala::stop_criteria<P> stop((P) tol, max_iter);
    *iters = hala::solve_cg(e, stop, vp, vi, vv,
                            [&](auto const &in, auto &out)->void{ hala::vcopy(e, in, out); }, vb, vx);
    return 0;
}

template<typename T>
int gmres(int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
          double tol, int max_outer, int restart, int *iters){
    hala::cpu_engine e;
    using P = typename hala::define_standard_precision<T>::value_type;
    view<const int> vp(pntr, (size_t) nrows + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vb((T const*) b, (size_t) nrows);
    view<T> vx((T*) x, (size_t) nrows);
    hala::stop_criteria<P> stop((P) tol, max_outer);
    *iters = hala::solve_gmres(e, stop, restart, vp, vi, vv,
                               [&](auto const &in, auto &out)->void{ hala::vcopy(e, in, out); }, vb, vx);
    return 0;
}

template<typename T>
int blas1(int op, int n, const void *alpha, const void *x, int incx, void *y, int incy, void *result){
    hala::cpu_engine e;
    using P = typename hala::define_standard_precision<T>::value_type;
    size_t lx = (n > 0) ? (size_t) (1 + (n - 1) * incx) : 0, ly = (n > 0) ? (size_t) (1 + (n - 1) * incy) : 0;
    view<const T> px((T const*) x, lx), cy((T const*) y, ly);
    view<T> py((T*) y, ly);
    switch(op){
        case 0: hala::vcopy(e, n, px, incx, py, incy); break;
        case 1: hala::axpy(e, n, rd<T>(alpha), px, incx, py, incy); break;
        case 2: hala::scal(e, n, rd<T>(alpha), py, incy); break;
        case 3: *reinterpret_cast<T*>(result) = hala::dot(e, n, px, incx, cy, incy); break;
        case 4: *reinterpret_cast<T*>(result) = hala::dotu(e, n, px, incx, cy, incy); break;
        case 5: *reinterpret_cast<P*>(result) = hala::norm2(e, n, px, incx); break;
        default: return 1;
    }
    return 0;
}

template<typename T>
int gemv(char trans, int M, int N, const void *alpha, const void *A, int lda, const void *x, const void *beta, void *y){
    hala::cpu_engine e;
    bool n = (trans == 'N' || trans == 'n');
    view<const T> vA((T const*) A, (size_t) lda * (size_t) N), vx((T const*) x, (size_t) (n ? N : M));
    view<T> vy((T*) y, (size_t) (n ? M : N));
    hala::gemv(e, trans, M, N, rd<T>(alpha), vA, lda, vx, 1, rd<T>(beta), vy, 1);
    return 0;
}

// sparse_trsv through cpu_triangular_matrix (sparse/hala_sparse_structs.hpp:257-269 -> sparse_trsv_array, hala_sparse_utils.hpp:283-335)
template<typename T>
int trsv(char uplo, char diag, char trans, int n, const void *alpha, int nnz, const int *pntr, const int *indx, const void *vals,
         const void *b, void *x){
    hala::cpu_engine e;
    view<const int> vp(pntr, (size_t) n + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vb((T const*) b, (size_t) n);
    view<T> vx((T*) x, (size_t) n);
    auto tri = hala::make_triangular_matrix(e, uplo, diag, vp, vi, vv);
    hala::sparse_trsv(trans, tri, rd<T>(alpha), vb, vx);
    return 0;
}
// make_ilu(cpu_engine) + apply (sparse/hala_sparse_ilu.hpp -> factorize_ilu_array / apply_ilu_array, hala_sparse_utils.hpp:228-274);
// also returns the factors themselves through factorize_ilu
template<typename T>
int ilu(int n, int nnz, const int *pntr, const int *indx, const void *vals, void *ilu_out, const void *x, void *r){
    hala::cpu_engine e;
    view<const int> vp(pntr, (size_t) n + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vx((T const*) x, (size_t) n);
    view<T> vr((T*) r, (size_t) n);
    std::vector<int> diag;
    std::vector<T> factors;
    hala::get_diagonal_index(vp, vi, diag);
    hala::factorize_ilu(vp, vi, vv, diag, factors);
    std::memcpy(ilu_out, factors.data(), sizeof(T) * (size_t) nnz);
    auto pre = hala::make_ilu(e, vp, vi, vv, 'N');
    pre.apply(vx, vr);
    return 0;
}

// hala::sparse_gemm(cpu_engine) -> sparse_gemm_array (sparse/hala_sparse_utils.hpp:120-160)
template<typename T>
int spmm(char transa, char transb, int M, int N, int K, const void *alpha, int nnz, const int *pntr, const int *indx, const void *vals,
         const void *B, int ldb, const void *beta, void *C, int ldc){
    hala::cpu_engine e;
    bool an = (transa == 'N' || transa == 'n'), bn = (transb == 'N' || transb == 'n');
    view<const int> vp(pntr, (size_t) (an ? M : K) + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vB((T const*) B, (size_t) ldb * (size_t) (bn ? N : K));
    view<T> vC((T*) C, (size_t) ldc * (size_t) N);
    hala::sparse_gemm(e, transa, transb, M, N, K, rd<T>(alpha), vp, vi, vv, vB, ldb, rd<T>(beta), vC, ldc);
    return 0;
}
// hala::solve_batch_cg(cpu_engine) with the identity preconditioner (hex/solvers/hala_solvers_cg_batch.hpp:68-151)
template<typename T>
int batch_cg(int nrows, int nnz, int nrhs, const int *pntr, const int *indx, const void *vals, const void *B, void *X, double tol, int max_iter, int *iters){
    hala::cpu_engine e;
    using P = typename hala::define_standard_precision<T>::value_type;
    view<const int> vp(pntr, (size_t) nrows + 1), vi(indx, (size_t) nnz);
    view<const T> vv((T const*) vals, (size_t) nnz), vB((T const*) B, (size_t) nrows * (size_t) nrhs);
    view<T> vX((T*) X, (size_t) nrows * (size_t) nrhs);
    hala::stop_criteria<P> stop((P) tol, max_iter);
    *iters = hala::solve_batch_cg(e, stop, vp, vi, vv,
                                  [&](auto const &in, auto &out)->void{ hala::vcopy(e, in, out); }, vB, vX);
    return 0;
}

#define DISPATCH(dtype, call) \
    switch(dtype){ \
        case 0: { using T = float; return call; } \
        case 1: { using T = double; return call; } \
        case 2: { using T = std::complex<float>; return call; } \
        case 3: { using T = std::complex<double>; return call; } \
        default: return 2; \
    }

}

extern "C" {

int RNAME(spmv)(int dtype, char trans, int M, int N, const void *alpha, int nnz, const int *pntr, const int *indx,
                const void *vals, const void *x, const void *beta, void *y){
    try{ DISPATCH(dtype, spmv<T>(trans, M, N, alpha, nnz, pntr, indx, vals, x, beta, y)) }catch(...){ return 3; }
}
int RNAME(cg)(int dtype, int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
              double tol, int max_iter, int *iters){
    try{ DISPATCH(dtype, cg<T>(nrows, nnz, pntr, indx, vals, b, x, tol, max_iter, iters)) }catch(...){ return 3; }
}
int RNAME(gmres)(int dtype, int nrows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x,
                 double tol, int max_outer, int restart, int *iters){
    try{ DISPATCH(dtype, gmres<T>(nrows, nnz, pntr, indx, vals, b, x, tol, max_outer, restart, iters)) }catch(...){ return 3; }
}
int RNAME(blas1)(int dtype, int op, int n, const void *alpha, const void *x, int incx, void *y, int incy, void *result){
    try{ DISPATCH(dtype, blas1<T>(op, n, alpha, x, incx, y, incy, result)) }catch(...){ return 3; }
}
int RNAME(gemv)(int dtype, char trans, int M, int N, const void *alpha, const void *A, int lda, const void *x,
                const void *beta, void *y){
    try{ DISPATCH(dtype, gemv<T>(trans, M, N, alpha, A, lda, x, beta, y)) }catch(...){ return 3; }
}
int RNAME(trsv)(int dtype, char uplo, char diag, char trans, int n, const void *alpha, int nnz, const int *pntr, const int *indx,
                const void *vals, const void *b, void *x){
    try{ DISPATCH(dtype, trsv<T>(uplo, diag, trans, n, alpha, nnz, pntr, indx, vals, b, x)) }catch(...){ return 3; }
}
int RNAME(ilu)(int dtype, int n, int nnz, const int *pntr, const int *indx, const void *vals, void *ilu_out, const void *x, void *r){
    try{ DISPATCH(dtype, ilu<T>(n, nnz, pntr, indx, vals, ilu_out, x, r)) }catch(...){ return 3; }
}
int RNAME(spmm)(int dtype, char transa, char transb, int M, int N, int K, const void *alpha, int nnz, const int *pntr, const int *indx,
                const void *vals, const void *B, int ldb, const void *beta, void *C, int ldc){
    try{ DISPATCH(dtype, spmm<T>(transa, transb, M, N, K, alpha, nnz, pntr, indx, vals, B, ldb, beta, C, ldc)) }catch(...){ return 3; }
}
int RNAME(batch_cg)(int dtype, int nrows, int nnz, int nrhs, const int *pntr, const int *indx, const void *vals, const void *B, void *X,
                    double tol, int max_iter, int *iters){
    try{ DISPATCH(dtype, batch_cg<T>(nrows, nnz, nrhs, pntr, indx, vals, B, X, tol, max_iter, iters)) }catch(...){ return 3; }
}
const char* RNAME(version)(){ return "LIBHALA/hala " HALA_VERSION_STRING " cpu_engine"; }

}
