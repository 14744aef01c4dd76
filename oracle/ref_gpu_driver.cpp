// TEST / BENCH INFRASTRUCTURE — not part of the product path.
//
// The UNMODIFIED reference's own GPU path (LIBHALA/hala v1.1.0 built with -DHALA_ENABLE_CUDA: cuSPARSE + cuBLAS behind
// gpu_engine) as "the kernel to beat on the same box" (SURVEY.md §2.1).  Compiled by oracle/Makefile from the
// reference headers where they lie under /root/reference; the output goes to oracle/_ref/libhala_ref_gpu.so.  Only bench.py's
// gpu_reference leg and scripts/ load it; nothing under hala_b200/ does.
//
//   refgpu_spmv -> hala::make_sparse_matrix(gpu_engine, ...).gemv   gpu/hala_cuda_sparse_general.hpp:264-277  (cusparseSpMV, ALG_DEFAULT)
//   refgpu_cg   -> hala::solve_cg(gpu_engine, ...)                  hex/solvers/hala_solvers_cg.hpp:232-246 -> solve_cg_core :92-156
//                                                                   through gpu/hala_gpu_blas1.hpp:102-245 (cublas dot/axpy/scal/nrm2/copy)
//   refgpu_gmres-> hala::solve_gmres(gpu_engine, ...)               hex/solvers/hala_solvers_gmres.hpp:127-230 (cublas gemv pair)
// All arrays are DEVICE pointers (the bench builds its matrices in HBM); dtype: 1 = double, 3 = complex<double>.
#include "hala.hpp"
#include "hala_solvers.hpp"

#include <chrono>
#include <complex>
#include <cstring>
#include <string>

namespace {
std::string g_err;

template<typename T>
int spmv_t(int rows, int cols, int nnz, const int *pntr, const int *indx, const void *vals, const void *x, void *y, int warmup, int reps, double *us){
    hala::gpu_engine engine(0);
    auto p = hala::wrap_gpu_array(pntr, (size_t) rows + 1);
    auto i = hala::wrap_gpu_array(indx, (size_t) nnz);
    auto v = hala::wrap_gpu_array((T const*) vals, (size_t) nnz);
    auto gx = hala::wrap_gpu_array((T const*) x, (size_t) cols);
    auto gy = hala::wrap_gpu_array((T*) y, (size_t) rows);
    auto matrix = hala::make_sparse_matrix(engine, rows, cols, nnz, p, i, v);
    size_t bsize = matrix.gemv_buffer_size('N', 1.0, gx, 0.0, gy);
    hala::gpu_vector<T> buffer(bsize / sizeof(T) + 1, engine.device());       // as solve_cg sizes it (:207-209)
    for(int k=0; k<warmup; k++) matrix.gemv('N', 1.0, gx, 0.0, gy, buffer);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, 0);
    for(int k=0; k<reps; k++) matrix.gemv('N', 1.0, gx, 0.0, gy, buffer);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *us = 1.0e3 * ms / reps;
    return 0;
}

template<typename T>
int cg_t(int rows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x, double tol, int max_iter, int *iters, double *seconds){
    using P = typename hala::define_standard_precision<T>::value_type;
    hala::gpu_engine engine(0);
    auto p = hala::wrap_gpu_array(pntr, (size_t) rows + 1);
    auto i = hala::wrap_gpu_array(indx, (size_t) nnz);
    auto v = hala::wrap_gpu_array((T const*) vals, (size_t) nnz);
    auto gb = hala::wrap_gpu_array((T const*) b, (size_t) rows);
    auto gx = hala::wrap_gpu_array((T*) x, (size_t) rows);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    *iters = hala::solve_cg(engine, hala::stop_criteria<P>((P) tol, max_iter), p, i, v,
                            [&](auto const &in, auto &out)->void{ hala::vcopy(engine, in, out); }, gb, gx);
    cudaDeviceSynchronize();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

template<typename T>
int gmres_t(int rows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x, double tol, int max_outer, int restart,
            int *iters, double *seconds){
    using P = typename hala::define_standard_precision<T>::value_type;
    hala::gpu_engine engine(0);
    auto p = hala::wrap_gpu_array(pntr, (size_t) rows + 1);
    auto i = hala::wrap_gpu_array(indx, (size_t) nnz);
    auto v = hala::wrap_gpu_array((T const*) vals, (size_t) nnz);
    auto gb = hala::wrap_gpu_array((T const*) b, (size_t) rows);
    auto gx = hala::wrap_gpu_array((T*) x, (size_t) rows);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    *iters = hala::solve_gmres(engine, hala::stop_criteria<P>((P) tol, max_outer), restart, p, i, v,
                               [&](auto const &in, auto &out)->void{ hala::vcopy(engine, in, out); }, gb, gx);
    cudaDeviceSynchronize();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

template<class F> int guarded(F f){
    try{ return f(); }
    catch(std::exception &e){ g_err = e.what(); return 1; }
    catch(...){ g_err = "unknown exception"; return 1; }
}
}

extern "C" {

const char* refgpu_last_error(){ return g_err.c_str(); }
const char* refgpu_version(){ return "LIBHALA/hala " HALA_VERSION_STRING " gpu_engine: cuSPARSE cusparseSpMV(ALG_DEFAULT) + cuBLAS level 1/2"; }

int refgpu_spmv(int dtype, int rows, int cols, int nnz, const int *pntr, const int *indx, const void *vals, const void *x, void *y,
                int warmup, int reps, double *us_per_product){
    return guarded([&]()->int{
        if (dtype == 1) return spmv_t<double>(rows, cols, nnz, pntr, indx, vals, x, y, warmup, reps, us_per_product);
        if (dtype == 3) return spmv_t<std::complex<double>>(rows, cols, nnz, pntr, indx, vals, x, y, warmup, reps, us_per_product);
        if (dtype == 0) return spmv_t<float>(rows, cols, nnz, pntr, indx, vals, x, y, warmup, reps, us_per_product);
        g_err = "dtype"; return 2;
    });
}
int refgpu_cg(int dtype, int rows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x, double tol, int max_iter,
              int *iters, double *seconds){
    return guarded([&]()->int{
        if (dtype == 1) return cg_t<double>(rows, nnz, pntr, indx, vals, b, x, tol, max_iter, iters, seconds);
        g_err = "dtype"; return 2;
    });
}
int refgpu_gmres(int dtype, int rows, int nnz, const int *pntr, const int *indx, const void *vals, const void *b, void *x, double tol, int max_outer,
                 int restart, int *iters, double *seconds){
    return guarded([&]()->int{
        if (dtype == 1) return gmres_t<double>(rows, nnz, pntr, indx, vals, b, x, tol, max_outer, restart, iters, seconds);
        g_err = "dtype"; return 2;
    });
}

}
