// Runs the REFERENCE'S OWN test bodies on the B200 backend.  Built in the build container only (it includes test headers
// from /root/reference/tests at compile time — nothing is copied into this repository); the binary lands in tests/_bin/ and
// is executed on the GPU box by tests/test_cpp_dropin.py.  The include path puts hala_b200/gpu/ where the reference's gpu/
// directory would be, so every hala:: call below reaches libhalab200.so through the unmodified wax / hex templates.
//
//   tests/cuda_core_tests.hpp   load_unload                      (gpu_vector / engine semantics)
//   tests/cuda_blas1_tests.hpp  norm2, dot, axpy, scal           (direct gpu_engine BLAS-1)
//   tests/cuda_blas2_tests.hpp  gemv                             (direct gpu_engine gemv)
//   tests/cuda_sparse_tests.hpp sparse_matvec                    (one-shot sparse_gemv, output resized)
//   tests/blas1_tests.hpp       test_vcopy, test_vswap, test_axpy, test_rscalar, test_scal, test_iamax, test_rotate
//                               with mixed_engine, engine + no-engine API, strides
//   tests/sparse_tests.hpp      test_sparse_gemv, test_sparse_trsv, test_sparse_trsm   with mixed_engine, N/T/C
//   tests/solvers_tests.hpp     test_cg, test_cg_batch, test_gmres   the reference's ILU-preconditioned solver tests, mixed_engine
//   tests/cuda_blas0_tests.hpp  transpose, test_geam, test_dgmm  (dense helpers of the batch solvers)
//   tests/sparse_tests.hpp      test_sparse_gemm, test_batch_axpy / dot / max_norm / scale   (SpMM and the wax batch templates on top of our primitives)
// plus solver checks written here: solve_cg / solve_gmres (reference templates, identity preconditioner) on gpu_engine and
// mixed_engine against cpu_engine, 4 scalar types.
#include "cuda_core_tests.hpp"
#include "cuda_blas0_tests.hpp"
#include "cuda_blas1_tests.hpp"
#include "cuda_blas2_tests.hpp"
#include "cuda_sparse_tests.hpp"
#include "solvers_tests.hpp"

namespace eng_api {      // the "engine as first argument" flavour of the shared test bodies
#undef mengine
#undef bind
#define mengine engine,
#define bind(x) (x)
#include "blas1_tests.hpp"
#include "sparse_tests.hpp"
}

template<typename T> std::vector<T> tvec(std::initializer_list<double> l){ std::vector<T> v; for (double d : l) v.push_back(hala::get_cast<T>(d)); return v; }

// 2-D 5-point Laplacian n x n, the smallest member of the BASELINE workload family
template<typename T> void lap2d(int n, std::vector<int> &pntr, std::vector<int> &indx, std::vector<T> &vals){
    pntr = {0};
    for(int i=0; i<n; i++) for(int j=0; j<n; j++){
        auto add = [&](int ii, int jj, double v){ if (ii >= 0 && ii < n && jj >= 0 && jj < n){ indx.push_back(ii * n + jj); vals.push_back(hala::get_cast<T>(v)); } };
        add(i-1, j, -1); add(i, j-1, -1); add(i, j, 4); add(i, j+1, -1); add(i+1, j, -1);
        pntr.push_back((int) indx.size());
    }
}

template<typename T> void solvers_on_engines(){
    current_test<T> tests("cg/gmres drop-in");
    using P = typename hala::define_standard_precision<T>::value_type;
    const int n = 24, N = n * n;
    const P tol = std::is_same<P, float>::value ? 1.E-4f : 1.E-9;
    std::vector<int> pntr, indx; std::vector<T> vals;
    lap2d<T>(n, pntr, indx, vals);
    std::vector<T> b(N, hala::get_cast<T>(1.0 / n)), xcpu, xgpu, xmix;

    hala::cpu_engine ecpu;
    hala::gpu_engine egpu(0);
    hala::mixed_engine emix(egpu);

    int it_cpu = hala::solve_cg(ecpu, hala::stop_criteria<P>(tol, 1000), pntr, indx, vals,
                                [&](auto const &in, auto &out)->void{ hala::vcopy(ecpu, in, out); }, b, xcpu);
    auto gp = egpu.load(pntr); auto gi = egpu.load(indx); auto gv = egpu.load(vals); auto gb = egpu.load(b);
    hala::gpu_vector<T> gx(egpu.device());
    int it_gpu = hala::solve_cg(egpu, hala::stop_criteria<P>(tol, 1000), gp, gi, gv,
                                [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, gb, gx);
    gx.unload(xgpu);
    int it_mix = hala::solve_cg(emix, hala::stop_criteria<P>(tol, 1000), pntr, indx, vals,
                                [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, b, xmix);
    hassert(std::abs(it_cpu - it_gpu) <= 2);
    hassert(std::abs(it_cpu - it_mix) <= 2);
    hassert(testvec(xgpu, xcpu, 1.E+6 * hala::norm2(xcpu)));      // both solved to tol: agree to ~tol, not to round-off
    hassert(testvec(xmix, xcpu, 1.E+6 * hala::norm2(xcpu)));

    // GMRES on a nonsymmetric variant (upwind-ish perturbation of the off-diagonals)
    for(int i=0; i<N; i++) for(int j=pntr[i]; j<pntr[i+1]; j++) if (indx[j] < i) vals[j] = hala::get_cast<T>(-1.5); else if (indx[j] > i) vals[j] = hala::get_cast<T>(-0.5);
    std::vector<T> ycpu, ygpu;
    int g_cpu = hala::solve_gmres(ecpu, hala::stop_criteria<P>(tol, 100), 20, pntr, indx, vals,
                                  [&](auto const &in, auto &out)->void{ hala::vcopy(ecpu, in, out); }, b, ycpu);
    gv.load(vals);
    hala::gpu_vector<T> gy(egpu.device());
    int g_gpu = hala::solve_gmres(egpu, hala::stop_criteria<P>(tol, 100), 20, gp, gi, gv,
                                  [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, gb, gy);
    gy.unload(ygpu);
    hassert(std::abs(g_cpu - g_gpu) <= 2);
    hassert(testvec(ygpu, ycpu, 1.E+6 * hala::norm2(ycpu)));
}


// Row f3 through the header layer: a gpu_sparse_matrix kept across products, op 'T' / 'C' on the cached transpose in its three modes,
// values rewritten in place between products (the view is non-owning, reference gpu/hala_cuda_sparse_general.hpp:186-190), against
// hala::sparse_gemv on the CPU engine.
template<typename T> struct mkval { static T get(double a, double){ return (T) a; } };
template<typename R> struct mkval<std::complex<R>> { static std::complex<R> get(double a, double b){ return std::complex<R>((R) a, (R) b); } };
template<typename T> void transposed_products(){
    current_test<T> tests("gpu sparse T/C cached");
    const int n = 9, N = n * n * n;
    std::vector<int> pntr(1, 0), indx;
    std::vector<T> vals;
    for(int k=0; k<n; k++) for(int j=0; j<n; j++) for(int i=0; i<n; i++){        // nonsymmetric 7-point stencil
        auto add = [&](int kk, int jj, int ii, double v){ if (kk>=0 && kk<n && jj>=0 && jj<n && ii>=0 && ii<n){ indx.push_back((kk*n+jj)*n+ii); vals.push_back(mkval<T>::get(v, 0.125 * v + 0.01 * ii)); } };
        add(k-1,j,i,-1.5); add(k,j-1,i,-1.25); add(k,j,i-1,-1.125); add(k,j,i,6.0 + 0.01 * i); add(k,j,i+1,-0.875); add(k,j+1,i,-0.75); add(k+1,j,i,-0.5);
        pntr.push_back((int) indx.size());
    }
    std::vector<T> x(N);
    for(int i=0; i<N; i++) x[i] = mkval<T>::get(std::sin(0.37 * i) + 0.25, std::cos(0.11 * i));
    hala::gpu_engine egpu(0);
    auto gp = egpu.load(pntr); auto gi = egpu.load(indx); auto gv = egpu.load(vals); auto gx = egpu.load(x);
    auto matrix = hala::make_sparse_matrix(egpu, N, gp, gi, gv);
    auto check = [&](char trans){
        std::vector<T> yref, y;
        hala::sparse_gemv(trans, N, N, 2.0, pntr, indx, vals, x, 0.0, yref);
        hala::gpu_vector<T> gy(egpu.device());
        matrix.gemv(trans, 2.0, gx, 0.0, gy);
        gy.unload(y);
        double scale = 0.0;                                                      // testvec sums |error| over the entries: relative to sum |y|
        for(auto const &v : yref) scale += std::abs(v);
        hassert(testvec(y, yref, scale));
    };
    for(int mode : {HB_TRANS_CHECKED, HB_TRANS_FROZEN, HB_TRANS_SCATTER, HB_TRANS_CHECKED}){
        matrix.set_transpose_mode(mode);
        check('T'); check('C'); check('N');
        for(auto &v : vals) v *= hala::get_cast<T>(1.03125);                      // the caller rewrites the values in place
        gv.load(vals);
        if (mode == HB_TRANS_FROZEN) matrix.values_changed();
        check('C'); check('T');
    }
}

int main(int argc, char**){
    verbose = (argc > 1);
    std::string name = "REFERENCE TESTS ON THE B200 BACKEND";
    begin_report(name);
    if (hala::gpu_device_count() < 1){ cout << "no CUDA device" << endl; return 2; }

    std::vector<std::function<void(void)>> direct = {
        []()->void{ load_unload<float>(); load_unload<double>(); load_unload<std::complex<float>>(); load_unload<std::complex<double>>(); },
        []()->void{ norm2<float>(); norm2<double>(); norm2<std::complex<float>>(); norm2<std::complex<double>>(); },
        []()->void{ dot<float>(); dot<double>(); dot<std::complex<float>>(); dot<std::complex<double>>(); },
        []()->void{ axpy<float>(); axpy<double>(); axpy<std::complex<float>>(); axpy<std::complex<double>>(); },
        []()->void{ scal<float>(); scal<double>(); scal<std::complex<float>>(); scal<std::complex<double>>(); },
        []()->void{ gemv<float>(); gemv<double>(); gemv<std::complex<float>>(); gemv<std::complex<double>>(); },
        []()->void{ sparse_matvec<float>(); sparse_matvec<double>(); sparse_matvec<std::complex<float>>(); sparse_matvec<std::complex<double>>(); },
        []()->void{ transposed_products<float>(); transposed_products<double>(); transposed_products<std::complex<float>>(); transposed_products<std::complex<double>>(); },
    };
    for(auto const &t : direct) perform(t);

    hala::gpu_engine ecuda(0);
    hala::mixed_engine emixed(ecuda);
    begin_report(std::string("shared test bodies, mixed_engine (engine API)"));
    std::vector<std::function<void(void)>> shared = {
        [&]()->void{ eng_api::test_vcopy<float, 0>(emixed); eng_api::test_vcopy<double, 0>(emixed); eng_api::test_vcopy<std::complex<float>, 0>(emixed); eng_api::test_vcopy<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_vswap<float, 0>(emixed); eng_api::test_vswap<double, 0>(emixed); eng_api::test_vswap<std::complex<float>, 0>(emixed); eng_api::test_vswap<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_iamax<float, 0>(emixed); eng_api::test_iamax<double, 0>(emixed); eng_api::test_iamax<std::complex<float>, 0>(emixed); eng_api::test_iamax<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_rotate<float, 0>(emixed); eng_api::test_rotate<double, 0>(emixed); },
        [&]()->void{ eng_api::test_axpy<float, 0>(emixed); eng_api::test_axpy<double, 0>(emixed); eng_api::test_axpy<std::complex<float>, 0>(emixed); eng_api::test_axpy<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_rscalar<float, 0>(emixed); eng_api::test_rscalar<double, 0>(emixed); eng_api::test_rscalar<std::complex<float>, 0>(emixed); eng_api::test_rscalar<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_scal<float, 0>(emixed); eng_api::test_scal<double, 0>(emixed); eng_api::test_scal<std::complex<float>, 0>(emixed); eng_api::test_scal<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_sparse_gemv<float, 0>(emixed); eng_api::test_sparse_gemv<double, 0>(emixed); eng_api::test_sparse_gemv<std::complex<float>, 0>(emixed); eng_api::test_sparse_gemv<std::complex<double>, 0>(emixed); },
    };
    for(auto const &t : shared) perform(t);

    begin_report(std::string("triangular solves and ILU-preconditioned solvers (reference tests, mixed_engine)"));
    std::vector<std::function<void(void)>> f1 = {
        [&]()->void{ eng_api::test_sparse_trsv<float, 0>(emixed); eng_api::test_sparse_trsv<double, 0>(emixed); eng_api::test_sparse_trsv<std::complex<float>, 0>(emixed); eng_api::test_sparse_trsv<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_sparse_trsm<float, 0>(emixed); eng_api::test_sparse_trsm<double, 0>(emixed); eng_api::test_sparse_trsm<std::complex<float>, 0>(emixed); eng_api::test_sparse_trsm<std::complex<double>, 0>(emixed); },
        [&]()->void{ test_cg<float>(emixed); test_cg<double>(emixed); test_cg<std::complex<float>>(emixed); test_cg<std::complex<double>>(emixed); },
        [&]()->void{ transpose<float>(); transpose<double>(); transpose<std::complex<float>>(); transpose<std::complex<double>>(); },
        [&]()->void{ test_geam<float>(); test_geam<double>(); test_geam<std::complex<float>>(); test_geam<std::complex<double>>(); },
        [&]()->void{ test_dgmm<float>(); test_dgmm<double>(); test_dgmm<std::complex<float>>(); test_dgmm<std::complex<double>>(); },
        [&]()->void{ eng_api::test_sparse_gemm<float, 0>(emixed); eng_api::test_sparse_gemm<double, 0>(emixed); eng_api::test_sparse_gemm<std::complex<float>, 0>(emixed); eng_api::test_sparse_gemm<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_batch_axpy<float, 0>(emixed); eng_api::test_batch_axpy<double, 0>(emixed); eng_api::test_batch_axpy<std::complex<float>, 0>(emixed); eng_api::test_batch_axpy<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_batch_dot<float, 0>(emixed); eng_api::test_batch_dot<double, 0>(emixed); eng_api::test_batch_dot<std::complex<float>, 0>(emixed); eng_api::test_batch_dot<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_batch_max_norm<float, 0>(emixed); eng_api::test_batch_max_norm<double, 0>(emixed); eng_api::test_batch_max_norm<std::complex<float>, 0>(emixed); eng_api::test_batch_max_norm<std::complex<double>, 0>(emixed); },
        [&]()->void{ eng_api::test_batch_scale<float, 0>(emixed); eng_api::test_batch_scale<double, 0>(emixed); eng_api::test_batch_scale<std::complex<float>, 0>(emixed); eng_api::test_batch_scale<std::complex<double>, 0>(emixed); },
        [&]()->void{ test_cg_batch<float>(emixed); test_cg_batch<double>(emixed); test_cg_batch<std::complex<float>>(emixed); test_cg_batch<std::complex<double>>(emixed); },
        [&]()->void{ test_gmres<float>(emixed); test_gmres<double>(emixed); test_gmres<std::complex<float>>(emixed); test_gmres<std::complex<double>>(emixed); },
    };
    for(auto const &t : f1) perform(t);

    begin_report(std::string("reference solver templates on gpu_engine / mixed_engine"));
    perform([]()->void{ solvers_on_engines<float>(); solvers_on_engines<double>(); solvers_on_engines<std::complex<float>>(); solvers_on_engines<std::complex<double>>(); });

    end_report(name);
    return test_result();
}
