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


// The fused iteration behind the UNCHANGED template calls (hala_b200/gpu/hala_gpu_solvers.hpp): what a HALA user gets by swapping the
// include path.  Counts the library's kernel launches per iteration (hb_ctx_launch_count), compares iteration counts with the same
// call on cpu_engine, and pins the corner cases of the preconditioner bridge.
template<typename T> void fused_front_doors(){
    current_test<T> tests("fused solve_cg/gmres");
    using P = typename hala::define_standard_precision<T>::value_type;
    const int n = 40, N = n * n;
    const P tol = std::is_same<P, float>::value ? 1.E-4f : 1.E-9;
    std::vector<int> pntr, indx; std::vector<T> vals;
    lap2d<T>(n, pntr, indx, vals);
    std::vector<T> b(N, hala::get_cast<T>(1.0 / n));
    hala::cpu_engine ecpu;
    hala::gpu_engine egpu(0);
    auto gp = egpu.load(pntr); auto gi = egpu.load(indx); auto gv = egpu.load(vals); auto gb = egpu.load(b);
    auto launches = [&]()->long long{ long long c = 0; hb_ctx_launch_count(egpu, &c); return c; };
    const hala::stop_criteria<P> stop(tol, 1000);

    // 1. identity preconditioner written as the reference's tests write it (a copy): 4 library launches + the copy per iteration
    std::vector<T> xcpu, x1, x2, x3, x4;
    int it_cpu = hala::solve_cg(ecpu, stop, pntr, indx, vals, [&](auto const &in, auto &out)->void{ hala::vcopy(ecpu, in, out); }, b, xcpu);
    hala::gpu_vector<T> gx(egpu.device());
    long long l0 = launches();
    int it1 = hala::solve_cg(egpu, stop, gp, gi, gv, [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, gb, gx);
    long long per_it = (launches() - l0) / it1;
    gx.unload(x1);
    hassert(std::abs(it1 - it_cpu) <= 2);
    hassert(per_it <= 5);                                           // spmv+dot, update, [copy], dot, direction (+ a few skipped past the stop)
    hassert(testvec(x1, xcpu, 1.E+6 * hala::norm2(xcpu)));

    // 2. hala::identity_preconditioner: the unpreconditioned fused iteration, 3 launches
    hala::gpu_vector<T> gx2(egpu.device());
    l0 = launches();
    int it2 = hala::solve_cg(egpu, stop, gp, gi, gv, hala::identity_preconditioner(), gb, gx2);
    per_it = (launches() - l0) / it2;
    gx2.unload(x2);
    hassert(std::abs(it2 - it_cpu) <= 2);
    hassert(per_it <= 3);
    hassert(testvec(x2, xcpu, 1.E+6 * hala::norm2(xcpu)));

    // 3. a real preconditioner: Jacobi on a matrix with a varying diagonal (element-wise divide = tbsv with bandwidth 0 on the GPU)
    std::vector<T> dvals = vals, diag(N);
    for(int i=0; i<N; i++) for(int j=pntr[i]; j<pntr[i+1]; j++) if (indx[j] == i){ dvals[j] = hala::get_cast<T>(4.0 + (i % 7) * 0.5); diag[i] = dvals[j]; }
    auto gdv = egpu.load(dvals); auto gdiag = egpu.load(diag);
    std::vector<T> xjc;
    int jc = hala::solve_cg(ecpu, stop, pntr, indx, dvals,
                            [&](auto const &in, auto &out)->void{ hala::vcopy(ecpu, in, out); for(int i=0; i<N; i++) out[i] /= diag[i]; }, b, xjc);
    hala::gpu_vector<T> gx3(egpu.device());
    int jg = hala::solve_cg(egpu, stop, gp, gi, gdv,
                            [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); hala::tbsv(egpu, 'U', 'N', 'N', N, 0, gdiag, out); }, gb, gx3);
    gx3.unload(x3);
    hassert(std::abs(jc - jg) <= 2);
    hassert(jg < it_cpu + 10);
    hassert(testvec(x3, xjc, 1.E+6 * hala::norm2(xjc)));

    // 4. a preconditioner that hands back storage of its own (move-assignment into the output) and one that throws
    hala::gpu_vector<T> gx4(egpu.device());
    int it4 = hala::solve_cg(egpu, stop, gp, gi, gv, [&](auto const &in, auto &out)->void{ out = egpu.vcopy(in); }, gb, gx4);
    gx4.unload(x4);
    hassert(std::abs(it4 - it_cpu) <= 2);
    hassert(testvec(x4, xcpu, 1.E+6 * hala::norm2(xcpu)));
    bool thrown = false;
    try{
        hala::gpu_vector<T> gx5(egpu.device());
        hala::solve_cg(egpu, stop, gp, gi, gv, [&](auto const&, auto&)->void{ throw std::runtime_error("precon says no"); }, gb, gx5);
    }catch(std::runtime_error &e){ thrown = (std::string(e.what()) == "precon says no"); }
    hassert(thrown);

    // 5. ILU-preconditioned CG through solve_cg_ilu(gpu_engine, ...) with a ready factor and on the fly, against cpu_engine
    std::vector<T> xic, xig;
    int ic = hala::solve_cg_ilu(ecpu, stop, pntr, indx, vals, b, xic);
    auto ilu = hala::make_ilu(egpu, gp, gi, gv, 'N');
    hala::gpu_vector<T> gxi(egpu.device());
    int ig = hala::solve_cg_ilu(egpu, stop, gp, gi, gv, ilu, gb, gxi);
    gxi.unload(xig);
    hassert(std::abs(ic - ig) <= 2);
    hassert(testvec(xig, xic, 1.E+6 * hala::norm2(xic)));
    hala::gpu_vector<T> gxf(egpu.device());
    int ig2 = hala::solve_cg_ilu(egpu, stop, gp, gi, gv, gb, gxf);
    hassert(ig2 == ig);

    // 6. GMRES: copy-lambda, identity tag and ILU (solve_gmres_ilu is the reference's own template: it must land on the fused overload)
    for(int i=0; i<N; i++) for(int j=pntr[i]; j<pntr[i+1]; j++) if (indx[j] < i) vals[j] = hala::get_cast<T>(-1.5); else if (indx[j] > i) vals[j] = hala::get_cast<T>(-0.5);
    gv.load(vals);
    const hala::stop_criteria<P> gstop(tol, 100);
    std::vector<T> ycpu, y1, y2, yic, yig;
    int g_cpu = hala::solve_gmres(ecpu, gstop, 20, pntr, indx, vals, [&](auto const &in, auto &out)->void{ hala::vcopy(ecpu, in, out); }, b, ycpu);
    hala::gpu_vector<T> gy1(egpu.device()), gy2(egpu.device()), gy3(egpu.device());
    int g1 = hala::solve_gmres(egpu, gstop, 20, gp, gi, gv, [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, gb, gy1);
    int g2 = hala::solve_gmres(egpu, gstop, 20, gp, gi, gv, hala::identity_preconditioner(), gb, gy2);
    gy1.unload(y1); gy2.unload(y2);
    hassert(std::abs(g1 - g_cpu) <= 2);
    hassert(g2 == g1);
    hassert(testvec(y1, ycpu, 1.E+6 * hala::norm2(ycpu)));
    hassert(testvec(y2, ycpu, 1.E+6 * hala::norm2(ycpu)));
    int gic = hala::solve_gmres_ilu(ecpu, gstop, 20, pntr, indx, vals, b, yic);
    int gig = hala::solve_gmres_ilu(egpu, gstop, 20, gp, gi, gv, gb, gy3);
    gy3.unload(yig);
    hassert(std::abs(gic - gig) <= 2);
    hassert(testvec(yig, yic, 1.E+6 * hala::norm2(yic)));

    // 7. the no-engine forms on vectors bound to the engine (bind_engine_vector): same fused iteration, same launch counts; the
    //    preconditioner sees engined vectors; one typed as the reference's std::function (preconditioner_noe) must still compile and solve
    indx.clear(); vals.clear();
    lap2d<T>(n, pntr, indx, vals);
    gv.load(vals);
    hala::gpu_vector<T> gx7(egpu.device()), gx8(egpu.device()), gx9(egpu.device()), gx10(egpu.device());
    auto bp = hala::bind_engine_vector(egpu, gp); auto bi = hala::bind_engine_vector(egpu, gi); auto bv = hala::bind_engine_vector(egpu, gv);
    auto bb = hala::bind_engine_vector(egpu, gb);
    auto bx7 = hala::bind_engine_vector(egpu, gx7); auto bx8 = hala::bind_engine_vector(egpu, gx8);
    auto bx9 = hala::bind_engine_vector(egpu, gx9); auto bx10 = hala::bind_engine_vector(egpu, gx10);
    std::vector<T> x7, x8, x9, x10;
    l0 = launches();
    int it7 = hala::solve_cg(stop, bp, bi, bv, [&](auto const &in, auto &out)->void{ hala::vcopy(in, out); }, bb, bx7);
    per_it = (launches() - l0) / it7;
    gx7.unload(x7);
    hassert(it7 == it1);
    hassert(per_it <= 5);
    hassert(testvec(x7, xcpu, 1.E+6 * hala::norm2(xcpu)));
    l0 = launches();
    int it8 = hala::solve_cg(stop, bp, bi, bv, hala::identity_preconditioner(), bb, bx8);
    per_it = (launches() - l0) / it8;
    gx8.unload(x8);
    hassert(it8 == it2);
    hassert(per_it <= 3);
    hassert(testvec(x8, xcpu, 1.E+6 * hala::norm2(xcpu)));
    hala::preconditioner_noe<decltype(bv)> typed = [&](auto const &in, auto &out)->void{ hala::vcopy(in, out); hala::scal(2.0, out); };
    int it9 = hala::solve_cg(stop, bp, bi, bv, typed, bb, bx9);       // a scaled identity, as a std::function: the reference's loop
    gx9.unload(x9);
    hassert(std::abs(it9 - it_cpu) <= 2);
    hassert(testvec(x9, xcpu, 1.E+6 * hala::norm2(xcpu)));
    int it10 = hala::solve_cg_ilu(stop, bp, bi, bv, ilu, bb, bx10);    // the reference's template: calls hala::solve_cg(stop, ...)
    gx10.unload(x10);
    hassert(it10 == ig);
    hassert(testvec(x10, xic, 1.E+6 * hala::norm2(xic)));
}

// Row f3, staging half: page-locked containers through load / unload / gpu_bind_vector and a mixed_engine solve.
template<typename T> void pinned_containers(){
    current_test<T> tests("pinned load/unload");
    using P = typename hala::define_standard_precision<T>::value_type;
    const int n = 32, N = n * n;
    std::vector<int> pntr, indx; std::vector<T> vals;
    lap2d<T>(n, pntr, indx, vals);
    hala::pinned_vector<int> ppntr(pntr.begin(), pntr.end()), pindx(indx.begin(), indx.end());
    hala::pinned_vector<T> pvals(vals.begin(), vals.end()), pb(N, hala::get_cast<T>(1.0 / n)), px, pback;
    hala::gpu_engine egpu(0);
    hala::mixed_engine emix(egpu);
    auto gv = egpu.load(pvals);                      // load from pinned memory, unload into a pinned container
    gv.unload(pback);
    hassert(pback.size() == pvals.size());
    bool same = true;
    for(size_t i=0; i<pvals.size(); i++) same = same && (pback[i] == pvals[i]);
    hassert(same);
    {   // gpu_bind_vector on a pinned container: loaded on construction, written back on destruction
        hala::pinned_vector<T> y(N, hala::get_cast<T>(2.0));
        {
            auto by = hala::gpu_bind_vector(egpu, y);
            hala::scal(egpu, 0.5, by);
        }
        bool halved = true;
        for(auto const &v : y) halved = halved && (std::abs(v - hala::get_cast<T>(1.0)) < 1.E-6);
        hassert(halved);
    }
    std::vector<T> b(N, hala::get_cast<T>(1.0 / n)), xref;
    const P tol = std::is_same<P, float>::value ? 1.E-4f : 1.E-9;
    int it_ref = hala::solve_cg(emix, hala::stop_criteria<P>(tol, 1000), pntr, indx, vals,
                                [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, b, xref);
    int it_pin = hala::solve_cg(emix, hala::stop_criteria<P>(tol, 1000), ppntr, pindx, pvals,
                                [&](auto const &in, auto &out)->void{ hala::vcopy(egpu, in, out); }, pb, px);
    hassert(it_ref == it_pin);
    std::vector<T> xpin(px.begin(), px.end());
    hassert(testvec(xpin, xref, hala::norm2(xref)));
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
    perform([]()->void{ fused_front_doors<float>(); fused_front_doors<double>(); fused_front_doors<std::complex<float>>(); fused_front_doors<std::complex<double>>(); });
    perform([]()->void{ pinned_containers<float>(); pinned_containers<double>(); pinned_containers<std::complex<float>>(); pinned_containers<std::complex<double>>(); });

    end_report(name);
    return test_result();
}
