// it/s of the fused CG iteration THROUGH the hala:: template API (the call an existing HALA program makes), next to the raw C entry
// point, on the BASELINE matrices: configs[0] (2-D 5-point Laplacian 1024^2) and configs[2] (3-D 7-point Laplacian n^3, n = 512 by
// default).  Built in the build container against the reference's headers (tests/cpp/Makefile), run on the GPU box by bench.py and
// tests/test_cpp_dropin.py.  Prints one JSON line per matrix.
//   template_bench [n3d=512] [iterations=50]
#include "hala.hpp"
#include "hala_solvers.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>

static void lap(int dims, int n, std::vector<int> &pntr, std::vector<int> &indx, std::vector<double> &vals){
    const long long N = dims == 2 ? (long long) n * n : (long long) n * n * n;
    pntr.assign(1, 0); pntr.reserve(N + 1);
    indx.clear(); vals.clear();
    indx.reserve((size_t) N * (dims == 2 ? 5 : 7)); vals.reserve((size_t) N * (dims == 2 ? 5 : 7));
    const int nz = dims == 2 ? 1 : n;
    for(int k=0; k<nz; k++) for(int j=0; j<n; j++) for(int i=0; i<n; i++){
        const long long row = ((long long) k * n + j) * n + i;
        auto add = [&](bool ok, long long col, double v){ if (ok){ indx.push_back((int) col); vals.push_back(v); } };
        add(k > 0, row - (long long) n * n, -1.0);
        add(j > 0, row - n, -1.0);
        add(i > 0, row - 1, -1.0);
        add(true, row, dims == 2 ? 4.0 : 6.0);
        add(i < n - 1, row + 1, -1.0);
        add(j < n - 1, row + n, -1.0);
        add(k < nz - 1, row + (long long) n * n, -1.0);
        pntr.push_back((int) indx.size());
    }
}

template<class F> static double seconds(hala::gpu_engine const &e, F f){
    e.synchronize();
    auto t0 = std::chrono::steady_clock::now();
    f();
    e.synchronize();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

static void run(const char *name, int dims, int n, int K, bool with_reference_loop){
    std::vector<int> pntr, indx; std::vector<double> vals;
    lap(dims, n, pntr, indx, vals);
    const size_t N = pntr.size() - 1;
    hala::gpu_engine engine(0);
    auto gp = engine.load(pntr); auto gi = engine.load(indx); auto gv = engine.load(vals);
    const size_t nnz = indx.size();
    { std::vector<int>().swap(pntr); std::vector<int>().swap(indx); std::vector<double>().swap(vals); }
    auto gb = engine.vector(N, 1.0 / std::sqrt((double) N));
    hala::gpu_vector<double> gx(engine.device());
    const hala::stop_criteria<double> stop(0.0, K + 1);            // tolerance 0: exactly K iterations (K + 1 operator applications)
    auto matrix = hala::make_sparse_matrix(engine, (int) N, gp, gi, gv);
    auto zero = [&](){ gx.resize(N); hala::set_zero(engine, N, gx); };
    auto launches = [&]()->long long{ long long c = 0; hb_ctx_launch_count(engine, &c); return c; };

    int it = 0; double res = 0;
    auto raw = [&](){ zero(); hala::check_hb(hb_cg(engine, matrix.csr(), gb.data(), gx.data(), 0.0, K + 1, &it, &res), "hb_cg"); };
    auto tag = [&](){ zero(); it = hala::solve_cg(engine, stop, gp, gi, gv, hala::identity_preconditioner(), gb, gx); };
    auto cpy = [&](){ zero(); it = hala::solve_cg(engine, stop, gp, gi, gv, [&](auto const &in, auto &out)->void{ hala::vcopy(engine, in, out); }, gb, gx); };
    // the no-engine form on engine-bound vectors with a copy lambda: unbound by hala_gpu_solvers.hpp, the fused iteration again
    auto noe = [&](){
        zero();
        auto p = hala::bind_engine_vector(engine, gp); auto i = hala::bind_engine_vector(engine, gi); auto v = hala::bind_engine_vector(engine, gv);
        auto bb = hala::bind_engine_vector(engine, gb); auto xx = hala::bind_engine_vector(engine, gx);
        it = hala::solve_cg(stop, p, i, v, [&](auto const &in, auto &out)->void{ hala::vcopy(in, out); }, bb, xx);
    };
    // the reference's own loop (solve_cg_core through BLAS-1 calls on this backend): what the template call ran before the overloads.
    // A preconditioner typed as the reference's std::function is the one call the overloads leave to it.
    auto ref = [&](){
        zero();
        auto p = hala::bind_engine_vector(engine, gp); auto i = hala::bind_engine_vector(engine, gi); auto v = hala::bind_engine_vector(engine, gv);
        auto bb = hala::bind_engine_vector(engine, gb); auto xx = hala::bind_engine_vector(engine, gx);
        hala::preconditioner_noe<decltype(v)> typed = [&](auto const &in, auto &out)->void{ hala::vcopy(in, out); };
        it = hala::solve_cg(stop, p, i, v, typed, bb, xx);
    };
    raw(); tag(); cpy(); noe();
    long long l0 = launches();
    const double t_raw = seconds(engine, raw); const long long l_raw = launches() - l0; l0 = launches();
    const double t_tag = seconds(engine, tag); const long long l_tag = launches() - l0; l0 = launches();
    const double t_cpy = seconds(engine, cpy); const long long l_cpy = launches() - l0; l0 = launches();
    const double t_noe = seconds(engine, noe); const long long l_noe = launches() - l0; l0 = launches();
    double t_ref = 0; long long l_ref = 0;
    if (with_reference_loop){ ref(); l0 = launches(); t_ref = seconds(engine, ref); l_ref = launches() - l0; }
    std::printf("{\"workload\": \"%s\", \"rows\": %zu, \"nnz\": %zu, \"iterations\": %d, "
                "\"hb_cg_its\": %.2f, \"template_identity_tag_its\": %.2f, \"template_copy_lambda_its\": %.2f, \"template_noengine_copy_lambda_its\": %.2f, "
                "\"reference_loop_on_our_blas1_its\": %.2f, "
                "\"launches_per_it\": {\"hb_cg\": %.2f, \"identity_tag\": %.2f, \"copy_lambda\": %.2f, \"noengine_copy_lambda\": %.2f, \"reference_loop\": %.2f}, "
                "\"tag_vs_hb_cg\": %.3f, \"copy_lambda_vs_hb_cg\": %.3f}\n",
                name, N, nnz, K, K / t_raw, K / t_tag, K / t_cpy, K / t_noe, with_reference_loop ? K / t_ref : 0.0,
                (double) l_raw / K, (double) l_tag / K, (double) l_cpy / K, (double) l_noe / K, (double) l_ref / K, t_raw / t_tag, t_raw / t_cpy);
    std::fflush(stdout);
}

int main(int argc, char **argv){
    const int n3 = argc > 1 ? std::atoi(argv[1]) : 512;
    const int K = argc > 2 ? std::atoi(argv[2]) : 50;
    if (hala::gpu_device_count() < 1){ std::printf("{\"error\": \"no CUDA device\"}\n"); return 2; }
    run("lap2d5-1024 fp64 CG (BASELINE configs[0])", 2, 1024, 10 * K, true);
    run((std::string("lap3d7-") + std::to_string(n3) + " fp64 CG (BASELINE configs[2])").c_str(), 3, n3, K, true);
    return 0;
}
