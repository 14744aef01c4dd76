#ifndef HALAB200_GPU_SOLVERS_HPP
#define HALAB200_GPU_SOLVERS_HPP
// The solver front doors of hex/solvers for gpu_engine, with the iteration on the device.
//
// The reference's solve_cg(engine, ...) (hex/solvers/hala_solvers_cg.hpp:232-246) binds its arguments to the engine and runs
// solve_cg_core (:92-156) through BLAS-1 calls: nine launches and three host round trips per iteration on a GPU.  The overloads
// below take `gpu_engine const&` where the reference's take `compute_engine const&`, so overload resolution prefers them for a
// gpu_engine (and for mixed_engine, whose overloads forward to engine.gpu(), :250-264, gmres:259-272) and existing calls
//     hala::solve_cg(engine, stop, pntr, indx, vals, precon, b, x);
//     hala::solve_cg_ilu(engine, stop, pntr, indx, vals, ilu, b, x);
//     hala::solve_gmres(engine, stop, restart, pntr, indx, vals, precon, b, x);     (and solve_gmres_ilu, which calls it)
// reach hb_pcg / hb_pgmres with no edit: SpMV fused with <p,Ap>, fused residual update + norm + stop test, the caller's
// preconditioner on the solver's own device arrays, <r,z>, fused solution + direction update — four launches of the library per
// iteration plus the preconditioner's, every scalar on the device, no host synchronisation that stalls the stream.
// This header is read before hex/solvers (wax/hala_lib_extensions.hpp:18 pulls in the gpu/ layer), so the overloads are in scope
// where the reference's own templates call solve_cg / solve_gmres unqualified; stop_criteria is only declared here.
#include "hala_gpu_ilu.hpp"
#include <exception>
#include <functional>

namespace hala{

template<typename precision> struct stop_criteria;          // defined in hex/solvers/hala_solvers_core.hpp:52-71
template<class cengine, class vec> struct engined_vector;   // defined in wax/hala_noengine_structs.hpp:153-216

//! extension: says "no preconditioner" in a way the gpu_engine solvers can see (a lambda that copies cannot be told from a real
//! preconditioner and costs the iteration a copy and a dot product).  Usable with every engine: it copies x to r.
struct identity_preconditioner{
    template<class VectorLikeX, class VectorLikeR> void operator()(VectorLikeX const &x, VectorLikeR &r) const{ vcopy(x, r); }
    template<typename T> void operator()(gpu_vector<T> const &x, gpu_vector<T> &r) const{ r = x; }
};

namespace b200_solvers{

//! Carries the caller's preconditioner across the C ABI: hb_precon_fn -> precon(gpu_vector const&, gpu_vector&) on views of the
//! solver's work arrays.  Exceptions are parked here and re-thrown by the front door once the C call has returned.
template<typename T, class Precon> struct precon_bridge{
    Precon &precon;
    gpu_engine const &engine;
    size_t n;
    std::exception_ptr error;

    static int call(void *self, const void *in_dev, void *out_dev){
        auto *me = static_cast<precon_bridge*>(self);
        try{
            const int device = me->engine.device();
            gpu_vector<T> const vin = gpu_vector<T>::view(device, const_cast<T*>(static_cast<T const*>(in_dev)), me->n);
            gpu_vector<T> vout = gpu_vector<T>::view(device, static_cast<T*>(out_dev), me->n);
            me->precon(vin, vout);
            if (vout.data() != static_cast<T*>(out_dev)){       // the preconditioner put its result into storage of its own
                if (vout.size() != me->n) throw std::runtime_error("hala::solve (gpu_engine): the preconditioner returned a vector of the wrong size");
                check_hb(hb_memcpy_async(me->engine, out_dev, vout.data(), me->n * sizeof(T), HB_D2D), "hala::solve (gpu_engine): preconditioner output");
                check_hb(hb_ctx_sync(me->engine), "hala::solve (gpu_engine): preconditioner output");      // vout is freed on return
            }
        }catch(...){
            me->error = std::current_exception();
            return 1;
        }
        return 0;
    }
};

template<class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeX, class VectorLikeB>
int prepare(gpu_engine const &engine, VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, VectorLikeB const &b, VectorLikeX &x){
    check_types(vals, x, b);
    check_types_int(pntr, indx);
    engine.check_gpu(pntr, indx, vals, b, x);
    static_assert(std::is_same<get_vdefault<gpu_engine, VectorLikeX>, gpu_vector<typename define_type<VectorLikeX>::value_type>>::value,
                  "the fused gpu_engine solvers hand gpu_vector views to the preconditioner: define_vdefault<gpu_engine, T> must stay gpu_vector<T>");
    const int num_rows = get_size_int(pntr) - 1;
    assert( num_rows > 0 );
    assert( check_size(b, num_rows) );
    if (get_size_int(x) < num_rows){        // as the reference: an x of the wrong size is resized and zeroed (:201-204)
        force_size(num_rows, x);
        set_zero(engine, (size_t) num_rows, x);
    }
    return num_rows;
}

//! Takes a preconditioner written for engine-bound vectors (preconditioner_noe, hala_solvers_core.hpp:139-142) to the gpu_vector
//! form of the fused solvers: the solver's arrays are moved, as views, into the owning engined_vector<gpu_engine, gpu_vector<T>>
//! objects the reference would hand over (decltype(new_vector(vals)), hala_solvers_cg.hpp:198).
template<typename T, class Precon> struct unbound_precon{
    Precon &precon;
    gpu_engine const &engine;
    void operator()(gpu_vector<T> const &in, gpu_vector<T> &out) const{
        engined_vector<gpu_engine, gpu_vector<T>> bin(engine), bout(engine);
        bin.vector() = gpu_vector<T>::view(engine.device(), const_cast<T*>(in.data()), in.size());
        bout.vector() = gpu_vector<T>::view(engine.device(), out.data(), out.size());
        precon(const_cast<engined_vector<gpu_engine, gpu_vector<T>> const&>(bin), bout);
        if (bout.vector().data() != out.data()) out = std::move(bout.vector());     // storage of its own: the bridge copies it back
    }
};
//! A preconditioner already typed as the reference's std::function (preconditioner_noe) matches the reference's no-engine solve_cg
//! exactly; partial ordering cannot rank the two templates then (gcc: ambiguous), so the overload below steps aside for it.
template<class F> struct is_std_function : std::false_type{};
template<class S> struct is_std_function<std::function<S>> : std::true_type{};

template<typename T, class Precon> unbound_precon<T, Precon> unbind(gpu_engine const &engine, Precon &precon){ return {precon, engine}; }
template<typename T> identity_preconditioner unbind(gpu_engine const&, identity_preconditioner&){ return {}; }

}

//! hala::solve_cg on gpu_engine: hb_pcg (identity_preconditioner: hb_cg).  Returns the number of operator applications.
template<class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeX, class VectorLikeB, class Precon>
int solve_cg(gpu_engine const &engine, stop_criteria<get_precision_type<VectorLikeV>> const &stop,
             VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, Precon precon, VectorLikeB const &b, VectorLikeX &x){
    const int num_rows = b200_solvers::prepare(engine, pntr, indx, vals, b, x);
    using T = typename define_type<VectorLikeX>::value_type;
    auto matrix = make_sparse_matrix(engine, num_rows, pntr, indx, vals);
    int iterations = 0;
    double residual = 0;
    if (std::is_same<Precon, identity_preconditioner>::value){
        check_hb(hb_cg(engine, matrix.csr(), get_data(b), get_data(x), (double) stop.tol, stop.max_iter, &iterations, &residual), "hala::solve_cg(gpu_engine)");
        return iterations;
    }
    b200_solvers::precon_bridge<T, Precon> bridge{precon, engine, (size_t) num_rows, nullptr};
    const int status = hb_pcg(engine, matrix.csr(), get_data(b), get_data(x), (double) stop.tol, stop.max_iter,
                              &b200_solvers::precon_bridge<T, Precon>::call, &bridge, &iterations, &residual);
    if (bridge.error) std::rethrow_exception(bridge.error);
    check_hb(status, "hala::solve_cg(gpu_engine)");
    return iterations;
}

//! hala::solve_cg(stop, ...) without an engine argument, on vectors bound to a gpu_engine (bind_engine_vector): the reference runs
//! solve_cg_core on them (hala_solvers_cg.hpp:181-226); here they are unbound and take the fused iteration above.  The no-engine
//! solve_cg_ilu (:295-305) calls hala::solve_cg(stop, ...) and lands here too; the no-engine solve_gmres already forwards to the
//! engine form (hala_solvers_gmres.hpp:305-320), i.e. to the overload below.  A preconditioner passed as a std::function object
//! (not a lambda or functor) keeps the reference's loop: see b200_solvers::is_std_function.
template<class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeX, class VectorLikeB, class Precon>
std::enable_if_t<!b200_solvers::is_std_function<Precon>::value, int>
solve_cg(stop_criteria<get_precision_type<engined_vector<gpu_engine, VectorLikeV>>> const &stop,
             engined_vector<gpu_engine, VectorLikeP> const &pntr, engined_vector<gpu_engine, VectorLikeI> const &indx,
             engined_vector<gpu_engine, VectorLikeV> const &vals, Precon precon,
             engined_vector<gpu_engine, VectorLikeB> const &b, engined_vector<gpu_engine, VectorLikeX> &x){
    assert( check_engines(pntr, indx, vals, b, x) );
    using T = typename engined_vector<gpu_engine, VectorLikeX>::value_type;
    gpu_engine const &engine = x.engine();
    return solve_cg(engine, stop, pntr.vector(), indx.vector(), vals.vector(), b200_solvers::unbind<T>(engine, precon), b.vector(), x.vector());
}

//! hala::solve_cg_ilu on gpu_engine with a ready factorisation: the ILU application runs between the fused kernels.
template<class VectorLikeP, class VectorLikeI, class VectorLikeV, class ILUclass, class VectorLikeX, class VectorLikeB>
int solve_cg_ilu(gpu_engine const &engine, stop_criteria<get_precision_type<VectorLikeV>> const &stop,
                 VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, ILUclass const &ilu, VectorLikeB const &b, VectorLikeX &x){
    static_assert(std::is_same<typename ILUclass::engine_type, gpu_engine>::value, "Using compatible compute engine and ILU preconditioner");
    return solve_cg(engine, stop, pntr, indx, vals, [&](auto const &inx, auto &outr)->void{ ilu.apply(inx, outr, 1); }, b, x);
}

//! hala::solve_gmres on gpu_engine: hb_pgmres (identity_preconditioner: hb_gmres).  Complex data is projected with the conjugate
//! transpose (the reference's 'T' at hala_solvers_gmres.hpp:48,69 breaks down on genuinely complex matrices).
template<class VectorLikeP, class VectorLikeI, class VectorLikeV, class VectorLikeX, class VectorLikeB, class Precon>
int solve_gmres(gpu_engine const &engine, stop_criteria<get_precision_type<VectorLikeV>> const &stop, int restart,
                VectorLikeP const &pntr, VectorLikeI const &indx, VectorLikeV const &vals, Precon precon, VectorLikeB const &b, VectorLikeX &x){
    const int num_rows = b200_solvers::prepare(engine, pntr, indx, vals, b, x);
    using T = typename define_type<VectorLikeX>::value_type;
    auto matrix = make_sparse_matrix(engine, num_rows, pntr, indx, vals);
    const int cproj = is_complex<T>::value ? 1 : 0;
    int iterations = 0;
    double residual = 0;
    if (std::is_same<Precon, identity_preconditioner>::value){
        check_hb(hb_gmres(engine, matrix.csr(), get_data(b), get_data(x), (double) stop.tol, stop.max_iter, restart, cproj, &iterations, &residual),
                 "hala::solve_gmres(gpu_engine)");
        return iterations;
    }
    b200_solvers::precon_bridge<T, Precon> bridge{precon, engine, (size_t) num_rows, nullptr};
    const int status = hb_pgmres(engine, matrix.csr(), get_data(b), get_data(x), (double) stop.tol, stop.max_iter, restart, cproj,
                                 &b200_solvers::precon_bridge<T, Precon>::call, &bridge, &iterations, &residual);
    if (bridge.error) std::rethrow_exception(bridge.error);
    check_hb(status, "hala::solve_gmres(gpu_engine)");
    return iterations;
}

}
#endif
