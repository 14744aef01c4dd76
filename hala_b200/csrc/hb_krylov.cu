// hb_krylov.cu — fused vector kernels of the CG / GMRES iterations and the tall-skinny gemv pair.
// Replaces, per CG iteration, cublas?axpy x2 + cublas?nrm2 + cublas?copy + cublas?dot + cublas?scal + cublas?axpy
// (reference hex/solvers/hala_solvers_cg.hpp:135-150 through gpu/hala_gpu_blas1.hpp) by two streaming passes, and per
// GMRES inner iteration cublas?gemv('T') + cublas?gemv('N') + cublas?nrm2 (hala_solvers_gmres.hpp:67-72,192 through
// gpu/hala_gpu_blas2.hpp:39-62) by one multi-dot pass and one multi-axpy+norm pass over the Krylov basis.
// Every scalar these kernels consume or produce lives in device memory; nothing here synchronises with the host.
#include "hb_common.cuh"
#include "hb_gs_pipe.cuh"
#include <cstdlib>

static constexpr int KR_THREADS = 256;

// ------------------------------------------------------------------------------------------------ CG state (device)
// zr[2] is double-buffered by iteration parity so that no kernel both reads and writes the same scalar.
template<typename T> struct cg_state {
    T zr[2];            // <r, z> of the current / next iteration (z == r: identity preconditioner)
    T pAp;              // <p, A p>
    double rnorm;       // ||r||_2 after the last update
    double tol;
    int iterations;     // operator applications so far (reference counter, starts at 1); frozen once done
    int max_iter;
    int done;           // set by the update kernel when the reference's stop test fires
    int pad;
};
struct cg_host_status { volatile int done; volatile int iterations; volatile double rnorm; };

// Schedule of one iteration (10 vector passes instead of the 11 of the textbook split, 18 of the reference's BLAS-1 calls):
//   spmv+dot   reads p, writes Ap                                  <p,Ap>
//   update     r -= a Ap ; rr = <r,r>          reads r, Ap, writes r      (a = <r,z>/<p,Ap>, read from the device state)
//   direction  x += a p ; p = r + b p          reads x, p, r, writes x, p (b = rr/<r,z>)
// The x update rides in the direction kernel because that kernel streams p anyway: x leaves the update kernel, one pass saved.
// Same arithmetic per element as solve_cg_core (hala_solvers_cg.hpp:135-148), so results are bit-identical to the 11-pass split.

// r -= a q ; rr = <r, r> ; then (last block) bookkeeping of solve_cg_core: iterations, stop test, zr_next
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS, 4) cg_update_kernel(int n, cg_state<T> *st, int parity, const T * __restrict__ q, T *r,
                                                               void *partials_v, unsigned int *ticket, cg_host_status *host, int precond){
    __shared__ double red[32];
    if (st->done) return;
    const T na = hneg(hdiv(st->zr[parity], st->pAp));
    double acc = 0.0;
    constexpr int U = 4;
    vec16<T> vr[U], vq[U];
    vec16<T> *r4 = reinterpret_cast<vec16<T>*>(r);
    const vec16<T> *q4 = reinterpret_cast<const vec16<T>*>(q);
    stream_sweep<T, VEC, U>((size_t) n,
        [&](int u, size_t i){ vr[u] = r4[i]; vq[u] = q4[i]; },
        [&](int u, size_t i){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++){ vr[u].v[k] = hfma(na, vq[u].v[k], vr[u].v[k]); acc += (double) habs2(vr[u].v[k]); }
            r4[i] = vr[u];
        },
        [&](size_t j){ T ri = hfma(na, q[j], r[j]); r[j] = ri; acc += (double) habs2(ri); });
    double *partials = reinterpret_cast<double*>(partials_v);
    double b = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_arrives(ticket)){
        double rr = sum_partials<double>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0){
            if (!precond) st->zr[parity ^ 1] = from_real<T>((real_t<T>) rr);    // preconditioned run: <r,z> comes from pcg_dot_kernel
            const double nrm = sqrt(rr);
            st->rnorm = nrm;
            const int it = st->iterations + 1;  // iterations++ of solve_cg_core: one more operator application
            st->iterations = it;
            // reference test: (it == max_iter) || (nrm < tol).  >= is identical for max_iter >= 2 and also terminates for
            // max_iter < 2 (where the reference would spin); a NaN residual stops as well instead of iterating forever.
            const int stop = (it >= st->max_iter) || (nrm < st->tol) || !(nrm == nrm);
            if (stop) st->done = 1;
            if (host){ host->rnorm = nrm; host->iterations = it; __threadfence_system(); if (stop) host->done = 1; }
        }
    }
}
// x += a p ; p = r + (zr_next / zr) p.   it_now = the operator-application count the update kernel of THIS iteration produced:
// when that update raised the stop flag the x update still has to happen (p is left alone); iterations enqueued after the
// stop (st->iterations frozen at an earlier count) do nothing.
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS, 4) cg_direction_kernel(int n, const cg_state<T> *st, int parity, int it_now, const T * __restrict__ r, T *p, T *x){
    const bool stopped = st->done != 0;
    if (stopped && st->iterations != it_now) return;
    const T a = hdiv(st->zr[parity], st->pAp);
    const T beta = hdiv(st->zr[parity ^ 1], st->zr[parity]);
    constexpr int U = 2;
    vec16<T> vx[U], vp[U], vr[U];
    vec16<T> *x4 = reinterpret_cast<vec16<T>*>(x), *p4 = reinterpret_cast<vec16<T>*>(p);
    const vec16<T> *r4 = reinterpret_cast<const vec16<T>*>(r);
    if (!stopped)
        stream_sweep<T, VEC, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vp[u] = p4[i]; vr[u] = r4[i]; },
            [&](int u, size_t i){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++){ vx[u].v[k] = hfma(a, vp[u].v[k], vx[u].v[k]); vp[u].v[k] = hfma(beta, vp[u].v[k], vr[u].v[k]); }
                x4[i] = vx[u]; p4[i] = vp[u];
            },
            [&](size_t j){ const T pj = p[j]; x[j] = hfma(a, pj, x[j]); p[j] = hfma(beta, pj, r[j]); });
    else
        stream_sweep<T, VEC, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vp[u] = p4[i]; },
            [&](int u, size_t i){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++) vx[u].v[k] = hfma(a, vp[u].v[k], vx[u].v[k]);
                x4[i] = vx[u];
            },
            [&](size_t j){ x[j] = hfma(a, p[j], x[j]); });
}

// Preconditioned CG (hb_pcg): zr[slot] = <r, z> (conjugated on r, as hala::dot) once the caller's preconditioner has produced z;
// with p_out != null (start of the solve) also p = z.  Skipped once the stop flag is up.
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS, 4) pcg_dot_kernel(int n, cg_state<T> *st, int slot, const T * __restrict__ r, const T * __restrict__ z, T *p_out,
                                                             void *partials_v, unsigned int *ticket){
    __shared__ T red[32];
    if (st->done) return;
    T acc = zero_of<T>();
    constexpr int U = 4;
    vec16<T> vr[U], vz[U];
    const vec16<T> *r4 = reinterpret_cast<const vec16<T>*>(r), *z4 = reinterpret_cast<const vec16<T>*>(z);
    vec16<T> *p4 = reinterpret_cast<vec16<T>*>(p_out);
    stream_sweep<T, VEC, U>((size_t) n,
        [&](int u, size_t i){ vr[u] = r4[i]; vz[u] = z4[i]; },
        [&](int u, size_t i){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++) acc = hfma(hconj(vr[u].v[k]), vz[u].v[k], acc);
            if (p_out) p4[i] = vz[u];
        },
        [&](size_t j){ const T zj = z[j]; acc = hfma(hconj(r[j]), zj, acc); if (p_out) p_out[j] = zj; });
    T *partials = reinterpret_cast<T*>(partials_v);
    T b = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_arrives(ticket)){
        T total = sum_partials<T>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0) st->zr[slot] = total;
    }
}

// setup: r = b - q (q = A x0), p = r, zr[0] = <r,r>; state initialised by the last block
template<typename T>
__global__ void __launch_bounds__(KR_THREADS) cg_setup_kernel(int n, cg_state<T> *st, double tol, int max_iter, const T * __restrict__ b,
                                                              const T * __restrict__ q, T *r, T *p, void *partials_v, unsigned int *ticket,
                                                              cg_host_status *host){
    __shared__ double red[32];
    double acc = 0.0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < (size_t) n; i += stride){
        T ri = hsub(b[i], q[i]);
        r[i] = ri; p[i] = ri;
        acc += (double) habs2(ri);
    }
    double *partials = reinterpret_cast<double*>(partials_v);
    double bs = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
    if (last_block_arrives(ticket)){
        double rr = sum_partials<double>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0){
            st->zr[0] = from_real<T>((real_t<T>) rr); st->zr[1] = zero_of<T>(); st->pAp = one_of<T>();
            st->rnorm = sqrt(rr); st->tol = tol; st->iterations = 1; st->max_iter = max_iter; st->done = 0; st->pad = 0;
            if (host){ host->rnorm = sqrt(rr); host->iterations = 1; host->done = 0; __threadfence_system(); }
        }
    }
}

// ------------------------------------------------------------------------------------------------ generic fused pieces (C ABI)
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS) axpy2_nrm2_kernel(int n, const T *a_dev, const T * __restrict__ p, const T * __restrict__ q,
                                                                T *x, T *r, void *partials_v, unsigned int *ticket, T *rr_dev){
    __shared__ double red[32];
    const T a = *a_dev, na = hneg(a);
    double acc = 0.0;
    constexpr int U = 2;
    vec16<T> vx[U], vp[U], vr[U], vq[U];
    vec16<T> *x4 = reinterpret_cast<vec16<T>*>(x), *r4 = reinterpret_cast<vec16<T>*>(r);
    const vec16<T> *p4 = reinterpret_cast<const vec16<T>*>(p), *q4 = reinterpret_cast<const vec16<T>*>(q);
    stream_sweep<T, VEC, U>((size_t) n,
        [&](int u, size_t i){ vx[u] = x4[i]; vp[u] = p4[i]; vr[u] = r4[i]; vq[u] = q4[i]; },
        [&](int u, size_t i){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++){
                vx[u].v[k] = hfma(a, vp[u].v[k], vx[u].v[k]);
                vr[u].v[k] = hfma(na, vq[u].v[k], vr[u].v[k]);
                acc += (double) habs2(vr[u].v[k]);
            }
            x4[i] = vx[u]; r4[i] = vr[u];
        },
        [&](size_t j){ x[j] = hfma(a, p[j], x[j]); T ri = hfma(na, q[j], r[j]); r[j] = ri; acc += (double) habs2(ri); });
    double *partials = reinterpret_cast<double*>(partials_v);
    double b = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_arrives(ticket)){
        double rr = sum_partials<double>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0) *rr_dev = from_real<T>((real_t<T>) rr);
    }
}
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS) xpby_kernel(int n, const T * __restrict__ r, const T *b_dev, T *p){
    const T beta = *b_dev;
    constexpr int U = 4;
    vec16<T> vp[U], vr[U];
    vec16<T> *p4 = reinterpret_cast<vec16<T>*>(p);
    const vec16<T> *r4 = reinterpret_cast<const vec16<T>*>(r);
    stream_sweep<T, VEC, U>((size_t) n,
        [&](int u, size_t i){ vp[u] = p4[i]; vr[u] = r4[i]; },
        [&](int u, size_t i){
            #pragma unroll
            for (int k = 0; k < vec16<T>::N; k++) vp[u].v[k] = hfma(beta, vp[u].v[k], vr[u].v[k]);
            p4[i] = vp[u];
        },
        [&](size_t j){ p[j] = hfma(beta, p[j], r[j]); });
}

// ------------------------------------------------------------------------------------------------ GMRES: multi-dot
// h[c] = sum_i op(W[i,c]) r[i] for c < k.  Columns are taken MD_KC at a time: each thread keeps MD_KC running sums in
// registers while it sweeps its rows (128-bit packets of r and of the MD_KC columns, all loads of a step issued before the
// first FMA), so a column costs one load and one FMA per element and the reductions happen once per chunk, not once per
// row block.  W is read exactly once; r once per chunk of columns ((k + ceil(k/KC)) N s bytes in all; r mostly from L2).
// Block partials -> the last block sums them in fixed order (bit-reproducible).
static constexpr int MD_KMAX = 64;         // columns handled per launch (restart <= 64 in one launch; more -> several launches)
template<typename T> __host__ __device__ constexpr int md_kc(){ return sizeof(T) == 16 ? 8 : 16; }

template<typename T, bool CONJ, bool VEC>
__global__ void __launch_bounds__(KR_THREADS, 2) multi_dot_kernel(long long rows, int k, const T * __restrict__ W, size_t ldw, const T * __restrict__ r,
                                                                  void *partials_v, unsigned int *ticket, T *h_out, const int *skip_flag){
    constexpr int KC = md_kc<T>();
    constexpr int NP = VEC ? vec16<T>::N : 1;                  // elements per packet
    __shared__ T wsum[KR_THREADS / 32][KC];
    if (skip_flag && *skip_flag) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T *partials = reinterpret_cast<T*>(partials_v);             // layout [block][k]
    const size_t stride = (size_t) gridDim.x * blockDim.x, gtid = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    const size_t npk = (size_t) rows / NP;
    for (int c0 = 0; c0 < k; c0 += KC){
        const int kc = min(KC, k - c0);
        T acc[KC];
        #pragma unroll
        for (int j = 0; j < KC; j++) acc[j] = zero_of<T>();
        const T *Wc = W + (size_t) c0 * ldw;
        for (size_t i = gtid; i < npk; i += stride){
            T rv[NP], wv[KC][NP];
            if (VEC){
                *reinterpret_cast<vec16<T>*>(rv) = reinterpret_cast<const vec16<T>*>(r)[i];
                #pragma unroll
                for (int j = 0; j < KC; j++)
                    if (j < kc) *reinterpret_cast<int4*>(wv[j]) = __ldcs(reinterpret_cast<const int4*>(Wc + (size_t) j * ldw) + i);
            }else{
                rv[0] = r[i];
                #pragma unroll
                for (int j = 0; j < KC; j++) if (j < kc) wv[j][0] = ld_stream(Wc + (size_t) j * ldw + i);
            }
            #pragma unroll
            for (int j = 0; j < KC; j++) if (j < kc){
                #pragma unroll
                for (int e = 0; e < NP; e++) acc[j] = hfma(CONJ ? hconj(wv[j][e]) : wv[j][e], rv[e], acc[j]);
            }
        }
        if (VEC){                                               // scalar tail behind the last whole packet
            for (size_t i = npk * NP + gtid; i < (size_t) rows; i += stride){
                const T ri = r[i];
                #pragma unroll
                for (int j = 0; j < KC; j++) if (j < kc){ const T w = Wc[(size_t) j * ldw + i]; acc[j] = hfma(CONJ ? hconj(w) : w, ri, acc[j]); }
            }
        }
        // one reduction for the whole chunk: shuffle every running sum down the warp, one shared slot per (warp, column),
        // a single barrier, then KC threads add the warps' values in fixed order
        #pragma unroll
        for (int j = 0; j < KC; j++) acc[j] = warp_sum(acc[j]);
        if (lane == 0){
            #pragma unroll
            for (int j = 0; j < KC; j++) wsum[warp][j] = acc[j];
        }
        __syncthreads();
        if (threadIdx.x < kc){
            T s = zero_of<T>();
            #pragma unroll
            for (int w = 0; w < KR_THREADS / 32; w++) s = hadd(s, wsum[w][threadIdx.x]);
            partials[(size_t) blockIdx.x * k + c0 + threadIdx.x] = s;
        }
        __syncthreads();
    }
    if (last_block_arrives(ticket)){
        // column c is summed by warp (c mod 8): lanes stride over the blocks' partials, shuffle reduction, fixed order
        for (int c = warp; c < k; c += KR_THREADS / 32){
            T a = zero_of<T>();
            for (int b = lane; b < (int) gridDim.x; b += 32) a = hadd(a, ld_cg_T(partials + (size_t) b * k + c));
            a = warp_sum(a);
            if (lane == 0) h_out[c] = a;
        }
    }
}

// r += scale * W h ; nrm2sq = sum |r_i|^2 of the updated r (same pass). h (k scalars) is read from device memory into shared.
// Two 128-bit packets of r per thread; the columns are taken eight at a time with all sixteen loads issued before the FMAs.
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS, 2) multi_axpy_nrm2_kernel(long long rows, int k, const T * __restrict__ W, size_t ldw, const T *h_dev,
                                                                        T *r, void *partials_v, unsigned int *ticket, T *nrm2sq_out, const int *skip_flag,
                                                                        T scale_h){
    constexpr int NP = VEC ? vec16<T>::N : 1, U = 2, CG = 8;
    __shared__ T h[MD_KMAX];
    __shared__ double red[32];
    if (skip_flag && *skip_flag) return;
    for (int c = threadIdx.x; c < k; c += blockDim.x) h[c] = hmul(scale_h, h_dev[c]);
    __syncthreads();
    double acc = 0.0;
    const size_t stride = (size_t) gridDim.x * blockDim.x, gtid = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    const size_t npk = (size_t) rows / NP;
    for (size_t i0 = gtid; i0 < npk; i0 += U * stride){
        const size_t i1 = i0 + stride;
        const bool two = i1 < npk;
        T t[U][NP];
        if (VEC){
            *reinterpret_cast<vec16<T>*>(t[0]) = reinterpret_cast<const vec16<T>*>(r)[i0];
            if (two) *reinterpret_cast<vec16<T>*>(t[1]) = reinterpret_cast<const vec16<T>*>(r)[i1];
        }else{
            t[0][0] = r[i0];
            if (two) t[1][0] = r[i1];
        }
        for (int c0 = 0; c0 < k; c0 += CG){
            T w[CG][U][NP];
            #pragma unroll
            for (int j = 0; j < CG; j++) if (c0 + j < k){
                const T *col = W + (size_t) (c0 + j) * ldw;
                if (VEC){
                    *reinterpret_cast<int4*>(w[j][0]) = __ldcs(reinterpret_cast<const int4*>(col) + i0);
                    if (two) *reinterpret_cast<int4*>(w[j][1]) = __ldcs(reinterpret_cast<const int4*>(col) + i1);
                }else{
                    w[j][0][0] = ld_stream(col + i0);
                    if (two) w[j][1][0] = ld_stream(col + i1);
                }
            }
            #pragma unroll
            for (int j = 0; j < CG; j++) if (c0 + j < k){
                const T hc = h[c0 + j];
                #pragma unroll
                for (int e = 0; e < NP; e++){
                    t[0][e] = hfma(w[j][0][e], hc, t[0][e]);
                    if (two) t[1][e] = hfma(w[j][1][e], hc, t[1][e]);
                }
            }
        }
        #pragma unroll
        for (int e = 0; e < NP; e++){ acc += (double) habs2(t[0][e]); if (two) acc += (double) habs2(t[1][e]); }
        if (VEC){
            reinterpret_cast<vec16<T>*>(r)[i0] = *reinterpret_cast<vec16<T>*>(t[0]);
            if (two) reinterpret_cast<vec16<T>*>(r)[i1] = *reinterpret_cast<vec16<T>*>(t[1]);
        }else{
            r[i0] = t[0][0];
            if (two) r[i1] = t[1][0];
        }
    }
    if (VEC){
        for (size_t i = npk * NP + gtid; i < (size_t) rows; i += stride){
            T ti = r[i];
            for (int c = 0; c < k; c++) ti = hfma(W[(size_t) c * ldw + i], h[c], ti);
            r[i] = ti; acc += (double) habs2(ti);
        }
    }
    if (nrm2sq_out){
        double *partials = reinterpret_cast<double*>(partials_v);
        double b = block_sum(acc, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            double rr = sum_partials<double>(partials, gridDim.x, 1, red);
            if (threadIdx.x == 0) *nrm2sq_out = from_real<T>((real_t<T>) rr);
        }
    }
}

// w_out = r / sqrt(real(*nrm2sq_dev))   (normalise + append to the basis in one pass; also rewrites r when r_out != null)
template<typename T, bool VEC>
__global__ void __launch_bounds__(KR_THREADS) scale_copy_kernel(long long rows, const T * __restrict__ r, const T *nrm2sq_dev, T *w_out, T *r_out){
    const real_t<T> inv = real_t<T>(1) / (real_t<T>) sqrt((double) hreal(*nrm2sq_dev));
    const T s = from_real<T>(inv);
    constexpr int NV = vec16<T>::N;
    vec16<T> pk[4];
    stream_sweep<T, VEC, 4>((size_t) rows,
        [&](int u, size_t p){ pk[u] = reinterpret_cast<const vec16<T>*>(r)[p]; },
        [&](int u, size_t p){
            #pragma unroll
            for (int k = 0; k < NV; k++) pk[u].v[k] = hmul(s, pk[u].v[k]);
            reinterpret_cast<vec16<T>*>(w_out)[p] = pk[u];
            if (r_out) reinterpret_cast<vec16<T>*>(r_out)[p] = pk[u];
        },
        [&](size_t i){
            const T v = hmul(s, r[i]);
            w_out[i] = v;
            if (r_out) r_out[i] = v;
        });
}

// ------------------------------------------------------------------------------------------------ general gemv (column-major)
// 'N': y = alpha A x + beta y  — one thread per row (coalesced over the column-major A), any strides
template<typename T>
__global__ void __launch_bounds__(KR_THREADS) gemv_n_kernel(int M, int N, scalar_arg<T> alpha_s, const T * __restrict__ A, size_t lda,
                                                            const T * __restrict__ x, long long incx, scalar_arg<T> beta_s, T *y, long long incy){
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < M; i += (long long) gridDim.x * blockDim.x){
        T sum = zero_of<T>();
        #pragma unroll 4
        for (int c = 0; c < N; c++) sum = hfma(A[(size_t) c * lda + i], x[c * incx], sum);
        T out = hmul(alpha, sum);
        if (use_beta) out = hfma(beta, y[i * incy], out);
        y[i * incy] = out;
    }
}
// 'T'/'C' finishing step after multi_dot produced raw sums in tmp: y[c] = alpha * tmp[c] + beta * y[c]
template<typename T>
__global__ void gemv_t_finish_kernel(int N, scalar_arg<T> alpha_s, const T *tmp, scalar_arg<T> beta_s, T *y, long long incy){
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x){
        T out = hmul(alpha, tmp[c]);
        if (use_beta) out = hfma(beta, y[c * incy], out);
        y[c * incy] = out;
    }
}
// 'T'/'C' with a strided x: one block per column (rare path: only the strided BLAS tests reach it)
template<typename T, bool CONJ>
__global__ void __launch_bounds__(KR_THREADS) gemv_t_strided_kernel(int M, int N, scalar_arg<T> alpha_s, const T * __restrict__ A, size_t lda,
                                                                    const T * __restrict__ x, long long incx, scalar_arg<T> beta_s, T *y, long long incy){
    __shared__ T red[32];
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    for (int c = blockIdx.x; c < N; c += gridDim.x){
        T sum = zero_of<T>();
        for (int i = threadIdx.x; i < M; i += blockDim.x){
            T a = A[(size_t) c * lda + i];
            sum = hfma(CONJ ? hconj(a) : a, x[i * incx], sum);
        }
        sum = block_sum(sum, red);
        if (threadIdx.x == 0){
            T out = hmul(alpha, sum);
            if (use_beta) out = hfma(beta, y[c * incy], out);
            y[c * incy] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------------ internal launchers (used by hb_solvers.cu)
static inline int kr_grid(const hb_ctx *ctx, long long n, int per_block){
    long long need = (n + per_block - 1) / per_block;
    long long cap = (long long) ctx->num_sms * 4;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

// streaming (TMA-staged) Gram-Schmidt kernels: one wave of 2 CTAs per SM
template<typename T, int MODE, bool CONJ>
static int launch_gs_pipe(hb_ctx *ctx, long long rows, int k, const T *W, size_t ldw, const T *r_in, T *r_out, const T *h_dev, T scale,
                          unsigned int *ticket, T *out, const int *skip){
    auto kern = gs_pipe_kernel<T, MODE, CONJ>;
    static per_device_flag configured;
    if (configured.first_time(ctx->device)){
        HB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gs_smem_bytes<T>()));
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    const long long per16 = 16 / sizeof(T), rows_al = rows - rows % per16;
    const long long ntiles = (rows_al + GS_R - 1) / GS_R;
    long long grid = (long long) ctx->num_sms * 2;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kern<<<(int) grid, GS_R, gs_smem_bytes<T>(), ctx->stream>>>(rows, rows_al, k, W, ldw, r_in, r_out, h_dev, scale, ctx->partials, ticket, out, skip);
    return HB_OK;
}
static bool gs_pipe_enabled(){
    static int on = -1;
    if (on < 0){ const char *e = getenv("HB_GS_PIPE"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

#define HB_MD_LAUNCH(CJ, VC) multi_dot_kernel<T, CJ, VC><<<grid, KR_THREADS, 0, ctx->stream>>>(rows, kk, Wc, ldw, (const T*) r, ctx->partials, ctx->tickets + 2, hc, skip)
int hb_multi_dot_internal(hb_ctx *ctx, int dtype, int conj, long long rows, int k, const void *W, size_t ldw, const void *r, void *h_dev, const int *skip){
    for (int c0 = 0; c0 < k; c0 += MD_KMAX){
        const int kk = (k - c0 < MD_KMAX) ? (k - c0) : MD_KMAX;
        HB_DISPATCH(dtype, {
            const T *Wc = (const T*) W + (size_t) c0 * ldw;
            T *hc = (T*) h_dev + c0;
            const bool vec = aligned16(Wc) && aligned16(r) && (ldw * sizeof(T)) % 16 == 0;
            // one wave of 2 CTAs per SM; a packet per thread per step
            int grid = hb_grid_for(ctx, (size_t) (rows > 0 ? rows : 1), KR_THREADS * (vec ? vec16<T>::N : 1), 2);
            const bool cj = conj && is_cplx<T>::value;
            if (vec && gs_pipe_enabled() && rows >= GS_R){
                int rc = cj ? launch_gs_pipe<T, 0, true>(ctx, rows, kk, Wc, ldw, (const T*) r, nullptr, nullptr, zero_of<T>(), ctx->tickets + 2, hc, skip)
                            : launch_gs_pipe<T, 0, false>(ctx, rows, kk, Wc, ldw, (const T*) r, nullptr, nullptr, zero_of<T>(), ctx->tickets + 2, hc, skip);
                if (rc != HB_OK) return rc;
            }
            else if (cj){ if (vec) HB_MD_LAUNCH(true, true); else HB_MD_LAUNCH(true, false); }
            else        { if (vec) HB_MD_LAUNCH(false, true); else HB_MD_LAUNCH(false, false); }
        });
        HB_LAUNCH_CHECK(ctx);
    }
    return HB_OK;
}

// r += scale * W h  (scale = -1 for the projection, +1 for krylov_combine); optional |r|^2 of the result
int hb_multi_axpy_internal(hb_ctx *ctx, int dtype, long long rows, int k, const void *W, size_t ldw, const void *h_dev, void *r,
                           void *nrm2sq_dev, double scale, const int *skip){
    for (int c0 = 0; c0 < k || c0 == 0; c0 += MD_KMAX){     // k == 0 still launches once: it delivers the norm
        const int kk = (k - c0 < MD_KMAX) ? (k - c0) : MD_KMAX;
        const bool last = (c0 + kk >= k);
        HB_DISPATCH(dtype, {
            const T *Wc = (const T*) W + (size_t) c0 * ldw;
            const bool vec = aligned16(r) && (kk == 0 || (aligned16(Wc) && (ldw * sizeof(T)) % 16 == 0));
            int grid = hb_grid_for(ctx, (size_t) (rows > 0 ? rows : 1), KR_THREADS * 2 * (vec ? vec16<T>::N : 1), 2);
            if (vec && kk > 0 && gs_pipe_enabled() && rows >= GS_R){
                int rc = launch_gs_pipe<T, 1, false>(ctx, rows, kk, Wc, ldw, (const T*) r, (T*) r, (const T*) h_dev + c0, from_real<T>((real_t<T>) scale),
                                                     ctx->tickets + 3, last ? (T*) nrm2sq_dev : nullptr, skip);
                if (rc != HB_OK) return rc;
            }
            else if (vec) multi_axpy_nrm2_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(rows, kk, Wc, ldw, (const T*) h_dev + c0,
                        (T*) r, ctx->partials, ctx->tickets + 3, last ? (T*) nrm2sq_dev : nullptr, skip, from_real<T>((real_t<T>) scale));
            else     multi_axpy_nrm2_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(rows, kk, Wc, ldw, (const T*) h_dev + c0,
                        (T*) r, ctx->partials, ctx->tickets + 3, last ? (T*) nrm2sq_dev : nullptr, skip, from_real<T>((real_t<T>) scale));
        });
        HB_LAUNCH_CHECK(ctx);
    }
    return HB_OK;
}

int hb_scale_copy_internal(hb_ctx *ctx, int dtype, long long rows, const void *r, const void *nrm2sq_dev, void *w_out, void *r_out){
    int grid = kr_grid(ctx, rows, KR_THREADS * 8);
    const bool vec = aligned16(r) && aligned16(w_out) && (!r_out || aligned16(r_out));
    HB_DISPATCH(dtype, {
        if (vec) scale_copy_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(rows, (const T*) r, (const T*) nrm2sq_dev, (T*) w_out, (T*) r_out);
        else     scale_copy_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(rows, (const T*) r, (const T*) nrm2sq_dev, (T*) w_out, (T*) r_out);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

// CG kernels, typed entry points
int hb_cg_setup_internal(hb_ctx *ctx, int dtype, int n, void *state, double tol, int max_iter, const void *b, const void *q, void *r, void *p, void *host){
    int grid = kr_grid(ctx, n, KR_THREADS * 4);
    HB_DISPATCH(dtype, (cg_setup_kernel<T><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (cg_state<T>*) state, tol, max_iter, (const T*) b, (const T*) q,
                        (T*) r, (T*) p, ctx->partials, ctx->tickets + 4, (cg_host_status*) host)));
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_pcg_dot_internal(hb_ctx *ctx, int dtype, int n, void *state, int slot, const void *r, const void *z, void *p_out){
    int grid = kr_grid(ctx, n, KR_THREADS * 8);
    const bool vec = aligned16(r) && aligned16(z) && aligned16(p_out);
    HB_DISPATCH(dtype, {
        if (vec) pcg_dot_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (cg_state<T>*) state, slot, (const T*) r, (const T*) z, (T*) p_out, ctx->partials, ctx->tickets + 4);
        else     pcg_dot_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (cg_state<T>*) state, slot, (const T*) r, (const T*) z, (T*) p_out, ctx->partials, ctx->tickets + 4);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_cg_update_internal(hb_ctx *ctx, int dtype, int n, void *state, int parity, const void *q, void *r, void *host, int precond){
    int grid = kr_grid(ctx, n, KR_THREADS * 8);
    const bool vec = aligned16(q) && aligned16(r);
    HB_DISPATCH(dtype, {
        if (vec) cg_update_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (cg_state<T>*) state, parity, (const T*) q, (T*) r,
                                                                                  ctx->partials, ctx->tickets + 4, (cg_host_status*) host, precond);
        else     cg_update_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (cg_state<T>*) state, parity, (const T*) q, (T*) r,
                                                                                   ctx->partials, ctx->tickets + 4, (cg_host_status*) host, precond);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_cg_direction_internal(hb_ctx *ctx, int dtype, int n, const void *state, int parity, int it_now, const void *r, void *p, void *x){
    int grid = kr_grid(ctx, n, KR_THREADS * 8);
    const bool vec = aligned16(r) && aligned16(p) && aligned16(x);
    HB_DISPATCH(dtype, {
        if (vec) cg_direction_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const cg_state<T>*) state, parity, it_now, (const T*) r, (T*) p, (T*) x);
        else     cg_direction_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const cg_state<T>*) state, parity, it_now, (const T*) r, (T*) p, (T*) x);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
size_t hb_cg_state_bytes(int dtype){
    switch (dtype){ case HB_F32: return sizeof(cg_state<float>); case HB_F64: return sizeof(cg_state<double>);
                    case HB_C32: return sizeof(cg_state<cplx<float>>); default: return sizeof(cg_state<cplx<double>>); }
}
size_t hb_cg_state_pap_offset(int dtype){
    switch (dtype){ case HB_F32: return offsetof(cg_state<float>, pAp); case HB_F64: return offsetof(cg_state<double>, pAp);
                    case HB_C32: return offsetof(cg_state<cplx<float>>, pAp); default: return offsetof(cg_state<cplx<double>>, pAp); }
}
size_t hb_cg_state_done_offset(int dtype){
    switch (dtype){ case HB_F32: return offsetof(cg_state<float>, done); case HB_F64: return offsetof(cg_state<double>, done);
                    case HB_C32: return offsetof(cg_state<cplx<float>>, done); default: return offsetof(cg_state<cplx<double>>, done); }
}

extern "C" {

int hb_multi_dot(hb_ctx *ctx, int dtype, int conj, int rows, int k, const void *W, size_t ldw, const void *r, void *h_dev){
    HB_ARG(ctx && h_dev, "null");
    HB_ARG(rows >= 0 && k >= 0, "negative size");
    if (k == 0) return HB_OK;
    HB_ARG(rows == 0 || (W && r), "null array");
    return hb_multi_dot_internal(ctx, dtype, conj, rows, k, W, ldw, r, h_dev, nullptr);
}
int hb_multi_axpy_nrm2(hb_ctx *ctx, int dtype, int rows, int k, const void *W, size_t ldw, const void *h_dev, void *r, void *nrm2sq_dev){
    HB_ARG(ctx, "null");
    HB_ARG(rows >= 0 && k >= 0, "negative size");
    HB_ARG(k == 0 || (W && h_dev), "null array");
    if (k == 0){
        // nothing to subtract: still deliver the norm
        if (nrm2sq_dev) return hb_multi_axpy_internal(ctx, dtype, rows, 0, r, 0, r, r, nrm2sq_dev, -1.0, nullptr);
        return HB_OK;
    }
    return hb_multi_axpy_internal(ctx, dtype, rows, k, W, ldw, h_dev, r, nrm2sq_dev, -1.0, nullptr);
}
int hb_axpy2_nrm2(hb_ctx *ctx, int dtype, int n, const void *a_dev, const void *p, const void *q, void *x, void *r, void *rr_dev){
    HB_ARG(ctx && a_dev && rr_dev, "null");
    HB_ARG(n >= 0, "negative size");
    int grid = kr_grid(ctx, n, KR_THREADS * 4);
    const bool vec = aligned16(p) && aligned16(q) && aligned16(x) && aligned16(r);
    HB_DISPATCH(dtype, {
        if (vec) axpy2_nrm2_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const T*) a_dev, (const T*) p, (const T*) q, (T*) x, (T*) r,
                                                                                   ctx->partials, ctx->tickets + 5, (T*) rr_dev);
        else     axpy2_nrm2_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const T*) a_dev, (const T*) p, (const T*) q, (T*) x, (T*) r,
                                                                                    ctx->partials, ctx->tickets + 5, (T*) rr_dev);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_xpby(hb_ctx *ctx, int dtype, int n, const void *r, const void *b_dev, void *p){
    HB_ARG(ctx && b_dev, "null");
    if (n <= 0) return HB_OK;
    int grid = kr_grid(ctx, n, KR_THREADS * 4);
    const bool vec = aligned16(r) && aligned16(p);
    HB_DISPATCH(dtype, {
        if (vec) xpby_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const T*) r, (const T*) b_dev, (T*) p);
        else     xpby_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(n, (const T*) r, (const T*) b_dev, (T*) p);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}

int hb_gemv(hb_ctx *ctx, int dtype, char trans, int M, int N, const void *alpha, const void *A, int lda,
            const void *x, int incx, const void *beta, void *y, int incy){
    HB_ARG(ctx && alpha && beta, "null");
    HB_ARG(M >= 0 && N >= 0 && lda >= (M > 1 ? M : 1), "bad dimensions");
    const int ny = hb_is_n(trans) ? M : N;
    if (ny == 0) return HB_OK;
    HB_DISPATCH(dtype, {
        scalar_arg<T> a = make_scalar<T>(ctx, alpha), b = make_scalar<T>(ctx, beta);
        if (hb_is_n(trans)){
            int grid = kr_grid(ctx, M, KR_THREADS);
            gemv_n_kernel<T><<<grid, KR_THREADS, 0, ctx->stream>>>(M, N, a, (const T*) A, (size_t) lda, (const T*) x, incx, b, (T*) y, incy);
            HB_LAUNCH_CHECK(ctx);
        }else{
            const bool cj = hb_is_c(trans) && is_cplx<T>::value;
            if (incx == 1 && N <= 1024 && M > 0){
                // tall-skinny: fused multi-dot over the columns, then the alpha/beta finish
                T *tmp = reinterpret_cast<T*>(reinterpret_cast<char*>(ctx->partials) + HB_PARTIAL_BYTES - 1024 * sizeof(T));
                int rc = hb_multi_dot_internal(ctx, dtype, cj ? 1 : 0, M, N, A, (size_t) lda, x, tmp, nullptr);
                if (rc != HB_OK) return rc;
                gemv_t_finish_kernel<T><<<(N + 127) / 128, 128, 0, ctx->stream>>>(N, a, tmp, b, (T*) y, incy);
                HB_LAUNCH_CHECK(ctx);
            }else{
                int grid = N < ctx->num_sms * 4 ? N : ctx->num_sms * 4;
                if (cj) gemv_t_strided_kernel<T, true><<<grid, KR_THREADS, 0, ctx->stream>>>(M, N, a, (const T*) A, (size_t) lda, (const T*) x, incx, b, (T*) y, incy);
                else    gemv_t_strided_kernel<T, false><<<grid, KR_THREADS, 0, ctx->stream>>>(M, N, a, (const T*) A, (size_t) lda, (const T*) x, incx, b, (T*) y, incy);
                HB_LAUNCH_CHECK(ctx);
            }
        }
    });
    return HB_OK;
}

}
