// hb_dist.cu — row-partitioned multi-GPU path: halo exchange, scalar all-reduce and the distributed CG iteration.
// One process per GPU; NCCL (over NVLink 5 / NVSwitch) is loaded at run time.  New work: the reference is single-device.
#include "hb_common.cuh"
#include "hb_peer.cuh"
#include "../../include/halab200_dist.h"
#include <nccl.h>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>

int hb_spmv_dot_internal(hb_ctx *ctx, const hb_csr *A, const void *x, void *y, void *dot_dev, const int *skip);
int hb_spmv_internal(hb_ctx *ctx, const hb_csr *A, const void *x, void *y, const int *skip);
int hb_spmv_variant(const hb_csr *A);

// ------------------------------------------------------------------------------------------------ NCCL, resolved lazily
namespace {
struct nccl_api {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char*  (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
nccl_api g_nccl;

int load_nccl(){
    if (g_nccl.lib) return HB_OK;
    // RTLD_NOLOAD first: reuse the libnccl a host framework (e.g. torch) already mapped, then fall back to the system one
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h){ hb_set_error(std::string("cannot load libnccl.so.2: ") + dlerror()); return HB_ERR_NCCL; }
    g_nccl.lib = h;
    #define HB_SYM(field, name) do { *(void**) (&g_nccl.field) = dlsym(h, name); if (!g_nccl.field){ hb_set_error(std::string("libnccl lacks ") + name); g_nccl.lib = nullptr; return HB_ERR_NCCL; } } while (0)
    HB_SYM(GetUniqueId, "ncclGetUniqueId"); HB_SYM(CommInitRank, "ncclCommInitRank"); HB_SYM(CommDestroy, "ncclCommDestroy");
    HB_SYM(GetErrorString, "ncclGetErrorString"); HB_SYM(AllReduce, "ncclAllReduce"); HB_SYM(AllGather, "ncclAllGather"); HB_SYM(Send, "ncclSend"); HB_SYM(Recv, "ncclRecv");
    HB_SYM(GroupStart, "ncclGroupStart"); HB_SYM(GroupEnd, "ncclGroupEnd");
    #undef HB_SYM
    return HB_OK;
}
int nccl_fail(ncclResult_t r, const char *what){
    hb_set_error(std::string(what) + " failed: " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
    return HB_ERR_NCCL;
}
#define HB_NCCL(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return nccl_fail(r__, #call); } while (0)
}

struct hb_dist {
    hb_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int n_owned = 0, n_ghost = 0;
    std::vector<int> neigh, send_count, recv_count;
    int send_total = 0;
    const int *send_idx = nullptr;      // device, caller-owned
    void *sendbuf = nullptr;            // device, send_total * 16 bytes
    size_t sendbuf_bytes = 0;
    // ---- peer-memory transport (hb_peer.cuh): exchange buffer = mailbox | p buffer 0 | p buffer 1, mapped by every peer
    int    peer_state = 0;              // 0 not tried for the current plan, 1 ready, -1 unavailable (NCCL path is used)
    size_t peer_es = 0;                 // element size the p buffers were laid out for
    void  *pbuf = nullptr;              // this rank's exchange buffer (cudaMalloc, exported through CUDA IPC)
    size_t pbuf_ext_bytes = 0;          // bytes of one p buffer
    void  *peer_base[HB_MAX_PEERS] = {};// rank q's exchange buffer as mapped into this process (q == rank: pbuf)
    peer_view *pv_dev = nullptr;
    // Sequence numbers of the peer protocol.  They MUST be equal on all ranks, and nothing a single host observes on its own
    // (how many batches it happened to enqueue before it saw the done flag) may enter them: every solve re-bases them with one
    // all-reduce(MAX) (peer_epoch_agree) and advances them by counts all ranks agree on.
    unsigned long long epoch = 0;       // global iteration number: halo flags and scalar slots carry epoch + 1
    unsigned long long vepoch = 0;      // number of peer vector all-reduces so far
    int    plan_version = 0;
    bool   peer_disabled = false;       // a wait timed out once: this communicator stays on NCCL, whatever plan comes next
    int    peer_fallbacks = 0;          // solves that gave up on the peer transport (time-out) and were redone over NCCL
    int    epoch_repairs = 0;           // peer_epoch_agree calls that found the ranks' counters different
};

hb_ctx* hb_dist_context(hb_dist *d){ return d->ctx; }
int hb_dist_owned(const hb_dist *d){ return d->n_owned; }
int hb_dist_ghosts(const hb_dist *d){ return d->n_ghost; }

// ------------------------------------------------------------------------------------------------ kernels
template<typename T> __global__ void pack_kernel(int n, const int * __restrict__ idx, const T * __restrict__ x, T *out){
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = x[idx[i]];
}

// distributed CG state: every scalar that crosses a kernel boundary is double-buffered by iteration parity, so no kernel
// reads a slot that another block of the same kernel writes
template<typename T> struct cg_dstate {
    T zr[2];
    T pAp;              // local, then all-reduced in place
    T rr;               // local ||r||^2 (as T, imaginary part 0), then all-reduced in place
    double rnorm, tol;
    int iterations, max_iter;
    int done[2];
    T pAp_local;        // peer transport: this rank's <p,Ap> partial as written by the SpMV+dot kernel
};
struct cg_dhost { volatile int done; volatile int iterations; volatile double rnorm; };

static constexpr int DK_THREADS = 256;

template<typename T>
__global__ void __launch_bounds__(DK_THREADS) dcg_setup_kernel(int n, cg_dstate<T> *st, double tol, int max_iter, const T * __restrict__ b,
                                                               const T * __restrict__ q, T *r, T *p, void *partials_v, unsigned int *ticket){
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
            st->rr = from_real<T>((real_t<T>) rr);      // all-reduced next, then dcg_begin_kernel turns it into zr[0]
            st->zr[0] = zero_of<T>(); st->zr[1] = zero_of<T>(); st->pAp = one_of<T>();
            st->rnorm = 0; st->tol = tol; st->iterations = 1; st->max_iter = max_iter; st->done[0] = 0; st->done[1] = 0;
        }
    }
}
template<typename T> __global__ void dcg_begin_kernel(cg_dstate<T> *st){ st->zr[0] = st->rr; st->rnorm = sqrt((double) hreal(st->rr)); }

template<typename T>
__global__ void __launch_bounds__(DK_THREADS) dcg_update_kernel(int n, cg_dstate<T> *st, int parity, const T * __restrict__ p, const T * __restrict__ q,
                                                                T *x, T *r, void *partials_v, unsigned int *ticket){
    __shared__ double red[32];
    if (st->done[parity]) return;
    const T a = hdiv(st->zr[parity], st->pAp);
    const T na = hneg(a);
    double acc = 0.0;
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < (size_t) n; i += stride){
        x[i] = hfma(a, p[i], x[i]);
        T ri = hfma(na, q[i], r[i]);
        r[i] = ri;
        acc += (double) habs2(ri);
    }
    double *partials = reinterpret_cast<double*>(partials_v);
    double b = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = b;
    if (last_block_arrives(ticket)){
        double rr = sum_partials<double>(partials, gridDim.x, 1, red);
        if (threadIdx.x == 0) st->rr = from_real<T>((real_t<T>) rr);
    }
}
// after the all-reduce of rr: stop test (every thread evaluates the same scalars), p = r + beta p, bookkeeping by one thread
template<typename T>
__global__ void __launch_bounds__(DK_THREADS) dcg_direction_kernel(int n, cg_dstate<T> *st, int parity, int it_now, const T * __restrict__ r, T *p,
                                                                   cg_dhost *host){
    if (st->done[parity]){
        if (blockIdx.x == 0 && threadIdx.x == 0) st->done[parity ^ 1] = 1;
        return;
    }
    const T rr = st->rr, zr = st->zr[parity];
    const double nrm = sqrt((double) hreal(rr));
    const int stop = (it_now >= st->max_iter) || (nrm < st->tol) || !(nrm == nrm);
    if (!stop){
        const T beta = hdiv(rr, zr);
        const size_t stride = (size_t) gridDim.x * blockDim.x;
        for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < (size_t) n; i += stride)
            p[i] = hfma(beta, p[i], r[i]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0){
        st->zr[parity ^ 1] = rr;
        st->done[parity ^ 1] = stop;
        st->iterations = it_now;
        st->rnorm = nrm;
        if (host){ host->rnorm = nrm; host->iterations = it_now; __threadfence_system(); if (stop) host->done = 1; }
    }
}

// ------------------------------------------------------------------------------------------------ peer-transport CG kernels
// One iteration g = three kernels and no collective call (protocol: hb_peer.cuh):
//   SpMV+dot   waits for the halo flags of g, gathers from p[g&1] = [owned | ghosts], publishes its <p,Ap> partial (channel PAP)
//   update     every block adds the W partials in rank order -> a = <r,z>/<p,Ap>; r -= a Ap; publishes its ||r||^2 partial (RR)
//   direction  every block adds the W partials -> stop test, b; FIRST stores the entries its neighbours need of the new p into
//              their ghost slots of buffer (g+1)&1 (recomputed from r and the old p, so no grid-wide dependency) and releases
//              the flags of g+1, THEN sweeps x += a p_old, p_new = r + b p_old.  p ping-pongs between two buffers, which is
//              what makes the early push legal and double-buffers the ghost slots for free.

__global__ void peer_halo_wait_kernel(const peer_view *pv, unsigned long long g, const int *skip){
    if (skip && *skip) return;
    peer_halo_wait(pv, g);
}
template<typename T> __global__ void peer_publish_kernel(const peer_view *pv, int channel, unsigned long long g, const T *value, const int *skip){
    if (skip && *skip) return;
    peer_publish<T>(pv, channel, g, *value);
}
// after the NCCL all-reduce of the setup's <r,r>: zr[0], ||r||, and the halo of the first direction p = r
template<typename T>
__global__ void __launch_bounds__(DK_THREADS) pcg_begin_kernel(cg_dstate<T> *st, const peer_view *pv, unsigned long long g,
                                                               const int * __restrict__ send_idx, const T * __restrict__ p, unsigned int *ticket){
    if (blockIdx.x == 0 && threadIdx.x == 0){ st->zr[0] = st->rr; st->rnorm = sqrt((double) hreal(st->rr)); }
    peer_halo_push<T>(pv, g, ticket, [&](int j){ return p[send_idx[j]]; });
}
template<typename T, bool VEC>
__global__ void __launch_bounds__(DK_THREADS, 4) pcg_update_kernel(int n, cg_dstate<T> *st, int parity, unsigned long long g, const T * __restrict__ q,
                                                                   T *r, void *partials_v, unsigned int *ticket, const peer_view *pv){
    __shared__ double red[32];
    __shared__ T s_pap;
    __shared__ double s_rr;
    if (st->done[parity]) return;
    if (threadIdx.x < 32){
        const T v = peer_wait_sum<T>(pv, HB_PEER_CH_PAP, g);
        if (threadIdx.x == 0) s_pap = v;
    }
    __syncthreads();
    const T pap = s_pap;
    const T na = hneg(hdiv(st->zr[parity], pap));
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
        if (threadIdx.x == 0){ st->pAp = pap; s_rr = rr; }
        __syncthreads();
        if (threadIdx.x < 32) peer_publish<double>(pv, HB_PEER_CH_RR, g, s_rr);
    }
}
template<typename T, bool VEC>
__global__ void __launch_bounds__(DK_THREADS, 4) pcg_direction_kernel(int n, cg_dstate<T> *st, int parity, unsigned long long g, int it_now,
                                                                      const T * __restrict__ r, const T * __restrict__ p_old, T *p_new, T *x,
                                                                      const peer_view *pv, const int * __restrict__ send_idx, cg_dhost *host,
                                                                      unsigned int *ticket){
    __shared__ double s_rr;
    if (st->done[parity]){
        if (blockIdx.x == 0 && threadIdx.x == 0) st->done[parity ^ 1] = 1;
        return;
    }
    if (threadIdx.x < 32){
        const double v = peer_wait_sum<double>(pv, HB_PEER_CH_RR, g);
        if (threadIdx.x == 0) s_rr = v;
    }
    __syncthreads();
    const double rrd = s_rr;
    const T rr = from_real<T>((real_t<T>) rrd), zr = st->zr[parity];
    const double nrm = sqrt(rrd);
    const int stop = (it_now >= st->max_iter) || (nrm < st->tol) || !(nrm == nrm);
    const T a = hdiv(zr, st->pAp);
    constexpr int U = 2;
    vec16<T> vx[U], vp[U], vr[U];
    vec16<T> *x4 = reinterpret_cast<vec16<T>*>(x), *pn4 = reinterpret_cast<vec16<T>*>(p_new);
    const vec16<T> *r4 = reinterpret_cast<const vec16<T>*>(r), *po4 = reinterpret_cast<const vec16<T>*>(p_old);
    if (!stop){
        const T beta = hdiv(rr, zr);
        peer_halo_push<T>(pv, g + 1, ticket, [&](int j){ const int i = send_idx[j]; return hfma(beta, p_old[i], r[i]); });
        stream_sweep<T, VEC, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vp[u] = po4[i]; vr[u] = r4[i]; },
            [&](int u, size_t i){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++){ vx[u].v[k] = hfma(a, vp[u].v[k], vx[u].v[k]); vp[u].v[k] = hfma(beta, vp[u].v[k], vr[u].v[k]); }
                x4[i] = vx[u]; pn4[i] = vp[u];
            },
            [&](size_t j){ const T pj = p_old[j]; x[j] = hfma(a, pj, x[j]); p_new[j] = hfma(beta, pj, r[j]); });
    }else if (nrm == nrm){                      // a NaN residual (peer time-out, breakdown) leaves x at the last good iterate
        stream_sweep<T, VEC, U>((size_t) n,
            [&](int u, size_t i){ vx[u] = x4[i]; vp[u] = po4[i]; },
            [&](int u, size_t i){
                #pragma unroll
                for (int k = 0; k < vec16<T>::N; k++) vx[u].v[k] = hfma(a, vp[u].v[k], vx[u].v[k]);
                x4[i] = vx[u];
            },
            [&](size_t j){ x[j] = hfma(a, p_old[j], x[j]); });
    }
    if (blockIdx.x == 0 && threadIdx.x == 0){
        st->zr[parity ^ 1] = rr;
        st->done[parity ^ 1] = stop;
        st->iterations = it_now;
        st->rnorm = nrm;
        if (host){ host->rnorm = nrm; host->iterations = it_now; __threadfence_system(); if (stop) host->done = 1; }
    }
}

// ------------------------------------------------------------------------------------------------ helpers
static ncclDataType_t real_dtype(int dtype){ return (dtype == HB_F32 || dtype == HB_C32) ? ncclFloat32 : ncclFloat64; }
static int reals_per_scalar(int dtype){ return (dtype == HB_C32 || dtype == HB_C64) ? 2 : 1; }
// blocks of a stand-alone halo push: two entries per thread, at most two blocks per SM
static int halo_grid(const hb_ctx *ctx, long long send_total){
    long long need = (send_total + 511) / 512, cap = (long long) ctx->num_sms * 2;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}
static int dgrid(const hb_ctx *ctx, long long n, int per_block){
    long long need = (n + per_block - 1) / per_block, cap = (long long) ctx->num_sms * 4;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

extern "C" int hb_dist_halo_exchange_nccl(hb_dist *d, int dtype, void *x_ext);
extern "C" int hb_dist_allreduce_sum_nccl(hb_dist *d, int dtype, void *dev_scalars, int count);

// ------------------------------------------------------------------------------------------------ peer transport: setup
namespace {
struct peer_info {                              // what every rank tells every other one (all-gathered through NCCL)
    cudaIpcMemHandle_t handle;                  // 64 B
    unsigned long long ext_bytes;               // bytes of one p buffer on that rank
    long long ok;                               // 0: that rank cannot take the peer path
    long long ghost_base[HB_MAX_PEERS];         // element index in that rank's p buffers where entries coming from rank q land (-1: none)
    char pad[256 - 64 - 16 - 8 * HB_MAX_PEERS];
};
static_assert(sizeof(peer_info) == 256, "peer_info layout");

bool peer_env_enabled(){
    const char *e = getenv("HB_DIST_PEER");
    return !(e && e[0] == '0');
}
bool peer_env_fused_spmv(){
    const char *e = getenv("HB_PEER_FUSED_SPMV");
    return !(e && e[0] == '0');
}
// in-place minimum over ranks of one int (agreement on "does everybody have it")
int agree_min(hb_dist *d, int *value){
    int *dv = reinterpret_cast<int*>(d->ctx->dscalars);
    HB_CUDA(cudaMemcpyAsync(dv, value, sizeof(int), cudaMemcpyHostToDevice, d->ctx->stream));
    HB_NCCL(g_nccl.AllReduce(dv, dv, 1, ncclInt32, ncclMin, d->comm, d->ctx->stream));
    HB_CUDA(cudaMemcpyAsync(value, dv, sizeof(int), cudaMemcpyDeviceToHost, d->ctx->stream));
    HB_CUDA(cudaStreamSynchronize(d->ctx->stream));
    return HB_OK;
}
void peer_release(hb_dist *d){
    for (int q = 0; q < HB_MAX_PEERS; q++){
        if (d->peer_base[q] && q != d->rank) cudaIpcCloseMemHandle(d->peer_base[q]);
        d->peer_base[q] = nullptr;
    }
    if (d->pbuf){ cudaFree(d->pbuf); d->pbuf = nullptr; }
    if (d->pv_dev){ cudaFree(d->pv_dev); d->pv_dev = nullptr; }
    d->peer_es = 0; d->pbuf_ext_bytes = 0;
    cudaGetLastError();
}
// Collective over all ranks.  HB_OK: the peer path is ready for element size `es`; HB_ERR_UNSUPPORTED: agreed by ALL ranks
// that it is not available (the callers then take the NCCL path); anything else is an error.
int peer_setup(hb_dist *d, size_t es){
    hb_range nvtx_range("hb_dist: peer transport set-up");
    if (d->peer_state == 1 && d->peer_es == es) return HB_OK;
    if (d->peer_state == -1) return HB_ERR_UNSUPPORTED;
    hb_ctx *ctx = d->ctx;
    const int W = d->world;
    int rc;
    if (d->peer_state == 1){                    // laid out for another element size: everybody leaves the old buffers first
        int one = 1;
        if ((rc = agree_min(d, &one)) != HB_OK) return rc;
        peer_release(d);
        d->peer_state = 0;
    }
    const size_t ext_elems = (size_t) d->n_owned + d->n_ghost;
    const size_t ext_bytes = ((es * ext_elems + 255) / 256) * 256;
    peer_info mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = (d->neigh.size() <= (size_t) HB_MAX_NEIGH) ? 1 : 0;
    mine.ext_bytes = ext_bytes;
    for (int q = 0; q < HB_MAX_PEERS; q++) mine.ghost_base[q] = -1;
    {
        long long roff = 0;
        for (size_t k = 0; k < d->neigh.size(); k++){
            if (d->recv_count[k] > 0) mine.ghost_base[d->neigh[k]] = (long long) d->n_owned + roff;
            roff += d->recv_count[k];
        }
    }
    // whole 2 MiB pages: small cudaMalloc requests are carved out of a shared block, and an IPC handle exports the block
    const size_t pbuf_bytes = ((HB_MAILBOX_BYTES + 2 * ext_bytes + (2u << 20) - 1) >> 21) << 21;
    if (mine.ok){
        if (cudaMalloc(&d->pbuf, pbuf_bytes) != cudaSuccess){ cudaGetLastError(); d->pbuf = nullptr; mine.ok = 0; }
        else if (cudaMemsetAsync(d->pbuf, 0, pbuf_bytes, ctx->stream) != cudaSuccess
                 || cudaIpcGetMemHandle(&mine.handle, d->pbuf) != cudaSuccess){ cudaGetLastError(); mine.ok = 0; }
    }
    // all-gather the 256-byte records
    char *stage = nullptr;
    HB_CUDA(cudaMalloc((void**) &stage, sizeof(peer_info) * (size_t) (W + 1)));
    std::vector<peer_info> all((size_t) W);
    cudaError_t ce = cudaMemcpyAsync(stage, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t nr = ncclSuccess;
    if (ce == cudaSuccess) nr = g_nccl.AllGather(stage, stage + sizeof(peer_info), sizeof(peer_info), ncclChar, d->comm, ctx->stream);
    if (ce == cudaSuccess && nr == ncclSuccess) ce = cudaMemcpyAsync(all.data(), stage + sizeof(peer_info), sizeof(peer_info) * (size_t) W, cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess && nr == ncclSuccess) ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(stage);
    if (nr != ncclSuccess) return nccl_fail(nr, "ncclAllGather (peer setup)");
    if (ce != cudaSuccess) return hb_cuda_fail(ce, "peer setup exchange");
    int ok = 1;
    for (int q = 0; q < W; q++) if (!all[q].ok) ok = 0;
    // map the peers' buffers
    if (ok){
        d->peer_base[d->rank] = d->pbuf;
        for (int q = 0; q < W && ok; q++){
            if (q == d->rank) continue;
            if (cudaIpcOpenMemHandle(&d->peer_base[q], all[q].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess){
                cudaGetLastError(); d->peer_base[q] = nullptr; ok = 0;
            }
        }
    }
    if ((rc = agree_min(d, &ok)) != HB_OK) return rc;
    if (!ok){
        peer_release(d);
        d->peer_state = -1;
        return HB_ERR_UNSUPPORTED;
    }
    peer_view pv;
    memset(&pv, 0, sizeof(pv));
    pv.rank = d->rank; pv.world = W; pv.nneigh = (int) d->neigh.size();
    {   // ~4 s of GPU clock by default; HB_PEER_TIMEOUT_MS shortens it (tests of the time-out / fallback path)
        const char *e = getenv("HB_PEER_TIMEOUT_MS");
        const double ms = e ? atof(e) : 4000.0;
        pv.timeout_clocks = (long long) ((ms > 1.0 ? ms : 1.0) * 2.0e6);
    }
    for (int q = 0; q < W; q++) pv.mail[q] = reinterpret_cast<peer_mailbox*>(d->peer_base[q]);
    int soff = 0;
    for (int k = 0; k < pv.nneigh; k++){
        const int nq = d->neigh[k];
        pv.neigh[k] = nq;
        pv.recv_from[k] = d->recv_count[k] > 0;
        pv.send_off[k] = soff;
        soff += d->send_count[k];
        for (int b = 0; b < 2; b++){
            pv.ghost_dst[k][b] = nullptr;
            if (d->send_count[k] > 0){
                if (all[nq].ghost_base[d->rank] < 0){ hb_set_error("exchange plans of two neighbours disagree"); peer_release(d); return HB_ERR_ARG; }
                pv.ghost_dst[k][b] = (char*) d->peer_base[nq] + HB_MAILBOX_BYTES + (size_t) b * all[nq].ext_bytes + (size_t) all[nq].ghost_base[d->rank] * es;
            }
        }
    }
    pv.send_off[pv.nneigh] = soff;
    HB_CUDA(cudaMalloc((void**) &d->pv_dev, sizeof(peer_view)));
    HB_CUDA(cudaMemcpy(d->pv_dev, &pv, sizeof(pv), cudaMemcpyHostToDevice));
    d->peer_es = es; d->pbuf_ext_bytes = ext_bytes;
    d->peer_state = 1;
    return HB_OK;
}

// All ranks leave with the same epoch / vepoch: the maximum over ranks (a rank whose counters ran ahead published nothing beyond
// what the others expect: flags only ever carry values <= the publishing rank's own counter).  One 16-byte all-reduce and one host
// synchronisation per solve; it also fences this solve's peer traffic from the previous one's.
int peer_epoch_agree(hb_dist *d){
    hb_ctx *ctx = d->ctx;
    unsigned long long *dv = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctx->dscalars) + 3072);
    unsigned long long *hv = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctx->hscalars) + 640);
    static const bool off = [](){ const char *e = getenv("HB_DEBUG_NO_EPOCH_AGREE"); return e && e[0] == '1'; }();
    if (off) return HB_OK;                          // test hook: reproduces the round-1 behaviour (per-host counting only)
    const unsigned long long mine[2] = {d->epoch, d->vepoch};
    hv[0] = mine[0]; hv[1] = mine[1];
    HB_CUDA(cudaMemcpyAsync(dv, hv, 16, cudaMemcpyHostToDevice, ctx->stream));
    HB_NCCL(g_nccl.AllReduce(dv, dv, 2, ncclUint64, ncclMax, d->comm, ctx->stream));
    HB_CUDA(cudaMemcpyAsync(hv, dv, 16, cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (hv[0] != mine[0] || hv[1] != mine[1]) d->epoch_repairs++;
    d->epoch = hv[0]; d->vepoch = hv[1];
    return HB_OK;
}
// maximum over ranks of the local mailbox's error flag (0 / 1), cleared on the way; enqueued behind the solve, one host sync
int peer_error_agree(hb_dist *d, int *any){
    hb_ctx *ctx = d->ctx;
    peer_mailbox *mail = reinterpret_cast<peer_mailbox*>(d->pbuf);
    int *dv = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->dscalars) + 3072 + 64);
    int *hv = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->hscalars) + 640 + 64);
    HB_CUDA(cudaMemcpyAsync(dv, &mail->error, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    HB_CUDA(cudaMemsetAsync(&mail->error, 0, sizeof(int), ctx->stream));
    HB_NCCL(g_nccl.AllReduce(dv, dv, 1, ncclInt32, ncclMax, d->comm, ctx->stream));
    HB_CUDA(cudaMemcpyAsync(hv, dv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    *any = *hv;
    return HB_OK;
}

// CG over the peer transport; see the kernel block above for the per-iteration protocol.
// *timed_out = 1 (on ALL ranks, agreed): a wait on a peer's flag gave up somewhere; x holds the last good iterate, *iters the
// operator applications spent so far, and the caller redoes the solve over NCCL.
int dist_cg_peer(hb_dist *d, const hb_csr *A, const void *b, void *x, double tol, int max_iter, int *iters, double *res, int *timed_out){
    hb_ctx *ctx = d->ctx;
    const int n = d->n_owned, dtype = A->dtype;
    const size_t es = hb_dtype_size(dtype);
    const size_t vec_bytes = ((es * (size_t) n + 255) / 256) * 256;
    void *arena = nullptr;
    int rc;
    *timed_out = 0;
    if ((rc = hb_ctx_workspace(ctx, 2 * vec_bytes + 256, &arena)) != HB_OK) return rc;       // r | Ap | state
    char *base = (char*) arena;
    void *r = base, *Ap = base + vec_bytes, *state = base + 2 * vec_bytes;
    char *pb[2] = {(char*) d->pbuf + HB_MAILBOX_BYTES, (char*) d->pbuf + HB_MAILBOX_BYTES + d->pbuf_ext_bytes};
    cg_dhost *hstat = reinterpret_cast<cg_dhost*>(reinterpret_cast<char*>(ctx->hscalars) + 512);
    void *hstat_dev = reinterpret_cast<char*>(ctx->hscalars_dev) + 512;
    hstat->done = 0; hstat->iterations = 0; hstat->rnorm = 0;
    const int grid = dgrid(ctx, n, DK_THREADS * 8);
    const bool fused = peer_env_fused_spmv() && hb_spmv_variant(A) == 3;
    const peer_view *pv = d->pv_dev;
    int trot = 0, twait = 0;
    if (fused && (rc = hb_csr_halo_order(ctx, A, &trot, &twait)) != HB_OK) return rc;
    {   // HB_PEER_HALO_DEFER=0: wait for the halo before the first gather of every CTA, natural tile order (A/B probe)
        const char *e = getenv("HB_PEER_HALO_DEFER");
        if (e && e[0] == '0'){ trot = 0; twait = 0; }
    }

    // the ranks' sequence numbers are re-based first (see hb_dist::epoch); g0 = the number of this solve's first iteration
    if ((rc = peer_epoch_agree(d)) != HB_OK) return rc;
    const unsigned long long g0 = d->epoch;

    // p0 <- x0 (owned) + halo over NCCL; Ap = A x0; r = b - Ap; p0 = r; <r,r> all-reduced over NCCL (this also fences the
    // previous solve's peer traffic from this one's); then the first halo push
    void *p0 = pb[g0 & 1];
    HB_CUDA(cudaMemcpyAsync(p0, x, es * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = hb_dist_halo_exchange_nccl(d, dtype, p0)) != HB_OK) return rc;
    if ((rc = hb_spmv_internal(ctx, A, p0, Ap, nullptr)) != HB_OK) return rc;
    HB_DISPATCH(dtype, {
        cg_dstate<T> *st = (cg_dstate<T>*) state;
        dcg_setup_kernel<T><<<dgrid(ctx, n, DK_THREADS * 4), DK_THREADS, 0, ctx->stream>>>(n, st, tol, max_iter, (const T*) b, (const T*) Ap, (T*) r, (T*) p0,
                                                                                           ctx->partials, ctx->tickets + 6);
        HB_LAUNCH_CHECK(ctx);
        if ((rc = hb_dist_allreduce_sum_nccl(d, dtype, &st->rr, 1)) != HB_OK) return rc;
        pcg_begin_kernel<T><<<halo_grid(ctx, d->send_total), DK_THREADS, 0, ctx->stream>>>(st, pv, g0, d->send_idx, (const T*) p0, ctx->tickets + 8);
        HB_LAUNCH_CHECK(ctx);
    });

    hb_prof_begin(ctx);
    cudaEvent_t ev[2];
    HB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    HB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    const int batch = 8;
    long long it = 0;
    int status = HB_OK;
    // How many batches get enqueued before this host sees the done flag depends on host timing and differs between ranks.  That is
    // harmless for the protocol — iterations past the stop are skipped by the done flag and publish nothing — as long as the count
    // never reaches d->epoch: g is derived from g0 and the iteration index only.
    for (long long bidx = 0; status == HB_OK; bidx++){
        for (int j = 0; j < batch && status == HB_OK; j++, it++){
            const int parity = (int) (it & 1);
            const unsigned long long g = g0 + (unsigned long long) it;
            void *p_old = pb[g & 1], *p_new = pb[(g + 1) & 1];
            const int it_now = (int) (it + 2 < 0x7fffffffLL ? it + 2 : 0x7fffffffLL);
            HB_DISPATCH(dtype, {
                cg_dstate<T> *st = (cg_dstate<T>*) state;
                const bool vec = aligned16(x) && aligned16(r) && aligned16(Ap) && aligned16(p_old) && aligned16(p_new);
                hb_prof_mark(ctx, it, 0);
                if (fused){
                    ctx->peer_hook = pv; ctx->peer_epoch = g; ctx->peer_trot = trot; ctx->peer_twait = twait;
                    status = hb_spmv_dot_internal(ctx, A, p_old, Ap, &st->pAp_local, &st->done[parity]);
                    ctx->peer_hook = nullptr;
                    if (status != HB_OK) break;
                }else{
                    peer_halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(pv, g, &st->done[parity]);
                    ctx->launches++;
                    if ((status = hb_spmv_dot_internal(ctx, A, p_old, Ap, &st->pAp_local, &st->done[parity])) != HB_OK) break;
                    peer_publish_kernel<T><<<1, 32, 0, ctx->stream>>>(pv, HB_PEER_CH_PAP, g, &st->pAp_local, &st->done[parity]);
                    ctx->launches++;
                }
                hb_prof_mark(ctx, it, 1);
                if (vec) pcg_update_kernel<T, true><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, g, (const T*) Ap, (T*) r, ctx->partials, ctx->tickets + 6, pv);
                else     pcg_update_kernel<T, false><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, g, (const T*) Ap, (T*) r, ctx->partials, ctx->tickets + 6, pv);
                ctx->launches++;
                hb_prof_mark(ctx, it, 2);
                if (vec) pcg_direction_kernel<T, true><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, g, it_now, (const T*) r, (const T*) p_old, (T*) p_new, (T*) x,
                                                                                             pv, d->send_idx, (cg_dhost*) hstat_dev, ctx->tickets + 8);
                else     pcg_direction_kernel<T, false><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, g, it_now, (const T*) r, (const T*) p_old, (T*) p_new, (T*) x,
                                                                                              pv, d->send_idx, (cg_dhost*) hstat_dev, ctx->tickets + 8);
                ctx->launches++;
                hb_prof_mark(ctx, it, 3);
            });
        }
        if (status != HB_OK) break;
        if (cudaPeekAtLastError() != cudaSuccess){ status = hb_cuda_fail(cudaGetLastError(), "kernel launch"); break; }
        if (cudaEventRecord(ev[bidx & 1], ctx->stream) != cudaSuccess){ status = hb_cuda_fail(cudaGetLastError(), "cudaEventRecord"); break; }
        if (bidx > 0){
            if (cudaEventSynchronize(ev[(bidx - 1) & 1]) != cudaSuccess){ status = hb_cuda_fail(cudaGetLastError(), "cudaEventSynchronize"); break; }
            if (hstat->done) break;
        }
    }
    {   // test hook: odd ranks count extra batches, as a host that saw the done flag late would have enqueued (and the done flag
        // skipped) them
        static const int skew = [](){ const char *e = getenv("HB_DEBUG_EPOCH_SKEW"); return e ? atoi(e) : 0; }();
        if (skew > 0 && (d->rank & 1)) it += (long long) batch * skew;
    }
    // the done flag stopped every rank at the same iteration: the (identical) number of operator applications is what the
    // sequence number advances by.  A solve that failed leaves the per-host count; the next peer_epoch_agree repairs it.
    int any_timeout = 0;
    if (status == HB_OK) status = peer_error_agree(d, &any_timeout);        // synchronises the stream
    else cudaStreamSynchronize(ctx->stream);
    if (status == HB_OK) hb_prof_collect(ctx, (long long) hstat->iterations - 1, 3);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    // HB_DEBUG_EPOCH_RULE=host: the round-1 rule (every host counts what it enqueued), kept as a test hook
    static const bool host_rule = [](){ const char *e = getenv("HB_DEBUG_EPOCH_RULE"); return e && e[0] == 'h'; }();
    if (status != HB_OK || any_timeout || host_rule) d->epoch = g0 + (unsigned long long) it;
    else d->epoch = g0 + (unsigned long long) hstat->iterations + 1;
    if (status != HB_OK) return status;
    if (iters) *iters = hstat->iterations;
    if (res) *res = hstat->rnorm;
    if (any_timeout){
        *timed_out = 1;
        hb_set_error("peer transport: a wait on a peer's flag timed out (rank " + std::to_string(d->rank) + ", epoch " + std::to_string(g0) + " + " +
                     std::to_string(hstat->iterations) + ")");
    }
    return HB_OK;
}
}

extern "C" {

int hb_dist_unique_id(void *id128){
    HB_ARG(id128, "null");
    int rc = load_nccl(); if (rc != HB_OK) return rc;
    ncclUniqueId id;
    HB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, HB_NCCL_ID_BYTES);
    return HB_OK;
}

int hb_dist_create(hb_ctx *ctx, int rank, int world, const void *id128, hb_dist **out){
    HB_ARG(ctx && id128 && out, "null");
    HB_ARG(world >= 1 && rank >= 0 && rank < world, "rank/world");
    int rc = load_nccl(); if (rc != HB_OK) return rc;
    HB_CUDA(cudaSetDevice(ctx->device));
    hb_dist *d = new hb_dist();
    d->ctx = ctx; d->rank = rank; d->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, HB_NCCL_ID_BYTES);
    ncclResult_t r = g_nccl.CommInitRank(&d->comm, world, id, rank);
    if (r != ncclSuccess){ delete d; return nccl_fail(r, "ncclCommInitRank"); }
    *out = d;
    return HB_OK;
}

int hb_dist_destroy(hb_dist *d){
    if (!d) return HB_OK;
    peer_release(d);
    if (d->sendbuf) cudaFree(d->sendbuf);
    if (d->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    delete d;
    return HB_OK;
}

int hb_dist_transport(const hb_dist *d, int *transport){
    HB_ARG(d && transport, "null");
    *transport = (d->peer_state == 1) ? HB_TRANSPORT_PEER : HB_TRANSPORT_NCCL;
    return HB_OK;
}

int hb_dist_debug_info(const hb_dist *d, unsigned long long *epoch, unsigned long long *vepoch, int *peer_fallbacks, int *epoch_repairs){
    HB_ARG(d, "dist is null");
    if (epoch) *epoch = d->epoch;
    if (vepoch) *vepoch = d->vepoch;
    if (peer_fallbacks) *peer_fallbacks = d->peer_fallbacks;
    if (epoch_repairs) *epoch_repairs = d->epoch_repairs;
    return HB_OK;
}

int hb_dist_info(const hb_dist *d, int *rank, int *world){
    HB_ARG(d, "dist is null");
    if (rank) *rank = d->rank;
    if (world) *world = d->world;
    return HB_OK;
}

int hb_dist_set_plan(hb_dist *d, int n_owned, int n_ghost, int nneigh, const int *neigh, const int *send_count, const int *recv_count,
                     const int *send_idx_dev){
    HB_ARG(d, "dist is null");
    HB_ARG(n_owned >= 0 && n_ghost >= 0 && nneigh >= 0, "negative size");
    HB_ARG(nneigh == 0 || (neigh && send_count && recv_count), "null plan arrays");
    if (d->peer_state == 1){                    // buffers were laid out for the previous plan; a new plan is set on all ranks together
        HB_CUDA(cudaStreamSynchronize(d->ctx->stream));
        int one = 1;
        int rc = agree_min(d, &one);
        if (rc != HB_OK) return rc;
        peer_release(d);
    }
    d->peer_state = d->peer_disabled ? -1 : 0; d->plan_version++;
    d->n_owned = n_owned; d->n_ghost = n_ghost;
    d->neigh.assign(neigh, neigh + nneigh);
    d->send_count.assign(send_count, send_count + nneigh);
    d->recv_count.assign(recv_count, recv_count + nneigh);
    long long st = 0, rt = 0;
    for (int k = 0; k < nneigh; k++){
        HB_ARG(neigh[k] >= 0 && neigh[k] < d->world && neigh[k] != d->rank, "neighbour rank");
        st += send_count[k]; rt += recv_count[k];
    }
    HB_ARG(rt == n_ghost, "recv counts do not add up to n_ghost");
    HB_ARG(st == 0 || send_idx_dev, "send_idx is null");
    d->send_total = (int) st;
    d->send_idx = send_idx_dev;
    const size_t need = (size_t) st * 16;
    if (need > d->sendbuf_bytes){
        if (d->sendbuf) HB_CUDA(cudaFree(d->sendbuf));
        HB_CUDA(cudaMalloc(&d->sendbuf, need));
        d->sendbuf_bytes = need;
    }
    return HB_OK;
}

int hb_dist_halo_exchange_nccl(hb_dist *d, int dtype, void *x_ext){
    HB_ARG(d && (x_ext || d->n_owned + d->n_ghost == 0), "null");
    hb_ctx *ctx = d->ctx;
    if (d->neigh.empty()) return HB_OK;
    const size_t es = hb_dtype_size(dtype);
    if (d->send_total > 0){
        int grid = dgrid(ctx, d->send_total, 256);
        HB_DISPATCH(dtype, (pack_kernel<T><<<grid, 256, 0, ctx->stream>>>(d->send_total, d->send_idx, (const T*) x_ext, (T*) d->sendbuf)));
        HB_LAUNCH_CHECK(ctx);
    }
    const int rps = reals_per_scalar(dtype);
    const ncclDataType_t nt = real_dtype(dtype);
    HB_NCCL(g_nccl.GroupStart());
    size_t soff = 0, roff = 0;
    for (size_t k = 0; k < d->neigh.size(); k++){
        if (d->send_count[k] > 0)
            HB_NCCL(g_nccl.Send((char*) d->sendbuf + soff * es, (size_t) d->send_count[k] * rps, nt, d->neigh[k], d->comm, ctx->stream));
        if (d->recv_count[k] > 0)
            HB_NCCL(g_nccl.Recv((char*) x_ext + ((size_t) d->n_owned + roff) * es, (size_t) d->recv_count[k] * rps, nt, d->neigh[k], d->comm, ctx->stream));
        soff += d->send_count[k]; roff += d->recv_count[k];
    }
    HB_NCCL(g_nccl.GroupEnd());
    return HB_OK;
}

int hb_dist_allreduce_sum_nccl(hb_dist *d, int dtype, void *dev_scalars, int count){
    HB_ARG(d && dev_scalars && count >= 0, "null");
    if (d->world == 1 || count == 0) return HB_OK;
    HB_NCCL(g_nccl.AllReduce(dev_scalars, dev_scalars, (size_t) count * reals_per_scalar(dtype), real_dtype(dtype), ncclSum, d->comm, d->ctx->stream));
    return HB_OK;
}

// Public forms: over peer memory once the transport of the current plan is up for this element size (hb_dist_cg / hb_dist_gmres
// bring it up collectively), over NCCL otherwise.
int hb_dist_halo_exchange(hb_dist *d, int dtype, void *x_ext){
    HB_ARG(d && (x_ext || d->n_owned + d->n_ghost == 0), "null");
    const size_t es = hb_dtype_size(dtype);
    if (!(d->peer_state == 1 && d->peer_es == es)) return hb_dist_halo_exchange_nccl(d, dtype, x_ext);
    if (d->neigh.empty()) return HB_OK;
    hb_ctx *ctx = d->ctx;
    const unsigned long long g = d->epoch++;
    char *pb = (char*) d->pbuf + HB_MAILBOX_BYTES + (size_t) (g & 1) * d->pbuf_ext_bytes;
    HB_DISPATCH(dtype, {
        peer_vec_push_kernel<T><<<halo_grid(ctx, d->send_total), 256, 0, ctx->stream>>>(d->pv_dev, g, d->send_idx, (const T*) x_ext, ctx->tickets + 8);
        ctx->launches++;
        const int grid = dgrid(ctx, d->n_ghost, 256 * 4);
        peer_vec_pull_kernel<T><<<grid, 256, 0, ctx->stream>>>(d->pv_dev, g, (const T*) pb + d->n_owned, (T*) x_ext + d->n_owned, d->n_ghost);
    });
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
int hb_dist_allreduce_sum(hb_dist *d, int dtype, void *dev_scalars, int count){
    HB_ARG(d && dev_scalars && count >= 0, "null");
    if (d->world == 1 || count == 0) return HB_OK;
    if (!(d->peer_state == 1 && count <= HB_PEER_VMAX)) return hb_dist_allreduce_sum_nccl(d, dtype, dev_scalars, count);
    hb_ctx *ctx = d->ctx;
    const unsigned long long v = d->vepoch++;
    HB_DISPATCH(dtype, (peer_allsum_kernel<T><<<1, 128, 0, ctx->stream>>>(d->pv_dev, v, (T*) dev_scalars, count)));
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
// brings the peer transport up for this element size when it is available (collective) and re-bases the ranks' sequence numbers;
// HB_OK either way unless something failed
int hb_dist_prepare_transport(hb_dist *d, int dtype){
    HB_ARG(d, "null");
    if (d->world > 1 && d->world <= HB_MAX_PEERS && peer_env_enabled()){
        int prc = peer_setup(d, hb_dtype_size(dtype));
        if (prc != HB_OK && prc != HB_ERR_UNSUPPORTED) return prc;
        if (prc == HB_OK && (prc = peer_epoch_agree(d)) != HB_OK) return prc;
    }
    return HB_OK;
}
// collective, after a solve that used the public halo / all-reduce forms: *timed_out = 1 on all ranks when a wait on a peer's flag
// gave up anywhere; the peer buffers are then given back and the communicator stays on NCCL (the caller redoes its solve)
int hb_dist_finish_transport(hb_dist *d, int *timed_out){
    HB_ARG(d && timed_out, "null");
    *timed_out = 0;
    if (d->peer_state != 1) return HB_OK;
    int any = 0, rc = peer_error_agree(d, &any);
    if (rc != HB_OK) return rc;
    if (any){
        hb_set_error("peer transport: a wait on a peer's flag timed out (rank " + std::to_string(d->rank) + ")");
        if (getenv("HB_DIST_VERBOSE")) fprintf(stderr, "[hb_dist rank %d] %s -> redoing the solve over NCCL\n", d->rank, hb_last_error());
        d->peer_fallbacks++;
        peer_release(d);
        d->peer_state = -1; d->peer_disabled = true;
        *timed_out = 1;
    }
    return HB_OK;
}

int hb_dist_spmv(hb_dist *d, const hb_csr *A, void *x_ext, void *y){
    HB_ARG(d && A, "null");
    HB_ARG(A->rows == d->n_owned && A->cols == d->n_owned + d->n_ghost, "matrix shape does not match the exchange plan");
    int rc = hb_dist_halo_exchange(d, A->dtype, x_ext);
    if (rc != HB_OK) return rc;
    return hb_spmv_internal(d->ctx, A, x_ext, y, nullptr);
}

int hb_dist_cg(hb_dist *d, const hb_csr *A, const void *b, void *x, double tol, int max_iter, int *iters, double *res){
    hb_range nvtx_range("hb_dist_cg");
    HB_ARG(d && A && b && x, "null");
    HB_ARG(A->rows == d->n_owned && A->cols == d->n_owned + d->n_ghost, "matrix shape does not match the exchange plan");
    hb_ctx *ctx = d->ctx;
    const int n = d->n_owned, dtype = A->dtype;
    const size_t es = hb_dtype_size(dtype);
    // transport: peer memory over NVLink when every rank can map every other one's exchange buffer (one NVLink domain),
    // otherwise NCCL send/recv + all-reduce.  The choice is agreed by all ranks inside peer_setup.
    if (d->world > 1 && d->world <= HB_MAX_PEERS && peer_env_enabled()){
        int prc = peer_setup(d, es);
        if (prc == HB_OK){
            int timed_out = 0, it_peer = 0;
            double res_peer = 0;
            prc = dist_cg_peer(d, A, b, x, tol, max_iter, &it_peer, &res_peer, &timed_out);
            if (prc != HB_OK) return prc;
            if (!timed_out){ if (iters) *iters = it_peer; if (res) *res = res_peer; return HB_OK; }
            // Agreed by all ranks: some wait on a peer's flag gave up.  x is the last good iterate on every rank; the peer buffers
            // are given back and this communicator stays on NCCL.  The solve restarts from x, so the iteration count is that of a
            // restarted CG: operator applications of both legs are reported together.
            if (getenv("HB_DIST_VERBOSE")) fprintf(stderr, "[hb_dist rank %d] %s -> redoing the solve over NCCL\n", d->rank, hb_last_error());
            d->peer_fallbacks++;
            peer_release(d);
            d->peer_state = -1; d->peer_disabled = true;
            const int spent = it_peer > 1 ? it_peer - 1 : 0;
            max_iter = max_iter > spent + 2 ? max_iter - spent : 2;
            int it2 = 0;
            prc = hb_dist_cg(d, A, b, x, tol, max_iter, &it2, res);
            if (iters) *iters = spent + it2;
            return prc;
        }
        if (prc != HB_ERR_UNSUPPORTED) return prc;
    }
    const size_t vec_bytes = ((es * (size_t) n + 255) / 256) * 256, ext_bytes = ((es * ((size_t) n + d->n_ghost) + 255) / 256) * 256;
    void *arena = nullptr;
    int rc;
    // r | Ap | p_ext | state
    if ((rc = hb_ctx_workspace(ctx, 2 * vec_bytes + ext_bytes + 256, &arena)) != HB_OK) return rc;
    char *base = (char*) arena;
    void *r = base, *Ap = base + vec_bytes, *p = base + 2 * vec_bytes, *state = base + 2 * vec_bytes + ext_bytes;
    cg_dhost *hstat = reinterpret_cast<cg_dhost*>(reinterpret_cast<char*>(ctx->hscalars) + 512);
    void *hstat_dev = reinterpret_cast<char*>(ctx->hscalars_dev) + 512;
    hstat->done = 0; hstat->iterations = 0; hstat->rnorm = 0;
    const int grid = dgrid(ctx, n, DK_THREADS * 4);

    // p_ext <- x0 (owned), ghosts exchanged; Ap = A x0; r = b - Ap; p = r; zr = all-reduced <r,r>
    HB_CUDA(cudaMemcpyAsync(p, x, es * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = hb_dist_halo_exchange(d, dtype, p)) != HB_OK) return rc;
    if ((rc = hb_spmv_internal(ctx, A, p, Ap, nullptr)) != HB_OK) return rc;
    HB_DISPATCH(dtype, {
        cg_dstate<T> *st = (cg_dstate<T>*) state;
        dcg_setup_kernel<T><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, tol, max_iter, (const T*) b, (const T*) Ap, (T*) r, (T*) p, ctx->partials, ctx->tickets + 6);
        HB_LAUNCH_CHECK(ctx);
        if ((rc = hb_dist_allreduce_sum(d, dtype, &st->rr, 1)) != HB_OK) return rc;
        dcg_begin_kernel<T><<<1, 1, 0, ctx->stream>>>(st);
        HB_LAUNCH_CHECK(ctx);
    });

    cudaEvent_t ev[2];
    HB_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    HB_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    hb_prof_begin(ctx);
    const int batch = 4;
    long long it = 0;
    int status = HB_OK;
    for (long long bidx = 0; status == HB_OK; bidx++){
        for (int j = 0; j < batch && status == HB_OK; j++, it++){
            const int parity = (int) (it & 1);
            HB_DISPATCH(dtype, {
                cg_dstate<T> *st = (cg_dstate<T>*) state;
                // every rank runs the collectives of every enqueued iteration, converged or not (the kernels around them
                // are skipped by the done flag); all ranks see the same flag, so the call sequences stay matched
                hb_prof_mark(ctx, it, 0);
                if ((status = hb_dist_halo_exchange(d, dtype, p)) != HB_OK) break;
                if ((status = hb_spmv_dot_internal(ctx, A, p, Ap, &st->pAp, &st->done[parity])) != HB_OK) break;
                hb_prof_mark(ctx, it, 1);
                if ((status = hb_dist_allreduce_sum(d, dtype, &st->pAp, 1)) != HB_OK) break;
                dcg_update_kernel<T><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, (const T*) p, (const T*) Ap, (T*) x, (T*) r, ctx->partials, ctx->tickets + 6);
                ctx->launches++;
                hb_prof_mark(ctx, it, 2);
                if ((status = hb_dist_allreduce_sum(d, dtype, &st->rr, 1)) != HB_OK) break;
                dcg_direction_kernel<T><<<grid, DK_THREADS, 0, ctx->stream>>>(n, st, parity, (int) (it + 2), (const T*) r, (T*) p, (cg_dhost*) hstat_dev);
                ctx->launches++;
                hb_prof_mark(ctx, it, 3);
            });
        }
        if (status != HB_OK) break;
        if (cudaEventRecord(ev[bidx & 1], ctx->stream) != cudaSuccess){ status = hb_cuda_fail(cudaGetLastError(), "cudaEventRecord"); break; }
        if (bidx > 0){
            if (cudaEventSynchronize(ev[(bidx - 1) & 1]) != cudaSuccess){ status = hb_cuda_fail(cudaGetLastError(), "cudaEventSynchronize"); break; }
            if (hstat->done) break;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    if (status != HB_OK) return status;
    if (e != cudaSuccess) return hb_cuda_fail(e, "cudaStreamSynchronize");
    hb_prof_collect(ctx, (long long) hstat->iterations - 1, 3);
    if (iters) *iters = hstat->iterations;
    if (res) *res = hstat->rnorm;
    return HB_OK;
}

}
