// hb_peer.cuh — peer-memory transport of the row-partitioned solvers: every rank of one NVLink/NVSwitch domain maps the
// others' exchange buffer (CUDA IPC) and the iteration kernels talk through it directly, so a CG iteration contains no
// collective call at all:
//   * halo:   the kernel that PRODUCES p stores the entries its neighbours need straight into their ghost slots
//             (remote stores over NVLink by the whole grid; the block that finishes last releases one flag per neighbour); the
//             SpMV that consumes them waits on those flags in-kernel, in front of its first tile that references a ghost column.
//   * scalar sums (<p,Ap>, ||r||^2): the last block of the producing kernel stores this rank's partial into a slot of every
//             peer's mailbox; every block of the consuming kernel waits for the W slots and adds them in rank order, so all
//             ranks (and all blocks) get the same bits — the iteration stays in lock step without a host or NCCL round trip.
// Slots and ghost buffers are double-buffered by the parity of a global iteration number g ("epoch"); a flag carries g + 1.
// Why two buffers suffice: rank X can only write epoch g + 2 after it consumed a value of epoch g + 1 from every peer Y, and
// Y published that value after the kernel in which it consumed epoch g (stream order).  The first exchanges of a solve go
// through NCCL (halo of x0, all-reduce of <r,r>), which also fences one solve from the next.
// The epoch MUST be the same number on every rank: hb_dist.cu derives it from counts all ranks agree on (iterations executed,
// collective calls made) and re-bases it with an all-reduce(MAX) at the start of every solve; a wait that times out poisons the
// sums with NaN, which stops every rank, and the solve is redone over NCCL.
#pragma once
#include "hb_common.cuh"

static constexpr int HB_MAX_PEERS   = 16;      // ranks of one NVLink domain (8 on an HGX B200 board)
static constexpr int HB_MAX_NEIGH   = 8;       // halo neighbours per rank on this path (1-D row blocks of a stencil: 2)
static constexpr int HB_HALO_BLOCKS = 32;      // flag slots per source rank in the mailbox (slot 0 is the one in use)
static constexpr int HB_PEER_CH_PAP = 0, HB_PEER_CH_RR = 1;

struct peer_slot { double v[2]; unsigned long long seq; unsigned long long pad; };              // 32 B
static constexpr int HB_PEER_VMAX = 66;        // scalars per vector all-reduce (GMRES: restart <= 64 Gram-Schmidt coefficients + the norm)
struct peer_vslot { double v[2 * HB_PEER_VMAX]; unsigned long long seq; unsigned long long pad; };
// mailbox at the start of every rank's exchange buffer
struct peer_mailbox {
    peer_slot slot[2][2][HB_MAX_PEERS];                             // [channel][epoch parity][source rank]
    unsigned long long halo_seq[HB_MAX_PEERS][HB_HALO_BLOCKS];      // [source rank][0] = epoch + 1 of that rank's last completed push (slots 1.. unused)
    int error;                                                      // set when a wait timed out (peer died / mis-sequenced)
    int pad[15];
    peer_vslot vslot[2][HB_MAX_PEERS];                              // [vector-sum epoch parity][source rank]
};
static constexpr size_t HB_MAILBOX_BYTES = 65536;
static_assert(sizeof(peer_mailbox) <= HB_MAILBOX_BYTES, "mailbox layout");

// what a kernel needs to reach its peers (lives in device memory, built by peer_setup)
struct peer_view {
    int rank, world, nneigh, pad;
    long long timeout_clocks;                   // a wait gives up after this many GPU clocks (default ~4 s; HB_PEER_TIMEOUT_MS)
    peer_mailbox *mail[HB_MAX_PEERS];           // mailbox of rank q as mapped here (q == rank: the local one)
    int neigh[HB_MAX_NEIGH];                    // rank of halo neighbour k
    int recv_from[HB_MAX_NEIGH];                // 1 when neighbour k sends us ghost entries
    int send_off[HB_MAX_NEIGH + 1];             // neighbour k receives send_idx[send_off[k] .. send_off[k+1])
    char *ghost_dst[HB_MAX_NEIGH][2];           // where OUR entries land in neighbour k's p buffer of parity 0/1 (mapped here)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v){
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p){
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double *p){
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// spin until *flag >= target; gives up after pv->timeout_clocks (~4 s of GPU clock by default: a peer that died must not hang
// this GPU): returns false
__device__ __forceinline__ bool peer_spin(const unsigned long long *flag, unsigned long long target, long long timeout_clocks){
    if (ld_acquire_sys(flag) >= target) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < target){
        __nanosleep(40);
        if (clock64() - t0 > timeout_clocks) return false;
    }
    return true;
}

template<typename T> __device__ __forceinline__ void slot_pack(T v, double &a, double &b);
template<> __device__ __forceinline__ void slot_pack<float>(float v, double &a, double &b){ a = (double) v; b = 0.0; }
template<> __device__ __forceinline__ void slot_pack<double>(double v, double &a, double &b){ a = v; b = 0.0; }
template<> __device__ __forceinline__ void slot_pack<cplx<float>>(cplx<float> v, double &a, double &b){ a = (double) v.re; b = (double) v.im; }
template<> __device__ __forceinline__ void slot_pack<cplx<double>>(cplx<double> v, double &a, double &b){ a = v.re; b = v.im; }
template<typename T> __device__ __forceinline__ T slot_unpack(double a, double b);
template<> __device__ __forceinline__ float  slot_unpack<float>(double a, double){ return (float) a; }
template<> __device__ __forceinline__ double slot_unpack<double>(double a, double){ return a; }
template<> __device__ __forceinline__ cplx<float>  slot_unpack<cplx<float>>(double a, double b){ return {(float) a, (float) b}; }
template<> __device__ __forceinline__ cplx<double> slot_unpack<cplx<double>>(double a, double b){ return {a, b}; }

// Called by (at least) the first warp of ONE block, all 32 lanes converged: lane q < world stores `value` into rank q's mailbox
// slot [channel][g & 1][our rank] and releases it with seq = g + 1.  `value` must be the same in every calling lane.
template<typename T> __device__ __forceinline__ void peer_publish(const peer_view *pv, int channel, unsigned long long g, T value){
    const int q = threadIdx.x;
    if (q < pv->world){
        peer_slot *s = &pv->mail[q]->slot[channel][g & 1][pv->rank];
        double a, b;
        slot_pack<T>(value, a, b);
        volatile double *sv = s->v;
        sv[0] = a; sv[1] = b;
        __threadfence_system();
        st_release_sys(&s->seq, g + 1);
    }
}
// Called by the first warp of a block (threadIdx.x < 32, converged): waits for the W partials of (channel, g) in the LOCAL
// mailbox and returns their sum in rank order (bit-identical on every rank and block) in all 32 lanes.  A timed-out wait marks
// the mailbox and yields NaN, which the solvers' stop test turns into an orderly stop on every rank.
template<typename T> __device__ __forceinline__ T peer_wait_sum(const peer_view *pv, int channel, unsigned long long g){
    const int q = threadIdx.x & 31, W = pv->world;
    peer_mailbox *mine = pv->mail[pv->rank];
    double a = 0.0, b = 0.0;
    bool ok = true;
    for (int s0 = q; s0 < W; s0 += 32){         // W <= 16: one trip
        const peer_slot *s = &mine->slot[channel][g & 1][s0];
        ok = peer_spin(&s->seq, g + 1, pv->timeout_clocks);
        a = ld_volatile_f64(&s->v[0]); b = ld_volatile_f64(&s->v[1]);
    }
    if (!__all_sync(0xffffffffu, ok)){
        if (q == 0) mine->error = 1;
        a = b = __longlong_as_double(0x7ff8000000000000LL);
    }
    double sa = 0.0, sb = 0.0;
    for (int s0 = 0; s0 < W; s0++){             // rank order, sequential: deterministic
        sa += __shfl_sync(0xffffffffu, a, s0);
        sb += __shfl_sync(0xffffffffu, b, s0);
    }
    return slot_unpack<T>(sa, sb);
}
// Called by the first warp of a block (converged): waits until every neighbour that sends us ghosts has released its flag for
// epoch g (one flag per neighbour: the LAST of the neighbour's pushing blocks releases it).  Returns false on time-out (mailbox marked).
__device__ __forceinline__ bool peer_halo_wait(const peer_view *pv, unsigned long long g){
    const int lane = threadIdx.x & 31;
    peer_mailbox *mine = pv->mail[pv->rank];
    bool ok = true;
    for (int k = lane; k < pv->nneigh; k += 32)
        if (pv->recv_from[k]) ok = peer_spin(&mine->halo_seq[pv->neigh[k]][0], g + 1, pv->timeout_clocks);
    ok = __all_sync(0xffffffffu, ok);
    if (!ok && lane == 0) mine->error = 1;
    return ok;
}
// Halo push by ALL blocks of a grid (every thread of every block calls it, converged per block).  value(j) = the new entry at local
// index send_idx[j]; entry j of the concatenated send list goes to neighbour k with send_off[k] <= j < send_off[k+1], into its ghost
// slot of buffer parity (g & 1).  The list is dealt over the whole grid (a few entries per thread: the push is a chain of dependent
// loads and a remote store, so 32 blocks walking 16 K entries each took ~50 us of the direction kernel at 8 GPUs; the whole grid takes
// a few).  Every block fences its stores at system scope and takes a ticket; the block that takes the last one releases ONE flag
// g + 1 per neighbour.  `ticket` is a device counter that is zero between kernels (re-armed by the last block).
template<typename T, typename F>
__device__ __forceinline__ void peer_halo_push(const peer_view *pv, unsigned long long g, unsigned int *ticket, F value){
    const int total = pv->send_off[pv->nneigh];
    if (total == 0) return;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x){
        int k = 0;
        while (j >= pv->send_off[k + 1]) k++;
        T *dst = reinterpret_cast<T*>(pv->ghost_dst[k][g & 1]) + (j - pv->send_off[k]);
        *dst = value(j);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0){
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1){
            *ticket = 0;
            __threadfence_system();
            for (int k = 0; k < pv->nneigh; k++)
                if (pv->send_off[k + 1] > pv->send_off[k]) st_release_sys(&pv->mail[pv->neigh[k]]->halo_seq[pv->rank][0], g + 1);
        }
    }
}

// In-place sum over ranks of `count` (<= HB_PEER_VMAX) scalars that live on the device: ONE block.  Every rank stores its values
// into slot [v & 1][rank] of every peer's mailbox (v = number of this all-reduce, same on all ranks), releases a flag, waits for
// the W flags in its own mailbox and adds the W vectors in rank order — the same bits on every rank.  Replaces ncclAllReduce for
// the Gram-Schmidt coefficients of the row-partitioned GMRES (latency of one NVLink round trip instead of a collective launch).
template<typename T>
__global__ void __launch_bounds__(128) peer_allsum_kernel(const peer_view *pv, unsigned long long v, T *vec, int count){
    const int W = pv->world, rank = pv->rank, t = threadIdx.x;
    peer_mailbox *mine = pv->mail[rank];
    if (t < count){
        double a, b;
        slot_pack<T>(vec[t], a, b);
        for (int q = 0; q < W; q++){
            volatile double *dst = pv->mail[q]->vslot[v & 1][rank].v;
            dst[2 * t] = a; dst[2 * t + 1] = b;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < W) st_release_sys(&pv->mail[t]->vslot[v & 1][rank].seq, v + 1);
    __shared__ int ok_s;
    if (t == 0) ok_s = 1;
    __syncthreads();
    if (t < W && !peer_spin(&mine->vslot[v & 1][t].seq, v + 1, pv->timeout_clocks)) ok_s = 0;
    __syncthreads();
    if (t < count){
        double sa = 0.0, sb = 0.0;
        for (int q = 0; q < W; q++){
            sa += ld_volatile_f64(&mine->vslot[v & 1][q].v[2 * t]);
            sb += ld_volatile_f64(&mine->vslot[v & 1][q].v[2 * t + 1]);
        }
        if (!ok_s) sa = sb = __longlong_as_double(0x7ff8000000000000LL);
        vec[t] = slot_unpack<T>(sa, sb);
    }
    if (t == 0 && !ok_s) mine->error = 1;
}
// Halo of an arbitrary vector v_ext = [owned | ghosts] (GMRES basis vectors): push = our boundary entries into the neighbours' ghost
// slots of exchange buffer (g & 1) + flags; pull = wait for the neighbours' flags, then copy our ghost slots behind the owned part.
template<typename T>
__global__ void __launch_bounds__(256) peer_vec_push_kernel(const peer_view *pv, unsigned long long g, const int * __restrict__ send_idx, const T * __restrict__ v,
                                                            unsigned int *ticket){
    peer_halo_push<T>(pv, g, ticket, [&](int j){ return v[send_idx[j]]; });
}
template<typename T>
__global__ void __launch_bounds__(256) peer_vec_pull_kernel(const peer_view *pv, unsigned long long g, const T *ghost_src, T *ghost_dst, int n_ghost){
    if (threadIdx.x < 32) peer_halo_wait(pv, g);
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ghost; i += gridDim.x * blockDim.x) ghost_dst[i] = ld_cg_T(ghost_src + i);
}
