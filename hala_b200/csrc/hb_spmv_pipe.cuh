// hb_spmv_pipe.cuh — the streaming CSR SpMV kernel: persistent CTAs, bulk-async (TMA engine) staging of the CSR stream
// into a shared-memory ring, sub-warp-per-row consumption.
//
// Why: a CSR row-vector kernel saturates the L1TEX tag stage long before HBM — every warp-level load of values/col_idx
// touches 4-5 cache lines because rows are 27*8 B apart (ncu: l1tex 68 %, DRAM 45 % on the 27-point Laplacian).  Here
// the matrix stream never enters L1TEX: it goes HBM -> L2 -> shared memory through cp.async.bulk (SASS: UBLKCP), and
// L1TEX is left with the only access that needs a cache, the gather of x.
//
// Work split: rows are cut into tiles of ROWS = THREADS/TPR consecutive rows; the tiles are cut into G contiguous pieces of
// equal non-zero count once per matrix (tile table built by hb_csr_create); CTA g streams its piece.
// Producer (thread 0): for tile t it issues three bulk copies into ring stage t % STAGES — the tile's slice of pntr, of
// col_idx and of values (one contiguous range each) — completing on the stage's mbarrier, STAGES-1 tiles ahead of the
// consumers; the two row-pointer bounds it needs are themselves read one step earlier, so the producer never waits.
// Consumers (all threads): TPR lanes per row, lane l walks entries rs+l, rs+l+TPR, ... in batches of four
// (all shared-memory reads, then all gathers, then the FMAs), shuffle-reduce, write y.  Adjacent lane groups own
// adjacent rows, so on banded matrices one gather instruction touches 2-3 lines.  One __syncthreads per tile releases
// the stage — and keeps the warps of a CTA on neighbouring rows at the same time, which is what makes one warp's x lines
// the next warp's L1 hits (handing the stages back through an mbarrier instead was measured: 1.7x slower).
// Tiles whose slice does not fit a stage are processed straight from global memory, long rows by the whole CTA; matrices
// with heavy-tailed row lengths do not take this form at all but the virtual-row one (VS, below).
#pragma once
#include "hb_common.cuh"
#include "hb_peer.cuh"
#include "hb_async.cuh"

// ring capacity per stage = THREADS * pipe_slots<T>() non-zeros; entries per lane per batch = the same number
template<typename T> __host__ __device__ constexpr int pipe_slots(){ return sizeof(T) == 16 ? 4 : 8; }
// the virtual-row form (VS) keeps its stages small: its gathers are random, every miss needs an L1 line to land in, and shared memory
// is carved out of the same array as L1 (see pipe_configure in hb_spmv.cu)
static constexpr int VS_SLOT_DIV = 1;   // (halved stages were tried: more barriers per non-zero, 351 us against 305 us on the power-law matrix)
template<typename T, bool VS> __host__ __device__ constexpr int pipe_slots_of(){ return VS ? pipe_slots<T>() / VS_SLOT_DIV : pipe_slots<T>(); }
static constexpr int PIPE_LONGROW = 2048;       // slow path: rows at least this long are reduced by the whole CTA
static constexpr int PIPE_WARPROW = 96;         // rows at least this long are reduced by a whole warp instead of TPR lanes

// cta_tiles[g] = first tile t (of `tile_rows` rows) whose first non-zero index is >= g*nnz/G ; cta_tiles[G] = ntiles
__global__ void csr_partition_kernel(int rows, int nnz, const int *pntr, int tile_rows, int G, int *cta_tiles){
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    const int ntiles = (rows + tile_rows - 1) / tile_rows;
    if (g == G){ cta_tiles[g] = ntiles; return; }
    const long long target = (long long) nnz * g / G;
    int lo = 0, hi = ntiles;
    while (lo < hi){
        int mid = lo + ((hi - lo) >> 1);
        if (pntr[(long long) mid * tile_rows] >= target) hi = mid; else lo = mid + 1;
    }
    cta_tiles[g] = lo;
}

// NBP > 0 turns the same kernel into the multi right-hand-side product of the batch solvers (hb_spmm): x is then the row-major
// ("interleaved") block Bt[col * ldx + kb], kb < NBP, so the NBP operands of one non-zero are ONE contiguous 128-bit-packet gather,
// and y the interleaved result Ct[row * ldy + kb]; the matrix stream is shared by all NBP right-hand sides.  The host only
// launches it when every tile fits its stage and no row is long enough for the warp / CTA paths.
// LPC ("lane per column", with NBP > 0): x and y are used where they lie — operand of non-zero (i, c) for right-hand side kb is
// x[c * ldx + kb * sxc], result y[i + kb * ldy] = alpha sum + beta y — and NBP adjacent lanes share one row, lane kb owning right-hand
// side kb: the NBP operands of a non-zero are one warp-level request of NBP adjacent lanes, adjacent rows (adjacent lane groups) gather
// adjacent columns, so one request touches 2-3 lines instead of one line per lane (the packet-per-lane form above is bound by the
// L1TEX tag stage: ncu 88 % l1tex, 44 % DRAM on the 27-point matrix), no shuffle reduction, and every row is summed left to right
// exactly as the reference does.  A tile of ROWS = THREADS / TPR rows is walked by THREADS / NBP groups, NBP / TPR rows each; `xpf`
// carries the number of live right-hand sides of this block (lanes beyond it compute on column 0 and store nothing).
// VS ("virtual split", heavy-tailed row lengths: hb_csr_create builds the tables once): `pntr` is then the row-pointer array of a VIRTUAL
// row set — rows longer than a segment length are cut into consecutive segments, indx / vals are the caller's arrays as they lie — and
// tiles come from a table: tile t = virtual rows [vs.trow[t], vs.trow[t+1]) (at most ROWS of them, starts are multiples of 4) whose
// non-zeros [vs.tnz[t], vs.tnz[t+1]) always fit one ring stage, so the global-memory paths for oversized tiles and the CTA-wide
// reduction of very long rows never run and no CTA is left alone with a 65 536-entry row.  A virtual row that is a whole row stores
// to y[vs.vmap[v]]; a segment stores its raw partial sum to vs.part[~vs.vmap[v]], and vsplit_combine_kernel adds the segments of each
// split row in order afterwards (deterministic; no atomics).
struct vsplit_view { const int *trow, *tnz, *vmap; void *part; int ntiles; };
static constexpr int VS_WARPROW = 96;           // VS: rows of this many entries and more are summed by their whole warp (16 / 24 / 32 / 48 measured: slower)

template<typename T, int THREADS, int TPR, int STAGES, bool DOT, int NBP = 0, bool LPC = false, bool VS = false>
__global__ void __launch_bounds__(THREADS, (VS ? 5 : (THREADS >= 256 ? 3 : 6))) spmv_pipe_kernel(int rows, int nnz, const int * __restrict__ pntr, const int * __restrict__ indx,
                                                            const T * __restrict__ vals, const T * __restrict__ x, T *y,
                                                            scalar_arg<T> alpha_s, scalar_arg<T> beta_s, const int * __restrict__ cta_tiles,
                                                            void *partials_v, unsigned int *ticket, T *dot_out, const int *skip_flag, int xpf, int cols,
                                                            const peer_view *pv, unsigned long long epoch, size_t ldx, size_t ldy, size_t sxc,
                                                            int trot, int twait, vsplit_view vs){
    static_assert(!VS || (NBP == 0 && !LPC), "the virtual-row form is the plain product");
    constexpr int ROWS = THREADS / TPR;
    constexpr int CAP  = THREADS * pipe_slots_of<T, VS>();  // non-zeros a stage can hold (after 4-alignment slack)
    constexpr int PIPE_UNR = pipe_slots_of<T, VS>();
    constexpr int PSL  = ROWS + 4;                          // ints of the pntr slice per stage
    constexpr size_t STAGE_BYTES = (size_t) (CAP + 4) * (sizeof(T) + sizeof(int)) + PSL * sizeof(int);
    static_assert(STAGE_BYTES % 16 == 0, "stage must stay 16-byte aligned");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[STAGES];
    constexpr int BND = 64;
    __shared__ int      bnd[BND], bnd1[BND];
    __shared__ int      brw[VS ? BND : 1], brw1[VS ? BND : 1];  // VS: first virtual row of the tile (bit 0: the tile holds long rows) / of the next tile
    __shared__ int      stage_a0[STAGES];                   // first staged non-zero index of the tile, or -1: not staged (slow path)
    __shared__ T        red[32];

    if (skip_flag && *skip_flag) return;
    // Row-partitioned run over peer memory: the ghost entries of x are being stored by the neighbours' direction kernels.  Only
    // rows that reference ghost columns depend on them, so the sweep is rotated by `trot` tiles (round-robin map) to start
    // behind the leading block of such rows and each CTA waits for the neighbours' flags right before its first tile at or beyond
    // virtual position `twait` (the first tile of the rotated order that touches a ghost): interior rows run while the halo is
    // still in flight, CTAs that own no boundary tile never wait.  trot = twait = 0 is a wait before the first gather.
    bool halo_pending = DOT && (pv != nullptr);
    [[maybe_unused]] const int vs_warprow = (xpf >> 8) > 0 ? (xpf >> 8) : VS_WARPROW;       // VS only; bits 8.. of xpf: probe override

    const int tid = threadIdx.x, sub = tid % TPR, grp = tid / TPR;
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    // Two tile-to-CTA maps.  Contiguous pieces of equal non-zero count (cta_tiles != null): best when row lengths vary wildly.
    // Round-robin (cta_tiles == null): CTA g takes tiles g, g+G, g+2G, ... so all CTAs sweep the matrix together and the window
    // of x they gather from stays small enough for L2 — on the 512^3 7-point problem the contiguous map re-reads x from DRAM
    // three times (ncu: 16.3 GB of traffic for 13.9 GB algorithmic), the sweep reads it once.
    const int ntiles_all = VS ? vs.ntiles : (rows + ROWS - 1) / ROWS;
    const int tstride = cta_tiles ? 1 : (int) gridDim.x;
    const int tile_begin = cta_tiles ? cta_tiles[blockIdx.x] : (int) blockIdx.x;
    const int ntile = cta_tiles ? (cta_tiles[blockIdx.x + 1] - tile_begin)
                                : ((int) blockIdx.x < ntiles_all ? (ntiles_all - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0);
    const int nnz4 = nnz & ~3, np4 = (rows + 1) & ~3;
    T dot_acc = zero_of<T>();

    auto stage_vals = [&](int s){ return reinterpret_cast<T*>(smem_raw + (size_t) s * STAGE_BYTES); };
    auto stage_cols = [&](int s){ return reinterpret_cast<int*>(smem_raw + (size_t) s * STAGE_BYTES + (size_t) (CAP + 4) * sizeof(T)); };
    auto stage_ptr  = [&](int s){ return reinterpret_cast<int*>(smem_raw + (size_t) s * STAGE_BYTES + (size_t) (CAP + 4) * (sizeof(T) + sizeof(int))); };

    if (ntile > 0){
        if (tid == 0){
            for (int s = 0; s < STAGES; s++) mbar_init(full + s, 1);
            fence_mbar_init();
        }
        __syncthreads();

        // Tile bounds ring: bnd[j % BND] = pntr[first row of local tile j] (j == ntile: end of the piece).  The producer needs
        // two of them per issue; a dependent global load there would put a DRAM round trip on every step's critical path
        // (all warps meet at the per-step barrier), so the ring is refilled 32 entries at a time with cp.async by the last
        // warp, a whole 16 steps before the data is waited for and 30 before it is used.
        // tile of step j; only the DOT form is ever rotated (the peer hooks ride on the fused SpMV + dot), the plain product keeps the
        // index arithmetic it had
        auto phys = [&](int j){
            int t = tile_begin + j * tstride;
            if constexpr (DOT && !VS){ t += trot; if (t >= ntiles_all) t -= ntiles_all; }
            return t;
        };
        auto bound_src  = [&](int j){ return VS ? vs.tnz + phys(j) : pntr + min((long long) phys(j) * ROWS, (long long) rows); };
        auto bound_src1 = [&](int j){ return VS ? vs.tnz + phys(j) + 1 : pntr + min(((long long) phys(j) + 1) * ROWS, (long long) rows); };
        for (int j = tid; j < BND && j < ntile; j += THREADS){
            bnd[j] = __ldg(bound_src(j)); bnd1[j] = __ldg(bound_src1(j));
            if constexpr (VS){ brw[j] = __ldg(vs.trow + phys(j)); brw1[j] = __ldg(vs.trow + phys(j) + 1); }
        }
        __syncthreads();
        auto tile_r0 = [&](int k){ if constexpr (VS) return brw[k % BND] & ~3; else return phys(k) * ROWS; };
        auto issue = [&](int k, int b0, int b1){             // thread 0 only; b0,b1 = non-zero bounds of tile k
            const int s = k % STAGES;
            const int r0 = tile_r0(k);
            T *sv = stage_vals(s); int *sc = stage_cols(s); int *sp = stage_ptr(s);
            // pntr slice [r0, r0 + PSL) clipped to the array; ragged end by hand
            const int pend = min(r0 + PSL, rows + 1);
            const int pbulk_end = min(pend & ~3, np4);
            const int npb = max(pbulk_end - r0, 0);
            for (int e = r0 + npb; e < pend; e++) sp[e - r0] = pntr[e];
            const int a0 = b0 & ~3;
            const bool fits = (b1 - a0) <= CAP;
            uint32_t bytes = (uint32_t) (npb * sizeof(int));
            int nb = 0;
            if (fits){
                const int cend4 = (b1 + 3) & ~3;
                const int bulk_end = min(cend4, nnz4);
                nb = max(bulk_end - a0, 0);
                for (int e = a0 + nb; e < min(cend4, nnz); e++){ sv[e - a0] = vals[e]; sc[e - a0] = indx[e]; }
                bytes += (uint32_t) (nb * (sizeof(T) + sizeof(int)));
            }
            stage_a0[s] = fits ? a0 : -1;
            fence_proxy_async();
            mbar_arrive_expect_tx(full + s, bytes);
            if (npb > 0) bulk_g2s(sp, pntr + r0, (uint32_t) (npb * sizeof(int)), full + s);
            if (nb > 0){
                bulk_g2s(sv, vals + a0, (uint32_t) (nb * sizeof(T)), full + s);
                bulk_g2s(sc, indx + a0, (uint32_t) (nb * sizeof(int)), full + s);
            }
        };
        if (tid == 0) for (int k = 0; k < STAGES - 1 && k < ntile; k++) issue(k, bnd[k % BND], bnd1[k % BND]);

        // VS (heavy-tailed rows): the gathers are dealt by NON-ZERO, not by row, and run one tile ahead of the sums.
        //   gathers(k):  thread t takes entries t, t + THREADS, ... of tile k (at most pipe_slots of them): column indices from the
        //                stage, then all operands of x in flight together — every lane of every warp carries the same number of
        //                gathers whatever the row lengths;
        //   products(k): when they are back, the staged values are overwritten by value * operand;            -- barrier --
        //   gathers(k+1) are issued NOW, and stay in flight while
        //   sums(k):     the rows of tile k are summed from the products in shared memory (TPR lanes per row in strides; rows of
        //                PIPE_WARPROW entries and more by their warp), y is written;                           -- barrier --
        // With rows walked by their own lanes (the general form) the longest row of a tile sets the pace: ~6 dependent gather round
        // trips per tile on the power-law matrix (626 us; cuSPARSE's merge-based csrmv_v3: 275 us).
        [[maybe_unused]] int vs_c[PIPE_UNR];
        [[maybe_unused]] T vs_x[PIPE_UNR];
        [[maybe_unused]] int vs_off = 0, vs_nz = 0;
        [[maybe_unused]] auto vs_gathers = [&](int kk){
            const int ss = kk % STAGES;
            mbar_wait(full + ss, (uint32_t) ((kk / STAGES) & 1));
            const int *spn = stage_ptr(ss), *scn = stage_cols(ss);
            const int r0n = tile_r0(kk), nr = min((brw1[kk % BND] & ~3) - r0n, rows - r0n);
            vs_off = spn[0] - stage_a0[ss];
            vs_nz = spn[nr] - spn[0];
            #pragma unroll
            for (int u = 0; u < PIPE_UNR; u++) if (tid + u * THREADS < vs_nz) vs_c[u] = scn[vs_off + tid + u * THREADS];
            #pragma unroll
            for (int u = 0; u < PIPE_UNR; u++) if (tid + u * THREADS < vs_nz) vs_x[u] = (xpf & 2) ? one_of<T>() : ld_ro(x + vs_c[u]);   // xpf: probe bits (HB_VS_PROBE)
        };
        if constexpr (VS){
            __syncthreads();                                    // stage_a0 of the first tiles, written by thread 0 above
            vs_gathers(0);
        }

        for (int k = 0; k < ntile; k++){
            const int s = k % STAGES;
            if (tid == 0 && k + STAGES - 1 < ntile) issue(k + STAGES - 1, bnd[(k + STAGES - 1) % BND], bnd1[(k + STAGES - 1) % BND]);
            if (tid >= THREADS - 32){                           // ring refill by the last warp (see above)
                if ((k & 31) == 0 && k > 0){
                    const int j = k + 32 + (tid & 31);
                    if (j < ntile){
                        cp_async4(&bnd[j % BND], bound_src(j)); cp_async4(&bnd1[j % BND], bound_src1(j));
                        if constexpr (VS){ cp_async4(&brw[j % BND], vs.trow + phys(j)); cp_async4(&brw1[j % BND], vs.trow + phys(j) + 1); }
                    }
                    cp_async_commit();
                }else if ((k & 31) == 16){
                    cp_async_wait_all();                        // published to thread 0 by this step's closing barrier
                }
            }
            if (DOT && halo_pending && tile_begin + k * tstride >= twait){ // block-uniform
                if (tid < 32) peer_halo_wait(pv, epoch);
                __syncthreads();
                halo_pending = false;
            }
            if constexpr (VS){
                T *sv = stage_vals(s);
                #pragma unroll
                for (int u = 0; u < PIPE_UNR; u++)              // products(k): waits for the gathers issued one tile ago
                    if (tid + u * THREADS < vs_nz) sv[vs_off + tid + u * THREADS] = hmul(sv[vs_off + tid + u * THREADS], vs_x[u]);
                fence_proxy_async();                            // generic-proxy writes into a stage that a bulk copy refills later
                __syncthreads();
                const int r0 = tile_r0(k), myrow = r0 + grp;
                const int *sp = stage_ptr(s);
                const int a0 = stage_a0[s];
                const bool live_row = grp < (brw1[k % BND] & ~3) - r0 && myrow < rows;
                int rs = 0, re = 0, vdst = 0;
                if (live_row){ rs = sp[grp]; re = sp[grp + 1]; if (sub == 0) vdst = __ldg(vs.vmap + myrow); }
                if (k + 1 < ntile) vs_gathers(k + 1);           // in flight during the sums below and the barrier
                T sum = zero_of<T>();
                const int end = re - a0;
                const bool wlong = (re - rs) >= vs_warprow;     // summing products is a chain of LDS + add: a warp takes over early
                if (!wlong && !(xpf & 1)){
                    T sum2 = zero_of<T>();                      // (four reads per step before the first add: measured 2 % slower)
                    int j = rs + sub - a0;
                    for (; j + TPR < end; j += 2 * TPR){ sum = hadd(sum, sv[j]); sum2 = hadd(sum2, sv[j + TPR]); }
                    if (j < end) sum = hadd(sum, sv[j]);
                    sum = hadd(sum, sum2);
                }
                unsigned todo = __ballot_sync(0xffffffffu, wlong && sub == 0 && !(xpf & 1));
                while (todo){                                   // long rows: by their warp, from the products (no gathers here)
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ls = __shfl_sync(0xffffffffu, rs, src) - a0, le = __shfl_sync(0xffffffffu, re, src) - a0;
                    T part = zero_of<T>(), part2 = zero_of<T>();
                    int j = ls + (tid & 31);
                    for (; j + 32 < le; j += 64){ part = hadd(part, sv[j]); part2 = hadd(part2, sv[j + 32]); }
                    if (j < le) part = hadd(part, sv[j]);
                    part = shfl_from(warp_sum(hadd(part, part2)), 0);
                    if ((tid & 31) == src) sum = part;
                }
                #pragma unroll
                for (int d = TPR / 2; d > 0; d >>= 1) sum = hadd(sum, shfl_down(sum, d));
                if (sub == 0 && live_row){
                    if (vdst < 0){
                        reinterpret_cast<T*>(vs.part)[~vdst] = sum;    // a segment: raw partial sum, combined afterwards
                    }else if (DOT){
                        y[vdst] = sum;
                        dot_acc = hfma(hconj(ld_ro(x + vdst)), sum, dot_acc);
                    }else{
                        T out = hmul(alpha, sum);
                        if (use_beta) out = hfma(beta, y[vdst], out);
                        y[vdst] = out;
                    }
                }
                __syncthreads();                                // stage s may be refilled
                continue;
            }
            mbar_wait(full + s, (uint32_t) ((k / STAGES) & 1));
            const int r0 = tile_r0(k);
            const int myrow = r0 + grp;
            const int *sp = stage_ptr(s);
            const int a0 = stage_a0[s];
            int rs = 0, re = 0;
            const bool live_row = myrow < rows;
            if (live_row){ rs = sp[grp]; re = sp[grp + 1]; }
            T sum = zero_of<T>();
            if constexpr (NBP > 0 && LPC){
                constexpr int GROUPS = THREADS / NBP, RPG = ROWS / GROUPS, MU = (sizeof(T) == 16 ? 4 : 8);
                static_assert(NBP >= TPR && ROWS % GROUPS == 0 && RPG >= 1, "a tile must be whole passes of the lane groups");
                const int kb = tid % NBP, g = tid / NBP;
                const bool live = kb < xpf;
                const T *sv = stage_vals(s); const int *sc = stage_cols(s);
                const T *xk = x + (live ? (size_t) kb * sxc : (size_t) 0);
                #pragma unroll 1
                for (int rr = 0; rr < RPG; rr++){
                    const int lr = g + rr * GROUPS, row = r0 + lr;
                    if (row >= rows) break;
                    const int ls = sp[lr] - a0, le = sp[lr + 1] - a0;
                    T acc = zero_of<T>();
                    for (int base = ls; base < le; base += MU){
                        int c[MU]; T v[MU], xv[MU];
                        #pragma unroll
                        for (int u = 0; u < MU; u++) if (base + u < le){ c[u] = sc[base + u]; v[u] = sv[base + u]; }
                        #pragma unroll
                        for (int u = 0; u < MU; u++) if (base + u < le) xv[u] = ld_ro(xk + (size_t) c[u] * ldx);
                        #pragma unroll
                        for (int u = 0; u < MU; u++) if (base + u < le) acc = hfma(v[u], xv[u], acc);
                    }
                    if (live){
                        T *cp = y + (size_t) row + (size_t) kb * ldy;
                        T out = hmul(alpha, acc);
                        if (use_beta) out = hfma(beta, *cp, out);
                        *cp = out;
                    }
                }
                __syncthreads();
                continue;
            }
            if constexpr (NBP > 0 && !LPC){
                constexpr int NV = vec16<T>::N, NPK = NBP / NV, MU = (NPK >= 4 ? 2 : 4);    // packets per entry; entries per lane in flight
                static_assert(NBP % NV == 0, "interleaved block must be whole 128-bit packets");
                T sums[NBP];
                #pragma unroll
                for (int e = 0; e < NBP; e++) sums[e] = zero_of<T>();
                const T *sv = stage_vals(s); const int *sc = stage_cols(s);
                const int end = re - a0;
                for (int base = rs + sub - a0; base < end; base += MU * TPR){
                    T v[MU]; vec16<T> pk[MU][NPK]; bool ok[MU];
                    #pragma unroll
                    for (int u = 0; u < MU; u++){
                        ok[u] = (base + u * TPR) < end;
                        if (ok[u]){
                            v[u] = sv[base + u * TPR];
                            const vec16<T> *bp = reinterpret_cast<const vec16<T>*>(x + (size_t) sc[base + u * TPR] * ldx);
                            #pragma unroll
                            for (int q = 0; q < NPK; q++) pk[u][q] = bp[q];
                        }
                    }
                    #pragma unroll
                    for (int u = 0; u < MU; u++) if (ok[u]){
                        #pragma unroll
                        for (int q = 0; q < NPK; q++){
                            #pragma unroll
                            for (int e = 0; e < NV; e++) sums[q * NV + e] = hfma(v[u], pk[u][q].v[e], sums[q * NV + e]);
                        }
                    }
                }
                #pragma unroll
                for (int e = 0; e < NBP; e++){
                    #pragma unroll
                    for (int d = TPR / 2; d > 0; d >>= 1) sums[e] = hadd(sums[e], shfl_down(sums[e], d));
                }
                if (sub == 0 && myrow < rows){
                    vec16<T> *cp = reinterpret_cast<vec16<T>*>(y + (size_t) myrow * ldy);
                    #pragma unroll
                    for (int q = 0; q < NPK; q++){
                        vec16<T> o;
                        #pragma unroll
                        for (int e = 0; e < NV; e++) o.v[e] = sums[q * NV + e];
                        cp[q] = o;
                    }
                }
                __syncthreads();
                continue;
            }
            if (a0 >= 0){
                const T *sv = stage_vals(s); const int *sc = stage_cols(s);
                // batches of PIPE_UNR entries per lane: all shared-memory reads, then all gathers, then the FMAs (two
                // accumulators), so one row of up to PIPE_UNR*TPR entries costs a single gather round trip
                const int end = re - a0;
                T sum2 = zero_of<T>();
                const bool wlong = (re - rs) >= PIPE_WARPROW;      // long rows: by the whole warp below, not by TPR lanes
                if (!wlong)
                for (int base = rs + sub - a0; base < end; base += PIPE_UNR * TPR){
                    int c[PIPE_UNR]; T v[PIPE_UNR], xv[PIPE_UNR]; bool ok[PIPE_UNR];
                    #pragma unroll
                    for (int u = 0; u < PIPE_UNR; u++){
                        ok[u] = (base + u * TPR) < end;
                        if (ok[u]){ c[u] = sc[base + u * TPR]; v[u] = sv[base + u * TPR]; }
                    }
                    #pragma unroll
                    for (int u = 0; u < PIPE_UNR; u++) if (ok[u]) xv[u] = ld_ro(x + c[u]);
                    // banded matrices: the tile `xpf` tiles ahead gathers the same columns shifted by xpf*ROWS rows — pull those
                    // lines into L1 now (speculative; a prefetch has no destination register and cannot stall; +4 % on the
                    // 7-point Laplacian, neutral elsewhere)
                    if (xpf > 0){
                        #pragma unroll
                        for (int u = 0; u < PIPE_UNR; u++) if (ok[u]){
                            const int cp = min(c[u] + xpf * ROWS, cols - 1);
                            asm volatile("prefetch.global.L1 [%0];" :: "l"(x + cp));
                        }
                    }
                    #pragma unroll
                    for (int u = 0; u < PIPE_UNR; u += 2){
                        if (ok[u]) sum = hfma(v[u], xv[u], sum);
                        if (ok[u + 1]) sum2 = hfma(v[u + 1], xv[u + 1], sum2);
                    }
                }
                sum = hadd(sum, sum2);
                // warp-cooperative pass over this warp's long rows (heavy-tailed row lengths): 32 lanes stride over the row in
                // shared memory, four entries per lane in flight, shuffle reduction; the owner lane (sub == 0) keeps the sum
                unsigned todo = __ballot_sync(0xffffffffu, wlong && sub == 0);
                while (todo){
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ls = __shfl_sync(0xffffffffu, rs, src) - a0, le = __shfl_sync(0xffffffffu, re, src) - a0;
                    T part = zero_of<T>(), part2 = zero_of<T>();
                    int j = ls + (tid & 31);
                    for (; j + 96 < le; j += 128){
                        const int c0 = sc[j], c1 = sc[j + 32], c2 = sc[j + 64], c3 = sc[j + 96];
                        const T x0 = ld_ro(x + c0), x1 = ld_ro(x + c1), x2 = ld_ro(x + c2), x3 = ld_ro(x + c3);
                        part = hfma(sv[j], x0, part); part2 = hfma(sv[j + 32], x1, part2);
                        part = hfma(sv[j + 64], x2, part); part2 = hfma(sv[j + 96], x3, part2);
                    }
                    for (; j < le; j += 32) part = hfma(sv[j], ld_ro(x + sc[j]), part);
                    part = shfl_from(warp_sum(hadd(part, part2)), 0);      // warp_sum leaves the total in lane 0
                    if ((tid & 31) == src) sum = part;
                }
            }else{
                // slice larger than a stage: straight from global memory — short rows by their TPR lanes, long rows by the whole
                // warp, very long rows (>= PIPE_LONGROW) by the whole CTA
                const bool wlong = (re - rs) >= PIPE_WARPROW;
                if (!wlong)
                    for (int j = rs + sub; j < re; j += TPR) sum = hfma(ld_stream(vals + j), ld_ro(x + __ldcs(indx + j)), sum);
                unsigned todo = __ballot_sync(0xffffffffu, wlong && sub == 0 && (re - rs) < PIPE_LONGROW);
                while (todo){
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ls = __shfl_sync(0xffffffffu, rs, src), le = __shfl_sync(0xffffffffu, re, src);
                    T part = zero_of<T>(), part2 = zero_of<T>();
                    int j = ls + (tid & 31);
                    for (; j + 96 < le; j += 128){
                        const int c0 = __ldcs(indx + j), c1 = __ldcs(indx + j + 32), c2 = __ldcs(indx + j + 64), c3 = __ldcs(indx + j + 96);
                        const T v0 = ld_stream(vals + j), v1 = ld_stream(vals + j + 32), v2 = ld_stream(vals + j + 64), v3 = ld_stream(vals + j + 96);
                        const T x0 = ld_ro(x + c0), x1 = ld_ro(x + c1), x2 = ld_ro(x + c2), x3 = ld_ro(x + c3);
                        part = hfma(v0, x0, part); part2 = hfma(v1, x1, part2); part = hfma(v2, x2, part); part2 = hfma(v3, x3, part2);
                    }
                    for (; j < le; j += 32) part = hfma(ld_stream(vals + j), ld_ro(x + __ldcs(indx + j)), part);
                    part = shfl_from(warp_sum(hadd(part, part2)), 0);      // warp_sum leaves the total in lane 0
                    if ((tid & 31) == src) sum = part;
                }
                const int rlast = min(r0 + ROWS, rows);         // (VS never gets here: every tile of the table fits its stage)
                for (int r = r0; r < rlast; r++){
                    const int ls = sp[r - r0], le = sp[r - r0 + 1];
                    if (le - ls < PIPE_LONGROW) continue;       // block-uniform: sp is shared
                    T part = zero_of<T>();
                    for (int j = ls + tid; j < le; j += THREADS) part = hfma(ld_stream(vals + j), ld_ro(x + __ldcs(indx + j)), part);
                    part = block_sum(part, red);                // contains __syncthreads; result in thread 0
                    if (tid == 0) red[0] = part;
                    __syncthreads();
                    if (myrow == r && sub == 0) sum = red[0];
                    __syncthreads();
                }
            }
            #pragma unroll
            for (int d = TPR / 2; d > 0; d >>= 1) sum = hadd(sum, shfl_down(sum, d));
            if (sub == 0 && myrow < rows){
                if (DOT){
                    y[myrow] = sum;
                    dot_acc = hfma(hconj(ld_ro(x + myrow)), sum, dot_acc);     // (asking for x[row] earlier, with the gathers, was measured: 3 % slower)
                }else{
                    T out = hmul(alpha, sum);
                    if (use_beta) out = hfma(beta, y[myrow], out);
                    y[myrow] = out;
                }
            }
            __syncthreads();                                    // stage s may be refilled
        }
    }
    if (DOT){
        T *partials = reinterpret_cast<T*>(partials_v);
        T b = block_sum(dot_acc, red);
        if (tid == 0) partials[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            T total = sum_partials<T>(partials, gridDim.x, 1, red);
            if (tid == 0){ *dot_out = total; red[0] = total; }
            if (pv){                                            // hand this rank's partial to every peer (hb_peer.cuh)
                __syncthreads();
                // a halo wait of this launch timed out: the product was formed from stale ghosts — hand out NaN, which stops
                // every rank at the next stop test instead of letting the iteration run on with wrong data
                const bool poisoned = *reinterpret_cast<volatile int*>(&pv->mail[pv->rank]->error) != 0;
                T out = red[0];
                if (poisoned) out = from_real<T>((real_t<T>) __longlong_as_double(0x7ff8000000000000LL));
                if (tid < 32) peer_publish<T>(pv, HB_PEER_CH_PAP, epoch, out);
            }
        }
    }
}

// VS tail: y[row] = alpha * (sum of the row's segment partials, in order) + beta * y[row] for the split rows; with DOT, y[row] = sum and
// *dot_out += sum over split rows of conj(x[row]) y[row] (added by the last block, after the streaming kernel wrote its own total).
// One warp per split row: lanes stride over the row's parts, shuffle reduction — the same order every time.
template<typename T, bool DOT>
__global__ void __launch_bounds__(256) vsplit_combine_kernel(int nsplit, const int * __restrict__ srow, const int * __restrict__ spart, const T * __restrict__ part,
                                                            scalar_arg<T> alpha_s, scalar_arg<T> beta_s, T *y, const T * __restrict__ x,
                                                            void *partials_v, unsigned int *ticket, T *dot_out, const int *skip_flag){
    __shared__ T red[32];
    if (skip_flag && *skip_flag) return;
    const T alpha = get_scalar(alpha_s), beta = get_scalar(beta_s);
    const bool use_beta = !hiszero(beta);
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    T dot_acc = zero_of<T>();
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nsplit; i += gridDim.x * wpb){
        const int p0 = spart[i], p1 = spart[i + 1], row = srow[i];
        T acc = zero_of<T>();
        for (int j = p0 + lane; j < p1; j += 32) acc = hadd(acc, part[j]);
        acc = warp_sum(acc);
        if (lane == 0){
            if (DOT){
                y[row] = acc;
                dot_acc = hfma(hconj(ld_ro(x + row)), acc, dot_acc);
            }else{
                T out = hmul(alpha, acc);
                if (use_beta) out = hfma(beta, y[row], out);
                y[row] = out;
            }
        }
    }
    if (DOT){
        T *partials = reinterpret_cast<T*>(partials_v);
        T b = block_sum(dot_acc, red);
        if (threadIdx.x == 0) partials[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            T total = sum_partials<T>(partials, gridDim.x, 1, red);
            if (threadIdx.x == 0) *dot_out = hadd(*dot_out, total);
        }
    }
}
