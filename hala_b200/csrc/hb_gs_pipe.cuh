// hb_gs_pipe.cuh — the Gram-Schmidt pair of GMRES as streaming kernels: h = W^H r (multi-dot) and r -= W h (+ ||r||^2).
//
// Both read a tall-skinny column-major panel W (rows x k, k <= 64) once.  Register-staged versions (hb_krylov.cu) stall at
// 2.4-3.3 TB/s on B200: every thread has to hold 16+ column packets in flight and the 112-128 registers that takes leave 16
// warps per SM, all of them waiting (ncu: no eligible warp 90 % of the cycles, DRAM 30-40 %).  Here the panel never passes
// through registers on its way in: one thread per CTA streams tiles of GS_R rows x KCH columns (one contiguous slice per
// column, plus the slice of r) into a GS_STAGES-deep shared-memory ring with cp.async.bulk completing on an mbarrier
// (SASS UBLKCP / SYNCS) — up to ~100 KB per CTA in flight with no register cost — and 256 consumer threads each own one row
// of the tile and read conflict-free columns from shared memory.
//   multi-dot : items ordered chunk-major (all tiles of column chunk 0, then chunk 1, ...): KCH running sums per thread,
//               reduced once per chunk (shuffle -> shared -> partials[block][k]); r is re-streamed per chunk (+1/KCH traffic).
//   multi-axpy: items ordered tile-major (all chunks of tile 0, then tile 1, ...): the row value stays in a register across
//               the chunks of its tile, r is read and written once.
// Needs 16-byte aligned W, r and column stride; anything else takes the register kernels.
#pragma once
#include "hb_common.cuh"
#include "hb_async.cuh"

static constexpr int GS_R = 256;            // rows per tile == threads per CTA
static constexpr int GS_STAGES = 3;
template<typename T> __host__ __device__ constexpr int gs_kch(){ return sizeof(T) == 16 ? 8 : 16; }
template<typename T> __host__ __device__ constexpr size_t gs_stage_bytes(){ return (size_t) (gs_kch<T>() + 1) * GS_R * sizeof(T); }
template<typename T> __host__ __device__ constexpr size_t gs_smem_bytes(){ return GS_STAGES * gs_stage_bytes<T>(); }

// MODE 0: multi-dot (h_out[c] = sum_i op(W[i,c]) r[i]);  MODE 1: multi-axpy (r += W (scale*h), optional sum |r|^2)
template<typename T, int MODE, bool CONJ>
__global__ void __launch_bounds__(GS_R, 2) gs_pipe_kernel(long long rows, long long rows_al, int k, const T * __restrict__ W, size_t ldw, const T *r_in, T *r_out,
                                                           const T *h_dev, T scale_h, void *partials_v, unsigned int *ticket, T *out, const int *skip_flag){
    constexpr int KCH = gs_kch<T>();
    constexpr size_t STAGE = gs_stage_bytes<T>();
    extern __shared__ __align__(128) unsigned char gs_smem[];
    __shared__ uint64_t full[GS_STAGES];
    __shared__ T wsum[GS_R / 32][KCH];
    __shared__ T hs[MODE == 1 ? 64 : 1];
    __shared__ double red[32];
    if (skip_flag && *skip_flag) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ntiles = (rows_al + GS_R - 1) / GS_R;
    const int nchunk = (k + KCH - 1) / KCH;
    // this CTA's tiles: b, b + G, b + 2G, ... (the whole grid sweeps the panel together)
    const long long my_tiles = (blockIdx.x < ntiles) ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const long long nitems = my_tiles * nchunk;
    if (MODE == 1){
        for (int c = tid; c < k; c += GS_R) hs[c] = hmul(scale_h, h_dev[c]);
    }
    if (tid == 0){
        for (int s = 0; s < GS_STAGES; s++) mbar_init(full + s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto item_tile  = [&](long long it){ return blockIdx.x + (MODE == 0 ? (it % my_tiles) : (it / nchunk)) * (long long) gridDim.x; };
    auto item_chunk = [&](long long it){ return (int) (MODE == 0 ? (it / my_tiles) : (it % nchunk)); };
    auto stage_w = [&](int s){ return reinterpret_cast<T*>(gs_smem + (size_t) s * STAGE); };                    // [KCH][GS_R]
    auto stage_r = [&](int s){ return reinterpret_cast<T*>(gs_smem + (size_t) s * STAGE) + (size_t) KCH * GS_R; };
    auto issue = [&](long long it){                             // thread 0 only
        const int s = (int) (it % GS_STAGES);
        const long long row0 = item_tile(it) * GS_R;
        const int chunk = item_chunk(it), c0 = chunk * KCH, kc = min(KCH, k - c0);
        const uint32_t nrow = (uint32_t) min((long long) GS_R, rows_al - row0);
        const uint32_t slice = nrow * (uint32_t) sizeof(T);
        const bool need_r = (MODE == 0) || chunk == 0;
        fence_proxy_async();
        mbar_arrive_expect_tx(full + s, slice * (uint32_t) (kc + (need_r ? 1 : 0)));
        T *sw = stage_w(s);
        for (int j = 0; j < kc; j++) bulk_g2s(sw + (size_t) j * GS_R, W + (size_t) (c0 + j) * ldw + row0, slice, full + s);
        if (need_r) bulk_g2s(stage_r(s), r_in + row0, slice, full + s);
    };
    if (tid == 0) for (long long it = 0; it < GS_STAGES - 1 && it < nitems; it++) issue(it);

    T acc[KCH];
    #pragma unroll
    for (int j = 0; j < KCH; j++) acc[j] = zero_of<T>();
    T t = zero_of<T>();
    double nrm = 0.0;
    T *partials = reinterpret_cast<T*>(partials_v);             // multi-dot: [block][k]
    for (long long it = 0; it < nitems; it++){
        const int s = (int) (it % GS_STAGES);
        if (tid == 0 && it + GS_STAGES - 1 < nitems) issue(it + GS_STAGES - 1);
        mbar_wait(full + s, (uint32_t) ((it / GS_STAGES) & 1));
        const long long row0 = item_tile(it) * GS_R;
        const int chunk = item_chunk(it), c0 = chunk * KCH, kc = min(KCH, k - c0);
        const bool live = row0 + tid < rows_al;
        const T *sw = stage_w(s);
        if (MODE == 0){
            if (live){
                const T rv = stage_r(s)[tid];
                #pragma unroll
                for (int j = 0; j < KCH; j++) if (j < kc){ const T w = sw[(size_t) j * GS_R + tid]; acc[j] = hfma(CONJ ? hconj(w) : w, rv, acc[j]); }
            }
            if (it % my_tiles == my_tiles - 1){                 // last tile of this chunk: hand the KCH sums over
                #pragma unroll
                for (int j = 0; j < KCH; j++) acc[j] = warp_sum(acc[j]);
                if (lane == 0){
                    #pragma unroll
                    for (int j = 0; j < KCH; j++) wsum[warp][j] = acc[j];
                }
                __syncthreads();
                if (tid < kc){
                    T sacc = zero_of<T>();
                    #pragma unroll
                    for (int w = 0; w < GS_R / 32; w++) sacc = hadd(sacc, wsum[w][tid]);
                    partials[(size_t) blockIdx.x * k + c0 + tid] = sacc;
                }
                #pragma unroll
                for (int j = 0; j < KCH; j++) acc[j] = zero_of<T>();
            }
        }else{
            if (live){
                if (chunk == 0) t = stage_r(s)[tid];
                #pragma unroll
                for (int j = 0; j < KCH; j++) if (j < kc) t = hfma(sw[(size_t) j * GS_R + tid], hs[c0 + j], t);
                if (chunk == nchunk - 1){ r_out[row0 + tid] = t; nrm += (double) habs2(t); }
            }
        }
        __syncthreads();                                        // stage s may be refilled
    }
    if (MODE == 0){
        if (my_tiles == 0){                                     // more CTAs than tiles: contribute zeros
            for (int c = tid; c < k; c += GS_R) partials[(size_t) blockIdx.x * k + c] = zero_of<T>();
        }
        if (last_block_arrives(ticket)){
            for (int c = warp; c < k; c += GS_R / 32){
                T a = zero_of<T>();
                for (int b = lane; b < (int) gridDim.x; b += 32) a = hadd(a, ld_cg_T(partials + (size_t) b * k + c));
                a = warp_sum(a);
                if (lane == 0){
                    for (long long i = rows_al; i < rows; i++){ // the < 16-byte tail behind the last whole packet
                        const T w = W[(size_t) c * ldw + i];
                        a = hfma(CONJ ? hconj(w) : w, r_in[i], a);
                    }
                    out[c] = a;
                }
            }
        }
    }else{
        if (blockIdx.x == 0 && rows_al + tid < rows){           // the < 16-byte tail behind the last whole packet
            const long long i = rows_al + tid;
            T tt = r_in[i];
            for (int c = 0; c < k; c++) tt = hfma(W[(size_t) c * ldw + i], hs[c], tt);
            r_out[i] = tt; nrm += (double) habs2(tt);
        }
    }
    if (MODE == 1 && out){
        double *dp = reinterpret_cast<double*>(partials_v);
        double b = block_sum(nrm, red);
        if (tid == 0) dp[blockIdx.x] = b;
        if (last_block_arrives(ticket)){
            double rr = sum_partials<double>(dp, gridDim.x, 1, red);
            if (tid == 0) *out = from_real<T>((real_t<T>) rr);
        }
    }
}
