// hb_spmm_band.cu — multi right-hand-side CSR product for BANDED matrices with the operand blocks staged in shared memory
// (reference: gpu_sparse_matrix::gemm -> cusparseSpMM, gpu/hala_cuda_sparse_general.hpp:284-332).
// EXPERIMENTAL, opt-in (HB_SPMM_PATH=band): first cut of DESIGN.md §9 item 1.
//
// Why: the interleaved streaming kernel (hb_spmv_pipe.cuh, NBP = 4) gathers one 32-byte operand block of B per non-zero through L1 and
// is bound by the L1TEX data stage at about one wavefront per gathered sector (ncu: 87 % busy, DRAM 44 %).  On stencil / banded
// matrices the columns that a tile of consecutive rows touches are a handful of contiguous runs (27-point: 9, 7-point: 5).  Here
//   * a one-time analysis (band_analyse_kernel) lists, per tile, those runs (at most BD_MAXRUNS, at most BD_SEGROWS rows in all) and
//     replaces every column index by a 16-bit offset into the tile's staged rows — the kernel streams values + 2-byte offsets
//     (10 B per non-zero in fp64 instead of 12) and never sees a column index;
//   * the producer thread of a persistent CTA bulk-copies (cp.async.bulk, mbarrier) the tile's slice of values and offsets AND its
//     runs of the interleaved operand Bt into a ring stage; consumers gather their operands from shared memory with 128-bit loads.
// Matrices whose tiles do not reduce to such runs are refused (the caller falls back to the interleaved kernel).
#include "hb_common.cuh"
#include "hb_async.cuh"
#include <algorithm>
#include <climits>
#include <cstdlib>

static constexpr int BD_THREADS = 256, BD_STAGES = 2, BD_NBP = 4, BD_MAXRUNS = 16, BD_GAP = 4, BD_SEGROWS = 672;
// tile descriptor, BD_DESC ints: [0] first non-zero, [1] one past the last, [2] runs, [3] staged rows, [4 + q] first column of run q,
// [20 + q] offset (in rows) of run q inside the staged block, [20 + runs] = staged rows
static constexpr int BD_DESC = 40;
template<typename T> __host__ __device__ constexpr int bd_cap(){ return sizeof(T) == 16 ? 1024 : 2048; }     // non-zeros per stage

struct hb_band {
    int state = 0;                    // 0 not analysed, 1 usable, -1 not banded (or out of memory)
    int tpr = 4, tile_rows = 64, ntiles = 0;
    int *desc = nullptr;
    unsigned short *soff = nullptr;
};
void hb_band_delete(hb_band *b){
    if (!b) return;
    if (b->desc) cudaFree(b->desc);
    if (b->soff) cudaFree(b->soff);
    delete b;
}

// ------------------------------------------------------------------------------------------------ analysis (once per matrix)
template<int CAP>
__global__ void __launch_bounds__(BD_THREADS) band_analyse_kernel(int rows, const int * __restrict__ pntr, const int * __restrict__ indx, int tile_rows, int ntiles,
                                                                  int *desc, unsigned short *soff, int *unfit){
    __shared__ int sorted[CAP];
    __shared__ int run_start[BD_MAXRUNS], run_off[BD_MAXRUNS + 1];
    __shared__ int nruns_s;
    const int tid = threadIdx.x;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x){
        const long long r0 = (long long) t * tile_rows, r1 = (r0 + tile_rows < rows) ? r0 + tile_rows : rows;
        const int b0 = pntr[r0], b1 = pntr[r1], n = b1 - b0;
        __syncthreads();
        if (n > CAP - 16){ if (tid == 0) atomicExch(unfit, 1); continue; }       // block-uniform
        for (int i = tid; i < CAP; i += BD_THREADS) sorted[i] = i < n ? indx[b0 + i] : INT_MAX;
        __syncthreads();
        for (int k = 2; k <= CAP; k <<= 1){
            for (int j = k >> 1; j > 0; j >>= 1){
                for (int i = tid; i < CAP; i += BD_THREADS){
                    const int l = i ^ j;
                    if (l > i){
                        const bool up = (i & k) == 0;
                        const int a = sorted[i], b = sorted[l];
                        if ((a > b) == up){ sorted[i] = b; sorted[l] = a; }
                    }
                }
                __syncthreads();
            }
        }
        if (tid == 0){
            int nr = 0, total = 0;
            bool ok = true;
            if (n > 0){
                int s = sorted[0], prev = s;
                for (int i = 1; i <= n && ok; i++){
                    const int c = (i < n) ? sorted[i] : INT_MAX;
                    if (i == n || c - prev > BD_GAP){                            // close the run [s, prev]
                        if (nr == BD_MAXRUNS){ ok = false; break; }
                        run_start[nr] = s; run_off[nr] = total; total += prev - s + 1; nr++;
                        s = c;
                    }
                    prev = c;
                }
            }
            if (!ok || total > BD_SEGROWS){ atomicExch(unfit, 1); nr = -1; }
            else run_off[nr] = total;
            nruns_s = nr;
        }
        __syncthreads();
        const int nr = nruns_s;
        if (nr < 0) continue;                                                    // block-uniform
        for (int i = tid; i < n; i += BD_THREADS){
            const int c = indx[b0 + i];
            int r = 0;
            for (int q = 1; q < nr; q++) if (run_start[q] <= c) r = q;
            soff[b0 + i] = (unsigned short) (run_off[r] + (c - run_start[r]));
        }
        int *d = desc + (size_t) t * BD_DESC;
        if (tid == 0){ d[0] = b0; d[1] = b1; d[2] = nr; d[3] = run_off[nr]; }
        if (tid < BD_MAXRUNS) d[4 + tid] = tid < nr ? run_start[tid] : 0;
        if (tid <= BD_MAXRUNS) d[20 + tid] = tid <= nr ? run_off[tid] : run_off[nr];
    }
}

// ------------------------------------------------------------------------------------------------ the product
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity){
    const long long t0 = clock64();
    for (;;){
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
        if (clock64() - t0 > 2000000000LL) return false;                         // ~1 s: a copy that never lands must not hang the device
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar){
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// Ct[row * NBP + kb] = sum_j a_(row, j) Bt[col_j * NBP + kb]; Bt / Ct interleaved blocks of BD_NBP right-hand sides.
// BD_THREADS consumer threads (TPR lanes per row) + one producer warp.  Per ring stage a `full` barrier (the producer's arrive + the
// bytes of its bulk copies) and an `empty` barrier (one arrive per consumer warp): no block-wide barrier in the loop, so a consumer
// warp that is done with a tile moves on to the next one while the others finish, and the producer refills a stage as soon as its
// last reader has left it.  The producer warp reads the tile descriptors itself (32 + 8 ints, one or two per lane, handed to lane 0
// by shuffles), one tile ahead of the copies it issues.
template<typename T, int TPR>
__global__ void __launch_bounds__(BD_THREADS + 32, 2) spmm_band_kernel(int rows, int nnz, const int * __restrict__ pntr, const T * __restrict__ vals,
                                                                       const unsigned short * __restrict__ soff, const int * __restrict__ desc, int ntiles,
                                                                       const T * __restrict__ Bt, T *Ct, int *err){
    constexpr int ROWS = BD_THREADS / TPR, CAP = bd_cap<T>(), NV = vec16<T>::N, NPK = BD_NBP / NV, MU = (sizeof(T) == 16 ? 2 : 4);
    constexpr size_t VAL_BYTES = (size_t) CAP * sizeof(T), OFF_BYTES = (size_t) CAP * 2, SEG_BYTES = (size_t) BD_SEGROWS * BD_NBP * sizeof(T);
    constexpr size_t STAGE_BYTES = VAL_BYTES + OFF_BYTES + SEG_BYTES;
    static_assert(VAL_BYTES % 16 == 0 && OFF_BYTES % 16 == 0 && SEG_BYTES % 16 == 0, "stage parts must stay 16-byte aligned");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[BD_STAGES], empty[BD_STAGES];
    __shared__ int      stage_a0[BD_STAGES];
    const int tid = threadIdx.x;
    const int my_ntile = (int) blockIdx.x < ntiles ? (ntiles - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
    if (my_ntile == 0) return;
    auto stage_vals = [&](int s){ return reinterpret_cast<T*>(smem_raw + (size_t) s * STAGE_BYTES); };
    auto stage_off  = [&](int s){ return reinterpret_cast<unsigned short*>(smem_raw + (size_t) s * STAGE_BYTES + VAL_BYTES); };
    auto stage_seg  = [&](int s){ return reinterpret_cast<T*>(smem_raw + (size_t) s * STAGE_BYTES + VAL_BYTES + OFF_BYTES); };
    if (tid == 0){
        for (int s = 0; s < BD_STAGES; s++){ mbar_init(full + s, 1); mbar_init(empty + s, BD_THREADS / 32); }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= BD_THREADS){
        // ---------------------------------------------------------------- producer warp
        const int lane = tid - BD_THREADS;
        const int nnz8 = nnz & ~7;
        auto tile_desc = [&](int k){ return desc + (size_t) ((int) blockIdx.x + k * (int) gridDim.x) * BD_DESC; };
        int d0 = __ldg(tile_desc(0) + lane), d1 = (lane < BD_DESC - 32) ? __ldg(tile_desc(0) + 32 + lane) : 0;
        for (int k = 0; k < my_ntile; k++){
            const int s = k % BD_STAGES;
            // this tile's descriptor out of the lanes' registers; the next one is requested before anything is waited for
            const int c0 = d0, c1 = d1;
            if (k + 1 < my_ntile){ d0 = __ldg(tile_desc(k + 1) + lane); d1 = (lane < BD_DESC - 32) ? __ldg(tile_desc(k + 1) + 32 + lane) : 0; }
            auto dget = [&](int i){ const int lo = __shfl_sync(0xffffffffu, c0, i & 31), hi = __shfl_sync(0xffffffffu, c1, i & 31); return i < 32 ? lo : hi; };
            const int b0 = dget(0), b1 = dget(1), nr = dget(2);
            if (k >= BD_STAGES){
                if (!mbar_wait_bounded(empty + s, (uint32_t) (((k / BD_STAGES) - 1) & 1))){ if (lane == 0) atomicExch(err, 1); return; }
            }
            T *sv = stage_vals(s); unsigned short *so = stage_off(s); T *sg = stage_seg(s);
            const int a0 = b0 & ~7, end8 = (b1 + 7) & ~7;
            const int nb = max(min(end8, nnz8) - a0, 0);
            uint32_t bytes = (uint32_t) (nb * (sizeof(T) + 2));
            int off_prev = dget(20);
            if (lane == 0){
                for (int e = a0 + nb; e < min(end8, nnz); e++){ sv[e - a0] = vals[e]; so[e - a0] = soff[e]; }   // ragged end of the arrays by hand
                stage_a0[s] = a0;
            }
            bytes += (uint32_t) ((dget(20 + nr) - off_prev) * BD_NBP * sizeof(T));
            if (lane == 0){
                fence_proxy_async();
                mbar_arrive_expect_tx(full + s, bytes);
                if (nb > 0){
                    bulk_g2s(sv, vals + a0, (uint32_t) (nb * sizeof(T)), full + s);
                    bulk_g2s(so, soff + a0, (uint32_t) (nb * 2), full + s);
                }
            }
            for (int q = 0; q < nr; q++){                       // warp-uniform trip count; only lane 0 issues
                const int start = dget(4 + q), off_next = dget(21 + q);
                if (lane == 0)
                    bulk_g2s(sg + (size_t) off_prev * BD_NBP, Bt + (size_t) start * BD_NBP, (uint32_t) ((off_next - off_prev) * BD_NBP * sizeof(T)), full + s);
                off_prev = off_next;
            }
        }
        return;
    }

    // -------------------------------------------------------------------- consumers
    const int sub = tid % TPR, grp = tid / TPR;
    // row bounds are read one tile ahead so that their latency is off the per-tile critical path
    int rs = 0, re = 0;
    {
        const long long row = (long long) blockIdx.x * ROWS + grp;
        if (row < rows){ rs = __ldg(pntr + row); re = __ldg(pntr + row + 1); }
    }
    for (int k = 0; k < my_ntile; k++){
        const int s = k % BD_STAGES;
        int nrs = 0, nre = 0;
        if (k + 1 < my_ntile){
            const long long row = ((long long) blockIdx.x + (long long) (k + 1) * gridDim.x) * ROWS + grp;
            if (row < rows){ nrs = __ldg(pntr + row); nre = __ldg(pntr + row + 1); }
        }
        if (!mbar_wait_bounded(full + s, (uint32_t) ((k / BD_STAGES) & 1))){ if (tid == 0) atomicExch(err, 1); return; }
        const long long myrow = ((long long) blockIdx.x + (long long) k * gridDim.x) * ROWS + grp;
        const int a0 = stage_a0[s];
        const T *sv = stage_vals(s); const unsigned short *so = stage_off(s);
        const vec16<T> *sg = reinterpret_cast<const vec16<T>*>(stage_seg(s));
        T sums[BD_NBP];
        #pragma unroll
        for (int e = 0; e < BD_NBP; e++) sums[e] = zero_of<T>();
        const int end = re - a0;
        for (int base = rs + sub - a0; base < end; base += MU * TPR){
            T v[MU]; vec16<T> pk[MU][NPK]; bool ok[MU];
            #pragma unroll
            for (int u = 0; u < MU; u++){
                ok[u] = (base + u * TPR) < end;
                if (ok[u]){
                    v[u] = sv[base + u * TPR];
                    const int o = so[base + u * TPR];
                    #pragma unroll
                    for (int q = 0; q < NPK; q++) pk[u][q] = sg[o * NPK + q];
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
        // every lane of the warp has read what it needs from the stage: hand it back to the producer
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(empty + s);
        #pragma unroll
        for (int e = 0; e < BD_NBP; e++){
            #pragma unroll
            for (int d = TPR / 2; d > 0; d >>= 1) sums[e] = hadd(sums[e], shfl_down(sums[e], d));
        }
        if (sub == 0 && myrow < rows){
            vec16<T> *cp = reinterpret_cast<vec16<T>*>(Ct + (size_t) myrow * BD_NBP);
            #pragma unroll
            for (int q = 0; q < NPK; q++){
                vec16<T> o;
                #pragma unroll
                for (int e = 0; e < NV; e++) o.v[e] = sums[q * NV + e];
                cp[q] = o;
            }
        }
        rs = nrs; re = nre;
    }
}

// ------------------------------------------------------------------------------------------------ host side
template<typename T> static size_t band_smem(){
    return BD_STAGES * ((size_t) bd_cap<T>() * (sizeof(T) + 2) + (size_t) BD_SEGROWS * BD_NBP * sizeof(T));
}
static int band_analyse(hb_ctx *ctx, const hb_csr *A, hb_band *b){
    b->tpr = A->mean_row_nnz > 15.0 ? 4 : 2;
    b->tile_rows = BD_THREADS / b->tpr;
    b->ntiles = (A->rows + b->tile_rows - 1) / b->tile_rows;
    int *unfit = nullptr;
    if (cudaMalloc((void**) &b->desc, sizeof(int) * (size_t) b->ntiles * BD_DESC) != cudaSuccess ||
        cudaMalloc((void**) &b->soff, sizeof(unsigned short) * ((size_t) A->nnz + 16)) != cudaSuccess ||
        cudaMalloc((void**) &unfit, sizeof(int)) != cudaSuccess){
        cudaGetLastError();
        if (unfit) cudaFree(unfit);
        return -1;
    }
    HB_CUDA(cudaMemsetAsync(unfit, 0, sizeof(int), ctx->stream));
    HB_CUDA(cudaMemsetAsync(b->soff, 0, sizeof(unsigned short) * ((size_t) A->nnz + 16), ctx->stream));
    const int grid = std::min(b->ntiles, ctx->num_sms * 4);
    if (hb_dtype_size(A->dtype) == 16) band_analyse_kernel<1024><<<grid, BD_THREADS, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, b->tile_rows, b->ntiles, b->desc, b->soff, unfit);
    else                               band_analyse_kernel<2048><<<grid, BD_THREADS, 0, ctx->stream>>>(A->rows, A->pntr, A->indx, b->tile_rows, b->ntiles, b->desc, b->soff, unfit);
    HB_LAUNCH_CHECK(ctx);
    int flag = 0;
    HB_CUDA(cudaMemcpyAsync(&flag, unfit, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    HB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(unfit);
    return flag ? -1 : 1;
}

// true when the matrix can take the staged-operand kernel (runs the analysis on first use)
bool hb_spmm_band_ok(hb_ctx *ctx, const hb_csr *A){
    if (!A->vec_aligned || A->nnz <= 0 || A->rows <= 0 || A->mean_row_nnz > 60.0) return false;
    hb_band *b = A->band_slot[0];
    if (!b){ b = new hb_band(); A->band_slot[0] = b; }
    if (b->state == 0){
        b->state = band_analyse(ctx, A, b);
        if (b->state < 0){
            if (b->desc) cudaFree(b->desc);
            if (b->soff) cudaFree(b->soff);
            b->desc = nullptr; b->soff = nullptr;
        }
    }
    return b->state == 1;
}

template<typename T, int TPR>
static int launch_band(hb_ctx *ctx, const hb_csr *A, const hb_band *b, const T *Bt, T *Ct){
    const size_t smem = band_smem<T>();
    auto k = spmm_band_kernel<T, TPR>;
    static int occ = -1;
    if (occ < 0){
        HB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int n = 0;
        HB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, BD_THREADS + 32, smem));
        occ = n < 1 ? 1 : n;
    }
    int *err = reinterpret_cast<int*>(reinterpret_cast<char*>(ctx->dscalars) + 2048 + 64);
    HB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), ctx->stream));
    const int grid = std::min(b->ntiles, ctx->num_sms * occ);
    k<<<grid, BD_THREADS + 32, smem, ctx->stream>>>(A->rows, A->nnz, A->pntr, (const T*) A->vals, b->soff, b->desc, b->ntiles, Bt, Ct, err);
    HB_LAUNCH_CHECK(ctx);
    return HB_OK;
}
// Ct = A Bt on interleaved blocks of 4 right-hand sides (Bt: cols x 4, Ct: rows x 4); hb_spmm_band_ok(ctx, A) must have returned true
int hb_spmm_band(hb_ctx *ctx, const hb_csr *A, const void *Bt, void *Ct){
    const hb_band *b = A->band_slot[0];
    HB_ARG(b && b->state == 1, "matrix has no band analysis");
    HB_DISPATCH(A->dtype, {
        if (b->tpr == 4) return launch_band<T, 4>(ctx, A, b, (const T*) Bt, (T*) Ct);
        return launch_band<T, 2>(ctx, A, b, (const T*) Bt, (T*) Ct);
    });
    return HB_OK;
}
