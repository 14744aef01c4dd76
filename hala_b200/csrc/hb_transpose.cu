// hb_transpose.cu — cached transposed copy behind op 'T' / 'C' of hb_spmv / hb_spmm.
// The reference runs y = alpha op(A) x + beta y for op = 'T' / 'C' through cusparseSpMV on the same CSR descriptor
// (gpu/hala_cuda_sparse_general.hpp:264-277; pinned by tests/sparse_tests.hpp:184-190); its CPU twin scatters row by row
// (sparse/hala_sparse_utils.hpp:110-117).  A scatter on the GPU is bound by the atomic units (one RED per non-zero), far
// below the HBM roofline, so the matrix object keeps — built on the first 'T'/'C' product, entirely on the device — the CSR
// of A^T:  t_pntr (cols + 1), t_indx (row of every entry), t_perm (position of every entry in the caller's arrays) and
// t_vals (values in transposed order, conjugated for 'C').  The product is then the ordinary op 'N' streaming kernel on it.
//   * Entries of one column are ordered by row (t_perm ascending), so a column is summed in the order in which the
//     reference's CPU scatter adds to y[column]; the build is deterministic although its fill uses atomics.
//   * The matrix object is a NON-owning view (gpu_sparse_matrix keeps raw pointers, :186-190,370-371): the caller may change
//     the values between two products.  Structure (pntr/indx) is taken as fixed for the life of the object, as the tile tables
//     of the op 'N' kernel already assume.  Values are handled by the transpose mode:
//       HB_TRANS_CHECKED (default)  every 'T'/'C' product first fingerprints the caller's value array (one streaming pass,
//                                   64-bit position-keyed multilinear sum) and re-gathers t_vals only when it differs from the
//                                   fingerprint of the cached copy;
//       HB_TRANS_FROZEN             the caller promises not to change values without hb_csr_values_changed(): no check at all;
//       HB_TRANS_SCATTER            no cached copy, atomic scatter (no extra memory).
//   * Memory: nnz * (8 + sizeof(T)) + 4 cols bytes, allocated on first use; when that allocation fails the object falls back
//     to the scatter kernel for good.
#include "hb_common.cuh"
#include <cstdlib>

struct hb_tcache {
    int   mode  = HB_TRANS_CHECKED;
    int   state = 0;                    // 0 not built, 1 built, -1 cannot be built (scatter instead)
    int  *pntr = nullptr, *indx = nullptr, *perm = nullptr;
    void *vals = nullptr;
    hb_csr *At = nullptr;               // op 'N' view of the arrays above
    char  op = 0;                       // what vals holds: 'T' plain, 'C' conjugated, 0 nothing yet
    bool  stale = true;                 // values must be gathered again whatever the fingerprint says
    unsigned long long *fp = nullptr;   // device: [0] fingerprint of the cached copy, [2] fingerprint of the caller's array now
    size_t bytes = 0;
};

static constexpr int TR_THREADS = 256;

// ------------------------------------------------------------------------------------------------ build: count, scan, fill, sort
__global__ void __launch_bounds__(TR_THREADS) tr_count_kernel(int nnz, const int * __restrict__ indx, int *cnt){
    for (long long j = blockIdx.x * (long long) blockDim.x + threadIdx.x; j < nnz; j += (long long) gridDim.x * blockDim.x)
        atomicAdd(cnt + __ldcs(indx + j), 1);
}

// exclusive scan of 1024 values held one per thread; returns the exclusive prefix, *total = sum of all (valid in every thread)
__device__ __forceinline__ int block_excl_scan_1024(int v, int *total){
    __shared__ int warp_tot[32];
    __shared__ int all_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
    #pragma unroll
    for (int d = 1; d < 32; d <<= 1){ int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0){
        int w = warp_tot[lane], wi = w;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1){ int t = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += t; }
        warp_tot[lane] = wi - w;
        if (lane == 31) all_tot = wi;
    }
    __syncthreads();
    const int excl = warp_tot[warp] + incl - v;
    *total = all_tot;
    __syncthreads();
    return excl;
}
// three-phase exclusive scan of v[0..n): per-chunk sums, scan of the (<= 1024) chunk sums, per-chunk scan with carry
__global__ void __launch_bounds__(1024) tr_scan_reduce_kernel(long long n, long long chunk, const int * __restrict__ v, int *chunk_tot){
    const long long b0 = blockIdx.x * chunk, b1 = (b0 + chunk < n) ? b0 + chunk : n;
    int s = 0;
    for (long long i = b0 + threadIdx.x; i < b1; i += 1024) s += v[i];
    int total;
    block_excl_scan_1024(s, &total);
    if (threadIdx.x == 0) chunk_tot[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) tr_scan_tops_kernel(int nchunks, int *chunk_tot){
    const int v = (int) threadIdx.x < nchunks ? chunk_tot[threadIdx.x] : 0;
    int total;
    const int e = block_excl_scan_1024(v, &total);
    if ((int) threadIdx.x < nchunks) chunk_tot[threadIdx.x] = e;
}
__global__ void __launch_bounds__(1024) tr_scan_apply_kernel(long long n, long long chunk, int *v, const int * __restrict__ chunk_off){
    const long long b0 = blockIdx.x * chunk, b1 = (b0 + chunk < n) ? b0 + chunk : n;
    int carry = chunk_off[blockIdx.x];
    for (long long base = b0; base < b1; base += 1024){
        const long long i = base + threadIdx.x;
        const int x = i < b1 ? v[i] : 0;
        int total;
        const int e = block_excl_scan_1024(x, &total);
        if (i < b1) v[i] = carry + e;
        carry += total;
    }
}

// perm[slot] = position of the entry, slots of one column handed out by an atomic cursor (arbitrary order, sorted afterwards)
__global__ void __launch_bounds__(TR_THREADS) tr_fill_kernel(int nnz, const int * __restrict__ indx, int *cursor, int *perm){
    for (long long j = blockIdx.x * (long long) blockDim.x + threadIdx.x; j < nnz; j += (long long) gridDim.x * blockDim.x)
        perm[atomicAdd(cursor + __ldcs(indx + j), 1)] = (int) j;
}

// columns of up to 32 entries: one thread sorts its column (insertion sort, adaptive: the atomic order is already nearly sorted);
// longer columns are queued for the block-wide sort below
static constexpr int TR_SHORT = 32;
__global__ void __launch_bounds__(TR_THREADS) tr_sort_short_kernel(int cols, const int * __restrict__ tp, int *perm, int *long_list, int *long_count){
    for (long long c = blockIdx.x * (long long) blockDim.x + threadIdx.x; c < cols; c += (long long) gridDim.x * blockDim.x){
        const int a = tp[c], L = tp[c + 1] - a;
        if (L <= 1) continue;
        if (L > TR_SHORT){ long_list[atomicAdd(long_count, 1)] = (int) c; continue; }
        int v[TR_SHORT];
        for (int k = 0; k < L; k++) v[k] = perm[a + k];
        bool moved = false;
        for (int k = 1; k < L; k++){
            const int key = v[k];
            int m = k - 1;
            while (m >= 0 && v[m] > key){ v[m + 1] = v[m]; m--; moved = true; }
            v[m + 1] = key;
        }
        if (moved) for (int k = 0; k < L; k++) perm[a + k] = v[k];
    }
}
// one CTA per long column: bitonic network with ascending comparators only (first step of every merge pairs i with its mirror
// i ^ (k - 1)), so the virtual +inf padding above L never moves and non-power-of-two lengths need no scratch
__global__ void __launch_bounds__(TR_THREADS) tr_sort_long_kernel(const int * __restrict__ long_list, const int * __restrict__ long_count,
                                                                  const int * __restrict__ tp, int *perm){
    const int nlong = *long_count;
    for (int q = blockIdx.x; q < nlong; q += gridDim.x){
        const int c = long_list[q];
        int *a = perm + tp[c];
        const unsigned int L = (unsigned int) (tp[c + 1] - tp[c]);
        unsigned long long n2 = 1;
        while (n2 < L) n2 <<= 1;
        for (unsigned long long k = 2; k <= n2; k <<= 1){
            for (unsigned long long j = k >> 1; j > 0; j >>= 1){
                const unsigned int mask = (unsigned int) ((j == (k >> 1)) ? (k - 1) : j);
                for (unsigned int i = threadIdx.x; i < L; i += TR_THREADS){
                    const unsigned int l = i ^ mask;
                    if (l > i && l < L){
                        const int x = a[i], y = a[l];
                        if (x > y){ a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        }
    }
}
// row of every entry of the transposed copy: the largest r with pntr[r] <= position
__global__ void __launch_bounds__(TR_THREADS) tr_rows_kernel(int nnz, int rows, const int * __restrict__ pntr, const int * __restrict__ perm, int *tindx){
    for (long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x; k < nnz; k += (long long) gridDim.x * blockDim.x){
        const int j = perm[k];
        int lo = 0, hi = rows;
        while (hi - lo > 1){
            const int mid = (int) (((long long) lo + hi) >> 1);
            if (__ldg(pntr + mid) <= j) lo = mid; else hi = mid;
        }
        tindx[k] = lo;
    }
}

// ------------------------------------------------------------------------------------------------ per product: fingerprint, gather
// 64-bit fingerprint of an array of 32-bit words: sum of word * key(position) modulo 2^64 with 64-bit ODD keys derived from a mixed
// packet index.  A change confined to one word always changes the sum (odd key, word delta < 2^32); a change of several words
// (one double already is two) is missed only when sum delta_i * key_i == 0 mod 2^64, about 2^-64 per update for keys that behave as
// random (with the 32-bit keys of round 1 that was 2^-32).  The keys are public: a caller who needs certainty rather than odds uses
// HB_TRANS_FROZEN + hb_csr_values_changed or HB_TRANS_SCATTER.
// Integer addition is associative: the block partials are added with one atomic per block and the result is still deterministic.
__device__ __forceinline__ unsigned long long tr_mix(unsigned long long packet){
    unsigned long long z = packet + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ unsigned long long tr_key(unsigned long long base, int w){
    const int r = 13 * w + 1;
    return ((base << r) | (base >> (64 - r))) | 1ull;
}

template<bool VEC>
__global__ void __launch_bounds__(TR_THREADS) tr_fingerprint_kernel(size_t nwords, const unsigned int * __restrict__ w, unsigned long long *out){
    unsigned long long h = 0;
    uint4 pk[4];
    stream_sweep<unsigned int, VEC, 4>(nwords,
        [&](int u, size_t p){ pk[u] = __ldcs(reinterpret_cast<const uint4*>(w) + p); },
        [&](int u, size_t p){
            const unsigned long long b = tr_mix(p);
            h += (unsigned long long) pk[u].x * tr_key(b, 0);
            h += (unsigned long long) pk[u].y * tr_key(b, 1);
            h += (unsigned long long) pk[u].z * tr_key(b, 2);
            h += (unsigned long long) pk[u].w * tr_key(b, 3);
        },
        [&](size_t i){ h += (unsigned long long) __ldcs(w + i) * tr_key(tr_mix(i >> 2), (int) (i & 3)); });
    #pragma unroll
    for (int d = 16; d > 0; d >>= 1) h += __shfl_down_sync(0xffffffffu, h, d);
    __shared__ unsigned long long red[TR_THREADS / 32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0){
        unsigned long long t = 0;
        #pragma unroll
        for (int k = 0; k < TR_THREADS / 32; k++) t += red[k];
        atomicAdd(out, t + 0x9E3779B97F4A7C15ull * (blockIdx.x == 0 ? (unsigned long long) nwords + 1 : 0ull));
    }
}

// t_vals[k] = op(vals[perm[k]]) — skipped (every block returns at once) when the fingerprints agree and nothing forces it
template<typename T, bool CONJ>
__global__ void __launch_bounds__(TR_THREADS) tr_gather_kernel(int nnz, const T * __restrict__ vals, const int * __restrict__ perm, T *tvals,
                                                               unsigned long long *fp, int force, unsigned int *ticket){
    if (!force && __ldcg(fp) == __ldcg(fp + 2)) return;
    for (long long k = blockIdx.x * (long long) blockDim.x + threadIdx.x; k < nnz; k += (long long) gridDim.x * blockDim.x){
        const T v = ld_ro(vals + __ldcs(perm + k));
        tvals[k] = CONJ ? hconj(v) : v;
    }
    if (last_block_arrives(ticket)){
        if (threadIdx.x == 0) fp[0] = __ldcg(fp + 2);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static void tcache_release(hb_tcache *tc){
    if (tc->At){ hb_csr_destroy(tc->At); tc->At = nullptr; }
    if (tc->pntr) cudaFree(tc->pntr);
    if (tc->indx) cudaFree(tc->indx);
    if (tc->perm) cudaFree(tc->perm);
    if (tc->vals) cudaFree(tc->vals);
    if (tc->fp)   cudaFree(tc->fp);
    tc->pntr = tc->indx = tc->perm = nullptr; tc->vals = nullptr; tc->fp = nullptr;
    tc->op = 0; tc->stale = true; tc->bytes = 0;
}

hb_tcache* hb_tcache_new(){
    hb_tcache *tc = new hb_tcache();
    const char *m = getenv("HB_TRANS_MODE");
    if (m) tc->mode = (m[0] == 's' || m[0] == '0') ? HB_TRANS_SCATTER : (m[0] == 'f' || m[0] == '2') ? HB_TRANS_FROZEN : HB_TRANS_CHECKED;
    return tc;
}
void hb_tcache_delete(hb_tcache *tc){
    if (!tc) return;
    tcache_release(tc);
    delete tc;
}

#define TR_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess){ cudaGetLastError(); return false; } } while (0)

// builds the structure of A^T on the device; false = out of memory (or another CUDA failure): nothing is left allocated
static bool tcache_build(hb_ctx *ctx, const hb_csr *A, hb_tcache *tc){
    const size_t es = hb_dtype_size(A->dtype);
    const long long n = (long long) A->cols + 1;
    int *cursor = nullptr, *long_list = nullptr, *chunk_tot = nullptr, *long_count = nullptr;
    const size_t long_cap = (size_t) A->nnz / (TR_SHORT + 1) + 1;
    auto fail = [&](){
        if (cursor) cudaFree(cursor);
        if (long_list) cudaFree(long_list);
        if (chunk_tot) cudaFree(chunk_tot);
        tcache_release(tc);
        cudaGetLastError();
        return false;
    };
    #define TR_B(call) do { if ((call) != cudaSuccess) return fail(); } while (0)
    TR_B(cudaMalloc((void**) &tc->pntr, sizeof(int) * (size_t) n));
    TR_B(cudaMalloc((void**) &tc->indx, sizeof(int) * (size_t) A->nnz));
    TR_B(cudaMalloc((void**) &tc->perm, sizeof(int) * (size_t) A->nnz));
    TR_B(cudaMalloc(&tc->vals, es * (size_t) A->nnz));
    TR_B(cudaMalloc((void**) &tc->fp, 4 * sizeof(unsigned long long)));
    TR_B(cudaMalloc((void**) &cursor, sizeof(int) * (size_t) n));
    TR_B(cudaMalloc((void**) &long_list, sizeof(int) * long_cap));
    TR_B(cudaMalloc((void**) &chunk_tot, sizeof(int) * 1025));
    long_count = chunk_tot + 1024;
    tc->bytes = sizeof(int) * ((size_t) n + 2 * (size_t) A->nnz) + es * (size_t) A->nnz;
    TR_B(cudaMemsetAsync(tc->pntr, 0, sizeof(int) * (size_t) n, ctx->stream));
    TR_B(cudaMemsetAsync(tc->fp, 0, 4 * sizeof(unsigned long long), ctx->stream));
    TR_B(cudaMemsetAsync(long_count, 0, sizeof(int), ctx->stream));
    const int gnz = hb_grid_for(ctx, (size_t) A->nnz, TR_THREADS * 4, 8), gcol = hb_grid_for(ctx, (size_t) A->cols, TR_THREADS, 8);
    tr_count_kernel<<<gnz, TR_THREADS, 0, ctx->stream>>>(A->nnz, A->indx, tc->pntr);
    // exclusive scan of the counts (cols + 1 entries, the last one zero -> t_pntr[cols] = nnz)
    long long chunk = (n + 1023) / 1024;
    chunk = (chunk + 1023) / 1024 * 1024;
    const int nchunks = (int) ((n + chunk - 1) / chunk);
    tr_scan_reduce_kernel<<<nchunks, 1024, 0, ctx->stream>>>(n, chunk, tc->pntr, chunk_tot);
    tr_scan_tops_kernel<<<1, 1024, 0, ctx->stream>>>(nchunks, chunk_tot);
    tr_scan_apply_kernel<<<nchunks, 1024, 0, ctx->stream>>>(n, chunk, tc->pntr, chunk_tot);
    TR_B(cudaMemcpyAsync(cursor, tc->pntr, sizeof(int) * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
    tr_fill_kernel<<<gnz, TR_THREADS, 0, ctx->stream>>>(A->nnz, A->indx, cursor, tc->perm);
    tr_sort_short_kernel<<<gcol, TR_THREADS, 0, ctx->stream>>>(A->cols, tc->pntr, tc->perm, long_list, long_count);
    tr_sort_long_kernel<<<ctx->num_sms * 4, TR_THREADS, 0, ctx->stream>>>(long_list, long_count, tc->pntr, tc->perm);
    tr_rows_kernel<<<gnz, TR_THREADS, 0, ctx->stream>>>(A->nnz, A->rows, A->pntr, tc->perm, tc->indx);
    ctx->launches += 8;
    TR_B(cudaPeekAtLastError());
    TR_B(cudaStreamSynchronize(ctx->stream));
    cudaFree(cursor); cudaFree(long_list); cudaFree(chunk_tot);
    cursor = long_list = chunk_tot = nullptr;
    if (hb_csr_create(ctx, A->dtype, A->cols, A->rows, A->nnz, tc->pntr, tc->indx, tc->vals, &tc->At) != HB_OK) return fail();
    #undef TR_B
    tc->state = 1; tc->op = 0; tc->stale = true;
    return true;
}

// The op 'N' matrix that stands for op(A), op = 'T' / 'C', with current values; *out = nullptr -> the caller scatters instead.
int hb_csr_transposed(hb_ctx *ctx, const hb_csr *A, char trans, const hb_csr **out){
    *out = nullptr;
    hb_tcache *tc = A->tc;
    if (!tc || tc->mode == HB_TRANS_SCATTER || tc->state < 0 || A->nnz == 0 || A->rows == 0 || A->cols == 0) return HB_OK;
    if (tc->state == 0 && !tcache_build(ctx, A, tc)){ tc->state = -1; return HB_OK; }
    const bool cplx_t = (A->dtype == HB_C32 || A->dtype == HB_C64);
    const char want = (hb_is_c(trans) && cplx_t) ? 'C' : 'T';
    const bool force = tc->stale || tc->op != want;
    if (tc->mode == HB_TRANS_CHECKED || force){
        if (tc->mode == HB_TRANS_CHECKED){
            HB_CUDA(cudaMemsetAsync(tc->fp + 2, 0, sizeof(unsigned long long), ctx->stream));
            const size_t nwords = (size_t) A->nnz * (hb_dtype_size(A->dtype) / 4);
            const int g = hb_grid_for(ctx, nwords, TR_THREADS * 16, 8);
            if (aligned16(A->vals)) tr_fingerprint_kernel<true><<<g, TR_THREADS, 0, ctx->stream>>>(nwords, (const unsigned int*) A->vals, tc->fp + 2);
            else                    tr_fingerprint_kernel<false><<<g, TR_THREADS, 0, ctx->stream>>>(nwords, (const unsigned int*) A->vals, tc->fp + 2);
            HB_LAUNCH_CHECK(ctx);
        }
        const int g = hb_grid_for(ctx, (size_t) A->nnz, TR_THREADS * 4, 8);
        HB_DISPATCH(A->dtype, {
            if (want == 'C') tr_gather_kernel<T, true><<<g, TR_THREADS, 0, ctx->stream>>>(A->nnz, (const T*) A->vals, tc->perm, (T*) tc->vals, tc->fp, force ? 1 : 0, ctx->tickets + 12);
            else             tr_gather_kernel<T, false><<<g, TR_THREADS, 0, ctx->stream>>>(A->nnz, (const T*) A->vals, tc->perm, (T*) tc->vals, tc->fp, force ? 1 : 0, ctx->tickets + 12);
        });
        HB_LAUNCH_CHECK(ctx);
        tc->op = want; tc->stale = false;
    }
    *out = tc->At;
    return HB_OK;
}

extern "C" {

int hb_csr_set_transpose_mode(hb_csr *A, int mode){
    HB_ARG(A && A->tc, "csr is null");
    HB_ARG(mode == HB_TRANS_SCATTER || mode == HB_TRANS_CHECKED || mode == HB_TRANS_FROZEN, "transpose mode");
    if (mode != A->tc->mode) A->tc->stale = true;           // the cached fingerprint may be out of date
    A->tc->mode = mode;
    if (mode == HB_TRANS_SCATTER && A->tc->state == 1){      // give the memory back
        HB_CUDA(cudaStreamSynchronize(A->ctx->stream));
        tcache_release(A->tc);
        A->tc->state = 0;
    }
    if (mode != HB_TRANS_SCATTER && A->tc->state < 0) A->tc->state = 0;   // try the allocation again
    return HB_OK;
}
int hb_csr_values_changed(hb_csr *A){
    HB_ARG(A && A->tc, "csr is null");
    A->tc->stale = true;
    return HB_OK;
}
int hb_csr_transpose_info(const hb_csr *A, int *mode, int *built, size_t *bytes){
    HB_ARG(A && A->tc, "csr is null");
    if (mode) *mode = A->tc->mode;
    if (built) *built = A->tc->state == 1 ? 1 : 0;
    if (bytes) *bytes = A->tc->state == 1 ? A->tc->bytes : 0;
    return HB_OK;
}

}
