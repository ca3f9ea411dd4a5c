// sort.cu -- K4: on-device ordering of the surviving rows.
//
// Replaces Comparison.hs:160-162 (sortBy (comparing (abs . compRatio))) together with the prepend
// reversal at :155/:169, i.e. abs(ratio) DESCENDING inside every comparison; ties -- arbitrary
// HashMap order in the reference -- are broken by marker-name rank ascending so the order is
// reproducible.  All comparisons are ordered by one stable least-significant-digit radix sort over
// the composite key  (comparison | key64 | name_rank), 8 bits per pass:
//     key64 = ~bits(abs ratio)   (abs ratio in [0,1] => IEEE bits are monotone; inverted = descending)
//           =  bits(p)           for FEHT_SORT_PVALUE_ASC
// A pass whose digit is identical in every key is skipped (decided from an OR/AND reduction of the
// key words), so a single comparison costs no "comparison" passes and small jobs cost few
// name-rank passes.  The sort permutes 4-byte indices; the 40-byte records are gathered once.
//
// Bound: HBM/L2 (per pass: read index + gathered key word, write index).
#include "common.cuh"

namespace feht {
namespace {

constexpr int kSortThreads = 256;
constexpr int kItemsPerThread = 8;
constexpr int kTile = kSortThreads * kItemsPerThread;  // 2048 items per block

__global__ void __launch_bounds__(256)
build_keys_kernel(const double *__restrict__ r_ratio, const double *__restrict__ r_p,
                  const int32_t *__restrict__ r_rank, const int32_t *__restrict__ r_comp,
                  int sort_key, int64_t n, uint32_t *__restrict__ k0, uint32_t *__restrict__ k1,
                  uint32_t *__restrict__ k2, uint32_t *__restrict__ k3,
                  uint32_t *__restrict__ perm, uint32_t *__restrict__ or_and) {
    uint32_t o[4] = {0, 0, 0, 0}, a[4] = {~0u, ~0u, ~0u, ~0u};
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        unsigned long long bits;
        if (sort_key == FEHT_SORT_PVALUE_ASC) {
            bits = (unsigned long long)__double_as_longlong(r_p[i]);
        } else {
            bits = ~(unsigned long long)__double_as_longlong(fabs(r_ratio[i]));
        }
        const uint32_t w0 = uint32_t(r_rank[i]);
        const uint32_t w1 = uint32_t(bits);
        const uint32_t w2 = uint32_t(bits >> 32);
        const uint32_t w3 = uint32_t(r_comp[i]);
        k0[i] = w0; k1[i] = w1; k2[i] = w2; k3[i] = w3;
        perm[i] = uint32_t(i);
        o[0] |= w0; o[1] |= w1; o[2] |= w2; o[3] |= w3;
        a[0] &= w0; a[1] &= w1; a[2] &= w2; a[3] &= w3;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        for (int s = 16; s > 0; s >>= 1) {
            o[w] |= __shfl_xor_sync(0xffffffffu, o[w], s);
            a[w] &= __shfl_xor_sync(0xffffffffu, a[w], s);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicOr(or_and + w, o[w]);
            atomicAnd(or_and + 4 + w, a[w]);
        }
    }
}

// Per-block digit histogram: block_hist[digit * n_blocks + block].
__global__ void __launch_bounds__(kSortThreads)
hist_kernel(const uint32_t *__restrict__ perm, const uint32_t *__restrict__ key, int shift, int64_t n,
            int n_blocks, uint32_t *__restrict__ block_hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = int64_t(blockIdx.x) * kTile;
#pragma unroll
    for (int it = 0; it < kItemsPerThread; ++it) {
        const int64_t i = base + it * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(key[perm[i]] >> shift) & 255u], 1u);
    }
    __syncthreads();
    block_hist[size_t(threadIdx.x) * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// Exclusive scan of block_hist (digit-major) in place; one block walks 4096 entries per step.
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t *__restrict__ data, int64_t n) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t chunk_total;
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024 * 4) {
        const int64_t i0 = base + int64_t(threadIdx.x) * 4;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? data[i0 + j] : 0u;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t inc = mine;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = warp_tot[lane];
            uint32_t winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;  // exclusive prefix of the warp totals
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        uint32_t run = carry + warp_tot[wid] + (inc - mine);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) data[i0 + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
}

// Stable scatter of one 8-bit digit.  Tile element order = (warp, iteration, lane).
__global__ void __launch_bounds__(kSortThreads)
scatter_kernel(const uint32_t *__restrict__ perm_in, uint32_t *__restrict__ perm_out,
               const uint32_t *__restrict__ key, int shift, int64_t n, int n_blocks,
               const uint32_t *__restrict__ block_off) {
    __shared__ unsigned int cnt[kSortThreads / 32][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kSortThreads / 32) * 256; i += kSortThreads)
        (&cnt[0][0])[i] = 0;
    __syncthreads();

    const int64_t warp_base = int64_t(blockIdx.x) * kTile + int64_t(wid) * (32 * kItemsPerThread);
    uint32_t idx[kItemsPerThread];
    uint32_t dig[kItemsPerThread];
    uint32_t rnk[kItemsPerThread];
#pragma unroll
    for (int it = 0; it < kItemsPerThread; ++it) {
        const int64_t i = warp_base + it * 32 + lane;
        const bool in = i < n;
        idx[it] = in ? perm_in[i] : 0u;
        dig[it] = in ? ((key[idx[it]] >> shift) & 255u) : 256u;  // 256 = out of range
    }
#pragma unroll
    for (int it = 0; it < kItemsPerThread; ++it) {
        const uint32_t d = dig[it];
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned lower = peers & ((1u << lane) - 1u);
        uint32_t base = 0;
        if (d < 256u) base = cnt[wid][d];
        __syncwarp();
        if (d < 256u && lower == 0) cnt[wid][d] = base + __popc(peers);
        __syncwarp();
        rnk[it] = base + __popc(lower);
    }
    __syncthreads();
    // exclusive prefix over the warps of the block, per digit, plus the global block offset
    {
        const int d = threadIdx.x;  // 256 threads = 256 digits
        unsigned run = block_off[size_t(d) * n_blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const unsigned c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kItemsPerThread; ++it) {
        if (dig[it] < 256u) perm_out[cnt[wid][dig[it]] + rnk[it]] = idx[it];
    }
}

__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t *__restrict__ perm, int64_t n, const int32_t *__restrict__ r_row,
              const int32_t *__restrict__ r_rank, const int4 *__restrict__ r_cnt,
              const double *__restrict__ r_p, const double *__restrict__ r_ratio,
              int32_t *__restrict__ o_row, int32_t *__restrict__ o_rank, int4 *__restrict__ o_cnt,
              double *__restrict__ o_p, double *__restrict__ o_ratio) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t s = perm[i];
        o_row[i] = r_row[s];
        o_rank[i] = r_rank[s];
        o_cnt[i] = r_cnt[s];
        o_p[i] = r_p[s];
        o_ratio[i] = r_ratio[s];
    }
}

inline int blocks_for(int64_t n, int per) {
    int64_t b = (n + per - 1) / per;
    if (b < 1) b = 1;
    if (b > 148 * 16) b = 148 * 16;
    return int(b);
}

}  // namespace

cudaError_t launch_build_keys(const double *r_ratio, const double *r_p, const int32_t *r_rank,
                              const int32_t *r_comp, int sort_key, int64_t n, SortWorkspace &ws,
                              cudaStream_t st) {
    static const uint32_t init[8] = {0, 0, 0, 0, ~0u, ~0u, ~0u, ~0u};
    cudaError_t e = cudaMemcpyAsync(ws.or_and, init, sizeof(init), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    build_keys_kernel<<<blocks_for(n, 256), 256, 0, st>>>(r_ratio, r_p, r_rank, r_comp, sort_key, n,
                                                         ws.keys[0], ws.keys[1], ws.keys[2],
                                                         ws.keys[3], ws.perm_a, ws.or_and);
    return cudaGetLastError();
}

cudaError_t radix_sort_perm(SortWorkspace &ws, int64_t n, cudaStream_t st, uint32_t **perm_out,
                            int *launches) {
    *perm_out = ws.perm_a;
    if (n <= 1) return cudaSuccess;
    uint32_t oa[8];
    cudaError_t e = cudaMemcpyAsync(oa, ws.or_and, sizeof(oa), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    const int n_blocks = int((n + kTile - 1) / kTile);
    uint32_t *in = ws.perm_a, *out = ws.perm_b;
    for (int w = 0; w < 4; ++w) {
        const uint32_t varying = oa[w] ^ oa[4 + w];  // bits that differ somewhere
        for (int b = 0; b < 4; ++b) {
            if (((varying >> (8 * b)) & 255u) == 0) continue;  // constant digit: order unchanged
            hist_kernel<<<n_blocks, kSortThreads, 0, st>>>(in, ws.keys[w], 8 * b, n, n_blocks,
                                                           ws.block_hist);
            scan_kernel<<<1, 1024, 0, st>>>(ws.block_hist, int64_t(256) * n_blocks);
            scatter_kernel<<<n_blocks, kSortThreads, 0, st>>>(in, out, ws.keys[w], 8 * b, n, n_blocks,
                                                              ws.block_hist);
            if (launches) *launches += 3;
            uint32_t *t = in; in = out; out = t;
        }
    }
    *perm_out = in;
    return cudaGetLastError();
}

cudaError_t launch_gather_results(const uint32_t *perm, int64_t n, const int32_t *r_row,
                                  const int32_t *r_rank, const int4 *r_cnt, const double *r_p,
                                  const double *r_ratio, int32_t *o_row, int32_t *o_rank,
                                  int4 *o_cnt, double *o_p, double *o_ratio, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gather_kernel<<<blocks_for(n, 256), 256, 0, st>>>(perm, n, r_row, r_rank, r_cnt, r_p, r_ratio,
                                                     o_row, o_rank, o_cnt, o_p, o_ratio);
    return cudaGetLastError();
}

}  // namespace feht
