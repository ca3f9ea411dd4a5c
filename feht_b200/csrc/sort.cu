// sort.cu -- K4: on-device ordering of the surviving rows.
//
// Replaces Comparison.hs:160-162 (sortBy (comparing (abs . compRatio))) together with the prepend
// reversal at :155/:169, i.e. abs(ratio) DESCENDING inside every comparison; ties -- arbitrary
// HashMap order in the reference -- are broken by marker-name rank ascending so the order is
// reproducible.  All comparisons are ordered by one stable least-significant-digit radix sort over
// the composite key  (comparison | key64 | name_rank), 8 bits per pass:
//     key64 = ~bits(abs ratio)   (abs ratio in [0,1] => IEEE bits are monotone; inverted = descending)
//           =  bits(p)           for FEHT_SORT_PVALUE_ASC
//
// Each pass is ONE kernel ("onesweep"): a CTA claims the next 4096-item tile from an atomic ticket,
// ranks its items stably per digit (warp match-any + per-warp counters in shared memory), publishes
// its per-digit counts and resolves the counts of all earlier tiles with a decoupled look-back over
// a status word per (tile, digit), then scatters (key64, index) to the final positions.  The digit
// histogram the NEXT pass needs is accumulated by the current pass while it has the keys in
// registers, so there is no separate histogram/scan kernel per pass (the first version spent 63 us
// per pass in a single-block scan; profiles/r1_launches.md).  A pass whose digit is identical in
// every key is skipped (OR/AND reduction of the key words), so a single comparison costs no
// "comparison" passes; when the survivors are emitted in name-rank order (dense output) the stable
// sort needs no name-rank passes either.  The sort moves (8-byte key, 4-byte index) pairs; the
// 40-byte records are gathered once at the end.
//
// Bound: HBM/L2 traffic, 24 B per item per pass.
#include "common.cuh"

namespace feht {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 4096
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

// digit of item for a pass: src 0 = byte of the carried 64-bit key, src 1 = byte of word[val]
__device__ __forceinline__ uint32_t digit_of(int src, int shift, unsigned long long k, uint32_t v,
                                             const uint32_t *__restrict__ word) {
    return src == 0 ? uint32_t(k >> shift) & 255u : (__ldg(word + v) >> shift) & 255u;
}

// global histogram of the first pass
__global__ void __launch_bounds__(256)
first_hist_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ val,
                  int64_t n, int src, int shift, const uint32_t *__restrict__ word,
                  uint32_t *__restrict__ hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x)
        atomicAdd(&h[digit_of(src, shift, key[i], val[i], word)], 1u);
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(hist + threadIdx.x, h[threadIdx.x]);
}

struct PassArgs {
    const unsigned long long *kin;
    const uint32_t *vin;
    unsigned long long *kout;
    uint32_t *vout;
    int64_t n;
    int src, shift;              // this pass's digit
    const uint32_t *word;        // gather source when src == 1
    const uint32_t *hist;        // [256] global digit counts of this pass
    uint32_t *ticket;            // tile ticket of this pass
    volatile uint32_t *status;   // [n_tiles][256]
    int has_next, nsrc, nshift;  // next pass's digit (its histogram is accumulated here)
    const uint32_t *nword;
    uint32_t *nhist;
};

__global__ void __launch_bounds__(kThreads) onesweep_pass_kernel(PassArgs A) {
    __shared__ unsigned int cnt[kWarps][256];
    __shared__ unsigned int goff[256];
    __shared__ unsigned int nh[256];
    __shared__ unsigned int wsum[kWarps];
    __shared__ unsigned int tile_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) tile_s = atomicAdd(A.ticket, 1u);
    nh[tid] = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) cnt[w][tid] = 0;
    __syncthreads();
    const unsigned int tile = tile_s;

    const int64_t warp_base = int64_t(tile) * kTile + int64_t(wid) * (32 * kItems);
    unsigned long long k[kItems];
    uint32_t v[kItems];
    uint32_t dig[kItems];
    uint32_t rnk[kItems];
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const int64_t i = warp_base + it * 32 + lane;
        const bool in = i < A.n;
        k[it] = in ? A.kin[i] : 0ull;
        v[it] = in ? A.vin[i] : 0u;
    }
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const bool in = warp_base + it * 32 + lane < A.n;
        dig[it] = in ? digit_of(A.src, A.shift, k[it], v[it], A.word) : 256u;
        if (A.has_next && in) atomicAdd(&nh[digit_of(A.nsrc, A.nshift, k[it], v[it], A.nword)], 1u);
    }
    // stable rank inside the warp's 32*kItems items, per digit
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
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

    // thread d owns digit d: prefix over the warps, digit base, look-back over earlier tiles
    {
        const int d = tid;
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const unsigned c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
        volatile uint32_t *st = A.status + size_t(tile) * 256 + d;
        *st = (tile == 0 ? kFlagPrefix : kFlagAggregate) | run;

        // exclusive scan of the global digit histogram across the 256 threads
        const unsigned hcount = A.hist[d];
        unsigned inc = hcount;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        unsigned wbase = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if (w < wid) wbase += wsum[w];
        const unsigned digit_base = wbase + inc - hcount;

        unsigned excl = 0;
        if (tile > 0) {
            // Decoupled look-back, 8 predecessor tiles per step: the status loads of a step are issued
            // together (one L2 latency per 8 tiles instead of per tile -- with a single wave of tiles
            // the serial walk was most of a pass), then consumed nearest-first, spinning only on a
            // status that is not published yet.
            int64_t t = int64_t(tile) - 1;
            bool done = false;
            while (!done) {
                uint32_t sv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int64_t tj = t - j;
                    sv[j] = tj >= 0 ? *(A.status + size_t(tj) * 256 + d) : (2u << 30);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (done) break;
                    uint32_t sj = sv[j];
                    while ((sj >> 30) == 0u) sj = *(A.status + size_t(t - j) * 256 + d);
                    excl += sj & kValueMask;
                    if (sj & kFlagPrefix) done = true;
                }
                t -= 8;
            }
            *st = kFlagPrefix | (excl + run);
        }
        goff[d] = digit_base + excl;
        if (A.has_next && nh[d]) atomicAdd(A.nhist + d, nh[d]);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        const uint32_t d = dig[it];
        if (d < 256u) {
            const uint32_t pos = goff[d] + cnt[wid][d] + rnk[it];
            A.kout[pos] = k[it];
            A.vout[pos] = v[it];
        }
    }
}

__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t *__restrict__ perm, int64_t n, int64_t row_base, const int32_t *__restrict__ r_row,
              const int32_t *__restrict__ r_rank, const int4 *__restrict__ r_cnt,
              const double *__restrict__ r_p, const double *__restrict__ r_ratio, ResultSoA o) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t s = perm[i];
        const int4 t = r_cnt[s];
        o.row[i] = row_base + r_row[s];
        o.rank[i] = r_rank[s];
        o.goa[i] = t.x; o.gob[i] = t.y; o.gta[i] = t.z; o.gtb[i] = t.w;
        o.p[i] = r_p[s];
        o.ratio[i] = r_ratio[s];
    }
}

// ---- K5: k-way merge of sorted shards held on this device, rank-based: the final
// position of an element is its index in its own shard plus, for every other shard, the number of
// that shard's elements that come before it in the total order (key bits, name rank, shard) -- one
// binary search per other shard.  No radix passes, deterministic, reads only sorted data.
constexpr int kMaxShards = 32;
struct MergeArgs {
    int n_shards;
    int sort_key;
    const double *key[kMaxShards];
    const int32_t *rank[kMaxShards];
    long long *dest[kMaxShards];
    long long n[kMaxShards];
    long long off[kMaxShards + 1];
};

__device__ __forceinline__ unsigned long long merge_key_bits(double v, int sort_key) {
    if (sort_key == FEHT_SORT_PVALUE_ASC) return (unsigned long long)__double_as_longlong(v);
    return ~((unsigned long long)__double_as_longlong(fabs(v)) & 0x7fffffffffffffffULL);
}

__global__ void __launch_bounds__(256) merge_rank_kernel(const __grid_constant__ MergeArgs A) {
    const long long total = A.off[A.n_shards];
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        int s = 0;
        while (g >= A.off[s + 1]) ++s;
        const long long i = g - A.off[s];
        const unsigned long long kb = merge_key_bits(A.key[s][i], A.sort_key);
        const uint32_t rk = uint32_t(A.rank[s][i]);
        long long pos = i;
        for (int t = 0; t < A.n_shards; ++t) {
            if (t == s) continue;
            // number of elements e of shard t with (kb_e, rk_e, t) < (kb, rk, s)
            long long lo = 0, hi = A.n[t];
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                const unsigned long long ke = merge_key_bits(A.key[t][mid], A.sort_key);
                const uint32_t re = uint32_t(A.rank[t][mid]);
                const bool before = ke < kb || (ke == kb && (re < rk || (re == rk && t < s)));
                if (before) lo = mid + 1; else hi = mid;
            }
            pos += lo;
        }
        A.dest[s][i] = pos;
    }
}

__global__ void __launch_bounds__(256)
scatter_rows_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst,
                    const long long *__restrict__ dest, int64_t n, int elem_bytes) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        const long long d = dest[i];
        if (elem_bytes == 8)
            reinterpret_cast<unsigned long long *>(dst)[d] = reinterpret_cast<const unsigned long long *>(src)[i];
        else if (elem_bytes == 4)
            reinterpret_cast<uint32_t *>(dst)[d] = reinterpret_cast<const uint32_t *>(src)[i];
        else
            for (int b = 0; b < elem_bytes; ++b) dst[size_t(d) * elem_bytes + b] = src[size_t(i) * elem_bytes + b];
    }
}

inline int blocks_for(int64_t n, int per) {
    int64_t b = (n + per - 1) / per;
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return int(b);
}

}  // namespace

size_t sort_status_words(int64_t n) { return size_t((n + kTile - 1) / kTile) * 256; }

cudaError_t radix_sort_perm(SortWorkspace &ws, int64_t n, const int32_t *r_rank, const int32_t *r_comp,
                            bool rank_presorted, cudaStream_t st, uint32_t **perm_out, int *launches) {
    *perm_out = ws.val_a;
    if (n <= 1) return cudaSuccess;
    uint32_t oa[8];
    cudaError_t e = cudaMemcpyAsync(oa, ws.or_and, sizeof(oa), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;

    struct Pass { int src, shift; const uint32_t *word; };
    Pass passes[16];
    int np = 0;
    auto varying = [&](int w, int b) { return (((oa[w] ^ oa[4 + w]) >> (8 * b)) & 255u) != 0; };
    if (!rank_presorted)
        for (int b = 0; b < 4; ++b)
            if (varying(0, b)) passes[np++] = Pass{1, 8 * b, reinterpret_cast<const uint32_t *>(r_rank)};
    for (int b = 0; b < 8; ++b)
        if (varying(1 + b / 4, b % 4)) passes[np++] = Pass{0, 8 * b, nullptr};
    for (int b = 0; b < 4; ++b)
        if (varying(3, b)) passes[np++] = Pass{1, 8 * b, reinterpret_cast<const uint32_t *>(r_comp)};
    if (np == 0) return cudaSuccess;

    const int n_tiles = int((n + kTile - 1) / kTile);
    const size_t status_words = size_t(n_tiles) * 256;
    // control block: [16][256] histograms, [16] tickets, then per-pass status arrays
    e = cudaMemsetAsync(ws.control, 0, (size_t(16) * 256 + 16) * 4 + size_t(np) * status_words * 4, st);
    if (e != cudaSuccess) return e;
    uint32_t *hist = ws.control;
    uint32_t *ticket = ws.control + 16 * 256;
    uint32_t *status = ws.control + 16 * 256 + 16;

    first_hist_kernel<<<blocks_for(n, 2048), 256, 0, st>>>(ws.key_a, ws.val_a, n, passes[0].src,
                                                           passes[0].shift, passes[0].word, hist);
    if (launches) ++*launches;
    unsigned long long *kin = ws.key_a, *kout = ws.key_b;
    uint32_t *vin = ws.val_a, *vout = ws.val_b;
    for (int p = 0; p < np; ++p) {
        PassArgs a{};
        a.kin = kin; a.vin = vin; a.kout = kout; a.vout = vout; a.n = n;
        a.src = passes[p].src; a.shift = passes[p].shift; a.word = passes[p].word;
        a.hist = hist + p * 256;
        a.ticket = ticket + p;
        a.status = status + size_t(p) * status_words;
        a.has_next = p + 1 < np;
        if (a.has_next) {
            a.nsrc = passes[p + 1].src; a.nshift = passes[p + 1].shift; a.nword = passes[p + 1].word;
            a.nhist = hist + (p + 1) * 256;
        }
        onesweep_pass_kernel<<<n_tiles, kThreads, 0, st>>>(a);
        if (launches) ++*launches;
        unsigned long long *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    *perm_out = vin;
    return cudaGetLastError();
}

cudaError_t launch_gather_results(const uint32_t *perm, int64_t n, int64_t row_base, const int32_t *r_row,
                                  const int32_t *r_rank, const int4 *r_cnt, const double *r_p,
                                  const double *r_ratio, const ResultSoA &o, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gather_kernel<<<blocks_for(n, 256), 256, 0, st>>>(perm, n, row_base, r_row, r_rank, r_cnt, r_p, r_ratio, o);
    return cudaGetLastError();
}

int merge_max_shards() { return kMaxShards; }

cudaError_t launch_merge_rank(int n_shards, const int64_t *n, const double *const *d_key,
                              const int32_t *const *d_rank, int sort_key, int64_t *const *d_dest,
                              cudaStream_t st) {
    MergeArgs a{};
    a.n_shards = n_shards;
    a.sort_key = sort_key;
    long long off = 0;
    for (int i = 0; i < n_shards; ++i) {
        a.key[i] = d_key[i];
        a.rank[i] = d_rank[i];
        a.dest[i] = reinterpret_cast<long long *>(d_dest[i]);
        a.n[i] = n[i];
        a.off[i] = off;
        off += n[i];
    }
    a.off[n_shards] = off;
    if (off == 0) return cudaSuccess;
    merge_rank_kernel<<<blocks_for(off, 256), 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const void *d_src, void *d_dst, const int64_t *d_dest, int64_t n,
                                int elem_bytes, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    scatter_rows_kernel<<<blocks_for(n, 256), 256, 0, st>>>(
        static_cast<const unsigned char *>(d_src), static_cast<unsigned char *>(d_dst),
        reinterpret_cast<const long long *>(d_dest), n, elem_bytes);
    return cudaGetLastError();
}

}  // namespace feht
