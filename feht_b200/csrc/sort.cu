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
// Two engines.  The default is ONE cooperative launch for all passes (coop_sort_kernel below: a chunk of up to
// 8192 keys per SM, or several chunks per SM for larger outputs).  The other, kept for cross-checks
// (FEHT_SORT=onesweep), runs each pass as one kernel ("onesweep"): a CTA claims the next 2048-item tile from an
// atomic ticket, ranks its items stably per digit (warp match-any + per-warp counters in shared memory), publishes
// its per-digit counts and resolves the counts of all earlier tiles with a decoupled look-back over a status
// word per (tile, digit), then scatters (key64, index) to the final positions; the digit histogram the NEXT pass
// needs is accumulated by the current pass while it has the keys in registers.  In both, a pass whose digit is
// identical in every key is skipped (OR/AND reduction of the key words), so a single comparison costs no
// "comparison" passes; when the survivors are emitted in name-rank order (dense output) the stable sort needs no
// name-rank passes either.  The sort moves (8-byte key, 4-byte index) pairs; the 32-byte survivor records are read
// once, through the permutation, by the kernel that writes the results (K3b, or gather_kernel for the p-value order).
//
// Bound: HBM/L2 traffic, 24 B per item per pass.
#include "common.cuh"

namespace feht {
namespace {

constexpr int kRankStride = int(sizeof(SurvivorRec) / 4);   // words between the name ranks of two survivors
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;  // 2048
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

// digit of item for a pass: src 0 = byte of the carried 64-bit key, src 1 = byte of word[val]
// (the name rank lives in the 32-byte survivor record: word stride 8; the comparison index has its own array)
__device__ __forceinline__ uint32_t digit_of(int src, int shift, unsigned long long k, uint32_t v,
                                             const uint32_t *__restrict__ word, int stride) {
    return src == 0 ? uint32_t(k >> shift) & 255u : (__ldg(word + size_t(v) * stride) >> shift) & 255u;
}

// global histogram of the first pass
__global__ void __launch_bounds__(256)
first_hist_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ val,
                  int64_t n, int src, int shift, const uint32_t *__restrict__ word, int stride,
                  uint32_t *__restrict__ hist) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x)
        atomicAdd(&h[digit_of(src, shift, key[i], val[i], word, stride)], 1u);
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
    int stride, nstride;         // its word stride (this pass / next pass)
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
        dig[it] = in ? digit_of(A.src, A.shift, k[it], v[it], A.word, A.stride) : 256u;
        if (A.has_next && in) atomicAdd(&nh[digit_of(A.nsrc, A.nshift, k[it], v[it], A.nword, A.nstride)], 1u);
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


// ---- single-wave variant: all passes in ONE cooperative launch.
// With at most kCoopCap items per SM the whole job is one wave, and the look-back of the tile-per-CTA
// kernel above degenerates into a serial walk over every earlier tile (nobody has a prefix yet): ncu
// showed 23-32 us per pass at 1M keys, 12% issue-active, the hottest line the look-back spin.  Here each
// CTA owns one contiguous chunk (<= 8192 items), ranks it, stages it in shared memory in digit order and
// writes each digit's run with consecutive threads on consecutive addresses.  The CTAs exchange their
// per-digit counts through a [grid][256] table of self-validating words (pass tag << 16 | count): a
// reader spins only on the words it needs, so there is no grid barrier in front of the prefix; the one
// barrier per pass (a counter) stands between a pass's writes and the next pass's reads.  Which passes
// run is decided on the device from the OR/AND words K3b left behind, so the host neither synchronises
// nor copies anything before the launch.  1M keys x 8 passes: 241 us (tile-per-CTA) -> 131 us + gather.
constexpr int kCoopThreads = 1024;
constexpr int kCoopWarps = kCoopThreads / 32;
constexpr int kCoopItems = 8;
constexpr int kCoopCap = kCoopThreads * kCoopItems;   // items per CTA

struct CoopArgs {
    unsigned long long *key[2];
    uint32_t *val[2];
    int64_t n;
    int chunk;                  // items per chunk
    int rounds;                 // ceil(chunk / kCoopThreads) <= kCoopItems
    int chunks;                 // chunks per CTA (multi-chunk variant; the single-chunk variant has one)
    const uint32_t *or_and;     // [8]: OR of (rank, key lo, key hi, comp), then AND of the same
    const uint32_t *r_rank, *r_comp;   // digit sources: &rec[0].rank (word stride 8), comparison index (stride 1)
    int rank_presorted;
    uint32_t *hist;             // [grid][256] per-CTA digit counts of the current pass: tag << 16 | count
    uint32_t *sel;              // out: which val buffer holds the permutation
    uint32_t tag_base;          // 32 * launch number: tags of different launches never coincide
    uint32_t *bar;              // grid barrier counter between passes, zeroed before the launch
    uint32_t *flags;            // [grid] (two-CTAs-per-SM variant): pass number whose counts the CTA has published, zeroed
};

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// polling load: volatile asm, or the compiler folds the re-load of a spin loop into the first load and
// drops the loop (a side-effect-free loop may be assumed to terminate)
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void grid_barrier(uint32_t *bar, unsigned &target) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire_u32(bar) < target) {}
        __threadfence();
    }
    __syncthreads();
}

// lanes of the warp holding the same 9-bit digit: nine ballots instead of MATCH.ANY
__device__ __forceinline__ unsigned peers_of(uint32_t d) {
#ifdef FEHT_SORT_MATCH_ANY
    return __match_any_sync(0xffffffffu, d);
#else
    unsigned peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; ++b) {
        const bool bit = (d >> b) & 1u;
        const unsigned m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
#endif
}

constexpr size_t kCoopDynSmem = size_t(kCoopCap) * (8 + 4 + 1);

// MULTI = false: one chunk per CTA (n <= grid * 8192), counts exchanged as tagged words, no barrier before the
// prefix.  MULTI = true: each CTA owns a contiguous range of `chunks` full chunks; a pass first counts the digits
// of the whole range (keys read once more), exchanges plain counts behind a grid barrier, then ranks, stages and
// writes chunk after chunk with a running per-digit offset -- three sweeps of the data per pass instead of the
// two of a look-back design, but every write is a run of one digit's items (32 per chunk on average = 256 B).
template <bool MULTI>
__global__ void __launch_bounds__(kCoopThreads, 1) coop_sort_kernel(const __grid_constant__ CoopArgs A) {
    // cnt: per-warp digit counters while ranking; afterwards the same 32 KB hold the partial sums of the
    // cross-CTA prefix (16 slices x {before, total} x 256 digits)
    __shared__ __align__(16) unsigned int cnt[kCoopWarps][256];
    __shared__ unsigned int lstart[256];   // chunk-local start of each digit in the staged order
    __shared__ unsigned int goff[256];     // global position of staged index 0 of each digit
    __shared__ unsigned int run_off[256];  // global position of this CTA's next item of each digit
    __shared__ unsigned int wsum[8];
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned long long *stage_k = reinterpret_cast<unsigned long long *>(dyn);
    uint32_t *stage_v = reinterpret_cast<uint32_t *>(dyn + size_t(kCoopCap) * 8);
    unsigned char *stage_d = dyn + size_t(kCoopCap) * 12;
    unsigned int(*part)[256] = cnt;        // [0..15] before, [16..31] total

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int n_chunks = MULTI ? A.chunks : 1;
    const int64_t cta_base = int64_t(cta) * A.chunk * n_chunks;
    const int rounds = A.rounds;
    const int wbase = wid * 32 * rounds;
    uint32_t oa[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) oa[j] = A.or_and[j];
    // candidate digits, least significant first: 4 name-rank bytes, 8 key bytes, 4 comparison bytes;
    // a digit that is the same in every key needs no pass
    unsigned todo = 0;
#pragma unroll
    for (int cand = 0; cand < 16; ++cand) {
        const int w = cand < 4 ? 0 : (cand < 8 ? 1 : (cand < 12 ? 2 : 3));
        if (((((oa[w] ^ oa[4 + w]) >> (8 * (cand & 3))) & 255u) != 0u) && !(cand < 4 && A.rank_presorted))
            todo |= 1u << cand;
    }
    uint32_t tag = A.tag_base;
    int cur = 0;
    unsigned target = 0;

    while (todo) {
        const int cand = __ffs(todo) - 1;
        todo &= todo - 1;
        const int src = (cand >= 4 && cand < 12) ? 0 : 1;
        const int shift = src == 0 ? 8 * (cand - 4) : 8 * (cand & 3);
        const uint32_t *word = cand < 4 ? A.r_rank : A.r_comp;
        const int wstride = cand < 4 ? kRankStride : 1;
        ++tag;

        const unsigned long long *kin = A.key[cur];
        const uint32_t *vin = A.val[cur];
        unsigned long long *kout = A.key[cur ^ 1];
        uint32_t *vout = A.val[cur ^ 1];

        if (MULTI) {
            // sweep 1: digit counts of the CTA's whole range
            for (int j = tid; j < kCoopWarps * 256; j += kCoopThreads) (&cnt[0][0])[j] = 0;
            __syncthreads();
            const int64_t lo = cta_base, hi = min(A.n, cta_base + int64_t(A.chunk) * n_chunks);
            if (src == 0) {
                // four keys in flight per thread (one load per trip left this sweep waiting on its own load: 14% of the
                // kernel's stall samples, ncu r3g)
                for (int64_t i0 = lo + tid; i0 < hi; i0 += 4 * kCoopThreads) {
                    unsigned long long kk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t i = i0 + int64_t(u) * kCoopThreads;
                        kk[u] = i < hi ? __ldcg(kin + i) : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i0 + int64_t(u) * kCoopThreads < hi) atomicAdd(&cnt[wid][uint32_t(kk[u] >> shift) & 255u], 1u);
                }
            } else {
                for (int64_t i = lo + tid; i < hi; i += kCoopThreads) {
                    const uint32_t d = (__ldg(word + size_t(__ldcg(vin + i)) * wstride) >> shift) & 255u;
                    atomicAdd(&cnt[wid][d], 1u);
                }
            }
            __syncthreads();
            if (tid < 256) {
                unsigned run = 0;
#pragma unroll 8
                for (int ww = 0; ww < kCoopWarps; ++ww) run += cnt[ww][tid];
                A.hist[size_t(cta) * 256 + tid] = run;
            }
            grid_barrier(A.bar, target);
        }

        for (int ch = 0; ch < n_chunks; ++ch) {
            const int64_t base = cta_base + int64_t(ch) * A.chunk;
            const int count = int(min(int64_t(A.chunk), max(int64_t(0), A.n - base)));
            if (MULTI) __syncthreads();   // the previous chunk's stage and cnt are consumed
            for (int j = tid; j < kCoopWarps * 256; j += kCoopThreads) (&cnt[0][0])[j] = 0;
            unsigned long long k[kCoopItems];
            uint32_t v[kCoopItems];
            uint32_t dr[kCoopItems];   // digit | rank inside the warp's items << 9
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                const int li = wbase + it * 32 + lane;
                const bool in = it < rounds && li < count;
                // .cg: the ping-pong buffers are rewritten by other SMs every pass
                k[it] = in ? __ldcg(kin + base + li) : 0ull;
                v[it] = in ? __ldcg(vin + base + li) : 0u;
            }
            __syncthreads();
            // stable rank of every item among the warp's items with the same digit (issuing all the MATCH ops
            // first was tried: it spills under the 64-register cap and came out 4% slower)
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                if (it < rounds) {
                    const bool in = wbase + it * 32 + lane < count;
                    const uint32_t d = in ? digit_of(src, shift, k[it], v[it], word, wstride) : 256u;
                    const unsigned peers = peers_of(d);
                    const unsigned lower = peers & ((1u << lane) - 1u);
                    uint32_t b = 0;
                    if (d < 256u) b = cnt[wid][d];
                    __syncwarp();
                    if (d < 256u && lower == 0) cnt[wid][d] = b + __popc(peers);
                    __syncwarp();
                    dr[it] = d | ((b + __popc(lower)) << 9);
                }
            }
            __syncthreads();
            unsigned my_run = 0;   // threads 0..255: items of digit tid in this chunk
            if (tid < 256) {
                unsigned run = 0;
#pragma unroll 8
                for (int ww = 0; ww < kCoopWarps; ++ww) {
                    const unsigned c = cnt[ww][tid];
                    cnt[ww][tid] = run;
                    run += c;
                }
                my_run = run;
                if (!MULTI) A.hist[size_t(cta) * 256 + tid] = (tag << 16) | run;   // the tag makes the word self-validating
                unsigned inc = run;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) wsum[wid] = inc;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                unsigned wb = 0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww)
                    if (ww < wid) wb += wsum[ww];
                lstart[tid] = wb + inc - run;
            }
            __syncthreads();
            // stage the chunk in digit order (stable): needs nothing from the other CTAs
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                if (it < rounds) {
                    const uint32_t d = dr[it] & 511u;
                    if (d < 256u) {
                        const uint32_t li = lstart[d] + cnt[wid][d] + (dr[it] >> 9);
                        stage_k[li] = k[it];
                        stage_v[li] = v[it];
                        stage_d[li] = (unsigned char)d;
                    }
                }
            }
            __syncthreads();
            if (ch == 0) {
                // counts of every digit in the CTAs before me and in all CTAs: thread = 4 digits x one of 16
                // slices of the CTA list, 128-bit loads, all of a thread's loads in flight together.  Single
                // chunk: no grid barrier, a word is valid once it carries this pass's tag, so a thread only
                // waits for the words it needs (one L2 round trip after the slowest CTA published).
                const int q = tid & 63, grp = tid >> 6;
                uint4 before = make_uint4(0, 0, 0, 0), total = make_uint4(0, 0, 0, 0);
                const uint4 *h4 = reinterpret_cast<const uint4 *>(A.hist);
                constexpr int kSlice = 10;   // 16 x 10 >= any grid this kernel is launched with
                uint4 h[kSlice];
#pragma unroll
                for (int j = 0; j < kSlice; ++j) {
                    const int c = grp + 16 * j;
                    h[j] = c < G ? __ldcg(h4 + size_t(c) * 64 + q) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int j = 0; j < kSlice; ++j) {
                    const int c = grp + 16 * j;
                    if (c < G) {
                        uint4 m = h[j];
                        if (!MULTI) {
                            while ((h[j].x >> 16) != tag || (h[j].y >> 16) != tag || (h[j].z >> 16) != tag ||
                                   (h[j].w >> 16) != tag)
                                h[j] = ld_relaxed_v4(h4 + size_t(c) * 64 + q);
                            m = make_uint4(h[j].x & 0xffffu, h[j].y & 0xffffu, h[j].z & 0xffffu, h[j].w & 0xffffu);
                        }
                        total.x += m.x; total.y += m.y; total.z += m.z; total.w += m.w;
                        if (c < cta) { before.x += m.x; before.y += m.y; before.z += m.z; before.w += m.w; }
                    }
                }
                *reinterpret_cast<uint4 *>(&part[grp][4 * q]) = before;
                *reinterpret_cast<uint4 *>(&part[16 + grp][4 * q]) = total;
                __syncthreads();
                if (tid < 256) {
                    unsigned before1 = 0, total1 = 0;
#pragma unroll
                    for (int g = 0; g < 16; ++g) {
                        before1 += part[g][tid];
                        total1 += part[16 + g][tid];
                    }
                    unsigned inc = total1;
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += t;
                    }
                    if (lane == 31) wsum[wid] = inc;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    unsigned wb = 0;
#pragma unroll
                    for (int ww = 0; ww < 8; ++ww)
                        if (ww < wid) wb += wsum[ww];
                    run_off[tid] = wb + inc - total1 + before1;
                }
            }
            if (tid < 256) {
                // (threads 0..255 own run_off[tid]: no other thread touches it)
                goff[tid] = run_off[tid] - lstart[tid];
                run_off[tid] += my_run;
            }
            __syncthreads();
            // consecutive threads write consecutive addresses inside each digit's run
            for (int li = tid; li < count; li += kCoopThreads) {
                const uint32_t pos = goff[stage_d[li]] + uint32_t(li);
                kout[pos] = stage_k[li];
                vout[pos] = stage_v[li];
            }
        }
        cur ^= 1;
        // the next pass reads what every CTA just wrote.  (One flag per CTA with every CTA polling all
        // flags was measured slower than the single counter: 0.176 vs 0.156 ms per 1M-key sort.)
        if (todo) grid_barrier(A.bar, target);
    }
    if (cta == 0 && tid == 0) *A.sel = uint32_t(cur);
}

// Multi-chunk variant with TWO 512-thread CTAs per SM (chunks of up to 4096 survivors): the phases of a chunk -- load,
// rank, stage, write -- are separated by CTA barriers, and with one 1024-thread CTA per SM nothing else ran while a
// phase waited on memory (ncu r3g: issue active 40%, stalls spread over global loads and the barrier).  Two CTAs
// interleave their phases.  Same algorithm and exchange as coop_sort_kernel<true>; FEHT_SORT_MULTI=1024 selects that one.
constexpr int kM2Threads = 512;
constexpr int kM2Warps = kM2Threads / 32;
constexpr int kM2Cap = kM2Threads * kCoopItems;   // 4096 items per chunk
constexpr size_t kM2DynSmem = size_t(kM2Cap) * (8 + 4 + 1);

__global__ void __launch_bounds__(kM2Threads, 2) coop_sort_multi2_kernel(const __grid_constant__ CoopArgs A) {
    __shared__ __align__(16) unsigned int cnt[kM2Warps][256];   // per-warp digit counters; then 8 x {before, total} x 256
    __shared__ unsigned int lstart[256], goff[256], run_off[256];
    __shared__ unsigned int wsum[8];
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned long long *stage_k = reinterpret_cast<unsigned long long *>(dyn);
    uint32_t *stage_v = reinterpret_cast<uint32_t *>(dyn + size_t(kM2Cap) * 8);
    unsigned char *stage_d = dyn + size_t(kM2Cap) * 12;
    unsigned int(*part)[256] = cnt;        // [0..7] before, [8..15] total

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;
    const int n_chunks = A.chunks;
    const int64_t cta_base = int64_t(cta) * A.chunk * n_chunks;
    const int rounds = A.rounds;
    const int wbase = wid * 32 * rounds;
    uint32_t oa[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) oa[j] = A.or_and[j];
    unsigned todo = 0;
#pragma unroll
    for (int cand = 0; cand < 16; ++cand) {
        const int w = cand < 4 ? 0 : (cand < 8 ? 1 : (cand < 12 ? 2 : 3));
        if (((((oa[w] ^ oa[4 + w]) >> (8 * (cand & 3))) & 255u) != 0u) && !(cand < 4 && A.rank_presorted))
            todo |= 1u << cand;
    }
    int cur = 0;
    unsigned target = 0;
    uint32_t pass_no = 0;

    while (todo) {
        ++pass_no;
        const int cand = __ffs(todo) - 1;
        todo &= todo - 1;
        const int src = (cand >= 4 && cand < 12) ? 0 : 1;
        const int shift = src == 0 ? 8 * (cand - 4) : 8 * (cand & 3);
        const uint32_t *word = cand < 4 ? A.r_rank : A.r_comp;
        const int wstride = cand < 4 ? kRankStride : 1;

        const unsigned long long *kin = A.key[cur];
        const uint32_t *vin = A.val[cur];
        unsigned long long *kout = A.key[cur ^ 1];
        uint32_t *vout = A.val[cur ^ 1];

        // sweep 1: digit counts of the CTA's whole range
        for (int j = tid; j < kM2Warps * 256; j += kM2Threads) (&cnt[0][0])[j] = 0;
        __syncthreads();
        {
            const int64_t lo = cta_base, hi = min(A.n, cta_base + int64_t(A.chunk) * n_chunks);
            if (src == 0) {
                for (int64_t i0 = lo + tid; i0 < hi; i0 += 8 * kM2Threads) {
                    unsigned long long kk[8];   // eight keys in flight per thread
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int64_t i = i0 + int64_t(u) * kM2Threads;
                        kk[u] = i < hi ? __ldcg(kin + i) : 0ull;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (i0 + int64_t(u) * kM2Threads < hi) atomicAdd(&cnt[wid][uint32_t(kk[u] >> shift) & 255u], 1u);
                }
            } else {
                for (int64_t i = lo + tid; i < hi; i += kM2Threads) {
                    const uint32_t d = (__ldg(word + size_t(__ldcg(vin + i)) * wstride) >> shift) & 255u;
                    atomicAdd(&cnt[wid][d], 1u);
                }
            }
        }
        __syncthreads();
        if (tid < 256) {
            unsigned run = 0;
#pragma unroll 8
            for (int ww = 0; ww < kM2Warps; ++ww) run += cnt[ww][tid];
            A.hist[size_t(cta) * 256 + tid] = run;
        }
        // No grid barrier here: the counts are published with a flag (release), and nothing needs the other CTAs'
        // counts before the first chunk has been ranked and staged -- the wait for the slowest sweep hides behind that
        // work (ncu r3p: 19% of the samples sat in the two grid barriers of a pass).
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(A.flags + cta), "r"(pass_no) : "memory");
        }

        // a chunk's keys and indices are loaded while the previous chunk is written out (k / v are free once the chunk
        // is staged): ncu r3o had 7% of the samples on the first use of these loads
        unsigned long long k[kCoopItems];
        uint32_t v[kCoopItems];
        auto load_chunk = [&](int ch) {
            const int64_t base = cta_base + int64_t(ch) * A.chunk;
            const int count = int(min(int64_t(A.chunk), max(int64_t(0), A.n - base)));
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                const int li = wbase + it * 32 + lane;
                const bool in = it < rounds && li < count;
                // .cg: the ping-pong buffers are rewritten by other SMs every pass
                k[it] = in ? __ldcg(kin + base + li) : 0ull;
                v[it] = in ? __ldcg(vin + base + li) : 0u;
            }
        };
        load_chunk(0);
        for (int ch = 0; ch < n_chunks; ++ch) {
            const int64_t base = cta_base + int64_t(ch) * A.chunk;
            const int count = int(min(int64_t(A.chunk), max(int64_t(0), A.n - base)));
            __syncthreads();   // the previous chunk's stage and cnt are consumed
            for (int j = tid; j < kM2Warps * 256; j += kM2Threads) (&cnt[0][0])[j] = 0;
            uint32_t dr[kCoopItems];   // digit | rank inside the warp's items << 9
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                if (it < rounds) {
                    const bool in = wbase + it * 32 + lane < count;
                    const uint32_t d = in ? digit_of(src, shift, k[it], v[it], word, wstride) : 256u;
                    // MATCH.ANY here: this kernel is bound by issue slots, not by the latency of one warp's chain (the nine
                    // ballots of peers_of cost 54 of the ~97 instructions of a round: C3 order 1.73 -> 1.58 ms; the
                    // single-chunk kernel, one latency chain per pass, is 8% faster with the ballots)
                    const unsigned peers = __match_any_sync(0xffffffffu, d);
                    const unsigned lower = peers & ((1u << lane) - 1u);
                    uint32_t b = 0;
                    if (d < 256u) b = cnt[wid][d];
                    __syncwarp();
                    if (d < 256u && lower == 0) cnt[wid][d] = b + __popc(peers);
                    __syncwarp();
                    dr[it] = d | ((b + __popc(lower)) << 9);
                }
            }
            __syncthreads();
            unsigned my_run = 0;   // threads 0..255: items of digit tid in this chunk
            if (tid < 256) {
                unsigned run = 0;
#pragma unroll 8
                for (int ww = 0; ww < kM2Warps; ++ww) {
                    const unsigned c = cnt[ww][tid];
                    cnt[ww][tid] = run;
                    run += c;
                }
                my_run = run;
                unsigned inc = run;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) wsum[wid] = inc;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                unsigned wb = 0;
#pragma unroll
                for (int ww = 0; ww < 8; ++ww)
                    if (ww < wid) wb += wsum[ww];
                lstart[tid] = wb + inc - run;
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < kCoopItems; ++it) {
                if (it < rounds) {
                    const uint32_t d = dr[it] & 511u;
                    if (d < 256u) {
                        const uint32_t li = lstart[d] + cnt[wid][d] + (dr[it] >> 9);
                        stage_k[li] = k[it];
                        stage_v[li] = v[it];
                        stage_d[li] = (unsigned char)d;
                    }
                }
            }
            __syncthreads();
            if (ch + 1 < n_chunks) load_chunk(ch + 1);
            if (ch == 0) {
                for (int c = tid; c < G; c += kM2Threads)
                    while (ld_acquire_u32(A.flags + c) != pass_no) {}
                __syncthreads();
                // counts of every digit in the CTAs before me and in all CTAs (plain counts, published with the flag
                // above): thread = 4 digits x one of 8 slices of the CTA list, eight 128-bit loads in flight
                const int q = tid & 63, grp = tid >> 6;
                uint4 before = make_uint4(0, 0, 0, 0), total = make_uint4(0, 0, 0, 0);
                const uint4 *h4 = reinterpret_cast<const uint4 *>(A.hist);
                for (int c0 = grp; c0 < G; c0 += 64) {
                    uint4 h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + 8 * j;
                        h[j] = c < G ? __ldcg(h4 + size_t(c) * 64 + q) : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + 8 * j;
                        total.x += h[j].x; total.y += h[j].y; total.z += h[j].z; total.w += h[j].w;
                        if (c < cta) { before.x += h[j].x; before.y += h[j].y; before.z += h[j].z; before.w += h[j].w; }
                    }
                }
                *reinterpret_cast<uint4 *>(&part[grp][4 * q]) = before;
                *reinterpret_cast<uint4 *>(&part[8 + grp][4 * q]) = total;
                __syncthreads();
                if (tid < 256) {
                    unsigned before1 = 0, total1 = 0;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        before1 += part[g][tid];
                        total1 += part[8 + g][tid];
                    }
                    unsigned inc = total1;
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc += t;
                    }
                    if (lane == 31) wsum[wid] = inc;
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    unsigned wb = 0;
#pragma unroll
                    for (int ww = 0; ww < 8; ++ww)
                        if (ww < wid) wb += wsum[ww];
                    run_off[tid] = wb + inc - total1 + before1;
                }
            }
            if (tid < 256) {
                goff[tid] = run_off[tid] - lstart[tid];
                run_off[tid] += my_run;
            }
            __syncthreads();
            for (int li = tid; li < count; li += kM2Threads) {
                const uint32_t pos = goff[stage_d[li]] + uint32_t(li);
                kout[pos] = stage_k[li];
                vout[pos] = stage_v[li];
            }
        }
        cur ^= 1;
        if (todo) grid_barrier(A.bar, target);
    }
    if (cta == 0 && tid == 0) *A.sel = uint32_t(cur);
}

__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t *__restrict__ perm_a, const uint32_t *__restrict__ perm_b,
              const uint32_t *__restrict__ sel, int64_t n, int64_t row_base, const SurvivorRec *__restrict__ rec,
              const double *__restrict__ r_p, ResultSoA o) {
    // (p-value order only: with the reference's |ratio| order K3b writes the final rows itself)
    const uint32_t *__restrict__ perm = (sel && *sel) ? perm_b : perm_a;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += int64_t(gridDim.x) * blockDim.x) {
        const uint32_t s = perm[i];
        const int4 *rp = reinterpret_cast<const int4 *>(rec + s);
        const int4 t = __ldg(rp), u = __ldg(rp + 1);
        o.row[i] = row_base + u.z;
        o.rank[i] = u.w;
        o.goa[i] = t.x; o.gob[i] = t.y; o.gta[i] = t.z; o.gtb[i] = t.w;
        o.p[i] = r_p[s];
        o.ratio[i] = __hiloint2double(u.y, u.x);
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

// ---- K5b: the same rank-based k-way merge for EVERY comparison of a sharded job in one launch, fused with the
// scatter of all result columns.  Shard s holds its ordered rows of all comparisons back to back (segment c =
// [seg_off_s[c], seg_off_s[c+1])); element j of segment c of shard s goes to
//     out_seg_off[c] + j + sum over the other shards t of #{e in segment c of t : (key_e, rank_e, t) < (key, rank, s)}
// -- one binary search per other shard inside that shard's segment, then the eight columns are written at once
// (no position array, no per-column scatter launches, no loop over comparisons on the host).
struct MergeAllShard {
    const long long *row;
    const int32_t *rank, *goa, *gob, *gta, *gtb;
    const double *p, *ratio;
    const long long *seg_off;     // device, [n_comps + 1]
    // row remap for shards filled by several uploads (multi-device context): local unit -> global unit
    const long long *map_local;   // device, [n_map + 1] local unit starts (null: rows are global already)
    const long long *map_global;  // device, [n_map] global unit starts
    int n_map, mult;              // mult = marker rows per unit (4 in SNP mode)
};
struct MergeAllArgs {
    int n_shards, n_comps, sort_key;
    long long row_base;           // added to remapped rows
    MergeAllShard sh[kMaxShards];
    long long flat_off[kMaxShards + 1];
    const long long *out_seg_off; // device, [n_comps + 1]
    ResultSoA out;
    long long *dest;              // scratch [total rows]: final position of every row, in shard-flat order (two-phase
                                  // variant: rank, then a tile-wise gather with coalesced writes); null = scatter at once
};

// a shard's local marker row -> the job's global row (shards filled by several uploads: segment table)
__device__ __forceinline__ long long merge_global_row(const MergeAllArgs &A, const MergeAllShard &S, long long row) {
    if (!S.map_local) return row;
    const long long unit = row / S.mult;
    int a = 0, b = S.n_map;
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (__ldg(S.map_local + mid) <= unit) a = mid; else b = mid;
    }
    return A.row_base + (__ldg(S.map_global + a) + (unit - __ldg(S.map_local + a))) * S.mult + (row - unit * S.mult);
}

// number of rows e of segment c of shard t with (key_e, rank_e, t) < (kb, rk, s), searched in [lo, hi)
__device__ __forceinline__ long long merge_lower_bound(const MergeAllArgs &A, int t, int s, unsigned long long kb, uint32_t rk,
                                                       long long lo, long long hi) {
    const MergeAllShard &T = A.sh[t];
    const double *tk = A.sort_key == FEHT_SORT_PVALUE_ASC ? T.p : T.ratio;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const unsigned long long ke = merge_key_bits(__ldg(tk + mid), A.sort_key);
        bool before = ke < kb;
        if (ke == kb) {
            const uint32_t re = uint32_t(__ldg(T.rank + mid));
            before = re < rk || (re == rk && t < s);
        }
        if (before) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// shard, index in the shard and comparison of flat row g
__device__ __forceinline__ void merge_locate(const MergeAllArgs &A, long long g, int &s, long long &i, int &c) {
    s = 0;
    while (g >= A.flat_off[s + 1]) ++s;
    i = g - A.flat_off[s];
    const long long *so = A.sh[s].seg_off;
    int lo_c = 0, hi_c = A.n_comps;      // the last c with seg_off[c] <= i
    while (hi_c - lo_c > 1) {
        const int mid = (lo_c + hi_c) >> 1;
        if (__ldg(so + mid) <= i) lo_c = mid; else hi_c = mid;
    }
    c = lo_c;
}

// Level 1: where does the FIRST row of every 256-row block stand in every other shard?  One thread per (block, shard):
// a full binary search each, for 1/256 of the rows.  The rows of a block are consecutive rows of a sorted segment, so
// the answers for the block's rows lie between its own entry and the next block's (merge_all_kernel).
constexpr int kMergeBlock = 256;
__global__ void __launch_bounds__(256) merge_bounds_kernel(const __grid_constant__ MergeAllArgs A, long long *__restrict__ bounds) {
    const long long total = A.flat_off[A.n_shards];
    const long long n_blocks = (total + kMergeBlock - 1) / kMergeBlock;
    const long long n_items = n_blocks * A.n_shards;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n_items;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long blk = idx / A.n_shards;
        const int t = int(idx - blk * A.n_shards);
        int s, c;
        long long i;
        merge_locate(A, blk * kMergeBlock, s, i, c);
        long long b = 0;
        if (t != s) {
            const unsigned long long kb = merge_key_bits(A.sort_key == FEHT_SORT_PVALUE_ASC ? A.sh[s].p[i] : A.sh[s].ratio[i], A.sort_key);
            b = merge_lower_bound(A, t, s, kb, uint32_t(A.sh[s].rank[i]), __ldg(A.sh[t].seg_off + c), __ldg(A.sh[t].seg_off + c + 1));
        }
        bounds[idx] = b;
    }
}

// Level 2: every row's final position.  A block whose 256 rows (and the next block's first row) belong to one sorted
// segment searches only the windows level 1 found -- a few hundred rows per other shard, 8-9 steps that stay in L1 --
// instead of whole segments (~23 dependent L2 round trips per other shard and row: 6.6 ms for 20M rows in 8 shards).
__global__ void __launch_bounds__(kMergeBlock) merge_all_kernel(const __grid_constant__ MergeAllArgs A, const long long *__restrict__ bounds) {
    __shared__ long long w_lo[kMaxShards], w_hi[kMaxShards];
    __shared__ int first_s, first_c, windowed;
    const long long total = A.flat_off[A.n_shards];
    const long long n_blocks = (total + kMergeBlock - 1) / kMergeBlock;
    for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const long long g = blk * kMergeBlock + threadIdx.x;
        const bool valid = g < total;
        int s = 0, c = 0;
        long long i = 0;
        if (valid) merge_locate(A, g, s, i, c);
        if (threadIdx.x == 0) { first_s = s; first_c = c; }
        __syncthreads();
        const bool same = !valid || (s == first_s && c == first_c);
        const int uniform = __syncthreads_and(same);
        if (threadIdx.x == 0) windowed = uniform;
        if (uniform && threadIdx.x < A.n_shards) {
            const int t = threadIdx.x;
            // window of shard t: [this block's entry, the next block's entry] when the next block starts in the same
            // segment, else up to the end of shard t's segment
            const long long g_next = (blk + 1) * kMergeBlock;
            bool next_same = false;
            if (g_next < total && g_next < A.flat_off[first_s + 1]) {
                const long long i_next = g_next - A.flat_off[first_s];
                next_same = i_next < __ldg(A.sh[first_s].seg_off + first_c + 1);
            }
            w_lo[t] = bounds[blk * A.n_shards + t];
            w_hi[t] = next_same ? bounds[(blk + 1) * A.n_shards + t] : __ldg(A.sh[t].seg_off + first_c + 1);
        }
        __syncthreads();
        if (valid) {
            const MergeAllShard &S = A.sh[s];
            const double kd = A.sort_key == FEHT_SORT_PVALUE_ASC ? S.p[i] : S.ratio[i];
            const unsigned long long kb = merge_key_bits(kd, A.sort_key);
            const uint32_t rk = uint32_t(S.rank[i]);
            long long pos = __ldg(A.out_seg_off + c) + (i - __ldg(S.seg_off + c));
            for (int t = 0; t < A.n_shards; ++t) {
                if (t == s) continue;
                const long long t0 = __ldg(A.sh[t].seg_off + c);
                const long long lo = windowed ? w_lo[t] : t0;
                const long long hi = windowed ? w_hi[t] : __ldg(A.sh[t].seg_off + c + 1);
                pos += merge_lower_bound(A, t, s, kb, rk, lo, hi) - t0;
            }
            if (A.dest) {   // two-phase: the columns move in merge_place_kernel
                A.dest[g] = pos;
            } else {
                A.out.row[pos] = merge_global_row(A, S, S.row[i]);
                A.out.rank[pos] = int32_t(rk);
                A.out.goa[pos] = S.goa[i]; A.out.gob[pos] = S.gob[i]; A.out.gta[pos] = S.gta[i]; A.out.gtb[pos] = S.gtb[i];
                A.out.p[pos] = S.p[i];
                A.out.ratio[pos] = S.ratio[i];
            }
        }
        __syncthreads();
    }
}

// Phase 2 of the two-phase variant.  Scattering row by row writes every output sector 1/n_shards full (a warp's 32
// consecutive rows of one shard land n_shards apart): at 8 shards the one-kernel variant spent 6.6 ms on 20M rows.
// Here a CTA owns a tile of kPlaceTile consecutive OUTPUT positions.  A shard's positions increase along the shard, so
// the rows that fall into the tile are one contiguous run per shard (two binary searches on `dest` per shard and
// tile); the runs are read with consecutive threads on consecutive rows, parked in shared memory at (position - tile
// start), and written out with consecutive threads on consecutive positions.
constexpr int kPlaceTile = 1024;
__global__ void __launch_bounds__(256) merge_place_kernel(const __grid_constant__ MergeAllArgs A) {
    __shared__ long long s_row[kPlaceTile];
    __shared__ double s_p[kPlaceTile], s_ratio[kPlaceTile];
    __shared__ int32_t s_rank[kPlaceTile], s_goa[kPlaceTile], s_gob[kPlaceTile], s_gta[kPlaceTile], s_gtb[kPlaceTile];
    __shared__ long long run_lo[kMaxShards];
    __shared__ int run_cum[kMaxShards + 1];
    const long long total = A.flat_off[A.n_shards];
    const long long n_tiles = (total + kPlaceTile - 1) / kPlaceTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long q0 = tile * kPlaceTile, q1 = min(total, q0 + kPlaceTile);
        // the run of every shard inside [q0, q1): first row with dest >= q0 .. first row with dest >= q1
        if (threadIdx.x < 2 * A.n_shards) {
            const int sh = threadIdx.x >> 1;
            const long long want = (threadIdx.x & 1) ? q1 : q0;
            const long long *d = A.dest + A.flat_off[sh];
            long long lo = 0, hi = A.flat_off[sh + 1] - A.flat_off[sh];
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (__ldg(d + mid) < want) lo = mid + 1; else hi = mid;
            }
            if (threadIdx.x & 1) s_row[sh] = lo; else run_lo[sh] = lo;     // (s_row doubles as scratch for the run ends)
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int cum = 0;
            for (int sh = 0; sh < A.n_shards; ++sh) { run_cum[sh] = cum; cum += int(s_row[sh] - run_lo[sh]); }
            run_cum[A.n_shards] = cum;
        }
        __syncthreads();
        const int n_in = run_cum[A.n_shards];      // == q1 - q0
        long long my_row[kPlaceTile / 256];
        int my_slot[kPlaceTile / 256];
#pragma unroll
        for (int it = 0; it < kPlaceTile / 256; ++it) {
            const int j = it * 256 + threadIdx.x;
            my_slot[it] = -1;
            if (j < n_in) {
                int sh = 0;
                while (j >= run_cum[sh + 1]) ++sh;
                const MergeAllShard &S = A.sh[sh];
                const long long i = run_lo[sh] + (j - run_cum[sh]);
                my_slot[it] = int(A.dest[A.flat_off[sh] + i] - q0);
                my_row[it] = merge_global_row(A, S, S.row[i]);
                const int k = my_slot[it];
                s_rank[k] = S.rank[i];
                s_goa[k] = S.goa[i]; s_gob[k] = S.gob[i]; s_gta[k] = S.gta[i]; s_gtb[k] = S.gtb[i];
                s_p[k] = S.p[i];
                s_ratio[k] = S.ratio[i];
            }
        }
        __syncthreads();    // run ends (kept in s_row) are consumed: the rows may go in
#pragma unroll
        for (int it = 0; it < kPlaceTile / 256; ++it)
            if (my_slot[it] >= 0) s_row[my_slot[it]] = my_row[it];
        __syncthreads();
        for (int k = threadIdx.x; k < int(q1 - q0); k += 256) {
            const long long q = q0 + k;
            A.out.row[q] = s_row[k];
            A.out.rank[q] = s_rank[k];
            A.out.goa[q] = s_goa[k]; A.out.gob[q] = s_gob[k]; A.out.gta[q] = s_gta[k]; A.out.gtb[q] = s_gtb[k];
            A.out.p[q] = s_p[k];
            A.out.ratio[q] = s_ratio[k];
        }
        __syncthreads();
    }
}

inline int blocks_for(int64_t n, int per) {
    int64_t b = (n + per - 1) / per;
    if (b < 1) b = 1;
    if (b > int64_t(sm_count()) * 8) b = int64_t(sm_count()) * 8;
    return int(b);
}

}  // namespace

size_t merge_bounds_words(int64_t total_rows, int n_shards) {
    return size_t((total_rows + kMergeBlock - 1) / kMergeBlock + 1) * size_t(n_shards > 0 ? n_shards : 1);
}

size_t sort_status_words(int64_t n) { return size_t((n + kTile - 1) / kTile) * 256; }

cudaError_t radix_sort_perm(SortWorkspace &ws, int64_t n, const SurvivorRec *rec, const int32_t *r_comp,
                            bool rank_presorted, cudaStream_t st, uint32_t **perm_out, int *launches) {
    const uint32_t *r_rank = reinterpret_cast<const uint32_t *>(&rec->rank);
    *perm_out = ws.val_a;
    if (n <= 1) return cudaSuccess;
    uint32_t oa[8];
    cudaError_t e = cudaMemcpyAsync(oa, ws.or_and, sizeof(oa), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;

    struct Pass { int src, shift; const uint32_t *word; int stride; };
    Pass passes[16];
    int np = 0;
    auto varying = [&](int w, int b) { return (((oa[w] ^ oa[4 + w]) >> (8 * b)) & 255u) != 0; };
    if (!rank_presorted)
        for (int b = 0; b < 4; ++b)
            if (varying(0, b)) passes[np++] = Pass{1, 8 * b, r_rank, kRankStride};
    for (int b = 0; b < 8; ++b)
        if (varying(1 + b / 4, b % 4)) passes[np++] = Pass{0, 8 * b, nullptr, 1};
    for (int b = 0; b < 4; ++b)
        if (varying(3, b)) passes[np++] = Pass{1, 8 * b, reinterpret_cast<const uint32_t *>(r_comp), 1};
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
                                                           passes[0].shift, passes[0].word, passes[0].stride, hist);
    if (launches) ++*launches;
    unsigned long long *kin = ws.key_a, *kout = ws.key_b;
    uint32_t *vin = ws.val_a, *vout = ws.val_b;
    for (int p = 0; p < np; ++p) {
        PassArgs a{};
        a.kin = kin; a.vin = vin; a.kout = kout; a.vout = vout; a.n = n;
        a.src = passes[p].src; a.shift = passes[p].shift; a.word = passes[p].word; a.stride = passes[p].stride;
        a.hist = hist + p * 256;
        a.ticket = ticket + p;
        a.status = status + size_t(p) * status_words;
        a.has_next = p + 1 < np;
        if (a.has_next) {
            a.nsrc = passes[p + 1].src; a.nshift = passes[p + 1].shift; a.nword = passes[p + 1].word;
            a.nstride = passes[p + 1].stride;
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

int64_t coop_sort_capacity(int num_sms) { return int64_t(num_sms) * kCoopCap; }
int coop_sort_max_grid() { return 160; }   // kSlice * 16 in the kernel
// rows of the [grid][256] count table: the two-CTAs-per-SM multi-chunk kernel runs 2 * num_sms CTAs
static size_t coop_rows(int num_sms) { return size_t(2) * size_t(num_sms); }
size_t coop_sort_control_words(int num_sms) { return coop_rows(num_sms) * 256 + 8 + coop_rows(num_sms); }

static bool sort_multi_two_ctas() {
    static const bool on = [] { const char *e = getenv("FEHT_SORT_MULTI"); return !(e && strcmp(e, "1024") == 0); }();
    return on;
}

cudaError_t coop_sort_perm(SortWorkspace &ws, int64_t n, const SurvivorRec *rec, const int32_t *r_comp,
                           bool rank_presorted, int num_sms, cudaStream_t st, int *launches) {
    // control block: [grid][256] digit counts, the buffer selector, the barrier counter
    uint32_t *sel = ws.coop + coop_rows(num_sms) * 256;
    cudaError_t e = cudaMemsetAsync(sel, 0, 2 * sizeof(uint32_t), st);
    if (e != cudaSuccess || n <= 1) return e;
    CoopArgs a{};
    a.key[0] = ws.key_a; a.key[1] = ws.key_b;
    a.val[0] = ws.val_a; a.val[1] = ws.val_b;
    a.n = n;
    const bool multi = n > coop_sort_capacity(num_sms);
    bool two = false;
    if (multi && sort_multi_two_ctas()) {
        // (the attribute first: the occupancy query counts the dynamic shared memory it is given)
        e = cudaFuncSetAttribute(coop_sort_multi2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kM2DynSmem));
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_sort_multi2_kernel, kM2Threads, kM2DynSmem);
        if (e != cudaSuccess) return e;
        two = per_sm >= 2;    // a cooperative launch needs every CTA resident
    }
    int grid;
    if (!multi) {
        // whole warps per CTA, as few CTAs idle as possible
        int64_t chunk = (n + num_sms - 1) / num_sms;
        chunk = (chunk + kCoopThreads - 1) / kCoopThreads * kCoopThreads;
        a.chunk = int(chunk);
        a.rounds = int(chunk / kCoopThreads);
        a.chunks = 1;
        grid = int((n + chunk - 1) / chunk);
    } else if (two) {
        // two 512-thread CTAs per SM, chunks of up to 4096: same rule, twice the CTAs
        const int64_t ctas = int64_t(2) * num_sms, cap = ctas * kM2Cap;
        a.chunks = int((n + cap - 1) / cap);
        int64_t chunk = (n + ctas * a.chunks - 1) / (ctas * a.chunks);
        chunk = (chunk + kM2Threads - 1) / kM2Threads * kM2Threads;
        a.chunk = int(chunk);
        a.rounds = int(chunk / kM2Threads);
        grid = int((n + chunk * a.chunks - 1) / (chunk * a.chunks));
    } else {
        // as many chunks per CTA as full-size chunks would need, then the chunk shrunk so that every SM gets work
        // (1.24M keys: 2 chunks of 5120 on 121 CTAs, not 2 x 8192 on 76)
        a.chunks = int((n + coop_sort_capacity(num_sms) - 1) / coop_sort_capacity(num_sms));
        int64_t chunk = (n + int64_t(num_sms) * a.chunks - 1) / (int64_t(num_sms) * a.chunks);
        chunk = (chunk + kCoopThreads - 1) / kCoopThreads * kCoopThreads;
        a.chunk = int(chunk);
        a.rounds = int(chunk / kCoopThreads);
        grid = int((n + chunk * a.chunks - 1) / (chunk * a.chunks));
    }
    a.or_and = ws.or_and;
    a.r_rank = reinterpret_cast<const uint32_t *>(&rec->rank);
    a.r_comp = reinterpret_cast<const uint32_t *>(r_comp);
    a.rank_presorted = rank_presorted ? 1 : 0;
    a.hist = ws.coop;
    a.sel = sel;
    a.bar = sel + 1;
    // Tags (single-chunk variant).  The control block is zeroed when it is allocated; a launch uses tags
    // base+1 .. base+16, never 0 and never those of the launches around it.  The sequence repeats after 2047
    // launches, so rows of the table that no launch in between has rewritten (a larger grid than any recent one,
    // or multi-chunk launches, which store plain counts) are cleared first.
    ws.coop_launches = ws.coop_launches % 2047u + 1u;
    a.tag_base = ws.coop_launches * 32u;
    if (multi) {
        ws.coop_clean_rows = 0;
    } else if (grid > ws.coop_clean_rows || ws.coop_launches == 1u) {
        e = cudaMemsetAsync(ws.coop, 0, size_t(num_sms) * 256 * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        ws.coop_clean_rows = num_sms;
    }
    // (set on every launch: the attribute belongs to the current device's copy of the function, and one process
    // may drive several GPUs)
    void *args[] = {&a};
    if (two) {
        a.flags = sel + 8;
        e = cudaMemsetAsync(a.flags, 0, coop_rows(num_sms) * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        e = cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(coop_sort_multi2_kernel), dim3(grid), dim3(kM2Threads),
                                        args, kM2DynSmem, st);
        if (launches) ++*launches;
        return e;
    }
    e = cudaFuncSetAttribute(multi ? coop_sort_kernel<true> : coop_sort_kernel<false>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(kCoopDynSmem));
    if (e != cudaSuccess) return e;
    const void *fn = multi ? reinterpret_cast<const void *>(coop_sort_kernel<true>)
                           : reinterpret_cast<const void *>(coop_sort_kernel<false>);
    e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kCoopThreads), args, kCoopDynSmem, st);
    if (launches) ++*launches;
    return e;
}

const uint32_t *coop_sort_selector(const SortWorkspace &ws, int num_sms) { return ws.coop + coop_rows(num_sms) * 256; }

cudaError_t launch_gather_results(const uint32_t *perm_a, const uint32_t *perm_b, const uint32_t *d_sel, int64_t n,
                                  int64_t row_base, const SurvivorRec *rec, const double *r_p, const ResultSoA &o,
                                  cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    gather_kernel<<<blocks_for(n, 256), 256, 0, st>>>(perm_a, perm_b, d_sel, n, row_base, rec, r_p, o);
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

cudaError_t launch_merge_all(int n_shards, int n_comps, int sort_key, const MergeShardView *shards,
                             const int64_t *d_out_seg_off, const ResultSoA &out, int64_t row_base, int64_t *d_dest_scratch,
                             int64_t *d_bounds_scratch, cudaStream_t st) {
    if (!d_bounds_scratch) return cudaErrorInvalidValue;
    if (n_shards < 0 || n_shards > kMaxShards) return cudaErrorInvalidValue;
    MergeAllArgs a{};
    a.n_shards = n_shards;
    a.n_comps = n_comps;
    a.sort_key = sort_key;
    a.row_base = row_base;
    long long off = 0;
    for (int i = 0; i < n_shards; ++i) {
        const MergeShardView &v = shards[i];
        MergeAllShard &d = a.sh[i];
        d.row = reinterpret_cast<const long long *>(v.res.row);
        d.rank = v.res.rank; d.goa = v.res.goa; d.gob = v.res.gob; d.gta = v.res.gta; d.gtb = v.res.gtb;
        d.p = v.res.p; d.ratio = v.res.ratio;
        d.seg_off = reinterpret_cast<const long long *>(v.d_seg_off);
        d.map_local = reinterpret_cast<const long long *>(v.d_map_local);
        d.map_global = reinterpret_cast<const long long *>(v.d_map_global);
        d.n_map = v.n_map;
        d.mult = v.mult > 0 ? v.mult : 1;
        a.flat_off[i] = off;
        off += v.n;
    }
    a.flat_off[n_shards] = off;
    a.out_seg_off = reinterpret_cast<const long long *>(d_out_seg_off);
    a.out = out;
    if (off == 0 || n_comps == 0) return cudaSuccess;
    // small outputs and two shards (one kernel is faster there: 0.70 against 1.00 ms for 20M rows): scatter at once; otherwise rank, then place
    a.dest = (d_dest_scratch && n_shards > 2 && off >= (1 << 16)) ? reinterpret_cast<long long *>(d_dest_scratch) : nullptr;
    const int64_t n_blocks = (off + kMergeBlock - 1) / kMergeBlock;
    long long *d_bounds = reinterpret_cast<long long *>(d_bounds_scratch);
    merge_bounds_kernel<<<blocks_for(n_blocks * n_shards, 256), 256, 0, st>>>(a, d_bounds);
    int64_t gb = n_blocks;
    if (gb > int64_t(sm_count()) * 16) gb = int64_t(sm_count()) * 16;
    merge_all_kernel<<<int(gb), kMergeBlock, 0, st>>>(a, d_bounds);
    if (a.dest) {
        int64_t tiles = (off + kPlaceTile - 1) / kPlaceTile;
        if (tiles > int64_t(sm_count()) * 8) tiles = int64_t(sm_count()) * 8;
        merge_place_kernel<<<int(tiles), 256, 0, st>>>(a);
    }
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
