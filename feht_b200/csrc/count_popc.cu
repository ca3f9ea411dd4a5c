// count_popc.cu -- K1: popcount contingency scan (few comparisons).
//
// Replaces Comparison.hs:46-68,116-125 (/root/reference/src): four countCharInVectorByIndices walks
// per (marker, comparison) become, per 128 bits of each bit-plane,
//     P1 += popc(value & M1)   N1 += popc(valid & M1)   P2 += popc(value & M2)   N2 += popc(valid & M2)
// so that goa = P1, gob = N1 - P1, gta = P2, gtb = N2 - P2 (bit-exact integers).
//
// Mapping (DESIGN.md "K1"): a row is split over L = 2^logL lanes, every lane owning NCOL 16-byte
// columns per chunk, so a warp covers 32/L rows at once and all 2*NCOL 128-bit loads of a lane are
// issued back to back before the first use (memory-level parallelism).  Loads are coalesced
// (L consecutive lanes read L*16 contiguous bytes of a row) and bypass L1 allocation: the planes
// are read exactly once.  The group masks are staged ONCE per CTA into shared memory with a TMA
// bulk copy (cp.async.bulk + mbarrier) and re-read from there for every row (zero-padded to the
// lane grid, so no bounds test in the inner loop).  L-lane partial counts are combined with
// shuffles; lane 0 of each row writes one int4.
//
// Bound: HBM.  Algorithmic bytes per marker row = 2 planes * 8 * ceil(S/64);
// the kernel reads 2 * 8 * W (W = that, rounded up to even) and writes 16 B per (row, pair-slot).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace feht {
namespace {

__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
        : "l"(p));
    return r;
}

// POPC issues at 16 lanes/clk/SM (XU pipe) and LOP3 at 64 (ALU pipe), and a plain popcount of
// every masked word needs 0.5 POPC per byte streamed -- ncu showed the XU pipe at 47% with DRAM at
// only 50% (profiles/r1_count_popc.md), i.e. the scan would turn POPC-bound before HBM-bound.
// So each count stream keeps a carry-save "ones" word: two masked words are folded into it with one
// 3:2 compressor (2 LOP3) and only the carry word (weight 2) is popcounted:
//     count = 2 * sum popc(carry) + popc(ones_final)
// which halves the POPC work (2 per 16 bytes per stream instead of 4) for 4 extra LOP3.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

struct CountStream {
    uint32_t ones;
    int twos;
    __device__ __forceinline__ void reset() { ones = 0u; twos = 0; }
    // fold the 128 bits (a & m) into the stream
    __device__ __forceinline__ void add(const uint4 &a, const uint4 &m) {
        const uint32_t x = a.x & m.x, y = a.y & m.y, z = a.z & m.z, w = a.w & m.w;
        const uint32_t c1 = lop3_maj(ones, x, y);
        ones = lop3_xor3(ones, x, y);
        const uint32_t c2 = lop3_maj(ones, z, w);
        ones = lop3_xor3(ones, z, w);
        twos += __popc(c1) + __popc(c2);
    }
    // the same, with the two popcounts added on the FMA pipe (IMAD by a register that holds 1): the ring kernel is
    // bound by the ALU pipe (AND / LOP3 / IADD3), the FMA pipe is idle
    __device__ __forceinline__ void add_fma(const uint4 &a, const uint4 &m, int one) {
        const uint32_t x = a.x & m.x, y = a.y & m.y, z = a.z & m.z, w = a.w & m.w;
        const uint32_t c1 = lop3_maj(ones, x, y);
        ones = lop3_xor3(ones, x, y);
        const uint32_t c2 = lop3_maj(ones, z, w);
        ones = lop3_xor3(ones, z, w);
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(twos) : "r"(__popc(c1)), "r"(one));
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(twos) : "r"(__popc(c2)), "r"(one));
    }
    // the same without a mask (the whole row segment counts)
    __device__ __forceinline__ void add_fma_all(const uint4 &a, int one) {
        const uint32_t c1 = lop3_maj(ones, a.x, a.y);
        ones = lop3_xor3(ones, a.x, a.y);
        const uint32_t c2 = lop3_maj(ones, a.z, a.w);
        ones = lop3_xor3(ones, a.z, a.w);
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(twos) : "r"(__popc(c1)), "r"(one));
        asm("mad.lo.s32 %0, %1, %2, %0;" : "+r"(twos) : "r"(__popc(c2)), "r"(one));
    }
    __device__ __forceinline__ int total() const { return 2 * twos + __popc(ones); }
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// TMA bulk copy global -> shared of `bytes` (multiple of 16) signalled on an mbarrier.
__device__ __forceinline__ void stage_masks_tma(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                                uint64_t *bar) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                     "r"(bytes)
                     : "memory");
        // bulk copies are limited in size per instruction; issue in <= 32 KiB pieces
        uint32_t done = 0;
        while (done < bytes) {
            uint32_t n = bytes - done;
            if (n > 32768u) n = 32768u;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                    "r"(smem_u32(static_cast<char *>(smem_dst) + done)),
                "l"(static_cast<const char *>(gmem_src) + done), "r"(n), "r"(smem_u32(bar))
                : "memory");
            done += n;
        }
    }
    // every thread waits for phase 0 to complete
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar))
            : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// binary planes
// ---------------------------------------------------------------------------------------------
template <int NCOL, int CB>
__global__ void __launch_bounds__(256)
count_popc_kernel(const uint4 *__restrict__ value, const uint4 *__restrict__ valid, int64_t n_rows,
                  int w4, int logL, int nchunk, int w4pad, const uint4 *__restrict__ masks,
                  int4 *__restrict__ counts) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint4 *sm = reinterpret_cast<uint4 *>(smem_raw);  // [CB][2][w4pad]

    const int slot0 = blockIdx.y * CB;
    stage_masks_tma(sm, masks + size_t(slot0) * 2 * w4pad, uint32_t(CB) * 2u * uint32_t(w4pad) * 16u,
                    &bar);

    const int lane = threadIdx.x & 31;
    const int L = 1 << logL;
    const int sub = lane & (L - 1);
    const int rsub = lane >> logL;
    const int rpw = 32 >> logL;  // rows per warp iteration
    const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t ngroups = (n_rows + rpw - 1) / rpw;

    for (int64_t g = gwarp; g < ngroups; g += nwarps) {
        const int64_t row = g * rpw + rsub;
        const bool ok = row < n_rows;
        const uint4 *vrow = value + row * w4;
        const uint4 *drow = valid + row * w4;
        CountStream st[CB][4];
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) st[c][q].reset();

        for (int ch = 0; ch < nchunk; ++ch) {
            uint4 v[NCOL], d[NCOL];
            const int col0 = ch * (NCOL << logL) + sub;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int col = col0 + (j << logL);
                if (ok && col < w4) {
                    v[j] = ldg_stream(vrow + col);
                    d[j] = ldg_stream(drow + col);
                } else {
                    v[j] = make_uint4(0, 0, 0, 0);
                    d[j] = make_uint4(0, 0, 0, 0);
                }
            }
            // Scheduling fence: without it ptxas sinks each load next to its first use (2 loads in
            // flight per thread, long-scoreboard bound at 50% of DRAM); memory operations are not
            // moved across a warp barrier, so all 2*NCOL loads are issued back to back.
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CB; ++c) {
                const uint4 *m1p = sm + (c * 2 + 0) * w4pad + col0;
                const uint4 *m2p = sm + (c * 2 + 1) * w4pad + col0;
#pragma unroll
                for (int j = 0; j < NCOL; ++j) {
                    const uint4 m1 = m1p[j << logL];
                    const uint4 m2 = m2p[j << logL];
                    st[c][0].add(v[j], m1);
                    st[c][1].add(d[j], m1);
                    st[c][2].add(v[j], m2);
                    st[c][3].add(d[j], m2);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            int acc[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int x = st[c][q].total();
                for (int o = L >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                acc[q] = x;
            }
            if (ok && sub == 0)
                counts[size_t(slot0 + c) * n_rows + row] = make_int4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SNP planes: per site four marker rows _a,_t,_c,_g (Table.hs:184-191), valid shared by all four.
// ---------------------------------------------------------------------------------------------
template <int NCOL, int CB>
__global__ void __launch_bounds__(256)
count_popc_snp_kernel(const uint4 *__restrict__ hi, const uint4 *__restrict__ valid,
                      const uint4 *__restrict__ lo, int64_t n_sites, int w4, int logL, int nchunk,
                      int w4pad, const uint4 *__restrict__ masks, int4 *__restrict__ counts) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    uint4 *sm = reinterpret_cast<uint4 *>(smem_raw);

    const int slot0 = blockIdx.y * CB;
    stage_masks_tma(sm, masks + size_t(slot0) * 2 * w4pad, uint32_t(CB) * 2u * uint32_t(w4pad) * 16u,
                    &bar);

    const int lane = threadIdx.x & 31;
    const int L = 1 << logL;
    const int sub = lane & (L - 1);
    const int rsub = lane >> logL;
    const int rpw = 32 >> logL;
    const int64_t gwarp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t ngroups = (n_sites + rpw - 1) / rpw;
    const int64_t n_rows = 4 * n_sites;

    for (int64_t g = gwarp; g < ngroups; g += nwarps) {
        const int64_t site = g * rpw + rsub;
        const bool ok = site < n_sites;
        // st[c][mask][0..3] = A, T, C, N
        CountStream st[CB][2][4];
#pragma unroll
        for (int c = 0; c < CB; ++c)
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) st[c][m][q].reset();

        for (int ch = 0; ch < nchunk; ++ch) {
            uint4 a[NCOL], t[NCOL], cc[NCOL], d[NCOL];
            const int col0 = ch * (NCOL << logL) + sub;
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int col = col0 + (j << logL);
                uint4 h = make_uint4(0, 0, 0, 0), l = h, v = h;
                if (ok && col < w4) {
                    h = ldg_stream(hi + site * w4 + col);
                    l = ldg_stream(lo + site * w4 + col);
                    v = ldg_stream(valid + site * w4 + col);
                }
                d[j] = v;
                a[j] = make_uint4(~h.x & ~l.x & v.x, ~h.y & ~l.y & v.y, ~h.z & ~l.z & v.z, ~h.w & ~l.w & v.w);
                t[j] = make_uint4(h.x & l.x & v.x, h.y & l.y & v.y, h.z & l.z & v.z, h.w & l.w & v.w);
                cc[j] = make_uint4(~h.x & l.x & v.x, ~h.y & l.y & v.y, ~h.z & l.z & v.z, ~h.w & l.w & v.w);
            }
            __syncwarp();  // scheduling fence (see count_popc_kernel)
#pragma unroll
            for (int c = 0; c < CB; ++c) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint4 *mp = sm + (c * 2 + m) * w4pad + col0;
#pragma unroll
                    for (int j = 0; j < NCOL; ++j) {
                        const uint4 mk = mp[j << logL];
                        st[c][m][0].add(a[j], mk);
                        st[c][m][1].add(t[j], mk);
                        st[c][m][2].add(cc[j], mk);
                        st[c][m][3].add(d[j], mk);
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            int acc[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int x = st[c][m][q].total();
                    for (int o = L >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    acc[m][q] = x;
                }
            if (ok && sub == 0) {
                int4 *out = counts + size_t(slot0 + c) * n_rows + 4 * site;
                const int n1 = acc[0][3], n2 = acc[1][3];
                const int g1 = n1 - acc[0][0] - acc[0][1] - acc[0][2];
                const int g2 = n2 - acc[1][0] - acc[1][1] - acc[1][2];
                out[0] = make_int4(acc[0][0], n1, acc[1][0], n2);  // _a
                out[1] = make_int4(acc[0][1], n1, acc[1][1], n2);  // _t
                out[2] = make_int4(acc[0][2], n1, acc[1][2], n2);  // _c
                out[3] = make_int4(g1, n1, g2, n2);                // _g
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SNP planes, TMA-fed ring (the kernel config 3 runs).  The register-staged kernel above needs 16 count streams
// per 128 bits for a 4-value category (32 registers) next to 15 outstanding 128-bit loads per lane, runs at one or
// two CTAs per SM and alternates between waiting for its loads and ~1000 ALU instructions per lane: 0.29 of the
// HBM peak.  Here the loads hold no registers: tiles of consecutive sites (three contiguous blocks, one per
// plane) are streamed into a ring of shared-memory stages with cp.async.bulk, one `full` mbarrier per stage, and
// the warps do nothing but LDS + AND / LOP3 / POPC.  Measured on config 3 (profiles/README.md r1l): 0.997 -> 0.48 ms
// = 3.9 TB/s, ALU pipe 81% busy, XU (POPC) 67%: the scan is bound by the integer pipes (16 masked popcount streams
// per word = ~150 ALU-pipe instructions and 32 POPC per 48 bytes of planes), no longer by latency.
// Counted per mask: H = popc(hi&M), Lo = popc(lo&M), T = popc(hi&lo&M), N = popc(valid&M) (hi, lo are subsets of
// valid); A=0 C=1 G=2 T=3 in (hi, lo) gives  _t = T, _c = Lo - T, _g = H - T, _a = N - H - Lo + T.
// What did not work (same file history): one producer lane per CTA posting every tile (0.4 us per tile of three
// bulk copies, serial: 0.71 ms whatever the number of consumer warps; two / four producer warps 0.56 / 0.49 ms);
// a third carry-save level (half the POPC, +2 LOP3 per stream and column): 0.52 ms.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    uint32_t done = 0;
    while (done < bytes) {  // <= 32 KiB per instruction
        uint32_t n = bytes - done;
        if (n > 32768u) n = 32768u;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                "r"(smem_u32(static_cast<char *>(smem_dst) + done)),
            "l"(static_cast<const char *>(gmem_src) + done), "r"(n), "r"(smem_u32(bar))
            : "memory");
        done += n;
    }
}

constexpr int kRingWarps = 16;       // consumer warps per CTA (+ 1 producer warp)
constexpr int kRingMaxStages = 64;

// Every warp is consumer AND producer: it claims the next tile of the CTA in ring order (a shared counter), waits
// for the tile's stage, counts, and then refills that very stage with the tile n_stages further on (lane 0:
// expect_tx + one bulk copy per plane), so there is no producer warp and no `empty` barrier.
// A tile is kiter * (32 >> LOGL) consecutive rows (SNP: sites): one contiguous block per plane.
// SNP = true: planes (hi, lo, valid), four counts per mask (H, Lo, T, N), CB <= 2 pair-slots, four allele rows out.
// SNP = false: planes (value, valid), two counts per mask (P, N), CB <= 4 pair-slots -- the path of the few-comparison
// scan when several pair-slots share one read of the planes (16 count streams make the register-staged kernel
// latency-bound: 0.65 ms for 4 pair-slots on config 2's matrix against 0.21 ms for one).
// The L lane partials of the NP * 2 counts are packed two per register (counts <= S < 65536) and combined by a
// reduce-scatter over the lanes of a row (NP - 1 + ... shuffles instead of 2 * NP * log2 L); the packed totals go
// through a few words of shared memory so that a few lanes write one record each.
// PART: the 2 * CB masks of the pass partition ALL samples (one category whose values cover every genome, the config-3
// shape): the last mask's counts are the row totals minus the other masks' -- its streams run without the AND and
// without the mask load (the ring kernel is bound by the ALU pipe: 64 of its 132 ALU instructions per 48 bytes of
// planes are mask ANDs; this drops 16 of them).
template <bool SNP, int CB, int LOGL, int NW, bool PART = false>
__global__ void __launch_bounds__(NW * 32, 1)
count_popc_ring_kernel(const uint4 *__restrict__ pl0, const uint4 *__restrict__ valid,
                       const uint4 *__restrict__ pl2, int64_t n_units, int w4, int w4pad, int kiter,
                       int n_stages, int one, const uint4 *__restrict__ masks, int4 *__restrict__ counts) {
    constexpr int L = 1 << LOGL;
    constexpr int RPW = 32 >> LOGL;
    constexpr int NPL = SNP ? 3 : 2;    // planes
    constexpr int NQ = SNP ? 4 : 2;     // counts per mask
    constexpr int NP = NQ * CB;         // packed values per unit: [c][q], low half = first mask, high = second
    constexpr int LOGNP = NP == 2 ? 1 : (NP == 4 ? 2 : 3);
    static_assert(NP == 2 || NP == 4 || NP == 8, "2, 4 or 8 packed counts per row");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;                          // masks
    __shared__ __align__(8) uint64_t full_bar[kRingMaxStages];     // bulk copies -> consumers (tx bytes)
    __shared__ uint32_t red[NW][RPW][NP];
    __shared__ uint32_t next_tile;                                 // tiles are claimed in ring order
    // seen[s] = number of times stage s has been handed back (= the phase of full_bar[s] that is armed or about to
    // be).  A parity wait only tells the current phase from the one before it, and tiles are claimed without
    // regard to what has landed: a warp could claim tile it + n_stages while the data of tile it (same stage) is
    // still in flight and its wait on the opposite parity would return at once.  So the claimer of a tile first
    // waits until the stage's previous tile has been consumed.
    __shared__ volatile uint32_t seen[kRingMaxStages];
    uint4 *sm_masks = reinterpret_cast<uint4 *>(smem_raw);         // [CB][2][w4pad]
    const int tile_units = kiter * RPW;
    const uint32_t plane_cols = uint32_t(tile_units) * uint32_t(w4);   // uint4 per plane block
    uint4 *ring = sm_masks + size_t(CB) * 2 * w4pad;               // [stage][NPL][tile_units][w4]

    const int slot0 = blockIdx.y * CB;
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full_bar[s], 1); seen[s] = 0u; }
        next_tile = 0;
    }
    // initialises `bar`, fences the inits, __syncthreads, copies the masks, all threads wait
    stage_masks_tma(sm_masks, masks + size_t(slot0) * 2 * w4pad, uint32_t(CB) * 2u * uint32_t(w4pad) * 16u,
                    &bar);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_units + tile_units - 1) / tile_units;
    const uint32_t ust = uint32_t(n_stages);

    // fills stage `it % n_stages` with tile `it` of this CTA (one lane; the stage must be free).
    // Stage layout: SNP hi, lo, valid; binary value, valid.
    auto fill = [&](uint32_t it) {
        const int64_t tile = int64_t(blockIdx.x) + int64_t(it) * gridDim.x;
        if (tile >= n_tiles) return;
        const uint32_t s = it % ust;
        const int64_t u0 = tile * tile_units;
        const int64_t nu = (n_units - u0 < tile_units) ? (n_units - u0) : tile_units;
        const uint32_t bytes = uint32_t(nu) * uint32_t(w4) * 16u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[s])),
                     "r"(uint32_t(NPL) * bytes)
                     : "memory");
        uint4 *dst = ring + size_t(s) * NPL * plane_cols;
        const size_t goff = size_t(u0) * size_t(w4);
        bulk_g2s(dst, pl0 + goff, bytes, &full_bar[s]);
        if (SNP) bulk_g2s(dst + plane_cols, pl2 + goff, bytes, &full_bar[s]);
        bulk_g2s(dst + size_t(NPL - 1) * plane_cols, valid + goff, bytes, &full_bar[s]);
    };
    // prologue: the first n_stages tiles, spread over the warps
    if (lane == 0)
        for (uint32_t it = uint32_t(warp); it < ust; it += NW) fill(it);

    const int sub = lane & (L - 1);
    const int rsub = lane >> LOGL;
    const int64_t n_rows = (SNP ? 4 : 1) * n_units;
    for (;;) {
        // the next tile of this CTA in ring order goes to whichever warp is free first, so stages come back in
        // (nearly) the order they were filled
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(&next_tile, 1u);
        it = __shfl_sync(0xffffffffu, it, 0);
        const int64_t tile = int64_t(blockIdx.x) + int64_t(it) * gridDim.x;
        if (tile >= n_tiles) break;
        const uint32_t s = it % ust, phase = it / ust;
        if (lane == 0)
            while (seen[s] < phase) {}
        __syncwarp();
        mbar_wait(&full_bar[s], phase & 1u);
        const int64_t u0 = tile * tile_units;
        const int nu = int((n_units - u0 < tile_units) ? (n_units - u0) : tile_units);
        const uint4 *p0 = ring + size_t(s) * NPL * plane_cols;      // SNP: hi, binary: value
        const uint4 *p1 = p0 + plane_cols;                          // SNP: lo (binary: valid)
        const uint4 *pv = p0 + size_t(NPL - 1) * plane_cols;        // valid
        for (int g = 0; g < nu; g += RPW) {
            const int ls = g + rsub;          // unit inside the tile
            const bool ok = ls < nu;
            // SNP: st[c][mask][0..3] = H, Lo, T, N; binary: st[c][mask][0..1] = P, N
            CountStream st[CB][2][NQ];
#pragma unroll
            for (int c = 0; c < CB; ++c)
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int q = 0; q < NQ; ++q) st[c][m][q].reset();
            if (ok) {
                const uint32_t base = uint32_t(ls) * uint32_t(w4);
#pragma unroll 2
                for (int col = sub; col < w4; col += L) {
                    // hi / lo / value are subsets of valid (every producer of the planes guarantees it, capi.cu)
                    const uint4 v = pv[base + col];
                    const uint4 h = p0[base + col];
                    uint4 l = h, t = h;
                    if (SNP) {
                        l = p1[base + col];
                        t = make_uint4(h.x & l.x, h.y & l.y, h.z & l.z, h.w & l.w);
                    }
#pragma unroll
                    for (int c = 0; c < CB; ++c)
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            if (PART && c == CB - 1 && m == 1) {   // row totals (every sample is in some mask)
                                st[c][m][0].add_fma_all(h, one);
                                if (SNP) {
                                    st[c][m][1].add_fma_all(l, one);
                                    st[c][m][2].add_fma_all(t, one);
                                }
                                st[c][m][NQ - 1].add_fma_all(v, one);
                                continue;
                            }
                            const uint4 mk = sm_masks[(c * 2 + m) * w4pad + col];
                            st[c][m][0].add_fma(h, mk, one);
                            if (SNP) {
                                st[c][m][1].add_fma(l, mk, one);
                                st[c][m][2].add_fma(t, mk, one);
                            }
                            st[c][m][NQ - 1].add_fma(v, mk, one);
                        }
                }
            }
            // packed lane partials, then reduce-scatter over the L lanes of the unit
            uint32_t pk[NP];
#pragma unroll
            for (int c = 0; c < CB; ++c)
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    int second = st[c][1][q].total();
                    if (PART && c == CB - 1) {   // last mask = totals - the others (a lane's partial sums are linear too)
#pragma unroll
                        for (int cc = 0; cc < CB; ++cc) {
                            second -= st[cc][0][q].total();
                            if (cc != CB - 1) second -= st[cc][1][q].total();
                        }
                    }
                    pk[c * NQ + q] = uint32_t(st[c][0][q].total()) | (uint32_t(second) << 16);
                }
            int n = NP;
#pragma unroll
            for (int o = L >> 1; o > 0; o >>= 1) {
                if (n > 1) {
                    const int half = n >> 1;
                    const bool up = (sub & o) != 0;
#pragma unroll
                    for (int i = 0; i < NP / 2; ++i)
                        if (i < half) {
                            const uint32_t send = up ? pk[i] : pk[i + half];
                            const uint32_t keep = up ? pk[i + half] : pk[i];
                            pk[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    n = half;
                } else {
                    pk[0] += __shfl_xor_sync(0xffffffffu, pk[0], o);
                }
            }
            // lane `sub` now holds packed values [first, first + n) of its unit
            {
                constexpr int kSteps = LOGL < LOGNP ? LOGL : LOGNP;   // scatter steps taken
                const int first = (sub >> (LOGL - kSteps)) * (NP >> kSteps);
                const bool writer = (sub & ((1 << (LOGL - kSteps)) - 1)) == 0;
                if (writer) {
#pragma unroll
                    for (int i = 0; i < (NP >> kSteps); ++i) red[warp][rsub][first + i] = pk[i];
                }
            }
            __syncwarp();
            if (SNP) {
                for (int r = sub; r < 4 * CB; r += L) {
                    const int c = r >> 2, row = r & 3;
                    if (ok) {
                        const uint32_t H = red[warp][rsub][c * 4 + 0], Lo = red[warp][rsub][c * 4 + 1],
                                       T = red[warp][rsub][c * 4 + 2], N = red[warp][rsub][c * 4 + 3];
                        // packed arithmetic would borrow across the halves: unpack first
                        const int n1 = int(N & 0xffffu), n2 = int(N >> 16);
                        const int h1 = int(H & 0xffffu), h2 = int(H >> 16);
                        const int l1 = int(Lo & 0xffffu), l2 = int(Lo >> 16);
                        const int t1 = int(T & 0xffffu), t2 = int(T >> 16);
                        int x1, x2;
                        if (row == 0) { x1 = n1 - h1 - l1 + t1; x2 = n2 - h2 - l2 + t2; }        // _a
                        else if (row == 1) { x1 = t1; x2 = t2; }                                  // _t
                        else if (row == 2) { x1 = l1 - t1; x2 = l2 - t2; }                        // _c
                        else { x1 = h1 - t1; x2 = h2 - t2; }                                      // _g
                        counts[size_t(slot0 + c) * n_rows + 4 * (u0 + ls) + row] = make_int4(x1, n1, x2, n2);
                    }
                }
            } else {
                for (int c = sub; c < CB; c += L) {
                    if (ok) {
                        const uint32_t P = red[warp][rsub][c * 2 + 0], N = red[warp][rsub][c * 2 + 1];
                        counts[size_t(slot0 + c) * n_rows + (u0 + ls)] =
                            make_int4(int(P & 0xffffu), int(N & 0xffffu), int(P >> 16), int(N >> 16));
                    }
                }
            }
            __syncwarp();
        }
        // every lane is done with the stage (the shuffles above consumed its loads): the warp that emptied it refills
        // it with the tile n_stages further on -- no producer warp, no `empty` barrier, and the latency of posting
        // a tile (expect_tx + the bulk copies, ~0.4 us for one thread) is spread over all warps
        __syncwarp();
        if (lane == 0) {
            seen[s] = phase + 1u;
            fill(it + ust);
        }
    }
}

// "whole category" record = sum of the category's value records: the values partition the category's genomes
// (Table.hs:111-115,243-247), so the total mask never has to be scanned
__global__ void __launch_bounds__(256)
sum_slots_kernel(int4 *__restrict__ counts, int64_t n_rows, int slot0, int n, int tot_slot) {
    for (int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; r < n_rows; r += int64_t(gridDim.x) * blockDim.x) {
        int p = 0, q = 0;
        for (int s = 0; s < n; ++s) {
            const int4 a = counts[size_t(slot0 + s) * size_t(n_rows) + size_t(r)];
            p += a.x + a.z;
            q += a.y + a.w;
        }
        counts[size_t(tot_slot) * size_t(n_rows) + size_t(r)] = make_int4(p, q, 0, 0);
    }
}

// Launch one instantiation as a persistent grid: exactly (resident CTAs per SM) x (SMs) CTAs, each
// looping over row groups, so there is no partial last wave (the first version launched 6 CTAs/SM
// against an occupancy of 4 and lost a quarter of the time in a half-empty second wave).
template <class Kernel, class... Args>
cudaError_t launch_persistent(Kernel kernel, int64_t want_ctas, int n_batches, size_t smem,
                              int num_sms, cudaStream_t st, Args... args) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    // all batches (blockIdx.y) are co-resident waves of the same kernel: share the SMs between them
    int64_t resident = int64_t(num_sms) * per_sm;
    int64_t gx = resident / n_batches;
    if (gx < num_sms) gx = num_sms;
    if (gx > want_ctas) gx = want_ctas;
    if (gx < 1) gx = 1;
    kernel<<<dim3(unsigned(gx), unsigned(n_batches)), 256, smem, st>>>(args...);
    return cudaGetLastError();
}

template <int CB, bool SNP>
cudaError_t dispatch(int ncol, int64_t want_ctas, int n_batches, size_t smem, int num_sms,
                     cudaStream_t st, const uint4 *p0, const uint4 *p1, const uint4 *p2, int64_t n,
                     int w4, const PopcPlan &pl, const uint4 *masks, int4 *counts) {
#define FEHT_CASE(NC)                                                                              \
    case NC:                                                                                       \
        if constexpr (SNP)                                                                         \
            return launch_persistent(count_popc_snp_kernel<NC, CB>, want_ctas, n_batches, smem,    \
                                     num_sms, st, p0, p1, p2, n, w4, pl.logL, pl.nchunk, pl.w4pad, \
                                     masks, counts);                                               \
        else                                                                                       \
            return launch_persistent(count_popc_kernel<NC, CB>, want_ctas, n_batches, smem,        \
                                     num_sms, st, p0, p1, n, w4, pl.logL, pl.nchunk, pl.w4pad,     \
                                     masks, counts);
    switch (ncol) {
        FEHT_CASE(1)
        FEHT_CASE(2)
        FEHT_CASE(3)
        FEHT_CASE(4)
        FEHT_CASE(5)
        FEHT_CASE(6)
        FEHT_CASE(7)
        FEHT_CASE(8)
        default:
            return cudaErrorInvalidValue;
    }
#undef FEHT_CASE
}

// ---- ring kernel launch.  Returns cudaErrorNotSupported when the shape does not suit it (every warp needs a tile of
// its own plus a few in flight, and the packed lane partials need S < 65536); the caller then takes the
// register-staged kernel.
template <bool SNP, int CB, int LOGL, int NW, bool PART = false>
cudaError_t launch_ring_t(const uint4 *p0, const uint4 *valid, const uint4 *p2, int64_t n_units, int w4,
                          const PopcPlan &pl, const uint4 *masks, int n_batches, int4 *counts, int num_sms,
                          cudaStream_t st) {
    constexpr size_t kSmemTotal = 208 * 1024;   // + up to 17.5 KB static (barriers, reduction scratch)
    constexpr int RPW = 32 >> LOGL;
    const size_t mask_bytes = size_t(CB) * 2 * pl.w4pad * 16;
    if (mask_bytes >= kSmemTotal) return cudaErrorNotSupported;
    const size_t unit_bytes = size_t(SNP ? 48 : 32) * w4;            // all planes of one row / site
    // tiles of ~8 KB: large enough for the bulk copies, small enough for NW + 8 stages
    int kiter = int(size_t(8192) / (unit_bytes * RPW));
    if (kiter < 1) kiter = 1;
    // small inputs: keep several tiles per warp
    while (kiter > 1 && (n_units / (int64_t(kiter) * RPW)) < int64_t(num_sms) * NW * 2) kiter >>= 1;
    const size_t stage_bytes = unit_bytes * RPW * kiter;
    int n_stages = int((kSmemTotal - mask_bytes) / stage_bytes);
    if (n_stages > kRingMaxStages) n_stages = kRingMaxStages;
    if (n_stages < NW + 4) return cudaErrorNotSupported;
    const int64_t n_tiles = (n_units + int64_t(kiter) * RPW - 1) / (int64_t(kiter) * RPW);
    const size_t smem = mask_bytes + size_t(n_stages) * stage_bytes;
    auto kernel = count_popc_ring_kernel<SNP, CB, LOGL, NW, PART>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    int64_t gx = num_sms / n_batches;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    kernel<<<dim3(unsigned(gx), unsigned(n_batches)), NW * 32, smem, st>>>(
        p0, valid, p2, n_units, w4, pl.w4pad, kiter, n_stages, 1, masks, counts);
    return cudaGetLastError();
}

inline bool ring_enabled() {   // FEHT_SNP_RING=0 / FEHT_POPC_RING=0: register-staged kernels only (cross-checks)
    static const bool v = [] {
        const char *e = getenv("FEHT_SNP_RING"), *f = getenv("FEHT_POPC_RING");
        return !(e && atoi(e) == 0) && !(f && atoi(f) == 0);
    }();
    return v;
}

template <bool SNP, int CB>
cudaError_t launch_ring_any(const uint4 *p0, const uint4 *valid, const uint4 *p2, int64_t n_units, int w4,
                            const PopcPlan &pl, const uint4 *masks, int n_batches, int4 *counts, int num_sms,
                            cudaStream_t st, bool part = false) {
    if (!ring_enabled() || int64_t(w4) * 128 >= 65536) return cudaErrorNotSupported;
    if constexpr (SNP && CB == 2) {
        static const bool part_on = [] { const char *e = getenv("FEHT_SNP_PART"); return !(e && atoi(e) == 0); }();
        if (part && part_on && n_batches == 1) {
            cudaError_t e = cudaErrorNotSupported;
            switch (pl.logL) {
#define FEHT_RINGP(LL)                                                                                                            \
    case LL:                                                                                                                      \
        e = launch_ring_t<SNP, CB, LL, kRingWarps, true>(p0, valid, p2, n_units, w4, pl, masks, n_batches, counts, num_sms, st);  \
        if (e == cudaErrorNotSupported)                                                                                           \
            e = launch_ring_t<SNP, CB, LL, 8, true>(p0, valid, p2, n_units, w4, pl, masks, n_batches, counts, num_sms, st);       \
        break;
                FEHT_RINGP(0) FEHT_RINGP(1) FEHT_RINGP(2) FEHT_RINGP(3) FEHT_RINGP(4) FEHT_RINGP(5)
#undef FEHT_RINGP
                default: break;
            }
            if (e != cudaErrorNotSupported) return e;
        }
    }
    cudaError_t e = cudaErrorNotSupported;
    switch (pl.logL) {
        // 16 warps; 8 when the stages of 16 do not fit (wide rows)
#define FEHT_RING(LL)                                                                                                     \
    case LL:                                                                                                              \
        e = launch_ring_t<SNP, CB, LL, kRingWarps>(p0, valid, p2, n_units, w4, pl, masks, n_batches, counts, num_sms, st); \
        if (e == cudaErrorNotSupported)                                                                                   \
            e = launch_ring_t<SNP, CB, LL, 8>(p0, valid, p2, n_units, w4, pl, masks, n_batches, counts, num_sms, st);      \
        return e;
        FEHT_RING(0) FEHT_RING(1) FEHT_RING(2) FEHT_RING(3) FEHT_RING(4) FEHT_RING(5)
#undef FEHT_RING
        default: return cudaErrorNotSupported;
    }
}

template <bool SNP>
cudaError_t launch_any(const uint64_t *p0, const uint64_t *p1, const uint64_t *p2, int64_t n,
                       int64_t W, const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                       int4 *d_counts, int num_sms, cudaStream_t st, int *launches, bool masks_partition = false) {
    if (n == 0 || n_slots == 0) return cudaSuccess;
    const int w4 = int(W / 2);
    const int rpw = 32 >> plan.logL;
    const int64_t ngroups = (n + rpw - 1) / rpw;
    const int64_t want = (ngroups + 7) / 8;  // CTAs if every warp did one group
    const uint4 *a = reinterpret_cast<const uint4 *>(p0);
    const uint4 *b = reinterpret_cast<const uint4 *>(p1);
    const uint4 *c = reinterpret_cast<const uint4 *>(p2);
    // The masks of a pass live in shared memory (2 x w4pad x 16 B per pair-slot): ~900k samples is the most one
    // pair-slot can have; whole categories of wider matrices go to the GEMM engine, which streams its one-hot tiles.
    constexpr size_t kSmemMax = 200 * 1024;
    if (size_t(2) * plan.w4pad * 16 > kSmemMax) return cudaErrorInvalidConfiguration;
    // Batches of pair-slots share one read of the planes: 4 at a time (binary) or 2 (SNP: 8 count streams per
    // slot), then 2, then single ones.  Several pair-slots, and SNP, go to the TMA-fed ring kernel when its stages fit;
    // one binary pair-slot stays with the register-staged kernel (the config-2 path, 0.92 of the HBM peak).
    const size_t rows_out = SNP ? size_t(4) * n : size_t(n);
    int done = 0;
    auto run = [&](auto cb_tag, int n_batches) -> cudaError_t {
        constexpr int CB = decltype(cb_tag)::value;
        const uint4 *m = d_masks + size_t(done) * 2 * plan.w4pad;
        int4 *out = d_counts + size_t(done) * rows_out;
        cudaError_t e = cudaErrorNotSupported;
        // (one binary pair-slot through the ring: 0.2215 ms on config 2 against 0.2126 ms for the register-staged kernel)
        if (SNP || CB > 1)
            e = launch_ring_any<SNP, CB>(a, b, c, n, w4, plan, m, n_batches, out, num_sms, st,
                                         masks_partition && done == 0 && n_batches * CB == n_slots);
        if (e == cudaErrorNotSupported)
            e = dispatch<CB, SNP>(plan.ncol, want, n_batches, size_t(CB) * 2 * plan.w4pad * 16, num_sms, st, a, b, c, n,
                                  w4, plan, m, out);
        if (e == cudaSuccess) {
            if (launches) ++*launches;
            done += n_batches * CB;
        }
        return e;
    };
    auto fits = [&](int cb) { return size_t(cb) * 2 * plan.w4pad * 16 <= kSmemMax; };
    if constexpr (!SNP) {
        if (n_slots - done >= 4 && fits(4)) {
            cudaError_t e = run(std::integral_constant<int, 4>{}, (n_slots - done) / 4);
            if (e != cudaSuccess) return e;
        }
    }
    if (n_slots - done >= 2 && fits(2)) {
        cudaError_t e = run(std::integral_constant<int, 2>{}, (n_slots - done) / 2);
        if (e != cudaSuccess) return e;
    }
    if (done < n_slots) {
        cudaError_t e = run(std::integral_constant<int, 1>{}, n_slots - done);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace

PopcPlan make_popc_plan(int64_t W) {
    const int w4 = int(W / 2);
    PopcPlan best{0, 1, 1, 1};
    long best_cost = -1;
    for (int logL = 0; logL <= 5; ++logL) {
        const int L = 1 << logL;
        if (L < 8 && w4 >= 8) continue;  // keep >= 128 contiguous bytes per row segment
        for (int ncol = 1; ncol <= 8; ++ncol) {
            const int per_chunk = L * ncol;
            const int nchunk = (w4 + per_chunk - 1) / per_chunk;
            const long slots = long(per_chunk) * nchunk;
            // cost: wasted lane-slots first, then fewer chunks, then wider rows-per-warp split
            const long cost = (slots - w4) * 1000 + nchunk * 10 + (5 - logL);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = PopcPlan{logL, ncol, nchunk, int(slots)};
            }
        }
    }
    return best;
}

cudaError_t launch_count_popc(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows,
                              int64_t W, const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                              int4 *d_counts, int num_sms, cudaStream_t st, int *launches) {
    return launch_any<false>(d_value, d_valid, nullptr, n_rows, W, plan, d_masks, n_slots, d_counts,
                             num_sms, st, launches);
}

cudaError_t launch_count_popc_snp(const uint64_t *d_hi, const uint64_t *d_valid,
                                  const uint64_t *d_lo, int64_t n_sites, int64_t W,
                                  const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                                  int4 *d_counts, int num_sms, cudaStream_t st, int *launches, bool masks_partition) {
    return launch_any<true>(d_hi, d_valid, d_lo, n_sites, W, plan, d_masks, n_slots, d_counts,
                            num_sms, st, launches, masks_partition);
}

cudaError_t launch_sum_slots(int4 *d_counts, int64_t n_rows, int slot0, int n, int tot_slot, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    int64_t b = (n_rows + 255) / 256;
    if (b > int64_t(sm_count()) * 16) b = int64_t(sm_count()) * 16;
    sum_slots_kernel<<<int(b), 256, 0, st>>>(d_counts, n_rows, slot0, n, tot_slot);
    return cudaGetLastError();
}

}  // namespace feht
