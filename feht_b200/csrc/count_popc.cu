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

// pair-slots counted per read of the planes (register budget: 8 count streams per slot for SNP)
template <bool SNP> constexpr int kBatch = SNP ? 2 : 4;

template <bool SNP>
cudaError_t launch_any(const uint64_t *p0, const uint64_t *p1, const uint64_t *p2, int64_t n,
                       int64_t W, const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                       int4 *d_counts, int num_sms, cudaStream_t st, int *launches) {
    if (n == 0 || n_slots == 0) return cudaSuccess;
    const int w4 = int(W / 2);
    const int rpw = 32 >> plan.logL;
    const int64_t ngroups = (n + rpw - 1) / rpw;
    const int64_t want = (ngroups + 7) / 8;  // CTAs if every warp did one group
    const uint4 *a = reinterpret_cast<const uint4 *>(p0);
    const uint4 *b = reinterpret_cast<const uint4 *>(p1);
    const uint4 *c = reinterpret_cast<const uint4 *>(p2);
    constexpr int B = kBatch<SNP>;
    // The masks of a pass live in shared memory (2 x w4pad x 16 B per pair-slot): ~900k samples is the most one
    // pair-slot can have; whole categories of wider matrices go to the GEMM engine, which streams its one-hot tiles.
    constexpr size_t kSmemMax = 200 * 1024;
    if (size_t(2) * plan.w4pad * 16 > kSmemMax) return cudaErrorInvalidConfiguration;
    // batches of B pair-slots share one read of the planes; the remainder goes one at a time
    int done = 0;
    const int nb = size_t(B) * 2 * plan.w4pad * 16 <= kSmemMax ? n_slots / B : 0;
    if (nb > 0) {
        const size_t smem = size_t(B) * 2 * plan.w4pad * 16;
        cudaError_t e = dispatch<B, SNP>(plan.ncol, want, nb, smem, num_sms, st, a, b, c, n, w4, plan,
                                         d_masks, d_counts);
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
        done = nb * B;
    }
    if (done < n_slots) {
        const size_t smem = size_t(2) * plan.w4pad * 16;
        const size_t rows_out = SNP ? size_t(4) * n : size_t(n);
        cudaError_t e = dispatch<1, SNP>(plan.ncol, want, n_slots - done, smem, num_sms, st, a, b, c, n,
                                         w4, plan, d_masks + size_t(done) * 2 * plan.w4pad,
                                         d_counts + size_t(done) * rows_out);
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
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
                                  int4 *d_counts, int num_sms, cudaStream_t st, int *launches) {
    return launch_any<true>(d_hi, d_valid, d_lo, n_sites, W, plan, d_masks, n_slots, d_counts,
                            num_sms, st, launches);
}

cudaError_t launch_sum_slots(int4 *d_counts, int64_t n_rows, int slot0, int n, int tot_slot, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    int64_t b = (n_rows + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    sum_slots_kernel<<<int(b), 256, 0, st>>>(d_counts, n_rows, slot0, n, tot_slot);
    return cudaGetLastError();
}

}  // namespace feht
