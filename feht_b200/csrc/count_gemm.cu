// count_gemm.cu -- K2: contingency counts of whole metadata categories as an int8 tcgen05 GEMM.
//
// Replaces the same reference code as K1 (Comparison.hs:46-68,116-125 under /root/reference/src) for
// the many-comparison case (getAllPermutations, Comparison.hs:319-331): a category with V values
// yields V(V-1)/2 pairs + V one-vs-rest comparisons, but every one of their 2x2 tables is a sum or
// difference of the per-value counts
//     P[row][v] = #{s : value(s) = v, call(row,s) = '1'}      N[row][v] = #{s : value(s) = v, call valid}
// i.e. two dense contractions  [rows x samples](bits) . [samples x values](one-hot)  -- a GEMM with
// M = marker rows, K = samples, N = value columns (+ one "whole category" column per category for
// the one-vs-rest tables), exact in s32.  This is the only use of tensor cores in the library.
//
// Mapping onto sm_100a (DESIGN.md "K2"):
//   * A (1 bit per call in HBM) is never expanded in memory.  A TMA tensor map streams a
//     [256 rows x 128 B] box of each bit-plane into 128B-swizzled shared memory; 8 "expander" warps
//     (one thread per marker row = one TMEM lane) read 16 B of their row, spread the bits over bytes
//     in registers with ONE AND per 4 calls (byte = bit << j; the matching B entries are
//     1 << (7-j), so every hit adds exactly 128 and the epilogue shifts the s32 sums right by 7)
//     and write them straight into TENSOR MEMORY with tcgen05.st, where tcgen05.mma takes its A
//     operand from ([a_tmem] form).  Shared-memory bandwidth is therefore spent on the 1-bit data
//     and the B tiles only.
//   * B (one-hot, int8, K-major) is baked by the host into the exact 128B-swizzled core-matrix
//     byte order the UMMA shared-memory descriptor expects, one contiguous N x 128 B block per
//     128-sample stage, so a plain 1-D bulk copy (cp.async.bulk, UBLKCP) per stage suffices; the
//     K order inside a stage is permuted to match the cheap bit expansion (see bake_b()).
//   * D: four accumulators (2 row tiles x {value, valid} plane) x N <= 64 columns live in TMEM (256 of the 512
//     columns; the rest are the A stage buffers).  tcgen05.mma.cta_group::1 (kind::i8: M=128, N, K=32; kind::mxf4: K=64)
//     is issued by single threads; completion is tracked with tcgen05.commit on mbarriers.  Every MMA accumulates: the
//     accumulators start at zero and the epilogue zeroes what it has read.  The epilogue reads them back with
//     tcgen05.ld and writes the int4 (P_even, N_even, P_odd, N_odd) pair-slot records K3 consumes -- the same layout
//     K1 produces, so K3/K4 do not care which engine counted.
//   * warp roles: 0-7 expand + epilogue, 8 TMA producer for the planes, 9 bulk-copy producer for B, 10-13 MMA
//     issuers (+ TMEM allocation).  Persistent CTAs, one per SM, static tile round-robin.
//
// What paces it (measured on config 4, 2M x 20k, N = 64; profiles/README.md r1c, r1s, r1t and the round-2 probes):
//   * the planes alone stream at 6.7 TB/s through the TMA boxes (1.49 ms: the kernel with the stage loop cut out);
//   * the hand-over skeleton -- every barrier, no expansion, no MMA -- already takes 1.92 ms (~440 cycles per
//     128-sample stage), and that does not move with the depth of any ring (A buffers 2/3/4, B ring 6/12 -- 3 or 4
//     slots are too few --, plane chunks as two 64 KB tiles or five 32 KB half tiles), nor with spinning on test_wait
//     instead of try_wait, nor with half the MMAs (a value-plane-only variant that counts the valid calls from the
//     missing ones: 2.08 ms without its scan);
//   * on top of it, a stage's MMAs cost the ISSUING THREAD ~530 cycles + ~48 per MMA, so round 1's four issuers (one
//     per accumulator, each on every stage) ran at 650 cycles per stage = 2.83 ms.  Round 2, mxf4: one issuer per
//     stage in flight (three, one per A buffer), each issuing all 8 MMAs of its stage: 2.28 ms = 4.4 TB/s of planes,
//     0.67 of the measured HBM peak; the i8 flavour keeps an issuer per accumulator (3.3 ms).
#include <cuda.h>

#include <cstring>

#include "common.cuh"

namespace feht {
namespace {

constexpr int kTileRows = 256;        // two M = 128 MMA tiles per CTA tile
constexpr int kChunkBytes = 128;      // bytes of a plane row per TMA box = 1024 samples
constexpr int kStageSamples = 128;    // samples per pipeline stage = 16 B of a plane row = 4 MMA K steps
constexpr int kStagesPerChunk = kChunkBytes * 8 / kStageSamples;  // 8
constexpr int kNMax = 64;             // value columns per pass
constexpr int kBStageBytes = kNMax * 128;
constexpr int kBStages = 6;
constexpr int kPlaneTileBytes = kTileRows * kChunkBytes;  // 32 KiB
constexpr int kExpandWarps = 8;
constexpr int kThreads = (kExpandWarps + 3) * 32;
#ifndef FEHT_MMA_WARPS
#define FEHT_MMA_WARPS 4
#endif
constexpr int kMmaWarps = FEHT_MMA_WARPS;                       // i8 / mxf4 kernels: warps reserved for MMA issue (kABufs of them issue: 3 mxf4, 2 i8)
constexpr int kThreads2 = (kExpandWarps + 2 + kMmaWarps) * 32;
constexpr int kTmemCols = 512;
constexpr int kTmemAccBase = 0;       // + (mt * 2 + plane) * 64
constexpr int kTmemABase = 256;       // + buf * 128 + (mt * 2 + plane) * 32

// shared memory carve-up (dynamic, 1024-byte aligned base)
constexpr int kSmemPlanes = 0;                                     // [2 buf][2 plane][32 KiB]
constexpr int kSmemB = kSmemPlanes + 4 * kPlaneTileBytes;          // [kBStages][8 KiB]
constexpr int kSmemBars = kSmemB + kBStages * kBStageBytes;        // mbarriers
constexpr int kNumBars = 2 + 2 + kBStages * 2 + 3 + 3 + 2;
constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;
constexpr int kSmemTotal = kSmemTmemPtr + 16 + 1024;               // + slack for manual alignment

struct GemmPass {
    int32_t slot_base;   // first pair-slot written by this pass
    int32_t n_pairs;     // pair-slots written (<= n_pad / 2)
    int32_t n_pad;       // GEMM N of this pass: multiple of 16, <= 64
    int32_t pad;
    int64_t b_offset;    // byte offset of the pass's baked B blocks: [n_stages][n_pad * 128]
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *map, int x, int y,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void *gsrc, uint32_t bytes,
                                             uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]; one thread issues for the CTA
__device__ __forceinline__ void mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same contraction with 4-bit operands: kind::mxf4 (e2m1 x e2m1, one UE8M0 scale per 32 K elements, here all
// 2^0), M = 128, K = 64, FP32 accumulators -- exact below 2^24 samples.  Half the operand bytes per sample, and
// K2 is bound by operand bytes (profiles/README.md r1g), so a stage of 128 samples needs 8 MMAs instead of 16.
__device__ __forceinline__ void mma_fp4_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t sf_tmem, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], [%1], %2, %3, [%4], [%4], p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(sf_tmem), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
        : "memory");
}

// 32 calls (one word of a bit-plane row) -> 8 TMEM columns of 4 uint8 each.  Column j, byte b holds
// call (8*b + j) as the value (bit << j): y_j = w & (0x01010101 << j), ONE logic op per column.
// bake_b() orders B's K dimension the same way and stores 1 << (7 - j) there, so a hit adds 128.
__device__ __forceinline__ void expand_store(uint32_t taddr, uint32_t w) {
    uint32_t y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = w & (0x01010101u << j);
    tmem_st8(taddr, y);
}

// FP4 flavour: 32 calls -> 4 TMEM columns of 8 e2m1 nibbles.  Column j, nibble t holds call 4*t + j as the nibble
// 1 << j for j < 3 (e2m1 0.5, 1.0, 2.0) and, shifted down, 1 (0.5) for j = 3 (bit 3 of a nibble is the sign);
// gemm_bake_b() stores the reciprocal (2.0, 1.0, 0.5, 2.0) at the matching K position, so every hit adds 1.0.
__device__ __forceinline__ void expand4_fp4(uint32_t (&y)[8], int at, uint32_t w) {
    y[at + 0] = w & 0x11111111u;
    y[at + 1] = w & 0x22222222u;
    y[at + 2] = w & 0x44444444u;
    y[at + 3] = (w >> 3) & 0x11111111u;
}
__device__ __forceinline__ void expand_store_fp4(uint32_t taddr, uint32_t w0, uint32_t w1) {
    uint32_t y[8];
    expand4_fp4(y, 0, w0);
    expand4_fp4(y, 4, w1);
    tmem_st8(taddr, y);
}

// UMMA shared-memory descriptor: K-major, 128B swizzle, 8-row core-matrix groups 1024 B apart
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFFu) >> 4);       // start address, 16-byte units
    d |= uint64_t(1) << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= uint64_t(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 B
    d |= uint64_t(1) << 46;                           // descriptor version (sm_100)
    d |= uint64_t(2) << 61;                           // SWIZZLE_128B
    return d;
}

// ------------------------------------------------------------------------------------ the kernel
// FP4 = false: kind::i8, one B block (128 samples) per A stage.  FP4 = true: kind::mxf4, A stages still hold 128
// samples (16 B of a plane row -> 16 TMEM columns per accumulator), one B block (a full 128-byte swizzle row = 256
// e2m1 values) serves two consecutive A stages.
template <bool FP4>
__global__ void __launch_bounds__(kThreads2, 1)
count_gemm_kernel(const __grid_constant__ CUtensorMap tm_value, const __grid_constant__ CUtensorMap tm_valid,
                  const uint8_t *__restrict__ bmat, const GemmPass *__restrict__ passes, int n_stages,
                  int64_t n_rows, int n_tiles, int4 *__restrict__ counts) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t bars = base + kSmemBars;
    // barrier indices
    const uint32_t a_full = bars + 0 * 8;             // [2]  planes landed in smem
    const uint32_t a_empty = bars + 2 * 8;            // [2]  expanders done with a plane buffer
    const uint32_t b_full = bars + 4 * 8;             // [kBStages]
    const uint32_t b_empty = b_full + kBStages * 8;   // [kBStages]
    const uint32_t ta_full = b_empty + kBStages * 8;  // [2]  A stage written to TMEM
    const uint32_t ta_empty = ta_full + 3 * 8;        // [2..3]  MMAs of the stage retired
    const uint32_t acc_full = ta_empty + 3 * 8;       // accumulators complete
    const uint32_t acc_empty = acc_full + 8;          // epilogue drained TMEM
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(base_ptr + kSmemTmemPtr);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmPass pass = passes[blockIdx.y];
    const int n_pad = pass.n_pad;
    const uint32_t b_bytes = uint32_t(n_pad) * 128u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i * 8, 1);
            mbar_init(a_empty + i * 8, kExpandWarps);
        }
        // one issuer handles a whole stage (all four accumulators): it alone hands the A buffer back; a B block serves
        // kAPerB consecutive stages, i.e. that many issuers
        // mxf4: one issuer handles a whole stage (it alone hands the A buffer back; a B block serves two stages = two
        // issuers; three issuers close a tile).  i8: four issuers, one per accumulator, each on every stage.
        for (int i = 0; i < 3; ++i) {
            mbar_init(ta_full + i * 8, kExpandWarps);
            mbar_init(ta_empty + i * 8, FP4 ? 1 : 4);
        }
        for (int i = 0; i < kBStages; ++i) {
            mbar_init(b_full + i * 8, 1);
            mbar_init(b_empty + i * 8, FP4 ? 2 : 4);
        }
        mbar_init(acc_full, FP4 ? 3 : 4);
        mbar_init(acc_empty, kExpandWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kExpandWarps + 2) {  // the MMA warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(const_cast<uint32_t *>(tmem_ptr_smem))), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int n_chunks = (n_stages + kStagesPerChunk - 1) / kStagesPerChunk;
    constexpr int kAPerB = FP4 ? 2 : 1;          // A stages served by one B block
    constexpr int kAColsQ = FP4 ? 16 : 32;       // TMEM columns of one accumulator's A operand per stage
    constexpr int kABuf = 4 * kAColsQ;           // TMEM columns of one A stage buffer
    constexpr int kABufs = FP4 ? 3 : 2;          // A stage buffers in TMEM: 256 + 3 * 64 + 8 / 256 + 2 * 128 columns
    constexpr int kTmemSF = kTmemABase + kABufs * kABuf;   // FP4: eight columns of unit scale factors (0x7F = 2^0)

    if (warp < kExpandWarps) {
        // every MMA accumulates (the issuers work on different stages of a tile at the same time, so none of them is
        // "the first"): the accumulators start at zero and the epilogue zeroes what it has read
        const uint32_t la = uint32_t((warp & 3) * 32) << 16;
        uint32_t zero[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) zero[j] = 0u;
        for (int c0 = 0; c0 < 128; c0 += 8) tmem_st8(tmem_base + la + kTmemAccBase + (threadIdx.x >> 7) * 128 + c0, zero);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp < kExpandWarps) {
        // =============================================== expanders + epilogue: thread = marker row
        const int r = threadIdx.x;                 // row inside the 256-row tile
        const int mt = r >> 7;                     // M tile
        const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
        const uint32_t row_off = uint32_t(r) * kChunkBytes;
        const uint32_t sw = uint32_t(r & 7);
        uint32_t chunk_ctr = 0, stage_ctr = 0, tile_ctr = 0;
        if (FP4 && warp < 4) {
            // every byte of these columns is the UE8M0 code of 1.0, whichever of them an MMA picks as its scales
            uint32_t one[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) one[j] = 0x7F7F7F7Fu;
            tmem_st8(tmem_base + lane_addr + kTmemSF, one);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_ctr) {
            for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                const uint32_t cb = chunk_ctr & 1;
                mbar_wait(a_full + cb * 8, (chunk_ctr >> 1) & 1);
                const unsigned char *pv = base_ptr + kSmemPlanes + (cb * 2 + 0) * kPlaneTileBytes + row_off;
                const unsigned char *pd = base_ptr + kSmemPlanes + (cb * 2 + 1) * kPlaneTileBytes + row_off;
                const int st_n = min(kStagesPerChunk, n_stages - ch * kStagesPerChunk);
                for (int st = 0; st < st_n; ++st, ++stage_ctr) {
                    const uint32_t phys = (uint32_t(st) ^ sw) << 4;  // TMA 128B swizzle: chunk ^= row % 8
                    const uint4 v = *reinterpret_cast<const uint4 *>(pv + phys);
                    const uint4 d = *reinterpret_cast<const uint4 *>(pd + phys);
                    const uint32_t tb = stage_ctr % kABufs;
                    mbar_wait(ta_empty + tb * 8, ((stage_ctr / kABufs) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + lane_addr + kTmemABase + tb * kABuf + (mt * 2) * kAColsQ;
                    if (FP4) {
                        expand_store_fp4(ta + 0, v.x, v.y);
                        expand_store_fp4(ta + 8, v.z, v.w);
                        expand_store_fp4(ta + 16 + 0, d.x, d.y);
                        expand_store_fp4(ta + 16 + 8, d.z, d.w);
                    } else {
                        expand_store(ta + 0, v.x);
                        expand_store(ta + 8, v.y);
                        expand_store(ta + 16, v.z);
                        expand_store(ta + 24, v.w);
                        expand_store(ta + 32 + 0, d.x);
                        expand_store(ta + 32 + 8, d.y);
                        expand_store(ta + 32 + 16, d.z);
                        expand_store(ta + 32 + 24, d.w);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ta_full + tb * 8);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(a_empty + cb * 8);
            }
            // ---- epilogue: accumulators -> pair-slot records
            mbar_wait(acc_full, tile_ctr & 1);
            tc_fence_after();
            const int64_t row = int64_t(tile) * kTileRows + r;
            const uint32_t acc_p = tmem_base + lane_addr + kTmemAccBase + (mt * 2 + 0) * 64;
            const uint32_t acc_n = tmem_base + lane_addr + kTmemAccBase + (mt * 2 + 1) * 64;
            for (int c0 = 0; c0 < n_pad; c0 += 16) {
                uint32_t P[16], N[16];
                tmem_ld16(acc_p + c0, P);
                tmem_ld16(acc_n + c0, N);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                {   // the next tile accumulates onto zeros
                    uint32_t zero[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) zero[j] = 0u;
                    tmem_st8(acc_p + c0, zero); tmem_st8(acc_p + c0 + 8, zero);
                    tmem_st8(acc_n + c0, zero); tmem_st8(acc_n + c0 + 8, zero);
                }
                if (row < n_rows) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int pair = (c0 >> 1) + j;
                        if (pair < pass.n_pairs) {
                            // i8: every hit added 128 (expand_store / gemm_bake_b); FP4: 1.0 in FP32
                            auto cnt = [](uint32_t x) { return FP4 ? __float2int_rn(__uint_as_float(x)) : int(x >> 7); };
                            counts[size_t(pass.slot_base + pair) * size_t(n_rows) + size_t(row)] =
                                make_int4(cnt(P[2 * j]), cnt(N[2 * j]), cnt(P[2 * j + 1]), cnt(N[2 * j + 1]));
                        }
                    }
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    } else if (warp == kExpandWarps) {
        // =============================================== TMA producer: bit-plane boxes
        if (lane == 0) {
            uint32_t chunk_ctr = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                    const uint32_t cb = chunk_ctr & 1;
                    mbar_wait(a_empty + cb * 8, ((chunk_ctr >> 1) & 1) ^ 1);
                    mbar_expect_tx(a_full + cb * 8, 2u * kPlaneTileBytes);
                    tma_load_2d(base + kSmemPlanes + (cb * 2 + 0) * kPlaneTileBytes, &tm_value,
                                ch * kChunkBytes, tile * kTileRows, a_full + cb * 8);
                    tma_load_2d(base + kSmemPlanes + (cb * 2 + 1) * kPlaneTileBytes, &tm_valid,
                                ch * kChunkBytes, tile * kTileRows, a_full + cb * 8);
                }
            }
        }
    } else if (warp == kExpandWarps + 1) {
        // =============================================== bulk-copy producer: baked one-hot B blocks
        if (lane == 0) {
            uint32_t stage_ctr = 0;
            const uint8_t *bsrc = bmat + pass.b_offset;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int n_bblocks = (n_stages + kAPerB - 1) / kAPerB;
                for (int s = 0; s < n_bblocks; ++s, ++stage_ctr) {
                    const uint32_t bs = stage_ctr % kBStages;
                    mbar_wait(b_empty + bs * 8, ((stage_ctr / kBStages) & 1) ^ 1);
                    mbar_expect_tx(b_full + bs * 8, b_bytes);
                    bulk_load_1d(base + kSmemB + bs * kBStageBytes, bsrc + size_t(s) * b_bytes, b_bytes,
                                 b_full + bs * 8);
                }
            }
        }
    } else {
        // =============================================== MMA issuers.  Every MMA accumulates (the accumulators start at zero
        // and the epilogue zeroes what it has read), so no issuer is "the first" of a tile.
        const int mw = warp - (kExpandWarps + 2);
        // instruction descriptor, N = n_pad, M = 128, both operands K-major.  i8: D s32, A/B unsigned 8-bit.
        // mxf4 (block-scaled layout): A/B e2m1 (format 1), UE8M0 scales, scale ids 0, K = 64.
        const uint32_t idesc = FP4 ? ((1u << 7) | (1u << 10) | (uint32_t(n_pad >> 3) << 17) | (1u << 23) |
                                      (uint32_t(128 >> 4) << 24))
                                   : ((2u << 4) | (uint32_t(n_pad >> 3) << 17) | (uint32_t(128 >> 4) << 24));
        if (FP4) {
            // mxf4: one issuer per STAGE in flight.  A stage's hand-over (two mbarrier waits, a fence, the issues, the
            // commits) costs an issuing thread ~530 cycles + ~48 per MMA whatever the tensor pipe does
            // (profiles/README.md r1s), so issuer w takes every third stage and issues ALL of its MMAs (2 M tiles x
            // 2 planes x 2 K steps): the fixed part is paid once per 8 MMAs and the issuers overlap each other's waits
            // (config 4: 2.83 -> 2.28 ms).  Stages of one tile accumulate into the same four accumulators from
            // different threads: sums of integers below 2^24 do not depend on the order, and the tensor pipe retires
            // one MMA at a time (counts bit-identical in every engine test).  As many issuers as A stage buffers,
            // issuer w <-> buffer w: its stages are consecutive phases of ITS ta_full barrier, and with a B ring of 6
            // it meets every B slot it uses once per ring cycle -- a parity wait cannot tell a phase from the one two
            // before it (four issuers on three buffers ran a ring ahead of the expanders, and hung).
            if (lane == 0 && mw < kABufs) {
                const uint32_t n_bblocks = uint32_t((n_stages + kAPerB - 1) / kAPerB);
                uint32_t tile_ctr = 0;
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_ctr) {
                    mbar_wait(acc_empty, (tile_ctr & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t tile_stage0 = tile_ctr * uint32_t(n_stages);
                    for (int s = int((uint32_t(mw) + uint32_t(kABufs) - tile_stage0 % kABufs) % kABufs); s < n_stages; s += kABufs) {
                        const uint32_t stage_ctr = tile_stage0 + uint32_t(s);   // stage_ctr % kABufs == mw
                        const uint32_t bblock_ctr = tile_ctr * n_bblocks + uint32_t(s / kAPerB);
                        const uint32_t tb = stage_ctr % kABufs;
                        const int half = s % kAPerB;                 // which part of the B block this A stage uses
                        const uint32_t bs = bblock_ctr % kBStages;
                        mbar_wait(ta_full + tb * 8, (stage_ctr / kABufs) & 1);
                        mbar_wait(b_full + bs * 8, (bblock_ctr / kBStages) & 1);
                        tc_fence_after();
                        const uint64_t bdesc = make_b_desc(base + kSmemB + bs * kBStageBytes);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)   // q = mt * 2 + plane
                                mma_fp4_ts(tmem_base + kTmemAccBase + q * 64,
                                           tmem_base + kTmemABase + tb * kABuf + q * kAColsQ + ks * 8,
                                           bdesc + uint64_t((half * 2 + ks) * 2), idesc, tmem_base + kTmemSF, 1u);
                        }
                        tc_commit(ta_empty + tb * 8);
                        tc_commit(b_empty + bs * 8);
                        // a trailing B block that serves one stage only still owes its barrier the second arrival
                        if (half == 0 && s == n_stages - 1) tc_commit(b_empty + bs * 8);
                    }
                    tc_commit(acc_full);   // arrives once this issuer's MMAs of the tile have retired
                }
            }
        } else {
            // i8: one issuer per ACCUMULATOR (2 M tiles x 2 planes), each on every stage -- with only two A stage
            // buffers (tensor memory is full: 256 accumulator + 2 x 128 operand columns) a stage-parallel scheme has
            // two issuers, which is slower (4.1 ms against 3.3 ms on config 4).
            if (lane == 0) {
                uint32_t stage_ctr = 0, tile_ctr = 0;
                for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_ctr) {
                    mbar_wait(acc_empty, (tile_ctr & 1) ^ 1);
                    tc_fence_after();
                    for (int s = 0; s < n_stages; ++s, ++stage_ctr) {
                        const uint32_t tb = stage_ctr % kABufs;
                        const uint32_t bs = stage_ctr % kBStages;
                        mbar_wait(ta_full + tb * 8, (stage_ctr / kABufs) & 1);
                        mbar_wait(b_full + bs * 8, (stage_ctr / kBStages) & 1);
                        tc_fence_after();
                        const uint64_t bdesc = make_b_desc(base + kSmemB + bs * kBStageBytes);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_i8_ts(tmem_base + kTmemAccBase + mw * 64,
                                      tmem_base + kTmemABase + tb * kABuf + mw * kAColsQ + ks * 8,
                                      bdesc + uint64_t(ks * 2), idesc, 1u);
                        tc_commit(ta_empty + tb * 8);
                        tc_commit(b_empty + bs * 8);
                    }
                    tc_commit(acc_full);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kExpandWarps + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------ third flavour
// kind::f8f6f4 with e5m2 operands: BOTH planes in ONE contraction.  value is a subset of valid, so a call is one of
// three things -- missing, valid and absent, valid and present -- and one 8-bit operand can carry it as
//     0.0 (0x00)        4096.0 (0x6C)        1.0 (0x3C)
// against a one-hot B of 1.0: the FP32 accumulator of a column then holds  P + 4096 * (N - P)  exactly as long as the
// column has at most 4,095 member samples (4095 * 4096 + 4095 = 2^24 - 1), and the epilogue splits it with a mask
// and a shift.  Larger value groups are spread over several columns whose sums the epilogue adds; the "whole
// category" record is the sum of the category's value records, so it needs no column of its own.  An e5m2 MMA costs
// what an int8 MMA costs (measured: 5.41 vs 5.25 ms on config 4 with the int8 data flow), and this flavour issues
// half of them per sample: one operand set instead of two.  What it costs instead: the expanders need two shifts, two
// masks and two multiplies per 4 calls (the int8 / fp4 flavours: one AND per plane), 2,450 warp instructions per
// 128-sample stage, and the kernel is bound by their issue (ALU pipe 58%, tensor pipe 22%; 16 expander warps instead
// of 8 changed 5.01 into 4.87 ms): config 4 counts in 4.87 ms against 3.85 ms for kind::mxf4, config 5 in 5.98 against
// 4.78.  Kept as FEHT_ENGINE_GEMM_F8 (bit-identical counts, tests/test_gpu_parity.py); AUTO does not pick it.
// A 2-bit interleaved plane layout (missing / absent / present per call) would bring the expansion down to a shift,
// a mask and one multiply by 0x18 (codes 0x18 = 2^-9 and 0x48 = 8 = 4096 x 2^-9) -- a layout change for every kernel.
constexpr int kF8NMax = 128;
constexpr int kF8BStages = 4;
constexpr int kF8BStageBytes = kF8NMax * 128;
constexpr int kF8ExpandWarps = 16;
constexpr int kF8Threads = (kF8ExpandWarps + 3) * 32;
constexpr int kF8ABufs = 4;                  // A stage buffers in TMEM: 2 x 128 accumulator + 4 x 64 operand columns
constexpr int kF8SmemBars = kSmemB + kF8BStages * kF8BStageBytes;
constexpr int kF8NumBars = 2 + 2 + kF8BStages * 2 + kF8ABufs * 2 + 2;
constexpr int kF8SmemTmemPtr = kF8SmemBars + kF8NumBars * 8;
constexpr int kF8SmemTotal = kF8SmemTmemPtr + 16 + 1024;

struct GemmPassF8 {
    int32_t n_cols;            // GEMM columns in use
    int32_t n_pad;             // N of the MMA: multiple of 16, <= 128
    int64_t b_offset;          // byte offset of the pass's baked B blocks: [n_stages][n_pad * 128]
    int32_t slot[kF8NMax];     // pair-slot (index into counts) the column's value belongs to
    int32_t tot_slot[kF8NMax]; // "whole category" slot to write after this column, or -1
    uint8_t flags[kF8NMax];    // 1: second value of its pair-slot; 2: last column of its value; 4: last column of its
                               // pair-slot; 8: last column of its category
};

__device__ __forceinline__ void mma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return r;
}
// 32 calls (one word of each plane) -> 8 TMEM columns of 4 e5m2 bytes; column j, byte b holds call 8*b + j (the
// order of expand_store, so B is baked at the same K positions)
__device__ __forceinline__ void expand_store_f8(uint32_t taddr, uint32_t v, uint32_t d) {
    uint32_t y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t x = (v >> j) & 0x01010101u, t = (d >> j) & 0x01010101u;
        y[j] = t * 0x6cu - x * 0x30u;     // valid: 0x6c (4096.0); valid and present: 0x3c (1.0)
    }
    tmem_st8(taddr, y);
}

__global__ void __launch_bounds__(kF8Threads, 1)
count_gemm_f8_kernel(const __grid_constant__ CUtensorMap tm_value, const __grid_constant__ CUtensorMap tm_valid,
                     const uint8_t *__restrict__ bmat, const GemmPassF8 *__restrict__ passes, int n_stages,
                     int64_t n_rows, int n_tiles, int4 *__restrict__ counts) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char *base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t bars = base + kF8SmemBars;
    const uint32_t a_full = bars;                          // [2]  planes landed in smem
    const uint32_t a_empty = a_full + 2 * 8;               // [2]  expanders done with a plane buffer
    const uint32_t b_full = a_empty + 2 * 8;               // [kF8BStages]
    const uint32_t b_empty = b_full + kF8BStages * 8;      // [kF8BStages]
    const uint32_t ta_full = b_empty + kF8BStages * 8;     // [kF8ABufs]  A stage written to TMEM
    const uint32_t ta_empty = ta_full + kF8ABufs * 8;      // [kF8ABufs]  MMAs of the stage retired
    const uint32_t acc_full = ta_empty + kF8ABufs * 8;
    const uint32_t acc_empty = acc_full + 8;
    volatile uint32_t *tmem_ptr_smem = reinterpret_cast<volatile uint32_t *>(base_ptr + kF8SmemTmemPtr);
    __shared__ int32_t s_slot[kF8NMax], s_tot[kF8NMax];
    __shared__ uint8_t s_flags[kF8NMax];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmPassF8 *pass = passes + blockIdx.y;
    const int n_pad = pass->n_pad, n_cols = pass->n_cols;
    const uint32_t b_bytes = uint32_t(n_pad) * 128u;
    for (int i = threadIdx.x; i < kF8NMax; i += blockDim.x) {
        s_slot[i] = pass->slot[i];
        s_tot[i] = pass->tot_slot[i];
        s_flags[i] = pass->flags[i];
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + i * 8, 1);
            mbar_init(a_empty + i * 8, kF8ExpandWarps);
        }
        for (int i = 0; i < kF8ABufs; ++i) {
            mbar_init(ta_full + i * 8, kF8ExpandWarps);
            mbar_init(ta_empty + i * 8, 1);
        }
        for (int i = 0; i < kF8BStages; ++i) {
            mbar_init(b_full + i * 8, 1);
            mbar_init(b_empty + i * 8, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kExpandWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kF8ExpandWarps + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(const_cast<uint32_t *>(tmem_ptr_smem))), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int n_chunks = (n_stages + kStagesPerChunk - 1) / kStagesPerChunk;
    constexpr int kABuf = 64;                    // TMEM columns of one A stage buffer: 2 M tiles x 32
    constexpr uint32_t kAcc = 0, kA = 256;

    if (warp < kF8ExpandWarps) {
        // =============================================== expanders (16 warps: two per 32 marker rows, each taking two of
        // the four words of a stage -- the e5m2 bytes cost ~6 instructions per 4 calls, three times the int8
        // flavour's, and 8 warps could not keep up with the MMAs) + epilogue (the first 8 of them): thread = marker row
        const int r = threadIdx.x & 255;
        const int half = threadIdx.x >> 8;
        const int mt = r >> 7;
        const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
        const uint32_t row_off = uint32_t(r) * kChunkBytes;
        const uint32_t sw = uint32_t(r & 7);
        uint32_t chunk_ctr = 0, stage_ctr = 0, tile_ctr = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_ctr) {
            for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                const uint32_t cb = chunk_ctr & 1;
                mbar_wait(a_full + cb * 8, (chunk_ctr >> 1) & 1);
                const unsigned char *pv = base_ptr + kSmemPlanes + (cb * 2 + 0) * kPlaneTileBytes + row_off;
                const unsigned char *pd = base_ptr + kSmemPlanes + (cb * 2 + 1) * kPlaneTileBytes + row_off;
                const int st_n = min(kStagesPerChunk, n_stages - ch * kStagesPerChunk);
                for (int st = 0; st < st_n; ++st, ++stage_ctr) {
                    const uint32_t phys = (uint32_t(st) ^ sw) << 4;
                    const uint4 v = *reinterpret_cast<const uint4 *>(pv + phys);
                    const uint4 d = *reinterpret_cast<const uint4 *>(pd + phys);
                    const uint32_t tb = stage_ctr % kF8ABufs;
                    mbar_wait(ta_empty + tb * 8, ((stage_ctr / kF8ABufs) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + lane_addr + kA + tb * kABuf + mt * 32 + half * 16;
                    expand_store_f8(ta + 0, half ? v.z : v.x, half ? d.z : d.x);
                    expand_store_f8(ta + 8, half ? v.w : v.y, half ? d.w : d.y);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ta_full + tb * 8);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(a_empty + cb * 8);
            }
            // ---- epilogue: P + 4096 * (N - P) per column -> pair-slot records, whole-category records
            if (half) continue;
            mbar_wait(acc_full, tile_ctr & 1);
            tc_fence_after();
            const int64_t row = int64_t(tile) * kTileRows + r;
            const bool live = row < n_rows;
            const uint32_t acc = tmem_base + lane_addr + kAcc + mt * kF8NMax;
            int aP = 0, aN = 0, p0 = 0, n0 = 0, p1 = 0, n1 = 0, tP = 0, tN = 0;
            for (int c = 0; c < n_cols; ++c) {
                const uint32_t x = tmem_ld1(acc + c);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int vv = __float2int_rn(__uint_as_float(x));
                aP += vv & 4095;
                aN += (vv & 4095) + (vv >> 12);
                const uint32_t f = s_flags[c];
                if (f & 2u) {
                    if (f & 1u) { p1 = aP; n1 = aN; } else { p0 = aP; n0 = aN; }
                    tP += aP; tN += aN;
                    aP = 0; aN = 0;
                }
                if (f & 4u) {
                    if (live) counts[size_t(s_slot[c]) * size_t(n_rows) + size_t(row)] = make_int4(p0, n0, p1, n1);
                    p0 = n0 = p1 = n1 = 0;
                }
                if (f & 8u) {
                    const int ts = s_tot[c];
                    if (ts >= 0 && live) counts[size_t(ts) * size_t(n_rows) + size_t(row)] = make_int4(tP, tN, 0, 0);
                    tP = 0; tN = 0;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    } else if (warp == kF8ExpandWarps) {
        // =============================================== TMA producer: bit-plane boxes
        if (lane == 0) {
            uint32_t chunk_ctr = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int ch = 0; ch < n_chunks; ++ch, ++chunk_ctr) {
                    const uint32_t cb = chunk_ctr & 1;
                    mbar_wait(a_empty + cb * 8, ((chunk_ctr >> 1) & 1) ^ 1);
                    mbar_expect_tx(a_full + cb * 8, 2u * kPlaneTileBytes);
                    tma_load_2d(base + kSmemPlanes + (cb * 2 + 0) * kPlaneTileBytes, &tm_value,
                                ch * kChunkBytes, tile * kTileRows, a_full + cb * 8);
                    tma_load_2d(base + kSmemPlanes + (cb * 2 + 1) * kPlaneTileBytes, &tm_valid,
                                ch * kChunkBytes, tile * kTileRows, a_full + cb * 8);
                }
            }
        }
    } else if (warp == kF8ExpandWarps + 1) {
        // =============================================== bulk-copy producer: baked one-hot B blocks
        if (lane == 0) {
            uint32_t stage_ctr = 0;
            const uint8_t *bsrc = bmat + pass->b_offset;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int s = 0; s < n_stages; ++s, ++stage_ctr) {
                    const uint32_t bs = stage_ctr % kF8BStages;
                    mbar_wait(b_empty + bs * 8, ((stage_ctr / kF8BStages) & 1) ^ 1);
                    mbar_expect_tx(b_full + bs * 8, b_bytes);
                    bulk_load_1d(base + kSmemB + bs * kF8BStageBytes, bsrc + size_t(s) * b_bytes, b_bytes,
                                 b_full + bs * 8);
                }
            }
        }
    } else {
        // =============================================== MMA issuer
        if (lane == 0) {
            // D f32, A / B e5m2, both K-major, N = n_pad, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(n_pad >> 3) << 17) |
                                   (uint32_t(128 >> 4) << 24);
            uint32_t stage_ctr = 0, tile_ctr = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_ctr) {
                mbar_wait(acc_empty, (tile_ctr & 1) ^ 1);
                tc_fence_after();
                for (int s = 0; s < n_stages; ++s, ++stage_ctr) {
                    const uint32_t tb = stage_ctr % kF8ABufs;
                    const uint32_t bs = stage_ctr % kF8BStages;
                    mbar_wait(ta_full + tb * 8, (stage_ctr / kF8ABufs) & 1);
                    mbar_wait(b_full + bs * 8, (stage_ctr / kF8BStages) & 1);
                    tc_fence_after();
                    const uint64_t bdesc = make_b_desc(base + kSmemB + bs * kF8BStageBytes);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                        for (int m = 0; m < 2; ++m)
                            mma_f8_ts(tmem_base + kAcc + m * kF8NMax, tmem_base + kA + tb * kABuf + m * 32 + ks * 8,
                                      bdesc + uint64_t(ks * 2), idesc, (s > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc_commit(ta_empty + tb * 8);
                    tc_commit(b_empty + bs * 8);
                }
                tc_commit(acc_full);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kF8ExpandWarps + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// [rows][W*8 bytes] plane -> boxes of [256 rows][128 bytes], 128B swizzle, zero fill outside
bool make_plane_map(CUtensorMap *map, const uint64_t *plane, int64_t n_rows, int64_t W) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cuuint64_t(W) * 8, cuuint64_t(n_rows)};
    const cuuint64_t strides[1] = {cuuint64_t(W) * 8};
    const cuuint32_t box[2] = {cuuint32_t(kChunkBytes), cuuint32_t(kTileRows)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint64_t *>(plane), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

int gemm_stage_count(int64_t S) { return int((S + kStageSamples - 1) / kStageSamples); }
// B blocks per pass: one per A stage (i8) or one per two A stages (fp4: a 128-byte row holds 256 e2m1 values)
int gemm_bblock_count(int64_t S, bool fp4) { return fp4 ? (gemm_stage_count(S) + 1) / 2 : gemm_stage_count(S); }
size_t gemm_pass_desc_bytes() { return sizeof(GemmPass); }
int gemm_max_pairs_per_pass() { return kNMax / 2; }

// One pass = up to 32 consecutive pair-slots (64 value columns).  columns[c] lists, for GEMM column c
// of the pass, which samples carry a 1 (built by capi.cu from the categories).  Writes the baked B
// blocks [n_bblocks][n_pad][128] (128B-swizzled, K permuted to the expander's order) at `out`.
void gemm_bake_b(const std::vector<std::vector<int32_t>> &columns, int n_pad, int64_t S, bool fp4, uint8_t *out) {
    const int n_blocks = gemm_bblock_count(S, fp4);
    std::memset(out, 0, size_t(n_blocks) * size_t(n_pad) * 128);
    for (size_t n = 0; n < columns.size(); ++n) {
        for (int32_t s : columns[n]) {
            if (!fp4) {
                const int64_t g = s / kStageSamples;       // stage
                const int w = int(s % kStageSamples);
                const int i = w >> 5, bit = w & 31;        // word of the 16-byte stage row, bit in the word
                const int b = bit >> 3, j = bit & 7;       // expander: column j, byte b holds call 8*b + j
                const int kk = 32 * i + 4 * j + b;         // K position inside the stage
                const size_t off = size_t(g) * size_t(n_pad) * 128 + n * 128 +
                                   size_t((((kk >> 4) ^ int(n & 7)) << 4) + (kk & 15));
                out[off] = uint8_t(1u << (7 - j));
            } else {
                const int64_t g = s / (2 * kStageSamples);               // B block = two A stages
                const int half = int((s / kStageSamples) & 1);
                const int w = int(s % kStageSamples);
                const int i = w >> 5, bit = w & 31;
                const int j = bit & 3, t = bit >> 2;                     // expand4_fp4: column 4*i + j, nibble t
                const int e = half * 128 + (4 * i + j) * 8 + t;          // e2m1 element inside the 256-value row
                const int kb = e >> 1;
                const size_t off = size_t(g) * size_t(n_pad) * 128 + n * 128 +
                                   size_t((((kb >> 4) ^ int(n & 7)) << 4) + (kb & 15));
                // the expander leaves 0.5, 1.0, 2.0, 0.5 for j = 0..3: B holds 2.0, 1.0, 0.5, 2.0 (e2m1 0x4, 0x2, 0x1, 0x4)
                static const uint8_t recip[4] = {0x4, 0x2, 0x1, 0x4};
                out[off] |= uint8_t(recip[j] << (4 * (e & 1)));
            }
        }
    }
}

void gemm_fill_pass(void *dst, int slot_base, int n_pairs, int n_pad, int64_t b_offset) {
    GemmPass p{};
    p.slot_base = slot_base;
    p.n_pairs = n_pairs;
    p.n_pad = n_pad;
    p.b_offset = b_offset;
    std::memcpy(dst, &p, sizeof(p));
}

// ---- e5m2 flavour: one GEMM column per (value, chunk of <= 4095 member samples)
size_t gemm_f8_pass_desc_bytes() { return sizeof(GemmPassF8); }
int gemm_f8_max_cols() { return kF8NMax; }
int gemm_f8_max_members() { return 4095; }

void gemm_f8_fill_pass(void *dst, const std::vector<GemmF8Column> &cols, int n_pad, int64_t b_offset) {
    GemmPassF8 p{};
    p.n_cols = int(cols.size());
    p.n_pad = n_pad;
    p.b_offset = b_offset;
    for (int i = 0; i < kF8NMax; ++i) { p.slot[i] = 0; p.tot_slot[i] = -1; p.flags[i] = 0; }
    for (size_t i = 0; i < cols.size(); ++i) {
        p.slot[i] = cols[i].slot;
        p.tot_slot[i] = cols[i].tot_slot;
        p.flags[i] = cols[i].flags;
    }
    std::memcpy(dst, &p, sizeof(p));
}

// B blocks [n_stages][n_pad][128]: e5m2 1.0 (0x3c) where column n has sample s, at the K position expand_store_f8
// gives that sample (the int8 flavour's order)
void gemm_f8_bake_b(const std::vector<GemmF8Column> &cols, int n_pad, int64_t S, uint8_t *out) {
    const int n_blocks = gemm_stage_count(S);
    std::memset(out, 0, size_t(n_blocks) * size_t(n_pad) * 128);
    for (size_t n = 0; n < cols.size(); ++n)
        for (int32_t s : cols[n].samples) {
            const int64_t g = s / kStageSamples;
            const int w = int(s % kStageSamples);
            const int i = w >> 5, bit = w & 31;
            const int b = bit >> 3, j = bit & 7;
            const int kk = 32 * i + 4 * j + b;
            out[size_t(g) * size_t(n_pad) * 128 + n * 128 + size_t((((kk >> 4) ^ int(n & 7)) << 4) + (kk & 15))] = 0x3c;
        }
}

cudaError_t launch_count_gemm_f8(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows, int64_t W,
                                 int64_t S, const uint8_t *d_bmat, const void *d_passes, int n_passes,
                                 int4 *d_counts, int num_sms, cudaStream_t st, int *launches) {
    if (n_rows == 0 || n_passes == 0) return cudaSuccess;
    CUtensorMap tm_value, tm_valid;
    if (!make_plane_map(&tm_value, d_value, n_rows, W) || !make_plane_map(&tm_valid, d_valid, n_rows, W))
        return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(count_gemm_f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF8SmemTotal);
    if (e != cudaSuccess) return e;
    const int n_tiles = int((n_rows + kTileRows - 1) / kTileRows);
    int gx = num_sms / n_passes;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    count_gemm_f8_kernel<<<dim3(unsigned(gx), unsigned(n_passes)), kF8Threads, kF8SmemTotal, st>>>(
        tm_value, tm_valid, d_bmat, static_cast<const GemmPassF8 *>(d_passes), gemm_stage_count(S), n_rows, n_tiles,
        d_counts);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_count_gemm(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows, int64_t W,
                              int64_t S, const uint8_t *d_bmat, const void *d_passes, int n_passes, bool fp4,
                              int4 *d_counts, int num_sms, cudaStream_t st, int *launches) {
    if (n_rows == 0 || n_passes == 0) return cudaSuccess;
    CUtensorMap tm_value, tm_valid;
    if (!make_plane_map(&tm_value, d_value, n_rows, W) || !make_plane_map(&tm_valid, d_valid, n_rows, W))
        return cudaErrorNotSupported;
    {   // per launch: the attribute belongs to the current device's copy of the kernel
        cudaError_t e = fp4 ? cudaFuncSetAttribute(count_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal)
                            : cudaFuncSetAttribute(count_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
        if (e != cudaSuccess) return e;
    }
    const int n_tiles = int((n_rows + kTileRows - 1) / kTileRows);
    int gx = num_sms / n_passes;
    if (gx < 1) gx = 1;
    if (gx > n_tiles) gx = n_tiles;
    // every pass gets its own set of persistent CTAs; with one pass this is one CTA per SM
    if (fp4)
        count_gemm_kernel<true><<<dim3(unsigned(gx), unsigned(n_passes)), kThreads2, kSmemTotal, st>>>(
            tm_value, tm_valid, d_bmat, static_cast<const GemmPass *>(d_passes), gemm_stage_count(S), n_rows, n_tiles,
            d_counts);
    else
        count_gemm_kernel<false><<<dim3(unsigned(gx), unsigned(n_passes)), kThreads2, kSmemTotal, st>>>(
            tm_value, tm_valid, d_bmat, static_cast<const GemmPass *>(d_passes), gemm_stage_count(S), n_rows, n_tiles,
            d_counts);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace feht
