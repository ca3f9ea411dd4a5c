// pack.cu -- K0: bring the data matrix into the packed HBM layout.
//
// Replaces the *layout* produced by Table.hs:160-224 (GeneVectorMap = HashMap GeneName (Vector Char),
// 4 bytes per call) with bit-planes: 1 bit per call + 1 valid bit (binary), 2-bit nucleotide + valid
// (SNP).  The character semantics are the reference's: Comparison.hs:53-56 counts chars equal to
// '1' and to '0'; every other character is missing (README.md:227-234).  SNP: Table.hs:194-205
// (toUpper; A/C/G/T are calls, anything else is missing).
//
// Bound: HBM/PCIe (1 byte in per call, 2 bits out).  One warp turns 32 characters into one 32-bit
// word of each plane with __ballot_sync; stores are 4-byte coalesced across the warps of a row.
#include "common.cuh"

namespace feht {
namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// grid.x over groups of rows, each warp walks the 32-character segments of one row.
__global__ void __launch_bounds__(256)
pack_chars_kernel(const uint8_t *__restrict__ chars, const int64_t *__restrict__ row_offsets,
                  int64_t n_rows, int64_t S, int64_t W, uint32_t *__restrict__ value,
                  uint32_t *__restrict__ valid) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t segs = 2 * W;  // 32-bit words per row per plane
    const int64_t total = n_rows * segs;
    for (int64_t item = warp; item < total; item += n_warps) {
        const int64_t row = item / segs;
        const int64_t seg = item - row * segs;
        const int64_t off = row_offsets[row];
        const int64_t len = row_offsets[row + 1] - off;
        const int64_t col = seg * 32 + lane;
        uint8_t c = 0;
        if (col < len && col < S) c = chars[off + col];
        const unsigned one = __ballot_sync(0xffffffffu, c == '1');
        const unsigned zero = __ballot_sync(0xffffffffu, c == '0');
        if (lane == 0) {
            value[row * segs + seg] = one;
            valid[row * segs + seg] = one | zero;
        }
    }
}

__global__ void __launch_bounds__(256)
pack_snp_chars_kernel(const uint8_t *__restrict__ chars, const int64_t *__restrict__ row_offsets,
                      int64_t n_sites, int64_t S, int64_t W, uint32_t *__restrict__ hi,
                      uint32_t *__restrict__ valid, uint32_t *__restrict__ lo, unsigned int *__restrict__ digit_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t segs = 2 * W;
    const int64_t total = n_sites * segs;
    for (int64_t item = warp; item < total; item += n_warps) {
        const int64_t row = item / segs;
        const int64_t seg = item - row * segs;
        const int64_t off = row_offsets[row];
        const int64_t len = row_offsets[row + 1] - off;
        const int64_t col = seg * 32 + lane;
        uint8_t c = 0;
        if (col < len && col < S) c = chars[off + col];
        // replaceChar leaves a literal '0' / '1' as it is, so it counts for all four allele rows (Table.hs:194-205);
        // the 2-bit planes cannot say that: flag it, the upload is rejected
        if (__ballot_sync(0xffffffffu, c == '0' || c == '1') && lane == 0) *digit_flag = 1u;
        if (c >= 'a' && c <= 'z') c = uint8_t(c - 'a' + 'A');  // toUpper (Table.hs:205)
        // A=0 C=1 G=2 T=3
        const bool isA = c == 'A', isC = c == 'C', isG = c == 'G', isT = c == 'T';
        const unsigned bhi = __ballot_sync(0xffffffffu, isG || isT);
        const unsigned blo = __ballot_sync(0xffffffffu, isC || isT);
        const unsigned bv = __ballot_sync(0xffffffffu, isA || isC || isG || isT);
        if (lane == 0) {
            hi[row * segs + seg] = bhi;
            lo[row * segs + seg] = blo;
            valid[row * segs + seg] = bv;
        }
    }
}

// One thread makes one 32-bit word of each plane (32 cells).
__global__ void __launch_bounds__(256)
synth_kernel(const feht_synth_spec *__restrict__ spec_g, int snp_mode, int64_t first_row,
             int64_t n_rows, int64_t S, int64_t W, uint32_t *__restrict__ p0,
             uint32_t *__restrict__ valid, uint32_t *__restrict__ lo) {
    __shared__ feht_synth_spec sp;
    for (int i = threadIdx.x; i < int(sizeof(feht_synth_spec) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&sp)[i] = reinterpret_cast<const uint32_t *>(spec_g)[i];
    __syncthreads();
    const int64_t segs = 2 * W;
    const int64_t total = n_rows * segs;
    for (int64_t item = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; item < total;
         item += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = item / segs;
        const int64_t seg = item - r * segs;
        const int64_t row = first_row + r;
        const uint64_t hr = mix64(sp.seed ^ (uint64_t(row) * 0x9E3779B97F4A7C15ULL));
        const bool planted = ((hr >> 8) % 1000000ULL) < sp.planted_per_million;
        const uint32_t thr_row = sp.freq_thr[hr & 255];
        uint32_t w0 = 0, wv = 0, wl = 0;
        const uint32_t major = uint32_t(hr >> 40) & 3u;
        const uint32_t minor = (major + 1u + uint32_t((hr >> 42) % 3ULL)) & 3u;
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
            const int64_t col = seg * 32 + b;
            if (col >= S) break;
            const uint64_t hc = mix64(hr + uint64_t(col));
            const bool missing = uint32_t(hc >> 32) < sp.missing_thr;
            if (missing) continue;
            wv |= 1u << b;
            if (!snp_mode) {
                const uint32_t thr = planted ? (col < sp.split_col ? sp.thr_hi : sp.thr_lo) : thr_row;
                if (uint32_t(hc) < thr) w0 |= 1u << b;
            } else {
                const uint32_t code = (uint32_t(hc) < thr_row) ? minor : major;
                w0 |= (code >> 1) << b;
                wl |= (code & 1u) << b;
            }
        }
        p0[r * segs + seg] = w0;
        valid[r * segs + seg] = wv;
        if (snp_mode) lo[r * segs + seg] = wl;
    }
}

// Copy rows from a caller stride to the library stride, zeroing padding words and bits >= S.
__global__ void __launch_bounds__(256)
repack_rows_kernel(const uint64_t *__restrict__ src, int64_t src_stride, uint64_t *__restrict__ dst,
                   int64_t W, int64_t S, int64_t n_rows, const uint64_t *__restrict__ and_with) {
    const int64_t total = n_rows * W;
    const int64_t full = S >> 6;
    const int rem = int(S & 63);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / W;
        const int64_t w = i - r * W;
        uint64_t v = 0;
        if (w < full) v = src[r * src_stride + w];
        else if (w == full && rem) v = src[r * src_stride + w] & ((1ULL << rem) - 1ULL);
        if (and_with) v &= and_with[i];  // a value bit only counts where the call is valid
        dst[i] = v;
    }
}

inline int grid_for(int64_t work_items, int per_block) {
    int64_t b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > int64_t(sm_count()) * 32) b = int64_t(sm_count()) * 32;
    return int(b);
}

}  // namespace

cudaError_t launch_pack_chars(const uint8_t *d_chars, const int64_t *d_row_offsets, int64_t n_rows,
                              int64_t S, int64_t W, uint64_t *d_value, uint64_t *d_valid,
                              cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    pack_chars_kernel<<<grid_for(n_rows * 2 * W, 8), 256, 0, st>>>(
        d_chars, d_row_offsets, n_rows, S, W, reinterpret_cast<uint32_t *>(d_value),
        reinterpret_cast<uint32_t *>(d_valid));
    return cudaGetLastError();
}

cudaError_t launch_pack_snp_chars(const uint8_t *d_chars, const int64_t *d_row_offsets,
                                  int64_t n_sites, int64_t S, int64_t W, uint64_t *d_hi,
                                  uint64_t *d_valid, uint64_t *d_lo, unsigned int *d_digit_flag, cudaStream_t st) {
    if (n_sites == 0) return cudaSuccess;
    pack_snp_chars_kernel<<<grid_for(n_sites * 2 * W, 8), 256, 0, st>>>(
        d_chars, d_row_offsets, n_sites, S, W, reinterpret_cast<uint32_t *>(d_hi),
        reinterpret_cast<uint32_t *>(d_valid), reinterpret_cast<uint32_t *>(d_lo), d_digit_flag);
    return cudaGetLastError();
}

cudaError_t launch_synth(const feht_synth_spec *d_spec, int snp_mode, int64_t first_row,
                         int64_t n_rows, int64_t S, int64_t W, uint64_t *d_p0, uint64_t *d_valid,
                         uint64_t *d_lo, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    synth_kernel<<<grid_for(n_rows * 2 * W, 256), 256, 0, st>>>(
        d_spec, snp_mode, first_row, n_rows, S, W, reinterpret_cast<uint32_t *>(d_p0),
        reinterpret_cast<uint32_t *>(d_valid), reinterpret_cast<uint32_t *>(d_lo));
    return cudaGetLastError();
}

cudaError_t launch_repack_rows(const uint64_t *d_src, int64_t src_stride, uint64_t *d_dst,
                               int64_t W, int64_t S, int64_t n_rows, const uint64_t *d_and_with,
                               cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    repack_rows_kernel<<<grid_for(n_rows * W, 256), 256, 0, st>>>(d_src, src_stride, d_dst, W, S,
                                                                  n_rows, d_and_with);
    return cudaGetLastError();
}

}  // namespace feht
