// text_load.cu -- N2 of SURVEY.md 8(f): the data file's text straight to bit-planes on the device.
//
// Replaces, for the body of the data file (every line after the header), Table.hs:267-279 (`BS.lines`, the
// '\r' filter, `BS.split d`) and :160-191 (first cell = marker name, the remaining cells concatenated: the
// character index in that concatenation IS the column number, so a cell like "1 " shifts the later columns,
// SURVEY.md finding 5), then the same character semantics as K0 (pack.cu): binary '1' -> value+valid, '0' ->
// valid, anything else missing; SNP toUpper, A=0 C=1 G=2 T=3 in (hi, lo), anything else missing
// (Table.hs:194-205).  The reference holds the result as 4 bytes per call in a HashMap of vectors; here a piece
// of text (whole lines, <= 64 MB) is copied to the device as it is and three small kernels find its lines:
//   text_count_kernel   newlines per 4 KB block
//   text_scan_kernel    exclusive scan of the block counts (one CTA)
//   text_starts_kernel  line_start[i] for every line of the piece
// and text_pack_kernel (one warp per line) walks the line 32 characters at a time: the first delimiter ends the
// name; after it a character is kept iff it is neither the delimiter nor '\r', its column is the running count of
// kept characters (ballot + popc), and the plane words are assembled with warp OR-reductions -- a chunk of 32
// characters touches at most two 32-bit words of a plane row.
//
// Bound: the PCIe link (the text is 2 bytes per call: 190 ms for config 2's 10 GB against ~30 ms of kernels).
#include "common.cuh"

namespace feht {
namespace {

constexpr int kTxtThreads = 256;
constexpr int kTxtBlockBytes = kTxtThreads * 16;

__device__ __forceinline__ uint4 load16_guarded(const uint8_t *t, int64_t off, int64_t len) {
    // 16 bytes at off (16-byte aligned piece base); bytes at or beyond len read as 0
    if (off + 16 <= len) return *reinterpret_cast<const uint4 *>(t + off);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; ++k)
        if (off + k < len) w[k >> 2] |= uint32_t(t[off + k]) << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t newline_flags(uint32_t x) {   // 0xff in every byte that is '\n'
    return __vcmpeq4(x, 0x0a0a0a0au);
}

__device__ __forceinline__ int count16(const uint4 &v) {
    return (__popc(newline_flags(v.x)) + __popc(newline_flags(v.y)) + __popc(newline_flags(v.z)) +
            __popc(newline_flags(v.w))) >> 3;
}

__global__ void __launch_bounds__(kTxtThreads)
text_count_kernel(const uint8_t *__restrict__ t, int64_t len, uint32_t *__restrict__ block_counts) {
    __shared__ int wsum[kTxtThreads / 32];
    const int64_t off = (int64_t(blockIdx.x) * kTxtThreads + threadIdx.x) * 16;
    int c = off < len ? count16(load16_guarded(t, off, len)) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < kTxtThreads / 32; ++w) s += wsum[w];
        block_counts[blockIdx.x] = uint32_t(s);
    }
}

// in place: block_counts[b] = lines before block b; block_counts[n_blocks] = newlines of the piece
__global__ void __launch_bounds__(1024) text_scan_kernel(uint32_t *__restrict__ block_counts, int n_blocks) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_counts[i] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < wid; ++w) wb += wsum[w];
        const uint32_t carry = carry_s;
        if (i < n_blocks) block_counts[i] = carry + wb + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wb + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[n_blocks] = carry_s;
}

// line_start[0] = 0; line_start[k] = position after the k-th newline
__global__ void __launch_bounds__(kTxtThreads)
text_starts_kernel(const uint8_t *__restrict__ t, int64_t len, const uint32_t *__restrict__ block_base,
                   int32_t *__restrict__ line_start) {
    __shared__ int wsum[kTxtThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t off = (int64_t(blockIdx.x) * kTxtThreads + threadIdx.x) * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (off < len) v = load16_guarded(t, off, len);
    const int c = count16(v);
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    int wb = 0;
    for (int w = 0; w < wid; ++w) wb += wsum[w];
    int k = int(block_base[blockIdx.x]) + wb + inc - c;   // newlines before this thread's bytes
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = 0;
    if (c) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (((w[j >> 2] >> (8 * (j & 3))) & 255u) == 0x0au) line_start[++k] = int32_t(off + j + 1);
    }
}

struct TextPackArgs {
    const uint8_t *t;
    const int32_t *line_start;   // [n_lines + 1]; line i = [line_start[i], line_start[i + 1] - 1)
    int64_t n_lines;
    uint8_t delim;
    int snp;
    int64_t S;
    int words;                   // 32-bit words per plane row (2 * W)
    uint32_t *p0, *valid, *lo;   // plane rows of the first line (binary: value / valid; SNP: hi / valid / lo)
    int32_t *name_len;           // [n_lines]: bytes of the first cell (a '\r' inside it included)
    // [0] min row length, [1] an empty line was seen, [2] SNP text holds a literal '0' / '1'
    unsigned long long *flags;
};

__global__ void __launch_bounds__(256) text_pack_kernel(const TextPackArgs A) {
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const uint32_t d = A.delim;
    for (int64_t line = warp; line < A.n_lines; line += n_warps) {
        const int64_t b = A.line_start[line], e = int64_t(A.line_start[line + 1]) - 1;
        // ---- the name: up to the first delimiter
        int64_t pos = b, name_end = e;
        unsigned any = 0;       // a character other than '\r' anywhere in the line
        for (; pos < e; pos += 32) {
            const bool in = pos + lane < e;
            const uint32_t c = in ? A.t[pos + lane] : 0u;
            const unsigned m = __ballot_sync(0xffffffffu, in && c == d);
            any |= __ballot_sync(0xffffffffu, in && c != 0x0du);
            if (m) { name_end = pos + __ffs(m) - 1; break; }
        }
        if (lane == 0) A.name_len[line] = int32_t(name_end - b);
        // ---- the row: every later character that is neither the delimiter nor '\r'
        uint32_t *r0 = A.p0 + size_t(line) * A.words, *rv = A.valid + size_t(line) * A.words;
        uint32_t *rl = A.snp ? A.lo + size_t(line) * A.words : nullptr;
        uint32_t a0 = 0, av = 0, al = 0;   // the word being assembled (same in every lane)
        int wcur = 0;
        int64_t kept = 0;
        unsigned lit = 0;
        for (pos = name_end + 1; pos < e; pos += 32) {
            const bool in = pos + lane < e;
            const uint32_t c = in ? A.t[pos + lane] : d;
            const bool keep = in && c != d && c != 0x0du;
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            const int64_t idx = kept + __popc(km & lt);
            bool b0, bv, bl = false;
            if (!A.snp) {
                b0 = c == '1';
                bv = b0 || c == '0';
            } else {
                const uint32_t u = c & 0xdfu;   // toUpper on the letters that matter (Table.hs:205)
                const bool isA = u == 'A', isC = u == 'C', isG = u == 'G', isT = u == 'T';
                b0 = isG || isT;
                bl = isC || isT;
                bv = isA || isC || isG || isT;
                lit |= __ballot_sync(0xffffffffu, keep && (c == '0' || c == '1'));
            }
            const bool inS = keep && idx < A.S;
            const int w = int(idx >> 5);
            const uint32_t bit = 1u << (idx & 31);
            const bool first = inS && w == wcur, second = inS && w == wcur + 1;
            const uint32_t f0 = __reduce_or_sync(0xffffffffu, first && b0 ? bit : 0u);
            const uint32_t fv = __reduce_or_sync(0xffffffffu, first && bv ? bit : 0u);
            const uint32_t s0 = __reduce_or_sync(0xffffffffu, second && b0 ? bit : 0u);
            const uint32_t sv = __reduce_or_sync(0xffffffffu, second && bv ? bit : 0u);
            uint32_t fl = 0, sl = 0;
            if (A.snp) {
                fl = __reduce_or_sync(0xffffffffu, first && bl ? bit : 0u);
                sl = __reduce_or_sync(0xffffffffu, second && bl ? bit : 0u);
            }
            a0 |= f0; av |= fv; al |= fl;
            kept += __popc(km);
            if ((kept >> 5) > wcur) {   // the chunk went over a word boundary (at most one: <= 32 kept characters)
                if (lane == 0 && wcur < A.words) {
                    r0[wcur] = a0; rv[wcur] = av;
                    if (rl) rl[wcur] = al;
                }
                a0 = s0; av = sv; al = sl;
                ++wcur;
            }
            any |= km;
        }
        // last (partial) word, then zero padding up to the row stride
        if (lane == 0 && wcur < A.words) {
            r0[wcur] = a0; rv[wcur] = av;
            if (rl) rl[wcur] = al;
        }
        for (int w = wcur + 1 + lane; w < A.words; w += 32) {
            r0[w] = 0u; rv[w] = 0u;
            if (rl) rl[w] = 0u;
        }
        if (lane == 0) {
            atomicMin(A.flags, (unsigned long long)kept);
            if (!any) atomicOr(A.flags + 1, 1ull);
            if (lit) atomicOr(A.flags + 2, 1ull);
        }
    }
}

}  // namespace

size_t text_scratch_words(size_t piece_bytes) { return (piece_bytes + kTxtBlockBytes - 1) / kTxtBlockBytes + 2; }

cudaError_t launch_text_count(const uint8_t *d_text, int64_t len, uint32_t *d_block_counts, cudaStream_t st) {
    const int nb = int((len + kTxtBlockBytes - 1) / kTxtBlockBytes);
    if (nb == 0) return cudaMemsetAsync(d_block_counts, 0, 4, st);
    text_count_kernel<<<nb, kTxtThreads, 0, st>>>(d_text, len, d_block_counts);
    text_scan_kernel<<<1, 1024, 0, st>>>(d_block_counts, nb);
    return cudaGetLastError();
}

cudaError_t launch_text_starts(const uint8_t *d_text, int64_t len, const uint32_t *d_block_base,
                               int32_t *d_line_start, cudaStream_t st) {
    const int nb = int((len + kTxtBlockBytes - 1) / kTxtBlockBytes);
    if (nb == 0) return cudaSuccess;
    text_starts_kernel<<<nb, kTxtThreads, 0, st>>>(d_text, len, d_block_base, d_line_start);
    return cudaGetLastError();
}

cudaError_t launch_text_pack(const uint8_t *d_text, const int32_t *d_line_start, int64_t n_lines, uint8_t delim,
                             int snp, int64_t S, int64_t W, uint64_t *p0, uint64_t *valid, uint64_t *lo,
                             int32_t *d_name_len, unsigned long long *d_flags, int num_sms, cudaStream_t st) {
    if (n_lines == 0) return cudaSuccess;
    TextPackArgs a{};
    a.t = d_text; a.line_start = d_line_start; a.n_lines = n_lines; a.delim = delim; a.snp = snp; a.S = S;
    a.words = int(2 * W);
    a.p0 = reinterpret_cast<uint32_t *>(p0);
    a.valid = reinterpret_cast<uint32_t *>(valid);
    a.lo = reinterpret_cast<uint32_t *>(lo);
    a.name_len = d_name_len;
    a.flags = d_flags;
    int64_t blocks = (n_lines + 7) / 8;
    if (blocks > int64_t(num_sms) * 8) blocks = int64_t(num_sms) * 8;
    text_pack_kernel<<<int(blocks), 256, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace feht
