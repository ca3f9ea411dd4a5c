// host_pack.cpp -- transport encoding of character rows on the host cores.
//
// feht_add_rows_chars / feht_add_snp_chars receive the matrix as the Haskell host holds it, one byte per
// call (Table.hs:160-205 narrowed from Char).  At 1 B/call the upload is bound by the PCIe link (53 GB/s
// measured = 94 ms for 1M x 5k), while the packed form is 2 bits per call.  So while the DMA engine moves raw
// pieces for K0 (pack.cu) to pack on the GPU, the otherwise idle host cores transcode pieces taken from the
// other end of the row range into the SAME bit-planes and upload those at a quarter of the bytes.  This is
// only the wire format: counting, testing, correcting, filtering and ordering happen on the GPU, and the two
// producers write bit-identical planes (tests/test_gpu_parity.py compares them).
//
// Character semantics are K0's: '1' -> value+valid, '0' -> valid, anything else missing (Comparison.hs:53-56,
// README.md:227-234); SNP: toUpper, A=0 C=1 G=2 T=3 in (hi, lo), anything else missing (Table.hs:194-205) -- except a
// literal '0' / '1', which the reference's replaceChar leaves alone (it then counts for all four allele rows): the SNP
// packers report those (return value) and the caller rejects the upload, as feht_add_rows_text does.
// Characters at index >= row length or >= S contribute nothing.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace feht {

bool host_pack_available() { return __builtin_cpu_supports("avx2") != 0; }

// portable versions (no AVX2): the same planes, one call at a time -- only feht_pack_chars_host uses them
void host_pack_binary_scalar(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                             uint64_t *value, uint64_t *valid) {
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *v = value + (r - r0) * W, *d = valid + (r - r0) * W;
        for (int64_t w = 0; w < W; ++w) { v[w] = 0; d[w] = 0; }
        for (int64_t j = 0; j < n; ++j) {
            const uint64_t bit = uint64_t(1) << (j & 63);
            if (src[j] == '1') { v[j >> 6] |= bit; d[j >> 6] |= bit; }
            else if (src[j] == '0') d[j >> 6] |= bit;
        }
    }
}
bool host_pack_snp_scalar(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                          uint64_t *hi, uint64_t *valid, uint64_t *lo) {
    bool digit = false;   // a literal '0' / '1' call: counts for all four alleles in the reference (Table.hs:194-205)
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *ph = hi + (r - r0) * W, *pv = valid + (r - r0) * W, *pl = lo + (r - r0) * W;
        for (int64_t w = 0; w < W; ++w) { ph[w] = 0; pv[w] = 0; pl[w] = 0; }
        for (int64_t j = 0; j < n; ++j) {
            uint8_t c = src[j];
            digit |= (c == '0' || c == '1');
            if (c >= 'a' && c <= 'z') c = uint8_t(c - 'a' + 'A');   // toUpper (Table.hs:205)
            const uint64_t bit = uint64_t(1) << (j & 63);
            if (c == 'G' || c == 'T') ph[j >> 6] |= bit;
            if (c == 'C' || c == 'T') pl[j >> 6] |= bit;
            if (c == 'A' || c == 'C' || c == 'G' || c == 'T') pv[j >> 6] |= bit;
        }
    }
    return digit;
}

namespace {

__attribute__((target("avx2"))) inline uint64_t eq_mask64(const __m256i a, const __m256i b, const __m256i c) {
    const uint32_t lo = uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(a, c)));
    const uint32_t hi = uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(b, c)));
    return uint64_t(lo) | (uint64_t(hi) << 32);
}

}  // namespace

__attribute__((target("avx2")))
static void host_pack_binary_avx2(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                      uint64_t *value, uint64_t *valid) {
    const __m256i one = _mm256_set1_epi8('1'), zero = _mm256_set1_epi8('0');
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *v = value + (r - r0) * W, *d = valid + (r - r0) * W;
        const int64_t full = n >> 6;
        for (int64_t w = 0; w < full; ++w) {
            const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64 * w));
            const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64 * w + 32));
            const uint64_t m1 = eq_mask64(a, b, one);
            // (plain stores: 8-byte streaming stores into the two planes were measured 45% slower, 16-byte ones 15%)
            v[w] = m1;
            d[w] = m1 | eq_mask64(a, b, zero);
        }
        int64_t w = full;
        if (n & 63) {
            uint64_t m1 = 0, m0 = 0;
            for (int64_t j = 64 * full; j < n; ++j) {
                m1 |= uint64_t(src[j] == '1') << (j & 63);
                m0 |= uint64_t(src[j] == '0') << (j & 63);
            }
            v[w] = m1;
            d[w] = m1 | m0;
            ++w;
        }
        for (; w < W; ++w) { v[w] = 0; d[w] = 0; }
    }
}

__attribute__((target("avx2")))
static bool host_pack_snp_avx2(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                   uint64_t *hi, uint64_t *valid, uint64_t *lo) {
    // toUpper on the four letters that matter: (c & 0xDF) == 'A' iff c is 'A' or 'a'
    const __m256i up = _mm256_set1_epi8(char(0xDF));
    const __m256i c0 = _mm256_set1_epi8('0'), c1 = _mm256_set1_epi8('1');
    uint64_t digits = 0;   // any literal '0' / '1' (raw characters, before the case fold)
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'),
                  cT = _mm256_set1_epi8('T');
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *ph = hi + (r - r0) * W, *pv = valid + (r - r0) * W, *pl = lo + (r - r0) * W;
        const int64_t full = n >> 6;
        for (int64_t w = 0; w < full; ++w) {
            const __m256i ra = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64 * w));
            const __m256i rb = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 64 * w + 32));
            digits |= eq_mask64(ra, rb, c0) | eq_mask64(ra, rb, c1);
            const __m256i a = _mm256_and_si256(ra, up);
            const __m256i b = _mm256_and_si256(rb, up);
            const uint64_t mA = eq_mask64(a, b, cA), mC = eq_mask64(a, b, cC), mG = eq_mask64(a, b, cG),
                           mT = eq_mask64(a, b, cT);
            ph[w] = mG | mT;
            pl[w] = mC | mT;
            pv[w] = mA | mC | mG | mT;
        }
        int64_t w = full;
        if (n & 63) {
            uint64_t mh = 0, ml = 0, mv = 0;
            for (int64_t j = 64 * full; j < n; ++j) {
                const uint8_t c = uint8_t(src[j] & 0xDF);
                digits |= uint64_t(src[j] == '0' || src[j] == '1');
                const uint64_t bit = uint64_t(1) << (j & 63);
                if (c == 'G' || c == 'T') mh |= bit;
                if (c == 'C' || c == 'T') ml |= bit;
                if (c == 'A' || c == 'C' || c == 'G' || c == 'T') mv |= bit;
            }
            ph[w] = mh; pl[w] = ml; pv[w] = mv;
            ++w;
        }
        for (; w < W; ++w) { ph[w] = 0; pl[w] = 0; pv[w] = 0; }
    }
    return digits != 0;
}


// AVX-512BW: one 64-byte load and one compare-into-mask per plane bit and word; the ragged end of a row is a
// masked load, so there is no scalar tail.  ~12% more bytes per core than the AVX2 loop on Sapphire Rapids
// (9.0 vs 8.0 GB/s per core against a 12.8 GB/s read-only ceiling), which is what the chars upload is bound by.
__attribute__((target("avx512f,avx512bw")))
static void host_pack_binary_avx512(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S,
                                    int64_t W, uint64_t *value, uint64_t *valid) {
    const __m512i one = _mm512_set1_epi8('1'), zero = _mm512_set1_epi8('0');
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *v = value + (r - r0) * W, *d = valid + (r - r0) * W;
        const int64_t full = n >> 6;
        for (int64_t w = 0; w < full; ++w) {
            const __m512i a = _mm512_loadu_si512(src + 64 * w);
            const uint64_t m1 = _mm512_cmpeq_epi8_mask(a, one);
            v[w] = m1;
            d[w] = m1 | _mm512_cmpeq_epi8_mask(a, zero);
        }
        int64_t w = full;
        if (n & 63) {
            const __mmask64 k = (uint64_t(1) << (n & 63)) - 1;
            const __m512i a = _mm512_maskz_loadu_epi8(k, src + 64 * full);
            const uint64_t m1 = _mm512_mask_cmpeq_epi8_mask(k, a, one);
            v[w] = m1;
            d[w] = m1 | _mm512_mask_cmpeq_epi8_mask(k, a, zero);
            ++w;
        }
        for (; w < W; ++w) { v[w] = 0; d[w] = 0; }
    }
}

__attribute__((target("avx512f,avx512bw")))
static bool host_pack_snp_avx512(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S,
                                 int64_t W, uint64_t *hi, uint64_t *valid, uint64_t *lo) {
    const __m512i up = _mm512_set1_epi8(char(0xDF));
    const __m512i c0 = _mm512_set1_epi8('0'), c1 = _mm512_set1_epi8('1');
    uint64_t digits = 0;
    const __m512i cA = _mm512_set1_epi8('A'), cC = _mm512_set1_epi8('C'), cG = _mm512_set1_epi8('G'),
                  cT = _mm512_set1_epi8('T');
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *src = chars + off[r];
        int64_t n = off[r + 1] - off[r];
        if (n > S) n = S;
        uint64_t *ph = hi + (r - r0) * W, *pv = valid + (r - r0) * W, *pl = lo + (r - r0) * W;
        const int64_t words = (n + 63) >> 6;
        for (int64_t w = 0; w < words; ++w) {
            const int64_t left = n - 64 * w;
            const __mmask64 k = left >= 64 ? ~uint64_t(0) : (uint64_t(1) << left) - 1;
            const __m512i raw = _mm512_maskz_loadu_epi8(k, src + 64 * w);
            digits |= _mm512_mask_cmpeq_epi8_mask(k, raw, c0) | _mm512_mask_cmpeq_epi8_mask(k, raw, c1);
            const __m512i a = _mm512_and_si512(raw, up);
            const uint64_t mA = _mm512_mask_cmpeq_epi8_mask(k, a, cA), mC = _mm512_mask_cmpeq_epi8_mask(k, a, cC),
                           mG = _mm512_mask_cmpeq_epi8_mask(k, a, cG), mT = _mm512_mask_cmpeq_epi8_mask(k, a, cT);
            ph[w] = mG | mT;
            pl[w] = mC | mT;
            pv[w] = mA | mC | mG | mT;
        }
        for (int64_t w = words; w < W; ++w) { ph[w] = 0; pl[w] = 0; pv[w] = 0; }
    }
    return digits != 0;
}

static bool use_avx512() {
    static const bool yes = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") &&
                            getenv("FEHT_HOST_PACK_AVX2") == nullptr;
    return yes;
}

void host_pack_binary(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                      uint64_t *value, uint64_t *valid) {
    if (use_avx512()) host_pack_binary_avx512(chars, off, r0, r1, S, W, value, valid);
    else host_pack_binary_avx2(chars, off, r0, r1, S, W, value, valid);
}
bool host_pack_snp(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                   uint64_t *hi, uint64_t *valid, uint64_t *lo) {
    if (use_avx512()) return host_pack_snp_avx512(chars, off, r0, r1, S, W, hi, valid, lo);
    return host_pack_snp_avx2(chars, off, r0, r1, S, W, hi, valid, lo);
}

}  // namespace feht
