// common.cuh -- shared declarations of libfeht_cuda (B200 / sm_100a).
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   binary mode : two bit-planes, value and valid, each [rows][W] 64-bit words, W = ceil(S/64)
//                 rounded up to an even number so that every row starts on a 16-byte boundary
//                 (128-bit loads).  Bit s%64 of word s/64 = sample s.  Padding bits are 0.
//   SNP mode    : three planes hi, valid, lo per site (2-bit nucleotide code A=0 C=1 G=2 T=3).
//   counts      : int4 per (pair-slot, marker row): (P_first, N_first, P_second, N_second) where
//                 P = #'1' and N = #('1' or '0') inside a column mask.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/feht_cuda.h"

namespace feht {

constexpr int kChooseN = 1000;          // FET.hs:105,109: hypergeometric branch only for l < 1000
constexpr int64_t kChooseEntries = 250500;  // sum_{n<1000} (n/2 + 1)

// index of C[n][k'] (k' <= n/2) in the folded triangular table
__host__ __device__ inline int choose_offset(int n) { return n + ((n - 1) * (n - 1)) / 4; }

// K3a works on GROUPS of comparisons that share a small table of per-row positive fractions.
// A fraction is P/N of one half of a pair-slot, or, for one-vs-rest, (T_P - P)/(T_N - N) with T the
// category's total slot.  A comparison is then two fraction ids.
struct FracDef {
    int32_t slot;       // pair-slot holding (P,N)
    int32_t sel;        // 0: (x,y) components, 1: (z,w) components
    int32_t tot_slot;   // >= 0: "rest of the category" = total slot (x,y) minus this value
    int32_t pad;
};
struct GroupComp {
    int32_t comp;       // index of the comparison in the run
    int16_t fa, fb;     // fraction ids (inside the group) of group one / group two
};
// One load of a pair-slot record feeds up to four fractions of the group: the plain fractions of
// its two halves and, for one-vs-rest, "category total minus this value" of each half.
struct SlotDef {
    int32_t slot;
    int16_t f_lo, f_hi;   // fraction ids of (x,y) / (z,w); -1 = not used by this group
    int16_t r_lo, r_hi;   // fraction ids of the one-vs-rest complements; -1 = none
    int32_t pad;
};
struct ScreenGroup {
    int32_t frac_off, n_frac;   // slice of the FracDef array (emit path: fraction id -> counts)
    int32_t comp_off, n_comp;   // slice of the GroupComp array
    int32_t slot_off, n_slot;   // slice of the SlotDef array (fraction fill)
    int32_t tot_slot;           // the category's total slot, -1 = none
    int32_t pad;
};

// SM count of the CURRENT device (cached per device; one process may drive several GPUs).  Launch grids are sized
// as multiples of it instead of a literal 148.
inline int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

struct Timings {
    double count_ms = 0, test_ms = 0, order_ms = 0, total_ms = 0;
    double launches = 0, count_launches = 0, count_bytes = 0;
};

// ---- kernels / launchers (defined in the .cu files) ----------------------------------------

// pack.cu
cudaError_t launch_pack_chars(const uint8_t *d_chars, const int64_t *d_row_offsets, int64_t n_rows,
                              int64_t S, int64_t W, uint64_t *d_value, uint64_t *d_valid,
                              cudaStream_t st);
// d_digit_flag: set to 1 when a literal '0' / '1' call is met (Table.hs:194-205 leaves those alone)
cudaError_t launch_pack_snp_chars(const uint8_t *d_chars, const int64_t *d_row_offsets,
                                  int64_t n_sites, int64_t S, int64_t W, uint64_t *d_hi,
                                  uint64_t *d_valid, uint64_t *d_lo, unsigned int *d_digit_flag, cudaStream_t st);
cudaError_t launch_synth(const feht_synth_spec *d_spec, int snp_mode, int64_t first_row,
                         int64_t n_rows, int64_t S, int64_t W, uint64_t *d_p0, uint64_t *d_valid,
                         uint64_t *d_lo, cudaStream_t st);
cudaError_t launch_repack_rows(const uint64_t *d_src, int64_t src_stride, uint64_t *d_dst,
                               int64_t W, int64_t S, int64_t n_rows, const uint64_t *d_and_with,
                               cudaStream_t st);

// host_pack.cpp: the same planes produced on the host cores (AVX2), rows [r0, r1) of the caller's buffer into
// [r1-r0][W] words per plane
bool host_pack_available();
void host_pack_binary(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                      uint64_t *value, uint64_t *valid);
// (the SNP packers return true when they met a literal '0' / '1' call, which the 2-bit planes cannot express)
bool host_pack_snp(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                   uint64_t *hi, uint64_t *valid, uint64_t *lo);
void host_pack_binary_scalar(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                             uint64_t *value, uint64_t *valid);
bool host_pack_snp_scalar(const uint8_t *chars, const int64_t *off, int64_t r0, int64_t r1, int64_t S, int64_t W,
                          uint64_t *hi, uint64_t *valid, uint64_t *lo);

// text_load.cu: lines of a piece of the data file's text, and their rows packed (N2 of SURVEY.md 8f)
size_t text_scratch_words(size_t piece_bytes);
// d_block_counts: per 4 KB block the lines before it (in place), [n_blocks] = newlines of the piece
cudaError_t launch_text_count(const uint8_t *d_text, int64_t len, uint32_t *d_block_counts, cudaStream_t st);
cudaError_t launch_text_starts(const uint8_t *d_text, int64_t len, const uint32_t *d_block_base,
                               int32_t *d_line_start, cudaStream_t st);
// d_flags: [0] min row length (init ~0), [1] an empty line, [2] literal '0'/'1' in SNP text
cudaError_t launch_text_pack(const uint8_t *d_text, const int32_t *d_line_start, int64_t n_lines, uint8_t delim,
                             int snp, int64_t S, int64_t W, uint64_t *p0, uint64_t *valid, uint64_t *lo,
                             int32_t *d_name_len, unsigned long long *d_flags, int num_sms, cudaStream_t st);

// count_popc.cu  (K1)
struct PopcPlan {
    int logL;      // lanes per row = 1 << logL
    int ncol;      // 16-byte columns per lane per chunk
    int nchunk;    // chunks per row
    int w4pad;     // padded 16-byte columns per mask row = (1<<logL) * ncol * nchunk
};
PopcPlan make_popc_plan(int64_t W);
// masks: [n_slots][2][w4pad] uint4 (zero padded).  counts: [n_slots][n_rows_out] int4.
cudaError_t launch_count_popc(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows,
                              int64_t W, const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                              int4 *d_counts, int num_sms, cudaStream_t st, int *launches);
cudaError_t launch_count_popc_snp(const uint64_t *d_hi, const uint64_t *d_valid,
                                  const uint64_t *d_lo, int64_t n_sites, int64_t W,
                                  const PopcPlan &plan, const uint4 *d_masks, int n_slots,
                                  int4 *d_counts, int num_sms, cudaStream_t st, int *launches,
                                  bool masks_partition = false);   // the 2 * n_slots masks cover every sample exactly once

// counts[tot_slot][row] = (sum of P, sum of N, 0, 0) over both halves of pair-slots [slot0, slot0 + n)
cudaError_t launch_sum_slots(int4 *d_counts, int64_t n_rows, int slot0, int n, int tot_slot, cudaStream_t st);

// count_gemm.cu  (K2)
int gemm_stage_count(int64_t S);
int gemm_bblock_count(int64_t S, bool fp4);
size_t gemm_pass_desc_bytes();
int gemm_max_pairs_per_pass();
// e5m2 flavour (both planes in one contraction, count_gemm.cu "third flavour"): a GEMM column = one value's samples,
// or a chunk of at most gemm_f8_max_members() of them
struct GemmF8Column {
    std::vector<int32_t> samples;
    int32_t slot = 0;        // pair-slot of the value
    int32_t tot_slot = -1;   // whole-category slot written after this column (last column of a category that has one)
    uint8_t flags = 0;       // 1 second value of the pair-slot, 2 last column of the value, 4 last of the pair-slot, 8 last of the category
};
size_t gemm_f8_pass_desc_bytes();
int gemm_f8_max_cols();
int gemm_f8_max_members();
void gemm_f8_fill_pass(void *dst, const std::vector<GemmF8Column> &cols, int n_pad, int64_t b_offset);
void gemm_f8_bake_b(const std::vector<GemmF8Column> &cols, int n_pad, int64_t S, uint8_t *out);
cudaError_t launch_count_gemm_f8(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows, int64_t W,
                                 int64_t S, const uint8_t *d_bmat, const void *d_passes, int n_passes,
                                 int4 *d_counts, int num_sms, cudaStream_t st, int *launches);
// columns[c] = samples carrying a 1 in GEMM column c of one pass; out: [n_stages][n_pad][128] bytes
void gemm_bake_b(const std::vector<std::vector<int32_t>> &columns, int n_pad, int64_t S, bool fp4, uint8_t *out);
void gemm_fill_pass(void *dst, int slot_base, int n_pairs, int n_pad, int64_t b_offset);
// counts for pair-slots [slot_base, ...) of every pass, same int4 layout as K1
cudaError_t launch_count_gemm(const uint64_t *d_value, const uint64_t *d_valid, int64_t n_rows, int64_t W,
                              int64_t S, const uint8_t *d_bmat, const void *d_passes, int n_passes, bool fp4,
                              int4 *d_counts, int num_sms, cudaStream_t st, int *launches);

// fet.cu  (K3)
// What the final gather needs about a surviving test: exactly one 32-byte sector.
struct __align__(16) SurvivorRec {
    int4 cnt;        // (goa, gob, gta, gtb)
    double ratio;
    int32_t row;     // marker row inside this context
    int32_t rank;    // name rank (tie-break of the order)
};
// ordered results, final types (what feht_result_fetch copies out)
struct ResultSoA {
    int64_t *row;
    int32_t *rank, *goa, *gob, *gta, *gtb;
    double *p, *ratio;
};
struct TestParams {
    const int4 *counts;       // [n_slots][n_rows]
    const ScreenGroup *groups;
    const FracDef *fracs;
    const GroupComp *gcomps;
    const SlotDef *slots;
    int n_groups;
    int max_frac, max_comp, max_slot;   // largest group (shared-memory sizing)
    int64_t n_rows;
    int n_comps;
    double ratio_filter;
    const int32_t *name_rank; // may be null: the rank is then the global row index, row_base + row
    int64_t row_base;
    // survivors: one 32-byte record each + the comparison index (a sort digit source)
    SurvivorRec *rec;
    int32_t *r_comp;
    // sort keys by |ratio|, built here when the order is the reference's (null: K3b builds p-value keys)
    unsigned long long *key;
    uint32_t *val;
    uint32_t *or_and;
    unsigned long long *seg_count;   // filtered output: [n_comps] survivors per comparison (zeroed by the caller)
    unsigned long long *cursor;      // filtered output: next free survivor slot (zeroed by the caller)
    long long capacity;              // filtered output: slots the survivor arrays hold
    int dense;                       // 1: every test survives, index = comp*n_rows + row
    int dense_by_rank;               // dense and name_rank is a permutation: index = comp*n_rows + rank
    int merge_comp;                  // 2..51 comparisons: the comparison index rides in the key's top byte (fet.cu)
    // filtered output: rows that can reach the ratio filter, per group of comparisons, listed by the prescreen kernel:
    // live_rows[g * n_rows ..], live_count[g] (zeroed by the launcher).  Null: the screen kernel walks every tile and
    // decides per tile itself.
    int32_t *live_rows;
    unsigned int *live_count;
};
// counts how many distinct in-range values rank[0..n) holds (== n iff it is a permutation of [0,n))
cudaError_t launch_count_distinct_ranks(const int32_t *rank, int64_t n, uint32_t *flags_words,
                                        unsigned long long *out_count, cudaStream_t st);
cudaError_t launch_test_emit(const TestParams &p, cudaStream_t st);
// dense output of ONE explicit comparison counted in pair-slot `slot`: thread = row, no shared state
cudaError_t launch_test_emit_single(const TestParams &p, int slot, cudaStream_t st);
// OR/AND accumulators of the sort-key words (rank, key lo, key hi, comparison): {0,0,0,0,~0,~0,~0,~0}
cudaError_t launch_init_or_and(uint32_t *d_or_and, cudaStream_t st);
// K3b over n survivors: p (+ Bonferroni).
//  * plain (perm_a == null): test i reads rec[i], writes d_p[i]; with d_key != null it also builds the p-value
//    sort keys + OR/AND words (FEHT_SORT_PVALUE_ASC);
//  * fused with the final gather (perm_a != null, the reference's |ratio| order): test i is the survivor
//    perm[i] (perm = *d_sel ? perm_b : perm_a), and every output column of row i is written, p included.
struct PvalueIo {
    const SurvivorRec *rec;
    const int32_t *r_comp;
    double *d_p;
    unsigned long long *d_key;
    uint32_t *d_val, *d_or_and;
    const uint32_t *perm_a, *perm_b, *d_sel;
    ResultSoA out;
    int64_t row_base;
    // N4, optional: exact tests with the same 2x2 table are evaluated once.  dd_keys: open-addressing
    // table of dd_mask + 1 packed tables (zeroed by the caller), dd_owner: the test that computes a slot's p,
    // dd_list / dd_count: the tests that wait for it (zeroed count); pvalue_fixup copies the owners' p afterwards.
    unsigned long long *dd_keys;
    uint32_t *dd_owner;
    uint32_t dd_mask;
    uint2 *dd_list;
    uint32_t *dd_count;
};
cudaError_t launch_pvalues(const PvalueIo &io, int64_t n, const double *d_choose, int correction,
                           double bonferroni_n, int one_tail, bool hyp_heavy, cudaStream_t st);
cudaError_t launch_eval_tables(const int4 *d_tables, int64_t n, const double *d_choose, int one_tail,
                               double *d_p, double *d_ratio, SurvivorRec *d_rec, cudaStream_t st);

// sort.cu  (K4)
struct SortWorkspace {
    unsigned long long *key_a = nullptr, *key_b = nullptr;  // 64-bit sort keys (ping-pong)
    uint32_t *val_a = nullptr, *val_b = nullptr;            // survivor indices (ping-pong)
    uint32_t *control = nullptr;  // [16][256] histograms, [16] tickets, [passes][tiles][256] status
    uint32_t *or_and = nullptr;   // 8 words
    uint32_t *coop = nullptr;     // cooperative sort: [num_sms][256] tagged CTA digit counts, [num_sms] flags, selector
    uint32_t coop_launches = 0;   // tag generator of the cooperative sort (host side)
    int coop_clean_rows = 0;      // rows of the count table known to hold only tags of recent launches
};
size_t sort_status_words(int64_t n);
// Sorts n survivors by (comp asc, key64 asc, rank asc); rank_presorted = the survivors already sit
// in name-rank order inside every comparison (the stable sort then needs no rank passes).
// Returns the permutation (device pointer inside ws) and adds the number of kernel launches.
cudaError_t radix_sort_perm(SortWorkspace &ws, int64_t n, const SurvivorRec *rec, const int32_t *r_comp,
                            bool rank_presorted, cudaStream_t st, uint32_t **perm_out, int *launches);
// Cooperative engine: every pass inside one launch, the pass list decided on the device (no host
// synchronisation); one chunk per CTA up to coop_sort_capacity, several beyond.  The permutation ends up in
// ws.val_a or ws.val_b as the device word coop_sort_selector() says (0 / 1).
int64_t coop_sort_capacity(int num_sms);
int coop_sort_max_grid();
size_t coop_sort_control_words(int num_sms);
cudaError_t coop_sort_perm(SortWorkspace &ws, int64_t n, const SurvivorRec *rec, const int32_t *r_comp,
                           bool rank_presorted, int num_sms, cudaStream_t st, int *launches);
const uint32_t *coop_sort_selector(const SortWorkspace &ws, int num_sms);
// permutation = d_sel && *d_sel ? perm_b : perm_a
cudaError_t launch_gather_results(const uint32_t *perm_a, const uint32_t *perm_b, const uint32_t *d_sel, int64_t n,
                                  int64_t row_base, const SurvivorRec *rec, const double *r_p, const ResultSoA &o,
                                  cudaStream_t st);
// K5: device k-way merge by ranking (dest[i][j] = final position of element j of shard i)
int merge_max_shards();
cudaError_t launch_merge_rank(int n_shards, const int64_t *n, const double *const *d_key,
                              const int32_t *const *d_rank, int sort_key, int64_t *const *d_dest,
                              cudaStream_t st);
cudaError_t launch_scatter_rows(const void *d_src, void *d_dst, const int64_t *d_dest, int64_t n,
                                int elem_bytes, cudaStream_t st);
// K5b: every comparison of a row-sharded job merged in ONE launch, all eight columns scattered by the same kernel.
// A shard's ordered rows of all comparisons lie back to back in `res` (n rows); d_seg_off = its [n_comps + 1] segment
// offsets on the device.  d_map_* (optional): local unit -> global unit table of a shard that was filled by several
// uploads (multi-device context); mult = marker rows per unit.
struct MergeShardView {
    ResultSoA res;
    int64_t n = 0;
    const int64_t *d_seg_off = nullptr;
    const int64_t *d_map_local = nullptr, *d_map_global = nullptr;
    int n_map = 0, mult = 1;
};
// d_dest_scratch: optional int64 [sum of the shards' rows]; with it, jobs of more than two shards rank first and then
// place the rows tile by tile with coalesced writes.
// d_bounds_scratch: int64 [merge_bounds_words(rows, shards)]: where every 256-row block starts in the other shards.
size_t merge_bounds_words(int64_t total_rows, int n_shards);
cudaError_t launch_merge_all(int n_shards, int n_comps, int sort_key, const MergeShardView *shards,
                             const int64_t *d_out_seg_off, const ResultSoA &out, int64_t row_base, int64_t *d_dest_scratch,
                             int64_t *d_bounds_scratch, cudaStream_t st);

// choose_table.cpp
const std::vector<double> &host_choose_table();

}  // namespace feht
