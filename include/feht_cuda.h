/*
 * feht_cuda.h -- C ABI of libfeht_cuda.so, the B200 (sm_100a) implementation of feht's
 * marker-discovery hot path.
 *
 * What it replaces in the reference (paths relative to /root/reference):
 *     app/Main.hs:26-29
 *         let resultMap       = generateResultMap (geneVectorMap parsedDataFile) cl
 *         let finalGroupComps = applyMultipleTestingCorrection (correction userArgs) resultMap
 *         let tableOfComps    = formatComparisonResultsAsTable (ratioFilter userArgs) finalGroupComps
 *     i.e. src/Comparison.hs:46-125 (count, fet, ratio), :238-264 (Bonferroni),
 *     :139-162 (ratio filter, |ratio|-descending order) and src/FET.hs:91-139.
 * The reference has no FFI of its own (src/StorableByteString.hs is unused); the Haskell host
 * binds these entry points with `foreign import ccall safe` -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns FEHT_OK (0) or a negative FEHT_E_* code; feht_last_error()
 *     returns the message (the host turns it into `error`, the reference's only failure mode);
 *   - plain pointers and sizes only; host pointers need only stay valid for the duration of
 *     the call (the library copies to the device before returning);
 *   - feht_create() makes a context that drives one GPU; feht_create_multi() makes ONE handle that shards the
 *     marker rows over several GPUs of the box and merges the ordered shards on the device, behind the same entry
 *     points (the Haskell host needs nothing else).  A host that runs one process per GPU (torchrun) makes one
 *     single-device context per process, moves the ordered shards to one GPU itself and merges them there with
 *     feht_merge_shards_device().  No context may be used from two threads at once.
 *   - there is no CPU fallback: without a usable sm_100-class device feht_create() fails.
 */
#ifndef FEHT_CUDA_H
#define FEHT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct feht_ctx feht_ctx;

enum {
    FEHT_OK = 0,
    FEHT_E_INVALID = -1,   /* bad argument / call order                                   */
    FEHT_E_CUDA = -2,      /* CUDA runtime error (message has the CUDA string)            */
    FEHT_E_BOUNDS = -3,    /* a group column index is outside a row (V.! in the reference) */
    FEHT_E_NOMEM = -4,
    FEHT_E_NODEVICE = -5   /* no CUDA device / not an sm_100-class device                  */
};

/* UserInput.hs:20  data Correction = Bonferroni | None */
enum { FEHT_CORRECTION_NONE = 0, FEHT_CORRECTION_BONFERRONI = 1 };

/* Ordering of each comparison's rows.  The reference orders by abs(ratio) descending
 * (Comparison.hs:155,160-162); ties are broken by name_rank ascending. */
enum { FEHT_SORT_ABS_RATIO_DESC = 0, FEHT_SORT_PVALUE_ASC = 1 };
/* FETMode (FET.hs:82); the reference's hot path always asks for TwoTail (Comparison.hs:51) */
enum { FEHT_FET_TWO_TAIL = 0, FEHT_FET_ONE_TAIL = 1 };

/* ---------------------------------------------------------------- context ---- */
/* Replaces: nothing (the reference is a pure function); owns the device state. */
int  feht_create(feht_ctx **out, int device);
/*
 * One handle for n_devices GPUs (north_star: "marker rows shard naturally across the 8 GPUs of one
 * box; each GPU counts, tests and sorts its shard; the final per-comparison merge is a k-way merge of the sorted
 * shards, with no other collectives").  Every upload call splits its rows contiguously over the devices (piece d of n
 * rows = [n*d/D, n*(d+1)/D)) and uploads the pieces in parallel; comparisons and options are replicated; feht_run
 * starts all shards side by side with Bonferroni's n = the job's row count (Comparison.hs:251), copies the ordered
 * shards to device_ids[0] (cudaMemcpyPeerAsync, NVLink between peers) and merges ALL comparisons there with one kernel
 * launch; feht_result_* then return the merged rows -- the same rows in the same order as one context holding
 * everything.  A device id may be repeated (several shards on one GPU).  Differences from a single-device context:
 * feht_set_stream is refused, feht_run_async blocks like feht_run.
 */
int  feht_create_multi(feht_ctx **out, const int *device_ids, int n_devices);
int  feht_num_devices(const feht_ctx *ctx);   /* shards behind the handle (1 for feht_create) */
void feht_destroy(feht_ctx *ctx);
/* Message of the last failure on ctx (or of the last failed feht_create if ctx == NULL). */
const char *feht_last_error(const feht_ctx *ctx);
/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = the context's own. */
int  feht_set_stream(feht_ctx *ctx, void *cuda_stream);
/* Library / kernel build identification, e.g. "feht_b200 0.1 sm_100a". */
const char *feht_version(void);

/* ---------------------------------------------------------------- data matrix ---- */
/* S = number of data-file columns (Table.hs:273,277).  Must be called first; drops any rows AND every comparison /
 * category (their column lists and per-sample value vectors belong to the old sample count). */
int feht_set_samples(feht_ctx *ctx, int64_t n_samples);
/* Optional: pre-size device storage for n_rows marker rows (binary) or sites (SNP). */
int feht_reserve_rows(feht_ctx *ctx, int64_t n_rows);
/* Drop all rows (keeps samples and comparisons). */
int feht_clear_rows(feht_ctx *ctx);

/*
 * Append marker rows exactly as the reference holds them: GeneVectorMap's `V.Vector Char`
 * (Table.hs:50, built :210-224) narrowed to bytes, row r = chars[row_offsets[r] .. row_offsets[r+1]).
 * '1' -> present, '0' -> absent, anything else -> missing (Comparison.hs:53-56; README.md:227-234).
 * Row length is the number of characters of the concatenated cells (Table.hs:170) and may differ
 * from S; a comparison that indexes past the end of a row fails feht_run with FEHT_E_BOUNDS,
 * as `V.!` would.  SNP-mode rows arrive here already expanded to _a/_t/_c/_g binary rows
 * (Table.hs:181-205), so this entry point is mode-agnostic.
 * name_rank[r] = rank of the marker name in byte order (tie-break of the final ordering);
 * NULL = the global row index.  The chars are packed on upload to 1 bit per call plus a
 * valid(=not missing) bit-plane.
 */
int feht_add_rows_chars(feht_ctx *ctx, const uint8_t *chars, const int64_t *row_offsets,
                        int64_t n_rows, const int32_t *name_rank);
/*
 * Uploads of 64 MB and more are split between two producers of the same bit-planes: the DMA engine moves
 * pieces raw (1 B/call) for the GPU pack kernel, and n_threads host threads transcode pieces from the other end
 * of the row range (AVX2) and upload them packed (2 bits/call), because the PCIe link, not the GPU, bounds this
 * call.  n_threads: 0 = DMA only, < 0 = default (env FEHT_HOST_PACK_THREADS, else every core).  Only the wire
 * format is produced on the host; there is no CPU path for counting / testing / ordering.
 */
int feht_set_host_threads(feht_ctx *ctx, int n_threads);
/*
 * The same transcoding as a plain host function (no context, no GPU): chars -> bit-planes of row_stride_words
 * 64-bit words per row, zero padded, on n_threads host threads (<= 0: every core; AVX2 when the CPU has it).
 * Binary mode: plane0 = value, plane1 = valid, plane2 unused.  SNP mode: plane0 = hi, plane1 = lo, plane2 = valid
 * -- the argument order of feht_add_rows_packed / feht_add_snp_packed, which take the result.
 */
int feht_pack_chars_host(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows, int64_t n_samples,
                         int snp_mode, int64_t row_stride_words, uint64_t *plane0, uint64_t *plane1,
                         uint64_t *plane2, int n_threads);
/* rows of the last chars upload that the host threads transcoded (diagnostics / tests) */
int64_t feht_last_upload_host_rows(const feht_ctx *ctx);

/*
 * The data file's text, straight to the device.  Replaces Table.hs:267-279 (`BS.lines`, the '\r'
 * filter, `BS.split d`) and Table.hs:160-191 (first cell = marker name, the remaining cells concatenated; the
 * character index in that concatenation is the column number) for the BODY of the file: `text` starts after the
 * header line.  Every line becomes one row (binary) or one site (snp_mode; expanded to _a,_t,_c,_g on the device,
 * Table.hs:184-205), in file order; lines are split at '\n', a final line without one counts.  The text is copied
 * in pieces of whole lines and parsed + packed by text_load.cu; nothing is parsed on the host.
 * Errors (nothing is appended): an empty line (the reference dies on `g:gnxs = xs`), a line longer than 64 MB,
 * SNP text with literal '0' / '1' calls (they count for all four alleles: use feht_add_rows_chars with the expanded
 * rows).  Marker names stay in the caller's text: feht_text_lines() gives, per line of the last call, the offset of
 * the line in `text` and the byte length of its first cell (a '\r' inside the name is counted; duplicate names are
 * the caller's to detect -- the reference keeps the last of them, Table.hs:210-224).  Name ranks for the tie-break
 * are set afterwards for all marker rows at once with feht_set_name_rank() (SNP: four per site, in the order
 * _a,_t,_c,_g); without it the rank is the global row index.
 */
int feht_add_rows_text(feht_ctx *ctx, const char *text, int64_t len, char delimiter, int snp_mode,
                       int64_t *n_lines_out);
int feht_text_lines(feht_ctx *ctx, int64_t *line_begin, int32_t *name_len, int64_t cap);
int feht_set_name_rank(feht_ctx *ctx, const int32_t *name_rank, int64_t n_marker_rows);

/*
 * Append rows already packed: bit s of word (s/64) of a row = sample s; value bit set iff the
 * call is '1', valid bit set iff it is '1' or '0'.  row_stride_words >= ceil(S/64) 64-bit words.
 * Benchmark-scale entry (the reference layout is 4 B per call and cannot hold configs 2-5).
 */
int feht_add_rows_packed(feht_ctx *ctx, const uint64_t *value_plane, const uint64_t *valid_plane,
                         int64_t n_rows, int64_t row_stride_words, const int32_t *name_rank);

/*
 * SNP mode, packed: per site a 2-bit nucleotide code (hi,lo planes; A=0, C=1, G=2, T=3) and a
 * valid plane (0 = not one of A/C/G/T -> missing, Table.hs:194-205).  Each site yields four
 * marker rows in the reference's order _a, _t, _c, _g (Table.hs:184-191); marker row index =
 * 4*site + {0:a, 1:t, 2:c, 3:g}.  name_rank has 4*n_sites entries (or NULL).
 * A context holds either binary rows or SNP sites, not both.
 */
int feht_add_snp_packed(feht_ctx *ctx, const uint64_t *hi_plane, const uint64_t *lo_plane,
                        const uint64_t *valid_plane, int64_t n_sites, int64_t row_stride_words,
                        const int32_t *name_rank);
/* SNP sites as characters (one char per sample per site, upper or lower case).  A literal '0' / '1' call -- which the
 * reference's replaceChar leaves alone, so that it counts for all four allele rows (Table.hs:194-205) -- cannot be
 * held in the 2-bit planes: the call fails with FEHT_E_INVALID and appends nothing (upload the expanded _a/_t/_c/_g
 * rows with feht_add_rows_chars instead; feht_pack_chars_host reports the same condition the same way). */
int feht_add_snp_chars(feht_ctx *ctx, const uint8_t *chars, const int64_t *row_offsets,
                       int64_t n_sites, const int32_t *name_rank);

/* Counter-based synthetic matrix generated on the device; the same integer
 * recipe is restated in oracle/feht_oracle.c so that a CPU checker can make any row. */
typedef struct {
    uint64_t seed;
    int64_t  n_samples;
    uint32_t missing_thr;         /* cell missing iff hash_hi < missing_thr (of 2^32)             */
    uint32_t planted_per_million; /* row is planted iff (row_hash >> 8) % 1e6 < planted_per_million */
    int64_t  split_col;           /* planted rows: cols < split_col use thr_hi, else thr_lo        */
    uint32_t thr_hi, thr_lo;
    uint32_t freq_thr[256];       /* per-row presence threshold, indexed by row_hash & 255         */
} feht_synth_spec;
/* Appends n_rows rows (binary) or sites (snp_mode != 0) with global indices first_row.. */
int feht_generate_synthetic(feht_ctx *ctx, const feht_synth_spec *spec, int snp_mode,
                            int64_t first_row, int64_t n_rows);

/* Read back the packed planes of rows [row0,row0+n) (tests / debugging).  plane: 0 value (SNP: hi),
 * 1 valid, 2 lo (SNP only).  out has n * feht_row_stride_words() words. */
int     feht_read_plane(feht_ctx *ctx, int plane, int64_t row0, int64_t n, uint64_t *out);
int64_t feht_row_stride_words(const feht_ctx *ctx);
int64_t feht_num_rows(const feht_ctx *ctx);      /* marker rows (SNP: 4 per site) */

/* ---------------------------------------------------------------- comparisons ---- */
/*
 * One comparison with explicit column-index lists = getListOfColumns of compGroup1/compGroup2
 * (Comparison.hs:57-58,61-68).  Groups may overlap (--one/--two naming different categories,
 * Comparison.hs:283-284).  Returns the comparison's index (>= 0) or a negative error.
 */
int64_t feht_add_comparison(feht_ctx *ctx, const int32_t *g1_cols, int64_t n1,
                            const int32_t *g2_cols, int64_t n2);
/*
 * A whole metadata category: value_of_sample[s] in [0,n_values) or -1 when sample s has no
 * metadata row / lacks the category.  Generates, as getAllPermutations does (Comparison.hs:319-331),
 * all pairs (i,j), i<j (getAllListPairs :362-364) followed, when n_values > 2, by one-vs-rest for
 * every value (getOva :334-344; "rest" = samples of the category with a different value,
 * Table.hs:243-247).  Order: pairs lexicographic, then one-vs-rest 0..n_values-1.
 * Returns the index of the first generated comparison, or a negative error.
 */
int64_t feht_add_category(feht_ctx *ctx, const int32_t *value_of_sample, int32_t n_values);
int     feht_clear_comparisons(feht_ctx *ctx);
int64_t feht_num_comparisons(const feht_ctx *ctx);
/* For comparison c: kind 0 explicit, 1 pair, 2 one-vs-rest; category index; v1; v2 (-1 for rest). */
int     feht_comparison_info(const feht_ctx *ctx, int64_t c, int32_t *kind, int32_t *category,
                             int32_t *v1, int32_t *v2);

/* Counting engine: 0 auto, 1 popcount scan (K1), 2 int8 tcgen05 GEMM (K2; categories only), 3 the same GEMM
 * with 4-bit operands (tcgen05 kind::mxf4, unit scales, FP32 accumulation: exact, half the operand bytes), 4 both
 * planes in one e5m2 GEMM (kind::f8f6f4: a call is 0, 1 or 4096, the FP32 sum P + 4096 (N - P) is exact for columns of
 * at most 4,095 samples; larger value groups take several columns; a category must fit 128 columns). */
enum { FEHT_ENGINE_AUTO = 0, FEHT_ENGINE_POPC = 1, FEHT_ENGINE_GEMM = 2, FEHT_ENGINE_GEMM_FP4 = 3, FEHT_ENGINE_GEMM_F8 = 4 };
int feht_set_count_engine(feht_ctx *ctx, int engine);

/* ---------------------------------------------------------------- run ---- */
/*
 * count -> fet/chi-square -> Bonferroni -> ratio filter -> order, for every comparison.
 *   correction     FEHT_CORRECTION_*                                  (Comparison.hs:238-242)
 *   ratio_filter   keep rows with ratio_filter <= abs(ratio)          (Comparison.hs:148)
 *   sort_key       FEHT_SORT_*                                        (Comparison.hs:160-162)
 *   bonferroni_n   the `length crxs` of Comparison.hs:251 = marker rows of the WHOLE job;
 *                  0 = the rows held by this context (single-GPU case).  A row-sharded job
 *                  passes the global row count to every shard; no collective is needed.
 */
int feht_run(feht_ctx *ctx, int correction, double ratio_filter, int sort_key,
             int64_t bonferroni_n);
/* The same work, enqueued on the context's stream without waiting for it: the call returns as soon as the
 * host has nothing left to do (immediately for ratio_filter <= 0, where the output size is known up front;
 * a filtered run still waits for the survivor counts).  feht_result_* / feht_last_timings synchronise.  A
 * host that pipelines several data sets (or the benchmark's back-to-back steps) uses this; the Haskell
 * shim calls the blocking feht_run. */
int feht_run_async(feht_ctx *ctx, int correction, double ratio_filter, int sort_key,
                   int64_t bonferroni_n);

/* FETMode of the p-values feht_run and feht_eval_tables compute (FET.hs:101-111): TwoTail (default, what
 * Comparison.hs:51 asks for) or OneTail = chi-square p / 2 for table totals >= 1000, else
 * `cumulative (hypergeometric m l k) goa`. */
int feht_set_fet_mode(feht_ctx *ctx, int mode);

/*
 * Emit only the comparisons [first, first + count) in the following runs; count < 0 = all comparisons again.
 * For outputs that do not fit one run (fewer than 2^30 result rows per run): the reference's default ratioFilter 0
 * keeps every test (Comparison.hs:148), so a category of V values over R marker rows yields (V(V-1)/2 + V) * R
 * rows.  The host walks the comparisons window by window -- set the window, feht_run, feht_result_fetch the
 * window's comparisons (the others report 0 rows), next window.  The contingency counts do not depend on the
 * window: the first window's run computes them, the following windows start at the test stage (until rows,
 * comparisons, the engine or the sample count change, or the window is dropped with count < 0).
 */
int feht_set_comparison_window(feht_ctx *ctx, int64_t first, int64_t count);

/* Surviving rows of comparison c, in final order. */
int64_t feht_result_count(const feht_ctx *ctx, int64_t comparison);
int64_t feht_result_total(const feht_ctx *ctx);
/*
 * Fetch comparison c's rows in final order into caller buffers of feht_result_count() entries
 * (any pointer may be NULL): the FETResult fields (FET.hs:84-89) + compRatio (Comparison.hs:37-40).
 * row = marker row index in order of insertion (+ row_base, see feht_set_row_base).
 * comparison = -1 fetches the rows of EVERY comparison, back to back in comparison order (feht_result_total()
 * entries; comparison c starts at the sum of feht_result_count() of the comparisons before it).
 */
int feht_result_fetch(feht_ctx *ctx, int64_t comparison, int64_t *row, int32_t *goa, int32_t *gob,
                      int32_t *gta, int32_t *gtb, double *p, double *ratio, int32_t *name_rank);
/* The same, into DEVICE buffers on the context's GPU (e.g. to hand a shard's sorted rows to a
 * GPU-to-GPU transport before feht_merge_sorted_device). */
int feht_result_fetch_device(feht_ctx *ctx, int64_t comparison, int64_t *d_row, int32_t *d_goa,
                             int32_t *d_gob, int32_t *d_gta, int32_t *d_gtb, double *d_p, double *d_ratio,
                             int32_t *d_name_rank);
/* Global index of this context's first row (row-sharded jobs); default 0 (set before feht_run). */
int feht_set_row_base(feht_ctx *ctx, int64_t row_base);

/* Device-side milliseconds of the last feht_run, by stage (CUDA events on the run's stream):
 * out[0] count, out[1] test+correct+filter, out[2] order, out[3] total, out[4] launches (count),
 * out[5] count-kernel launches, out[6] algorithmic bytes read by the count stage,
 * out[7] engine that counted the categories (FEHT_ENGINE_POPC, FEHT_ENGINE_GEMM or FEHT_ENGINE_GEMM_FP4).
 * Multi-device handle: out[0..2] are the slowest shard's stages, out[3] = slowest shard + merge, out[8] = the merge
 * (peer copies + K5b) on the first device; launches and bytes are summed over the shards. */
int feht_last_timings(const feht_ctx *ctx, double *out, int n);
/* Sums over every run since the last reset (waits for the enqueued runs): out[0] runs, out[1] count ms,
 * out[2] test ms, out[3] order ms, out[4] total ms, out[5] kernel launches, out[6] count-kernel launches. */
int feht_timings_accumulated(feht_ctx *ctx, double *out, int n, int reset);

/*
 * Host k-way merge of sorted shards of one comparison: shard i has n[i] rows with
 * keys ratio[i][..] (or p[i][..] when sort_key = FEHT_SORT_PVALUE_ASC) and name_rank[i][..];
 * writes, in merged order, (shard, index) pairs.  Pure host code.
 */
int feht_merge_sorted(int n_shards, const int64_t *n, const double *const *key,
                      const int32_t *const *name_rank, int sort_key, int32_t *out_shard,
                      int64_t *out_index);

/*
 * The same k-way merge on the device (K5), for shards that were moved to this context's GPU
 * (NVLink / NCCL): all pointers are device pointers; d_dest[i][j] receives the final position of
 * element j of shard i in the merged order of all sum(n) rows.  feht_scatter_rows_device then moves
 * any per-row field (elem_bytes each) of a shard to its merged place: d_dst[d_dest[j]] = d_src[j].
 * At most 32 shards.
 */
int feht_merge_sorted_device(feht_ctx *ctx, int n_shards, const int64_t *n, const double *const *d_key,
                             const int32_t *const *d_name_rank, int sort_key, int64_t *const *d_dest);
int feht_scatter_rows_device(feht_ctx *ctx, const void *d_src, void *d_dst, const int64_t *d_dest,
                             int64_t n, int32_t elem_bytes);

/*
 * The merge of a row-sharded job in one call (K5b), for hosts that run one process per GPU: every shard's ordered rows
 * of ALL comparisons, back to back in the order feht_result_fetch_device returns them comparison by comparison, have
 * been moved to this context's GPU; seg_off (HOST array, n_comparisons + 1 entries, seg_off[0] = 0) says where each
 * comparison's rows start inside a shard's columns.  One kernel launch ranks every row against the other shards'
 * segments of its comparison (binary searches on the total order: key bits, name rank, shard) and writes all eight
 * columns to their merged places.  `out`: device columns with room for the sum of all shards' rows; out->seg_off
 * (HOST, n_comparisons + 1) receives the merged offsets.  At most 32 shards.
 */
typedef struct {
    const int64_t *row;
    const int32_t *goa, *gob, *gta, *gtb;
    const double  *p, *ratio;
    const int32_t *name_rank;
    const int64_t *seg_off;
} feht_shard_rows;
int feht_merge_shards_device(feht_ctx *ctx, int n_shards, int64_t n_comparisons, const feht_shard_rows *shards,
                             int sort_key, const feht_shard_rows *out);

/* Host-built table of the reference's `choose` values C[n][k'], n < 1000, k' <= n/2 (the
 * arithmetic of math-functions 0.2.1.0), as uploaded to the device; for tests.
 * Returns the number of entries; copies min(cap, entries) doubles into out. */
int64_t feht_choose_table(double *out, int64_t cap);
/* Device evaluation of p for explicit 2x2 tables (tests the K3 arithmetic directly). */
int feht_eval_tables(feht_ctx *ctx, const int32_t *goa, const int32_t *gob, const int32_t *gta,
                     const int32_t *gtb, int64_t n, double *p_out, double *ratio_out);

#ifdef __cplusplus
}
#endif
#endif /* FEHT_CUDA_H */
