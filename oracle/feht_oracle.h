/*
 * feht_oracle.h -- CPU restatement of feht's marker-discovery hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under feht_b200/ may include, link or call
 * this.  Allowed users: tests/, __graft_entry__.smoke(), and bench.py's
 * cpu_baseline / --impl reference legs (as the checker / baseline, never as the
 * product path).
 *
 * Parity status: PINNED.  The arithmetic is checked bit-for-bit against the 8
 * known-answer values of /root/reference/test/unit/FETSpec.hs:13-100, the
 * counting cases of test/unit/ComparisonSpec.hs:43-57 and the published rows
 * of README.md:121-137,179,219-222 (tests/test_oracle_golden.py).
 *
 * The reference is Haskell (no GHC in this image, so it cannot be compiled
 * into oracle/_ref).  Its p-value arithmetic lives in two un-vendored
 * third-party packages pinned by /root/reference/cabal.config:
 *     statistics      ==0.13.3.0   (cabal.config:2025)
 *     math-functions  ==0.2.1.0    (cabal.config:1381)
 * whose published algorithms are restated here; each function cites the
 * reference call site (file:line under /root/reference) it serves.
 */
#ifndef FEHT_ORACLE_H
#define FEHT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- [3P] math-functions 0.2.1.0, Numeric.SpecFunctions / Numeric.Polynomial.Chebyshev ---- */
double fo_chebyshev_broucke(double x, const double *cs, int n);
double fo_log1p(double x);
double fo_log_gamma_correction(double x);
double fo_log_beta(double a, double b);
double fo_log_gamma(double x);                 /* AS 245 */
double fo_choose(int n, int k);
double fo_incomplete_gamma(double p, double x); /* AS 239 */

/* ---- [3P] statistics 0.13.3.0 ---- */
double fo_hypergeometric_probability(int m, int l, int k, int n);
double fo_hypergeometric_cumulative(int m, int l, int k, double x);
double fo_chi_squared1_compl_cumulative(double x);

/* ---- FET.hs ---- */
#define FO_TWO_TAIL 0
#define FO_ONE_TAIL 1
double fo_chi_square(int64_t a, int64_t b, int64_t c, int64_t d);          /* FET.hs:127-139 */
double fo_fet(int64_t goa, int64_t gob, int64_t gta, int64_t gtb, int mode); /* FET.hs:91-125 */

/* ---- Comparison.hs ---- */
/* Comparison.hs:116-125.  Returns -1 if an index is outside [0,len) (V.! would error). */
int64_t fo_count_char_in_vector_by_indices(const uint8_t *v, int64_t len, uint8_t match,
                                           const int32_t *idx, int64_t n_idx);
double fo_calculate_ratio(int64_t goa, int64_t gob, int64_t gta, int64_t gtb); /* Comparison.hs:71-82 */
double fo_bonferroni(double p, int64_t n);                                     /* Comparison.hs:255-264 */

/* Table.hs:194-205 */
uint8_t fo_replace_char(uint8_t match_char, uint8_t current_char);

/*
 * One comparison over a block of marker rows: Comparison.hs:86-111 (count, fet, ratio),
 * 238-264 (Bonferroni with n = bonferroni_n), no filter, no sort.
 * chars/row_offsets: row r is chars[row_offsets[r] .. row_offsets[r+1]).
 * Outputs are arrays of n_rows.  Returns 0, or -1 on an out-of-range column index.
 * correction: 0 none, 1 bonferroni.
 */
int fo_run_comparison(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows,
                      const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2,
                      int correction, int64_t bonferroni_n,
                      int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
                      double *p, double *ratio);

/* Same, rows [row_begin,row_end) sharded over n_threads pthreads (cpu_baseline "all cores" leg). */
int fo_run_comparison_mt(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows,
                         const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2,
                         int correction, int64_t bonferroni_n,
                         int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
                         double *p, double *ratio, int n_threads);

/*
 * Comparison.hs:148 (filter ratio_filter <= |ratio|) and 155/160-162 (|ratio| descending),
 * ties broken by name_rank ascending (name_rank == NULL: by row index).
 * Writes the surviving row indices, in final print order, to out_rows (capacity n_rows);
 * returns how many survived.
 */
int64_t fo_filter_and_order(const double *ratio, const int32_t *name_rank, int64_t n_rows,
                            double ratio_filter, int64_t *out_rows);

/* When non-zero (default), fo_choose memoises C[n][k'] for n < 1000; set 0 for the
 * reference-style cost profile (every `choose` recomputed, as Numeric.SpecFunctions does). */
void fo_set_choose_cache(int enabled);

/* ---- synthetic inputs: counter-based, so any row can be made anywhere ---- */
typedef struct {
    uint64_t seed;
    int64_t  n_samples;
    uint32_t missing_thr;        /* cell missing iff h_miss < missing_thr (of 2^32)           */
    uint32_t planted_per_million;/* row is "planted" iff h_row % 1000000 < planted_per_million */
    int64_t  split_col;          /* planted rows: cols < split_col use thr_hi else thr_lo      */
    uint32_t thr_hi, thr_lo;
    uint32_t freq_thr[256];      /* per-row presence threshold table (quantiles of the row-frequency law) */
} fo_synth_spec;

/* Binary-mode cell for (row, col): returns '1', '0' or '-'. */
uint8_t fo_synth_binary_cell(const fo_synth_spec *sp, int64_t row, int64_t col);
/* SNP-mode cell: returns one of 'A','C','G','T','-'. */
uint8_t fo_synth_snp_cell(const fo_synth_spec *sp, int64_t row, int64_t col);
/* Fill n_rows x n_samples chars (row-major, no separators) for global rows row0.. */
void fo_synth_fill(const fo_synth_spec *sp, int snp_mode, int64_t row0, int64_t n_rows, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
