// multi.cuh -- the multi-device context behind feht_create_multi.
//
// One handle, several GPUs: marker rows of every upload are split contiguously over the listed devices, each shard
// is an ordinary single-device context ("kid") that counts, tests and orders its rows on its own stream, the ordered
// shards are moved to the first device with cudaMemcpyPeerAsync (NVLink) and merged there by ONE launch of K5b
// (sort.cu: merge_all_kernel) for all comparisons.  The only cross-row quantity, Bonferroni's n
// (Comparison.hs:251), is the job's total row count -- a host scalar handed to every shard; there is no collective.
// capi.cu forwards every entry point of include/feht_cuda.h here when the handle is a multi-device one.
#pragma once

#include "common.cuh"

namespace feht {

struct MultiState;

namespace multi {

MultiState *create(const int *device_ids, int n, std::string *err);   // null on failure (err says why)
void destroy(MultiState *m);
const std::string &last_error(const MultiState *m);

int set_samples(MultiState *m, int64_t n_samples);
int reserve_rows(MultiState *m, int64_t n_rows);
int clear_rows(MultiState *m);
int add_chars(MultiState *m, const uint8_t *chars, const int64_t *row_offsets, int64_t n, const int32_t *name_rank, int snp);
int add_packed(MultiState *m, const uint64_t *p0, const uint64_t *p1, const uint64_t *p2, int64_t n, int64_t stride,
               const int32_t *name_rank, int snp);
int add_text(MultiState *m, const char *text, int64_t len, char delimiter, int snp_mode, int64_t *n_lines_out);
int text_lines(MultiState *m, int64_t *line_begin, int32_t *name_len, int64_t cap);
int set_name_rank(MultiState *m, const int32_t *name_rank, int64_t n_marker_rows);
int generate_synthetic(MultiState *m, const feht_synth_spec *spec, int snp_mode, int64_t first_row, int64_t n_rows);
int read_plane(MultiState *m, int plane, int64_t row0, int64_t n, uint64_t *out);
int64_t row_stride_words(const MultiState *m);
int64_t num_rows(const MultiState *m);
int set_host_threads(MultiState *m, int n_threads);
int64_t last_upload_host_rows(const MultiState *m);

int64_t add_comparison(MultiState *m, const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2);
int64_t add_category(MultiState *m, const int32_t *value_of_sample, int32_t n_values);
int clear_comparisons(MultiState *m);
int64_t num_comparisons(const MultiState *m);
int comparison_info(const MultiState *m, int64_t c, int32_t *kind, int32_t *category, int32_t *v1, int32_t *v2);
int set_count_engine(MultiState *m, int engine);
int set_fet_mode(MultiState *m, int mode);
int set_row_base(MultiState *m, int64_t row_base);
int set_comparison_window(MultiState *m, int64_t first, int64_t count);

int run(MultiState *m, int correction, double ratio_filter, int sort_key, int64_t bonferroni_n);
int64_t result_count(const MultiState *m, int64_t comparison);
int64_t result_total(const MultiState *m);
int fetch(MultiState *m, int64_t comparison, int64_t *row, int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
          double *p, double *ratio, int32_t *name_rank, bool to_device);
int last_timings(MultiState *m, double *out, int n);
int timings_accumulated(MultiState *m, double *out, int n, int reset);
int eval_tables(MultiState *m, const int32_t *goa, const int32_t *gob, const int32_t *gta, const int32_t *gtb, int64_t n,
                double *p_out, double *ratio_out);
int n_devices(const MultiState *m);

}  // namespace multi
}  // namespace feht

// internal accessors of a single-device context (capi.cu), not part of the public header
extern "C" int feht_internal_results(feht_ctx *ctx, feht::ResultSoA *out, const int64_t **seg_off_host, int64_t *n_total,
                                     int *device);
