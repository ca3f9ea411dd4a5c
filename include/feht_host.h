/*
 * feht_host.h -- C ABI of the C++ host mirror (feht_b200/host/feht_host.cpp, inside libfeht_cuda.so).
 *
 * The reference host is Haskell (app/Main.hs, src/UserInput.hs, src/Table.hs, src/Comparison.hs under
 * /root/reference).  Where no GHC is available this layer provides the same program: parse the data and
 * metadata files, plan the comparisons, run the CUDA hot path (include/feht_cuda.h) and format the blocks
 * exactly as `feht` prints them.  Rows SURVEY.md 8(f) N1 (formatter), N2 (loader), N3 (planner).
 */
#ifndef FEHT_HOST_H
#define FEHT_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * app/Main.hs:13-31.  metadata / data are the raw file contents (-i / -d), `one` / `two` the --one / --two
 * strings ("all all" by default, UserInput.hs:52-61), delimiter the -l character, snp_mode = -m snp,
 * correction FEHT_CORRECTION_* (-c), ratio_filter (-f), device the CUDA device, engine FEHT_ENGINE_*.
 * On success *out is a malloc'ed buffer (free with feht_host_free) holding exactly what feht writes to
 * stdout: one "[#- ... -#]" block per comparison with surviving rows, each followed by an empty line
 * (BS.putStrLn), then "Done".  Comparison blocks come in generation order and rows with equal |ratio| in
 * marker-name order (the reference's order there is HashMap iteration order).
 * On failure returns a negative FEHT_E_* code and writes the reference's `error` message into err.
 */
int feht_host_run(const char *metadata, int64_t meta_len, const char *data, int64_t data_len,
                  const char *one, const char *two, char delimiter, int snp_mode, int correction,
                  double ratio_filter, int device, int engine, char **out, int64_t *out_len, char *err,
                  int errlen);

/*
 * Parsing + planning only (no GPU): Table.hs:267-279, :80-115, :229-247; Comparison.hs:271-364.
 * *out: "columns C rows R\n" then per comparison "comparison\n<description>g1 <cols>\ng2 <cols>\n"
 * (column numbers ascending).  For tests of the host logic.
 */
int feht_host_plan(const char *metadata, int64_t meta_len, const char *data, int64_t data_len,
                   const char *one, const char *two, char delimiter, int snp_mode, char **out,
                   int64_t *out_len, char *err, int errlen);

/* GHC `show :: Double -> String` (Comparison.hs:180-181); returns the length or -1 if buf is too small. */
int feht_host_show_double(double x, char *buf, int buflen);

void feht_host_free(char *p);

#ifdef __cplusplus
}
#endif
#endif /* FEHT_HOST_H */
