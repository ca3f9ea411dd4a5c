// multi.cu -- multi-device context: row sharding over several GPUs of one box behind the C ABI (feht_create_multi).
//
// Replaces, for a host that drives several GPUs, the same app/Main.hs:26-29 as the single-device context: the
// reference's loop is one pure fold over the marker rows (Comparison.hs:86-111), and every (row, comparison) test is
// independent, so rows shard with no data-path exchange.  What this file adds to N single-device contexts:
//   * uploads are split contiguously over the devices (piece d of an n-row upload = rows [n*d/D, n*(d+1)/D)) and go up
//     in parallel, one host thread per device;
//   * feht_run starts every shard on its own device / stream (one host thread each: a filtered run waits for its
//     survivor counts), with Bonferroni's n = the job's row count (Comparison.hs:251) as a scalar;
//   * the ordered shards are copied to the first device (cudaMemcpyPeerAsync: NVLink when the devices are peers) and
//     merged there for ALL comparisons by one launch of K5b (sort.cu), which also scatters the eight result columns;
//   * feht_result_* then serve the merged rows, in exactly the order a single context holding all rows produces
//     (the total order (comparison, key bits, name rank) does not depend on the sharding).
// A device may be listed more than once (several shards on one GPU): that is how the path is tested on a one-GPU box.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <thread>

#include "multi.cuh"

namespace feht {

namespace {
struct Seg { int64_t local0, global0, n; };   // units (rows, or sites in SNP mode) of one upload held by one shard
}

struct MultiState {
    std::vector<feht_ctx *> kids;
    std::vector<int> dev;
    int root = 0;
    std::string err;
    int64_t S = 0;
    int snp = -1;
    int64_t total_units = 0;
    std::vector<std::vector<Seg>> segs;
    std::vector<int64_t> kid_units;
    int have_rank = -1;          // -1 no rows yet, 0 default (global row index), 1 given by the caller
    bool auto_ranked = false;    // default ranks were written into the shards explicitly (several uploads)
    bool auto_rank_valid = false;
    int64_t row_base = 0;
    int host_threads = -1;
    int64_t last_upload_host_rows = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // one copy stream (+ event) per shard on the root device: the peer copies from different GPUs run side by side
    std::vector<cudaStream_t> copy_st;
    std::vector<cudaEvent_t> copy_ev;
    // staged copies of the other devices' ordered results, and the merged rows, on the root device
    struct Cols { void *p[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; size_t cap = 0; };
    std::vector<Cols> staged;
    Cols merged;
    int64_t *d_seg = nullptr;    // [(kids + 1)][n_comps + 1] segment offsets; d_map: remap tables
    size_t d_seg_cap = 0;
    int64_t *d_map = nullptr;
    size_t d_map_cap = 0;
    int64_t *d_dest = nullptr;   // final positions of all rows (two-phase K5b)
    size_t d_dest_cap = 0;
    std::vector<int64_t> seg_off;
    int64_t n_total = 0;
    bool has_results = false;
    double tm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // count, test, order, total, launches, count launches, bytes, engine, merge ms
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // runs, count, test, order, total, launches, count launches, merge ms
    std::vector<int64_t> text_line_begin;
    std::vector<int32_t> text_name_len;
};

namespace {

constexpr size_t kColBytes[8] = {8, 4, 4, 4, 4, 4, 8, 8};   // row, rank, goa, gob, gta, gtb, p, ratio

int fail(MultiState *m, int code, const char *fmt, ...) {
    char buf[768];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    m->err = buf;
    return code;
}

#define MCU(m, expr)                                                                                  \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return fail((m), FEHT_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

int D(const MultiState *m) { return int(m->kids.size()); }
int mult(const MultiState *m) { return m->snp == 1 ? 4 : 1; }

// error of kid d -> error of the handle
int kid_fail(MultiState *m, int d, int rc) {
    return fail(m, rc, "device %d (shard %d): %s", m->dev[size_t(d)], d, feht_last_error(m->kids[size_t(d)]));
}

// run fn(d) for every shard, one host thread per shard beyond the first; returns the first failing shard or -1
template <class F>
int for_all_kids(MultiState *m, std::vector<int> &rc, F fn) {
    const int n = D(m);
    rc.assign(size_t(n), 0);
    std::vector<std::thread> pool;
    for (int d = 1; d < n; ++d) pool.emplace_back([&, d] { rc[size_t(d)] = fn(d); });
    rc[0] = fn(0);
    for (std::thread &t : pool) t.join();
    for (int d = 0; d < n; ++d)
        if (rc[size_t(d)] < 0) return d;
    return -1;
}

void drop_rows(MultiState *m) {
    for (feht_ctx *k : m->kids) feht_clear_rows(k);
    for (auto &s : m->segs) s.clear();
    std::fill(m->kid_units.begin(), m->kid_units.end(), 0);
    m->total_units = 0;
    m->snp = -1;
    m->have_rank = -1;
    m->auto_ranked = m->auto_rank_valid = false;
    m->has_results = false;
}

int check_mode(MultiState *m, int snp) {
    if (m->S <= 0) return fail(m, FEHT_E_INVALID, "feht_set_samples must be called first");
    if (m->snp == -1) m->snp = snp;
    if (m->snp != snp) return fail(m, FEHT_E_INVALID, "a context holds either binary rows or SNP sites, not both");
    return FEHT_OK;
}

int check_rank_mode(MultiState *m, const int32_t *name_rank) {
    const int given = name_rank ? 1 : 0;
    if (m->have_rank == -1) m->have_rank = given;
    if (m->have_rank != given) return fail(m, FEHT_E_INVALID, "name_rank must be given for all rows or for none");
    return FEHT_OK;
}

void note_upload(MultiState *m, const std::vector<int64_t> &cnt) {
    int64_t g = m->total_units;
    for (int d = 0; d < D(m); ++d) {
        const int64_t n = cnt[size_t(d)];
        if (n > 0) {
            std::vector<Seg> &v = m->segs[size_t(d)];
            if (!v.empty() && v.back().global0 + v.back().n == g && v.back().local0 + v.back().n == m->kid_units[size_t(d)])
                v.back().n += n;
            else
                v.push_back(Seg{m->kid_units[size_t(d)], g, n});
            m->kid_units[size_t(d)] += n;
        }
        g += n;
    }
    m->total_units = g;
    m->has_results = false;
    m->auto_rank_valid = false;
}

// default ranks (= global marker row index + row_base) of shard d's rows [local unit u0, u0 + n), as explicit values
std::vector<int32_t> default_ranks(const MultiState *m, int d, int64_t u0, int64_t n) {
    const int mu = mult(m);
    std::vector<int32_t> r(size_t(n) * size_t(mu));
    for (const Seg &s : m->segs[size_t(d)]) {
        const int64_t a = std::max(u0, s.local0), b = std::min(u0 + n, s.local0 + s.n);
        for (int64_t u = a; u < b; ++u)
            for (int j = 0; j < mu; ++j)
                r[size_t(u - u0) * size_t(mu) + size_t(j)] = int32_t(m->row_base + (s.global0 + (u - s.local0)) * mu + j);
    }
    return r;
}

int ensure_cols(MultiState *m, MultiState::Cols &c, size_t n) {
    if (n <= c.cap && c.p[0]) return FEHT_OK;
    for (int i = 0; i < 8; ++i) { if (c.p[i]) cudaFree(c.p[i]); c.p[i] = nullptr; }
    c.cap = 0;
    const size_t want = std::max<size_t>(n + n / 8, 1);
    for (int i = 0; i < 8; ++i) MCU(m, cudaMalloc(&c.p[i], want * kColBytes[i]));
    c.cap = want;
    return FEHT_OK;
}

ResultSoA as_soa(const MultiState::Cols &c) {
    ResultSoA r;
    r.row = static_cast<int64_t *>(c.p[0]); r.rank = static_cast<int32_t *>(c.p[1]);
    r.goa = static_cast<int32_t *>(c.p[2]); r.gob = static_cast<int32_t *>(c.p[3]);
    r.gta = static_cast<int32_t *>(c.p[4]); r.gtb = static_cast<int32_t *>(c.p[5]);
    r.p = static_cast<double *>(c.p[6]); r.ratio = static_cast<double *>(c.p[7]);
    return r;
}

const void *soa_col(const ResultSoA &r, int i) {
    switch (i) {
        case 0: return r.row; case 1: return r.rank; case 2: return r.goa; case 3: return r.gob;
        case 4: return r.gta; case 5: return r.gtb; case 6: return r.p; default: return r.ratio;
    }
}

}  // namespace

namespace multi {

MultiState *create(const int *device_ids, int n, std::string *err) {
    auto bail = [&](MultiState *m, const std::string &why) -> MultiState * {
        if (err) *err = why;
        if (m) destroy(m);
        return nullptr;
    };
    if (!device_ids || n < 1 || n > 32) return bail(nullptr, "feht_create_multi: between 1 and 32 devices");
    MultiState *m = new (std::nothrow) MultiState();
    if (!m) return bail(nullptr, "out of host memory");
    m->root = device_ids[0];
    for (int i = 0; i < n; ++i) {
        feht_ctx *k = nullptr;
        const int rc = feht_create(&k, device_ids[i]);
        if (rc != FEHT_OK) return bail(m, std::string("feht_create_multi: device ") + std::to_string(device_ids[i]) + ": " + feht_last_error(nullptr));
        m->kids.push_back(k);
        m->dev.push_back(device_ids[i]);
    }
    m->segs.resize(size_t(n));
    m->kid_units.assign(size_t(n), 0);
    m->staged.resize(size_t(n));
    cudaError_t e = cudaSetDevice(m->root);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&m->ev[i]);
    m->copy_st.assign(size_t(n), nullptr);
    m->copy_ev.assign(size_t(n), nullptr);
    for (int i = 1; i < n && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&m->copy_st[size_t(i)], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->copy_ev[size_t(i)], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return bail(m, std::string("feht_create_multi: ") + cudaGetErrorString(e));
    // direct peer access root <- others makes cudaMemcpyPeerAsync a single NVLink transfer (without it the runtime
    // stages through the host, still correct)
    for (int i = 1; i < n; ++i) {
        if (m->dev[size_t(i)] == m->root) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, m->root, m->dev[size_t(i)]) == cudaSuccess && can) {
            const cudaError_t pe = cudaDeviceEnablePeerAccess(m->dev[size_t(i)], 0);
            if (pe != cudaSuccess) cudaGetLastError();   // already enabled is fine
        }
    }
    // the host cores that transcode chars uploads are shared by the shards
    const int cores = int(std::thread::hardware_concurrency());
    for (feht_ctx *k : m->kids) feht_set_host_threads(k, std::max(1, cores / n - 1));
    return m;
}

void destroy(MultiState *m) {
    if (!m) return;
    for (feht_ctx *k : m->kids) feht_destroy(k);
    cudaSetDevice(m->root);
    if (m->st) cudaStreamSynchronize(m->st);
    for (auto &c : m->staged)
        for (void *p : c.p) if (p) cudaFree(p);
    for (void *p : m->merged.p) if (p) cudaFree(p);
    if (m->d_seg) cudaFree(m->d_seg);
    if (m->d_map) cudaFree(m->d_map);
    if (m->d_dest) cudaFree(m->d_dest);
    for (cudaEvent_t e : m->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : m->copy_ev) if (e) cudaEventDestroy(e);
    for (cudaStream_t q : m->copy_st) if (q) cudaStreamDestroy(q);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
}

const std::string &last_error(const MultiState *m) { return m->err; }
int n_devices(const MultiState *m) { return D(m); }

int set_samples(MultiState *m, int64_t n_samples) {
    for (int d = 0; d < D(m); ++d)
        if (int rc = feht_set_samples(m->kids[size_t(d)], n_samples)) return kid_fail(m, d, rc);
    m->S = n_samples;
    drop_rows(m);
    return FEHT_OK;
}

int reserve_rows(MultiState *m, int64_t n_rows) {
    if (m->S <= 0) return fail(m, FEHT_E_INVALID, "feht_set_samples must be called first");
    const int n = D(m);
    for (int d = 0; d < n; ++d)
        if (int rc = feht_reserve_rows(m->kids[size_t(d)], n_rows * (d + 1) / n - n_rows * d / n)) return kid_fail(m, d, rc);
    return FEHT_OK;
}

int clear_rows(MultiState *m) {
    drop_rows(m);
    return FEHT_OK;
}

int add_chars(MultiState *m, const uint8_t *chars, const int64_t *row_offsets, int64_t n, const int32_t *name_rank, int snp) {
    if (int rc = check_mode(m, snp)) return rc;
    if (n < 0 || (n > 0 && (!chars || !row_offsets))) return fail(m, FEHT_E_INVALID, "null rows");
    if (n == 0) return FEHT_OK;
    if (int rc = check_rank_mode(m, name_rank)) return rc;
    const int nd = D(m), mu = snp ? 4 : 1;
    std::vector<int64_t> cnt(static_cast<size_t>(nd));
    for (int d = 0; d < nd; ++d) cnt[size_t(d)] = n * (d + 1) / nd - n * d / nd;
    const bool synth = !name_rank && m->auto_ranked;   // earlier uploads already carry explicit default ranks
    int64_t g0 = m->total_units;
    std::vector<std::vector<int32_t>> auto_r(static_cast<size_t>(nd));
    if (synth)
        for (int d = 0; d < nd; ++d) {
            auto_r[size_t(d)].resize(size_t(cnt[size_t(d)]) * size_t(mu));
            for (size_t i = 0; i < auto_r[size_t(d)].size(); ++i) auto_r[size_t(d)][i] = int32_t(m->row_base + g0 * mu + int64_t(i));
            g0 += cnt[size_t(d)];
        }
    std::vector<int> rc;
    const int bad = for_all_kids(m, rc, [&](int d) -> int {
        const int64_t r0 = n * d / nd, c = cnt[size_t(d)];
        if (c == 0) return FEHT_OK;
        const int32_t *rk = name_rank ? name_rank + r0 * mu : (synth ? auto_r[size_t(d)].data() : nullptr);
        return snp ? feht_add_snp_chars(m->kids[size_t(d)], chars, row_offsets + r0, c, rk)
                   : feht_add_rows_chars(m->kids[size_t(d)], chars, row_offsets + r0, c, rk);
    });
    if (bad >= 0) {
        const int code = kid_fail(m, bad, rc[size_t(bad)]);
        // the shards that succeeded hold rows the others do not: a partial upload is not kept
        m->err += " -- the upload was dropped; the context now holds no rows";
        drop_rows(m);
        return code;
    }
    m->last_upload_host_rows = 0;
    for (feht_ctx *k : m->kids) m->last_upload_host_rows += std::max<int64_t>(0, feht_last_upload_host_rows(k));
    note_upload(m, cnt);
    return FEHT_OK;
}

int add_packed(MultiState *m, const uint64_t *p0, const uint64_t *p1, const uint64_t *p2, int64_t n, int64_t stride,
               const int32_t *name_rank, int snp) {
    if (int rc = check_mode(m, snp)) return rc;
    if (n < 0 || stride < (m->S + 63) / 64) return fail(m, FEHT_E_INVALID, "row_stride_words < ceil(S/64)");
    if (n == 0) return FEHT_OK;
    if (!p0 || !p1 || (snp && !p2)) return fail(m, FEHT_E_INVALID, "null plane");
    if (int rc = check_rank_mode(m, name_rank)) return rc;
    const int nd = D(m), mu = snp ? 4 : 1;
    std::vector<int64_t> cnt(static_cast<size_t>(nd));
    for (int d = 0; d < nd; ++d) cnt[size_t(d)] = n * (d + 1) / nd - n * d / nd;
    const bool synth = !name_rank && m->auto_ranked;
    std::vector<std::vector<int32_t>> auto_r(static_cast<size_t>(nd));
    if (synth) {
        int64_t g0 = m->total_units;
        for (int d = 0; d < nd; ++d) {
            auto_r[size_t(d)].resize(size_t(cnt[size_t(d)]) * size_t(mu));
            for (size_t i = 0; i < auto_r[size_t(d)].size(); ++i) auto_r[size_t(d)][i] = int32_t(m->row_base + g0 * mu + int64_t(i));
            g0 += cnt[size_t(d)];
        }
    }
    std::vector<int> rc;
    const int bad = for_all_kids(m, rc, [&](int d) -> int {
        const int64_t r0 = n * d / nd, c = cnt[size_t(d)];
        if (c == 0) return FEHT_OK;
        const int32_t *rk = name_rank ? name_rank + r0 * mu : (synth ? auto_r[size_t(d)].data() : nullptr);
        const size_t o = size_t(r0) * size_t(stride);
        return snp ? feht_add_snp_packed(m->kids[size_t(d)], p0 + o, p1 + o, p2 + o, c, stride, rk)
                   : feht_add_rows_packed(m->kids[size_t(d)], p0 + o, p1 + o, c, stride, rk);
    });
    if (bad >= 0) {
        const int code = kid_fail(m, bad, rc[size_t(bad)]);
        m->err += " -- the upload was dropped; the context now holds no rows";
        drop_rows(m);
        return code;
    }
    note_upload(m, cnt);
    return FEHT_OK;
}

int add_text(MultiState *m, const char *text, int64_t len, char delimiter, int snp_mode, int64_t *n_lines_out) {
    const int snp = snp_mode ? 1 : 0;
    if (int rc = check_mode(m, snp)) return rc;
    if (len < 0 || (len > 0 && !text)) return fail(m, FEHT_E_INVALID, "null text");
    if (m->have_rank == 1)
        return fail(m, FEHT_E_INVALID, "text rows get their name ranks afterwards (feht_set_name_rank); the context "
                    "already holds rows with ranks");
    m->text_line_begin.clear();
    m->text_name_len.clear();
    if (n_lines_out) *n_lines_out = 0;
    if (len == 0) return FEHT_OK;
    // pieces of whole lines, cut at the first newline at or after every len * d / D mark
    const int nd = D(m);
    std::vector<int64_t> cut(static_cast<size_t>(nd) + 1, len);
    cut[0] = 0;
    for (int d = 1; d < nd; ++d) {
        int64_t at = std::max(cut[size_t(d) - 1], len * d / nd);
        if (at < len) {
            const void *nl = memchr(text + at, '\n', size_t(len - at));
            at = nl ? static_cast<const char *>(nl) - text + 1 : len;
        }
        cut[size_t(d)] = at;
    }
    std::vector<int64_t> lines(static_cast<size_t>(nd), 0);
    std::vector<int> rc;
    const int bad = for_all_kids(m, rc, [&](int d) -> int {
        const int64_t b0 = cut[size_t(d)], b1 = cut[size_t(d) + 1];
        if (b1 <= b0) return FEHT_OK;
        return feht_add_rows_text(m->kids[size_t(d)], text + b0, b1 - b0, delimiter, snp_mode, &lines[size_t(d)]);
    });
    if (bad >= 0) {
        const int code = kid_fail(m, bad, rc[size_t(bad)]);
        m->err += " -- the upload was dropped; the context now holds no rows";
        drop_rows(m);
        return code;
    }
    int64_t total = 0;
    for (int d = 0; d < nd; ++d) {
        const int64_t nl = lines[size_t(d)];
        if (nl == 0) continue;
        std::vector<int64_t> b(static_cast<size_t>(nl));
        std::vector<int32_t> nm(static_cast<size_t>(nl));
        if (int r = feht_text_lines(m->kids[size_t(d)], b.data(), nm.data(), nl)) return kid_fail(m, d, r);
        for (int64_t &x : b) x += cut[size_t(d)];
        m->text_line_begin.insert(m->text_line_begin.end(), b.begin(), b.end());
        m->text_name_len.insert(m->text_name_len.end(), nm.begin(), nm.end());
        total += nl;
    }
    if (total > 0 && m->have_rank == -1) m->have_rank = 0;
    if (total > 0 && m->auto_ranked) m->auto_rank_valid = false;
    note_upload(m, lines);
    if (n_lines_out) *n_lines_out = total;
    return FEHT_OK;
}

int text_lines(MultiState *m, int64_t *line_begin, int32_t *name_len, int64_t cap) {
    const int64_t n = int64_t(m->text_line_begin.size());
    if (cap < n) return fail(m, FEHT_E_INVALID, "feht_text_lines: %lld lines, room for %lld", (long long)n, (long long)cap);
    if (n > 0 && line_begin) std::memcpy(line_begin, m->text_line_begin.data(), size_t(n) * 8);
    if (n > 0 && name_len) std::memcpy(name_len, m->text_name_len.data(), size_t(n) * 4);
    return FEHT_OK;
}

int set_name_rank(MultiState *m, const int32_t *name_rank, int64_t n_marker_rows) {
    if (!name_rank) return fail(m, FEHT_E_INVALID, "null name_rank");
    const int mu = mult(m);
    if (n_marker_rows != m->total_units * mu)
        return fail(m, FEHT_E_INVALID, "feht_set_name_rank: %lld ranks for %lld marker rows", (long long)n_marker_rows,
                    (long long)(m->total_units * mu));
    if (n_marker_rows == 0) return FEHT_OK;
    for (int d = 0; d < D(m); ++d) {
        const int64_t nu = m->kid_units[size_t(d)];
        if (nu == 0) continue;
        std::vector<int32_t> local(size_t(nu) * size_t(mu));
        for (const Seg &s : m->segs[size_t(d)])
            std::memcpy(local.data() + size_t(s.local0) * size_t(mu), name_rank + size_t(s.global0) * size_t(mu),
                        size_t(s.n) * size_t(mu) * 4);
        if (int rc = feht_set_name_rank(m->kids[size_t(d)], local.data(), nu * mu)) return kid_fail(m, d, rc);
    }
    m->have_rank = 1;
    m->auto_ranked = m->auto_rank_valid = false;
    m->has_results = false;
    return FEHT_OK;
}

int generate_synthetic(MultiState *m, const feht_synth_spec *spec, int snp_mode, int64_t first_row, int64_t n_rows) {
    if (!spec) return fail(m, FEHT_E_INVALID, "null spec");
    const int snp = snp_mode ? 1 : 0;
    if (int rc = check_mode(m, snp)) return rc;
    if (n_rows < 0) return fail(m, FEHT_E_INVALID, "n_rows < 0");
    if (m->have_rank == 1) return fail(m, FEHT_E_INVALID, "synthetic rows carry no name_rank");
    if (n_rows == 0) return FEHT_OK;
    m->have_rank = 0;
    const int nd = D(m);
    std::vector<int64_t> cnt(static_cast<size_t>(nd));
    for (int d = 0; d < nd; ++d) cnt[size_t(d)] = n_rows * (d + 1) / nd - n_rows * d / nd;
    std::vector<int> rc;
    const int bad = for_all_kids(m, rc, [&](int d) -> int {
        if (cnt[size_t(d)] == 0) return FEHT_OK;
        return feht_generate_synthetic(m->kids[size_t(d)], spec, snp_mode, first_row + n_rows * d / nd, cnt[size_t(d)]);
    });
    if (bad >= 0) {
        const int code = kid_fail(m, bad, rc[size_t(bad)]);
        drop_rows(m);
        return code;
    }
    note_upload(m, cnt);
    return FEHT_OK;
}

int read_plane(MultiState *m, int plane, int64_t row0, int64_t n, uint64_t *out) {
    if (!out) return fail(m, FEHT_E_INVALID, "null out");
    if (row0 < 0 || n < 0 || row0 + n > m->total_units) return fail(m, FEHT_E_INVALID, "row range");
    const int64_t W = row_stride_words(m);
    for (int d = 0; d < D(m); ++d)
        for (const Seg &s : m->segs[size_t(d)]) {
            const int64_t a = std::max(row0, s.global0), b = std::min(row0 + n, s.global0 + s.n);
            if (a >= b) continue;
            if (int rc = feht_read_plane(m->kids[size_t(d)], plane, s.local0 + (a - s.global0), b - a, out + size_t(a - row0) * size_t(W)))
                return kid_fail(m, d, rc);
        }
    return FEHT_OK;
}

int64_t row_stride_words(const MultiState *m) { return feht_row_stride_words(m->kids[0]); }
int64_t num_rows(const MultiState *m) { return m->total_units * mult(m); }

int set_host_threads(MultiState *m, int n_threads) {
    m->host_threads = n_threads;
    const int nd = D(m);
    const int cores = int(std::thread::hardware_concurrency());
    for (feht_ctx *k : m->kids)
        feht_set_host_threads(k, n_threads < 0 ? std::max(1, cores / nd - 1) : (n_threads == 0 ? 0 : std::max(1, n_threads / nd)));
    return FEHT_OK;
}

int64_t last_upload_host_rows(const MultiState *m) { return m->last_upload_host_rows; }

int64_t add_comparison(MultiState *m, const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2) {
    int64_t idx = -1;
    for (int d = 0; d < D(m); ++d) {
        const int64_t r = feht_add_comparison(m->kids[size_t(d)], g1, n1, g2, n2);
        if (r < 0) return kid_fail(m, d, int(r));
        idx = r;
    }
    m->has_results = false;
    return idx;
}

int64_t add_category(MultiState *m, const int32_t *value_of_sample, int32_t n_values) {
    int64_t idx = -1;
    for (int d = 0; d < D(m); ++d) {
        const int64_t r = feht_add_category(m->kids[size_t(d)], value_of_sample, n_values);
        if (r < 0) return kid_fail(m, d, int(r));
        idx = r;
    }
    m->has_results = false;
    return idx;
}

int clear_comparisons(MultiState *m) {
    for (feht_ctx *k : m->kids) feht_clear_comparisons(k);
    m->has_results = false;
    return FEHT_OK;
}

int64_t num_comparisons(const MultiState *m) { return feht_num_comparisons(m->kids[0]); }

int comparison_info(const MultiState *m, int64_t c, int32_t *kind, int32_t *category, int32_t *v1, int32_t *v2) {
    return feht_comparison_info(m->kids[0], c, kind, category, v1, v2);
}

int set_count_engine(MultiState *m, int engine) {
    for (int d = 0; d < D(m); ++d)
        if (int rc = feht_set_count_engine(m->kids[size_t(d)], engine)) return kid_fail(m, d, rc);
    return FEHT_OK;
}

int set_fet_mode(MultiState *m, int mode) {
    for (int d = 0; d < D(m); ++d)
        if (int rc = feht_set_fet_mode(m->kids[size_t(d)], mode)) return kid_fail(m, d, rc);
    m->has_results = false;
    return FEHT_OK;
}

int set_row_base(MultiState *m, int64_t row_base) {
    m->row_base = row_base;
    m->auto_rank_valid = false;
    return FEHT_OK;
}

int set_comparison_window(MultiState *m, int64_t first, int64_t count) {
    for (int d = 0; d < D(m); ++d)
        if (int rc = feht_set_comparison_window(m->kids[size_t(d)], first, count)) return kid_fail(m, d, rc);
    m->has_results = false;
    return FEHT_OK;
}

int run(MultiState *m, int correction, double ratio_filter, int sort_key, int64_t bonferroni_n) {
    m->has_results = false;
    const int nd = D(m), mu = mult(m);
    const int64_t n_comps = num_comparisons(m);
    if (m->S <= 0) return fail(m, FEHT_E_INVALID, "feht_set_samples must be called first");
    const int64_t R = m->total_units * mu;
    if (bonferroni_n <= 0) bonferroni_n = R;
    bool remap = false;
    for (int d = 0; d < nd; ++d) remap = remap || m->segs[size_t(d)].size() > 1;
    if (m->have_rank != 1 && m->row_base + R > int64_t(0x7fffffff))
        return fail(m, FEHT_E_INVALID, "row_base + rows exceeds 2^31-1: pass name_rank (the default rank is the global row index)");
    // shards filled by several uploads hold non-contiguous global rows: their default ranks (the global row index)
    // are written explicitly, and the merge maps shard rows back to global rows
    if (m->have_rank != 1 && remap && !m->auto_rank_valid) {
        for (int d = 0; d < nd; ++d) {
            const int64_t nu = m->kid_units[size_t(d)];
            if (nu == 0) continue;
            const std::vector<int32_t> r = default_ranks(m, d, 0, nu);
            if (int rc = feht_set_name_rank(m->kids[size_t(d)], r.data(), nu * mu)) return kid_fail(m, d, rc);
        }
        m->auto_ranked = m->auto_rank_valid = true;
    }
    for (int d = 0; d < nd; ++d) {
        const std::vector<Seg> &v = m->segs[size_t(d)];
        const int64_t base = (remap || v.empty()) ? 0 : m->row_base + v[0].global0 * mu;
        feht_set_row_base(m->kids[size_t(d)], base);
    }
    std::vector<int> rc;
    const int bad = for_all_kids(m, rc, [&](int d) -> int {
        return feht_run(m->kids[size_t(d)], correction, ratio_filter, sort_key, bonferroni_n);
    });
    if (bad >= 0) return kid_fail(m, bad, rc[size_t(bad)]);

    // ---- stage times: the shards run side by side, the job waits for the slowest
    double t[8];
    for (double &x : m->tm) x = 0;
    for (int d = 0; d < nd; ++d) {
        if (int r = feht_last_timings(m->kids[size_t(d)], t, 8)) return kid_fail(m, d, r);
        for (int i = 0; i < 4; ++i) m->tm[i] = std::max(m->tm[i], t[i]);
        m->tm[4] += t[4]; m->tm[5] += t[5]; m->tm[6] += t[6];
        m->tm[7] = t[7];
    }

    // ---- merge on the root device
    MCU(m, cudaSetDevice(m->root));
    std::vector<MergeShardView> views(static_cast<size_t>(nd));
    std::vector<const int64_t *> seg_host(static_cast<size_t>(nd), nullptr);
    m->seg_off.assign(size_t(n_comps) + 1, 0);
    int64_t total = 0;
    for (int d = 0; d < nd; ++d) {
        int device = 0;
        int64_t n = 0;
        ResultSoA res{};
        if (int r = feht_internal_results(m->kids[size_t(d)], &res, &seg_host[size_t(d)], &n, &device)) return kid_fail(m, d, r);
        views[size_t(d)].res = res;
        views[size_t(d)].n = n;
        views[size_t(d)].mult = mu;
        total += n;
        for (int64_t c = 0; c < n_comps; ++c)
            m->seg_off[size_t(c) + 1] += seg_host[size_t(d)][c + 1] - seg_host[size_t(d)][c];
    }
    for (int64_t c = 0; c < n_comps; ++c) m->seg_off[size_t(c) + 1] += m->seg_off[size_t(c)];
    m->n_total = total;
    MCU(m, cudaEventRecord(m->ev[0], m->st));
    if (total > 0 && n_comps > 0) {
        // segment offsets of every shard and of the output, one upload
        const size_t per = size_t(n_comps) + 1;
        std::vector<int64_t> hseg(per * size_t(nd + 1));
        for (int d = 0; d < nd; ++d) std::memcpy(hseg.data() + per * size_t(d), seg_host[size_t(d)], per * 8);
        std::memcpy(hseg.data() + per * size_t(nd), m->seg_off.data(), per * 8);
        if (hseg.size() > m->d_seg_cap) {
            if (m->d_seg) cudaFree(m->d_seg);
            m->d_seg = nullptr;
            MCU(m, cudaMalloc(&m->d_seg, hseg.size() * 8));
            m->d_seg_cap = hseg.size();
        }
        MCU(m, cudaMemcpyAsync(m->d_seg, hseg.data(), hseg.size() * 8, cudaMemcpyHostToDevice, m->st));
        // remap tables
        std::vector<int64_t> hmap;
        std::vector<size_t> map_at(static_cast<size_t>(nd), 0);
        if (remap) {
            for (int d = 0; d < nd; ++d) {
                const std::vector<Seg> &v = m->segs[size_t(d)];
                map_at[size_t(d)] = hmap.size();
                for (const Seg &s : v) hmap.push_back(s.local0);
                hmap.push_back(m->kid_units[size_t(d)]);
                for (const Seg &s : v) hmap.push_back(s.global0);
            }
            if (hmap.size() > m->d_map_cap) {
                if (m->d_map) cudaFree(m->d_map);
                m->d_map = nullptr;
                MCU(m, cudaMalloc(&m->d_map, hmap.size() * 8));
                m->d_map_cap = hmap.size();
            }
            MCU(m, cudaMemcpyAsync(m->d_map, hmap.data(), hmap.size() * 8, cudaMemcpyHostToDevice, m->st));
        }
        for (int d = 0; d < nd; ++d) {
            MergeShardView &v = views[size_t(d)];
            v.d_seg_off = m->d_seg + per * size_t(d);
            if (remap && !m->segs[size_t(d)].empty()) {
                v.n_map = int(m->segs[size_t(d)].size());
                v.d_map_local = m->d_map + map_at[size_t(d)];
                v.d_map_global = v.d_map_local + v.n_map + 1;
            }
            if (m->dev[size_t(d)] != m->root && v.n > 0) {
                // the shard's ordered columns travel GPU to GPU (the kid's run has finished: feht_run is blocking)
                if (int r = ensure_cols(m, m->staged[size_t(d)], size_t(v.n))) return r;
                cudaStream_t cs = m->copy_st[size_t(d)];
                MCU(m, cudaStreamWaitEvent(cs, m->ev[0], 0));      // after the merge's start mark (and earlier merges)
                for (int i = 0; i < 8; ++i)
                    MCU(m, cudaMemcpyPeerAsync(m->staged[size_t(d)].p[i], m->root, soa_col(v.res, i), m->dev[size_t(d)],
                                               size_t(v.n) * kColBytes[i], cs));
                MCU(m, cudaEventRecord(m->copy_ev[size_t(d)], cs));
                MCU(m, cudaStreamWaitEvent(m->st, m->copy_ev[size_t(d)], 0));
                v.res = as_soa(m->staged[size_t(d)]);
            }
        }
        if (int r = ensure_cols(m, m->merged, size_t(total))) return r;
        const size_t need = size_t(total) + merge_bounds_words(total, nd);   // positions, then block bounds
        if (need > m->d_dest_cap) {
            if (m->d_dest) cudaFree(m->d_dest);
            m->d_dest = nullptr;
            MCU(m, cudaMalloc(&m->d_dest, (need + need / 8) * 8));
            m->d_dest_cap = need + need / 8;
        }
        MCU(m, launch_merge_all(nd, int(n_comps), sort_key, views.data(), m->d_seg + per * size_t(nd), as_soa(m->merged),
                                m->row_base, m->d_dest, m->d_dest + total, m->st));
        m->tm[4] += 1;
    }
    MCU(m, cudaEventRecord(m->ev[1], m->st));
    MCU(m, cudaStreamSynchronize(m->st));   // (the host vectors above die here)
    float ms = 0;
    MCU(m, cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
    m->tm[8] = ms;
    m->tm[3] += ms;
    m->acc[0] += 1; m->acc[1] += m->tm[0]; m->acc[2] += m->tm[1]; m->acc[3] += m->tm[2]; m->acc[4] += m->tm[3];
    m->acc[5] += m->tm[4]; m->acc[6] += m->tm[5]; m->acc[7] += ms;
    m->has_results = true;
    return FEHT_OK;
}

int64_t result_count(const MultiState *m, int64_t comparison) {
    if (!m->has_results || comparison < 0 || comparison + 1 >= int64_t(m->seg_off.size())) return FEHT_E_INVALID;
    return m->seg_off[size_t(comparison) + 1] - m->seg_off[size_t(comparison)];
}

int64_t result_total(const MultiState *m) { return m->has_results ? m->n_total : FEHT_E_INVALID; }

int fetch(MultiState *m, int64_t comparison, int64_t *row, int32_t *goa, int32_t *gob, int32_t *gta, int32_t *gtb,
          double *p, double *ratio, int32_t *name_rank, bool to_device) {
    if (!m->has_results) return fail(m, FEHT_E_INVALID, "no results: call feht_run first");
    if (comparison < -1 || comparison + 1 >= int64_t(m->seg_off.size())) return fail(m, FEHT_E_INVALID, "comparison index");
    MCU(m, cudaSetDevice(m->root));
    const int64_t o = comparison < 0 ? 0 : m->seg_off[size_t(comparison)];
    const size_t nn = comparison < 0 ? size_t(m->n_total) : size_t(m->seg_off[size_t(comparison) + 1] - o);
    if (nn == 0) return FEHT_OK;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    void *dst[8] = {row, name_rank, goa, gob, gta, gtb, p, ratio};
    for (int i = 0; i < 8; ++i)
        if (dst[i])
            MCU(m, cudaMemcpyAsync(dst[i], static_cast<const char *>(m->merged.p[i]) + size_t(o) * kColBytes[i],
                                   nn * kColBytes[i], kind, m->st));
    MCU(m, cudaStreamSynchronize(m->st));
    return FEHT_OK;
}

int last_timings(MultiState *m, double *out, int n) {
    for (int i = 0; i < n && i < 9; ++i) out[i] = m->tm[i];
    return FEHT_OK;
}

int timings_accumulated(MultiState *m, double *out, int n, int reset) {
    for (int i = 0; i < n && i < 8; ++i) out[i] = m->acc[i];
    if (reset)
        for (double &a : m->acc) a = 0;
    return FEHT_OK;
}

int eval_tables(MultiState *m, const int32_t *goa, const int32_t *gob, const int32_t *gta, const int32_t *gtb, int64_t n,
                double *p_out, double *ratio_out) {
    if (int rc = feht_eval_tables(m->kids[0], goa, gob, gta, gtb, n, p_out, ratio_out)) return kid_fail(m, 0, rc);
    return FEHT_OK;
}

}  // namespace multi
}  // namespace feht
