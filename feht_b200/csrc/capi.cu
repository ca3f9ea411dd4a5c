// capi.cu -- the C ABI of libfeht_cuda.so (include/feht_cuda.h): context, uploads, run, fetch.
//
// The run replaces app/Main.hs:26-29 of the reference: generateResultMap (Comparison.hs:86-111),
// applyMultipleTestingCorrection (:238-264) and the filter + ordering half of
// formatComparisonResultsAsTable (:139-162).  Everything heavy happens in the kernels
// (pack.cu K0, count_popc.cu K1, count_gemm.cu K2, fet.cu K3, sort.cu K4); this file owns device
// memory, turns column lists / categories into bit masks and slot descriptors, sequences the
// launches on one stream and times the stages with CUDA events.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <limits>
#include <mutex>
#include <queue>
#include <thread>

#include "common.cuh"
#include "multi.cuh"

using namespace feht;

namespace {

thread_local std::string g_create_error;

struct Comp {
    int kind = 0;      // 0 explicit, 1 pair, 2 one-vs-rest
    int cat = -1;
    int v1 = -1, v2 = -1;
    std::vector<int32_t> g1, g2;  // explicit only
    int64_t max_col = -1;
};

struct Cat {
    std::vector<int32_t> value_of_sample;
    int n_values = 0;
    int64_t max_col = -1;
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
};

}  // namespace

struct feht_ctx {
    // multi-device handle (feht_create_multi): every entry point forwards to multi.cu; the rest of this struct is
    // then just a context on the first device (it serves feht_merge_sorted_device / feht_scatter_rows_device)
    MultiState *multi = nullptr;
    int device = 0;
    int num_sms = 148;
    cudaStream_t own_stream = nullptr, st = nullptr, copy_stream = nullptr;
    // stage events of a run, two sets used alternately: feht_run_async re-records a set only after the run
    // before last has been timed, so consecutive runs can be enqueued without a host synchronisation
    cudaEvent_t evr[2][5] = {{nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr, nullptr}};
    struct PendingRun { bool active = false, fused = false; double launches = 0, count_launches = 0, count_bytes = 0; } pend[2];
    int ev_slot = 0;
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // runs, count/test/order/total ms, launches, count launches
    cudaEvent_t ev_copied[4] = {nullptr, nullptr, nullptr, nullptr}, ev_used[4] = {nullptr, nullptr, nullptr, nullptr};
    // host transcoder (host_pack.cpp): worker threads pack pieces of a chars upload while the DMA engine moves
    // the others raw; each worker owns two pinned output buffers (events order their reuse)
    int host_threads = -1;                 // -1: FEHT_HOST_PACK_THREADS or all cores; 0: off
    cudaStream_t pack_stream = nullptr;    // H2D of the host-packed pieces
    uint64_t *pack_pinned = nullptr;       // [workers][2][kHostPackBufWords]
    int pack_workers_alloc = 0;
    std::vector<cudaEvent_t> pack_ev;      // [workers][2]
    int64_t last_upload_host_rows = 0;     // rows of the last chars upload the host cores packed
    std::string err;

    int64_t S = 0, W = 0;
    int snp = -1;  // -1 no rows yet, 0 binary, 1 snp
    int64_t n_units = 0, cap_units = 0;  // rows (binary) or sites (snp)
    uint64_t *plane[3] = {nullptr, nullptr, nullptr};  // value|hi, valid, lo
    DevBuf<int32_t> rank;
    int have_rank = -1;  // -1 unknown, 0 none, 1 given
    int64_t min_row_len = std::numeric_limits<int64_t>::max();
    int64_t row_base = 0;

    std::vector<Comp> comps;
    std::vector<Cat> cats;
    int engine = FEHT_ENGINE_AUTO;
    int fet_mode = FEHT_FET_TWO_TAIL;
    // feht_set_comparison_window: a run emits the comparisons [win_first, win_first + win_count) only (win_count < 0:
    // all).  While a window is set the counts of the last run are kept (counts_ready) and the next window starts at K3.
    int64_t win_first = 0, win_count = -1;
    bool counts_ready = false;
    int counts_engine = -1;

    double *d_choose = nullptr;
    DevBuf<unsigned char> stage[2];

    // scratch for a run
    DevBuf<uint4> masks;
    DevBuf<uint8_t> gemm_b;          // baked one-hot B blocks of all GEMM passes
    DevBuf<unsigned char> gemm_passes;
    int gemm_n_passes = 0;
    bool gemm_ready = false;         // gemm_b / gemm_passes match the current categories
    bool gemm_fp4 = false;           // ... and were baked for the 4-bit (kind::mxf4) flavour
    bool gemm_f8 = false;            // ... or for the e5m2 one (both planes in one GEMM)
    bool plan_ready = false;         // masks / groups / fracs / gcomps / sdefs on the device match the comparisons
    bool hyp_heavy = false;          // ... and: some comparison is all exact tests (K3b flavour)
    int plan_groups = 0, plan_max_frac = 0, plan_max_comp = 0, plan_max_slot = 0;
    int last_engine = 0;             // engine the last run counted categories with
    DevBuf<int4> counts;
    DevBuf<ScreenGroup> groups;
    DevBuf<FracDef> fracs;
    DevBuf<GroupComp> gcomps;
    DevBuf<SlotDef> sdefs;
    DevBuf<unsigned long long> seg_count;  // [2][n_comps]: counts, cursors
    DevBuf<int32_t> live_rows;             // [n_groups][rows]: rows the prescreen lets through (filtered runs)
    DevBuf<unsigned int> live_count;       // [n_groups]
    int64_t survivor_hint = 0;             // survivors of the last filtered run (sizes the next run's arrays)
    DevBuf<int32_t> r_comp, o_rank, o_goa, o_gob, o_gta, o_gtb;
    DevBuf<SurvivorRec> r_rec;
    DevBuf<int64_t> o_row;
    DevBuf<double> r_p, o_p, o_ratio;
    SortWorkspace sw;
    DevBuf<unsigned long long> sw_key_a, sw_key_b;
    DevBuf<uint32_t> sw_val_a, sw_val_b, sw_control, sw_oa, sw_coop;
    int rank_is_perm = -1;  // -1 unknown, 0 no, 1 name_rank is a permutation of [0, rows)

    bool has_results = false;
    int64_t n_surv = 0;
    std::vector<int64_t> seg_off;
    Timings tm;

    // text upload (text_load.cu): scratch on the device, and the lines of the last call for the host
    DevBuf<unsigned long long> dd_keys;    // N4 dedup of exact tests (fet.cu): table, owners, waiting list, its length
    DevBuf<uint32_t> dd_owner, dd_count;
    DevBuf<uint2> dd_list;
    DevBuf<int64_t> merge_dest;            // feht_merge_shards_device: final positions (two-phase K5b)
    DevBuf<unsigned int> snp_digit_flag;   // feht_add_snp_chars: K0 met a literal '0' / '1'
    DevBuf<uint32_t> txt_counts;
    DevBuf<int32_t> txt_starts, txt_name_len;
    DevBuf<unsigned long long> txt_flags;
    std::vector<int64_t> text_line_begin;
    std::vector<int32_t> text_name_len;
};

namespace {

int fail(feht_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

#define CU(c, expr)                                                                                \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail((c), FEHT_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                                 \
    } while (0)

// multi-device handles: forward the call, keep the message where feht_last_error looks for it
#define MULTI(c, call)                                                     \
    do {                                                                   \
        if ((c) && (c)->multi) {                                           \
            auto _r = (call);                                              \
            if (_r < 0) (c)->err = multi::last_error((c)->multi);          \
            return _r;                                                     \
        }                                                                  \
    } while (0)

template <class T>
cudaError_t ensure(DevBuf<T> &b, size_t n) {
    if (n <= b.cap && b.p) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc(&b.p, n * sizeof(T));
    if (e == cudaSuccess) b.cap = n;
    return e;
}

template <class T>
void release(DevBuf<T> &b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

int use_device(feht_ctx *c) {
    CU(c, cudaSetDevice(c->device));
    return FEHT_OK;
}

// grow the planes (and rank) to hold `need` units, keeping the contents
int grow_rows(feht_ctx *c, int64_t need, int n_planes) {
    if (need <= c->cap_units) {
        bool all = true;
        for (int i = 0; i < n_planes; ++i) all = all && c->plane[i];
        if (all) return FEHT_OK;
    }
    int64_t cap = std::max<int64_t>(need, c->cap_units + c->cap_units / 2);
    if (cap < 1) cap = 1;
    const size_t words = size_t(cap) * size_t(c->W);
    for (int i = 0; i < n_planes; ++i) {
        uint64_t *np = nullptr;
        CU(c, cudaMalloc(&np, words * 8));
        if (c->plane[i] && c->n_units > 0)
            CU(c, cudaMemcpyAsync(np, c->plane[i], size_t(c->n_units) * c->W * 8,
                                  cudaMemcpyDeviceToDevice, c->st));
        if (c->plane[i]) {
            CU(c, cudaStreamSynchronize(c->st));
            cudaFree(c->plane[i]);
        }
        c->plane[i] = np;
    }
    c->cap_units = cap;
    return FEHT_OK;
}

int append_rank(feht_ctx *c, const int32_t *name_rank, int64_t n_marker_rows, int64_t first_marker) {
    const int given = name_rank ? 1 : 0;
    if (c->have_rank == -1) c->have_rank = given;
    if (c->have_rank != given)
        return fail(c, FEHT_E_INVALID, "name_rank must be given for all rows or for none");
    if (!given) return FEHT_OK;
    const size_t need = size_t(first_marker + n_marker_rows);
    if (need > c->rank.cap) {
        DevBuf<int32_t> nb;
        CU(c, ensure(nb, std::max(need, c->rank.cap + c->rank.cap / 2)));
        if (c->rank.p && first_marker > 0)
            CU(c, cudaMemcpyAsync(nb.p, c->rank.p, size_t(first_marker) * 4, cudaMemcpyDeviceToDevice, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        release(c->rank);
        c->rank = nb;
    }
    CU(c, cudaMemcpyAsync(c->rank.p + first_marker, name_rank, size_t(n_marker_rows) * 4,
                          cudaMemcpyHostToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}

int64_t marker_rows(const feht_ctx *c) { return c->snp == 1 ? 4 * c->n_units : c->n_units; }

int check_mode(feht_ctx *c, int snp) {
    if (c->S <= 0) return fail(c, FEHT_E_INVALID, "feht_set_samples must be called first");
    if (c->snp == -1) c->snp = snp;
    if (c->snp != snp)
        return fail(c, FEHT_E_INVALID, "a context holds either binary rows or SNP sites, not both");
    return FEHT_OK;
}

constexpr size_t kStageBytes = size_t(96) << 20;
// chars upload (add_chars_common): piece = about 8 MB of chars (one row if a row is longer); the DMA path moves up to
// 3 consecutive pieces that fit a 24 MB staging slot (4 slots in one 96 MB buffer); a host worker packs one piece
// into a pinned buffer of at most 4 MB.  (2 MB pieces, meant to keep the packed buffers in the last-level cache until
// the DMA engine has read them, were slower: 52 vs 48 ms -- the workers then take 97% of the rows and the link idles.)
constexpr size_t kPieceBytes = size_t(8) << 20;
constexpr int kDmaPiecesMax = 3, kDmaSlots = 4;
constexpr size_t kDmaSlotBytes = kStageBytes / kDmaSlots;
constexpr int kPackBufsMax = 2;
constexpr size_t kHostPackBufWords = size_t(4) << 17;   // 4 MB of 64-bit words
constexpr size_t kHostPackMinBytes = size_t(64) << 20;  // smaller uploads are not worth waking the cores for
// AUTO engine: categories go to the GEMM from this many pair-slots on (measured crossover)
constexpr int kGemmAutoMinSlots = 8;
// ... and from 4 pair-slots on for matrices of at least this many rows: K2 streams the planes at ~4.1 TB/s whatever the
// number of value columns (<= 64), K1 at 4.4 TB/s for two pair-slots but 2.6 TB/s for three or four (16 count streams)
// and 1.7 TB/s for six (profiles/README.md round 2, c2cat3 .. c2cat12)
constexpr int kGemmAutoMinSlotsLarge = 4;
constexpr int64_t kGemmAutoLargeRows = 16384;
constexpr bool kGemmAutoFp4 = true;   // 1.37x the int8 flavour on C4 / C5, bit-identical counts (profiles/README.md r1j)

// Upload `total` bytes from host in pieces chosen by `next_piece` and hand each staged piece to
// `consume(dev_ptr, piece_index)` on the context stream; two staging buffers so that the copy of
// piece i+1 overlaps the kernel on piece i.
template <class Next, class Consume>
int staged_upload(feht_ctx *c, Next next_piece, Consume consume) {
    for (int b = 0; b < 2; ++b) CU(c, ensure(c->stage[b], kStageBytes));
    int i = 0;
    for (;; ++i) {
        const int b = i & 1;
        const void *src = nullptr;
        size_t bytes = 0;
        if (!next_piece(i, &src, &bytes)) break;
        if (bytes > kStageBytes) return fail(c, FEHT_E_INVALID, "internal: staging piece too large");
        if (i >= 2) CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_used[b], 0));
        if (bytes)
            CU(c, cudaMemcpyAsync(c->stage[b].p, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
        CU(c, cudaStreamWaitEvent(c->st, c->ev_copied[b], 0));
        int rc = consume(c->stage[b].p, i);
        if (rc) return rc;
        CU(c, cudaEventRecord(c->ev_used[b], c->st));
    }
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}

int add_chars_common(feht_ctx *c, const uint8_t *chars, const int64_t *row_offsets, int64_t n,
                     const int32_t *name_rank, int snp) {
    if (!c) return FEHT_E_INVALID;
    if (int rc = use_device(c)) return rc;
    if (int rc = check_mode(c, snp)) return rc;
    if (n < 0 || (n > 0 && (!chars || !row_offsets))) return fail(c, FEHT_E_INVALID, "null rows");
    if (n == 0) return FEHT_OK;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t len = row_offsets[r + 1] - row_offsets[r];
        if (len < 0) return fail(c, FEHT_E_INVALID, "row_offsets must be non-decreasing");
        c->min_row_len = std::min(c->min_row_len, len);
    }
    const int n_planes = snp ? 3 : 2;
    if (int rc = grow_rows(c, c->n_units + n, n_planes)) return rc;
    const int mult = snp ? 4 : 1;
    if (int rc = append_rank(c, name_rank, n * mult, c->n_units * mult)) return rc;

    // Pieces = runs of whole rows with at most kPieceBytes of chars + offsets and kHostPackBufWords of packed
    // output.  Two producers fill the SAME bit-planes from opposite ends of the piece list:
    //  * the DMA path (this thread): up to kDmaPieces consecutive pieces are copied raw on the copy stream into
    //    one of four staging slots and packed by K0 on the context stream (events order slot reuse);
    //  * the host transcoder (worker threads, host_pack.cpp): a worker packs a piece with AVX2 into one of its
    //    two pinned buffers and uploads the planes at 2 bits per call.
    // The PCIe link carries 1 B/call for the first kind and 0.25 B/call for the second, so the upload runs at
    // 53 GB/s + 3/4 of whatever the cores pack instead of 53 GB/s (measured 110 GB/s on 16 cores alone).
    std::vector<int64_t> piece_first;
    {
        int64_t r = 0;
        while (r < n) {
            piece_first.push_back(r);
            size_t bytes = 0;
            int64_t q = r;
            while (q < n) {
                const size_t add = size_t(row_offsets[q + 1] - row_offsets[q]) + 8;
                if (add + 128 > kDmaSlotBytes)
                    return fail(c, FEHT_E_INVALID, "a single row exceeds the staging buffer");
                if (q > r && bytes + add > kPieceBytes) break;
                if (size_t(q + 1 - r) * size_t(c->W) * size_t(n_planes) > kHostPackBufWords) break;
                bytes += add;
                ++q;
            }
            if (q == r) return fail(c, FEHT_E_INVALID, "a single row exceeds the packing buffer");
            r = q;
        }
        piece_first.push_back(n);
    }
    const int64_t n_pieces = int64_t(piece_first.size()) - 1;
    const int64_t first_unit = c->n_units;
    const int64_t total_chars = row_offsets[n] - row_offsets[0];

    // ---- host transcoder workers
    int workers = 0;
    if (host_pack_available() && total_chars >= int64_t(kHostPackMinBytes) && n_pieces >= 8) {
        workers = c->host_threads;
        if (workers < 0) {
            const char *e = getenv("FEHT_HOST_PACK_THREADS");
            workers = e ? atoi(e) : int(std::thread::hardware_concurrency());
        }
        workers = std::max(0, std::min(workers, 64));
        if (int64_t(workers) > n_pieces / 2) workers = int(n_pieces / 2);
    }
    if (workers > c->pack_workers_alloc) {
        if (c->pack_pinned) { cudaFreeHost(c->pack_pinned); c->pack_pinned = nullptr; }
        CU(c, cudaHostAlloc(reinterpret_cast<void **>(&c->pack_pinned), size_t(workers) * kPackBufsMax * kHostPackBufWords * 8,
                            cudaHostAllocDefault));
        while (int(c->pack_ev.size()) < kPackBufsMax * workers) {
            cudaEvent_t ev = nullptr;
            CU(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            c->pack_ev.push_back(ev);
        }
        c->pack_workers_alloc = workers;
    }
    constexpr int kPackBufs = kPackBufsMax;
    std::mutex mu;
    int64_t lo = 0, hi = n_pieces - 1;
    auto claim_back = [&]() -> int64_t {
        std::lock_guard<std::mutex> g(mu);
        return lo <= hi ? hi-- : -1;
    };
    auto claim_front = [&](int64_t *count) -> int64_t {   // consecutive pieces that fit one staging slot
        std::lock_guard<std::mutex> g(mu);
        if (lo > hi) return -1;
        const int64_t first = lo;
        int64_t cnt = 0;
        while (lo <= hi && cnt < kDmaPiecesMax) {
            const size_t bytes = size_t(row_offsets[piece_first[size_t(lo) + 1]] - row_offsets[piece_first[size_t(first)]]) +
                                 size_t(piece_first[size_t(lo) + 1] - piece_first[size_t(first)] + 1) * 8 + 128;
            if (cnt > 0 && bytes > kDmaSlotBytes) break;
            ++lo;
            ++cnt;
        }
        *count = cnt;
        return first;
    };
    std::atomic<int> worker_err{0};
    std::atomic<int> host_digits{0};   // SNP: a worker met a literal '0' / '1'
    std::atomic<int64_t> host_rows{0};
    if (snp) {
        CU(c, ensure(c->snp_digit_flag, 1));
        CU(c, cudaMemsetAsync(c->snp_digit_flag.p, 0, sizeof(unsigned int), c->st));
    }
    std::vector<std::thread> pool;
    pool.reserve(size_t(workers));
    for (int t = 0; t < workers; ++t) {
        pool.emplace_back([&, t] {
            if (cudaSetDevice(c->device) != cudaSuccess) { worker_err = 1; return; }
            bool used[kPackBufsMax] = {false, false};
            int k = 0;
            for (;;) {
                const int64_t i = claim_back();
                if (i < 0) break;
                uint64_t *buf = c->pack_pinned + (size_t(t) * kPackBufsMax + size_t(k)) * kHostPackBufWords;
                cudaEvent_t ev = c->pack_ev[size_t(t) * kPackBufsMax + size_t(k)];
                if (used[k] && cudaEventSynchronize(ev) != cudaSuccess) { worker_err = 1; return; }
                const int64_t r0 = piece_first[size_t(i)], r1 = piece_first[size_t(i) + 1];
                const size_t words = size_t(r1 - r0) * size_t(c->W);
                const int64_t u0 = first_unit + r0;
                if (snp) {
                    if (host_pack_snp(chars, row_offsets, r0, r1, c->S, c->W, buf, buf + words, buf + 2 * words)) host_digits = 1;
                } else
                    host_pack_binary(chars, row_offsets, r0, r1, c->S, c->W, buf, buf + words);
                cudaError_t e = cudaSuccess;
                for (int pl = 0; pl < n_planes && e == cudaSuccess; ++pl)
                    e = cudaMemcpyAsync(c->plane[pl] + u0 * c->W, buf + size_t(pl) * words, words * 8,
                                        cudaMemcpyHostToDevice, c->pack_stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev, c->pack_stream);
                if (e != cudaSuccess) { worker_err = 1; return; }
                used[k] = true;
                k = (k + 1) % kPackBufs;
                host_rows += r1 - r0;
            }
        });
    }

    // ---- DMA path.  Layout of a staged slot: [offsets (rows+1) x int64, rebased to 0][chars...].
    auto join_all = [&] { for (std::thread &th : pool) th.join(); };
    std::vector<std::vector<int64_t>> rebased;   // alive until the final synchronize
    rebased.reserve(size_t(n_pieces) + 2);
    cudaError_t derr = ensure(c->stage[0], kStageBytes);
    for (int64_t k = 0; derr == cudaSuccess; ++k) {
        const int b = int(k % kDmaSlots);
        // host-side wait (not a stream wait): the DMA path claims work only when it can start on it, so the
        // workers get whatever it cannot take
        if (k >= kDmaSlots && (derr = cudaEventSynchronize(c->ev_used[b])) != cudaSuccess) break;
        int64_t cnt = 0;
        const int64_t i = claim_front(&cnt);
        if (i < 0) break;
        unsigned char *dev = c->stage[0].p + size_t(b) * kDmaSlotBytes;
        const int64_t r0 = piece_first[size_t(i)], r1 = piece_first[size_t(i + cnt)];
        const int64_t rows = r1 - r0;
        rebased.emplace_back(size_t(rows) + 1);
        std::vector<int64_t> &off = rebased.back();
        for (int64_t r = 0; r <= rows; ++r) off[size_t(r)] = row_offsets[r0 + r] - row_offsets[r0];
        const size_t off_bytes = (size_t(rows) + 1) * 8;
        const size_t off_pad = (off_bytes + 63) & ~size_t(63);
        const size_t char_bytes = size_t(row_offsets[r1] - row_offsets[r0]);
        if ((derr = cudaMemcpyAsync(dev, off.data(), off_bytes, cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) break;
        if (char_bytes && (derr = cudaMemcpyAsync(dev + off_pad, chars + row_offsets[r0], char_bytes,
                                                  cudaMemcpyHostToDevice, c->copy_stream)) != cudaSuccess) break;
        if ((derr = cudaEventRecord(c->ev_copied[b], c->copy_stream)) != cudaSuccess) break;
        if ((derr = cudaStreamWaitEvent(c->st, c->ev_copied[b], 0)) != cudaSuccess) break;
        const int64_t u0 = first_unit + r0;
        if (snp)
            derr = launch_pack_snp_chars(dev + off_pad, reinterpret_cast<const int64_t *>(dev), rows, c->S, c->W,
                                         c->plane[0] + u0 * c->W, c->plane[1] + u0 * c->W, c->plane[2] + u0 * c->W,
                                         c->snp_digit_flag.p, c->st);
        else
            derr = launch_pack_chars(dev + off_pad, reinterpret_cast<const int64_t *>(dev), rows, c->S, c->W,
                                     c->plane[0] + u0 * c->W, c->plane[1] + u0 * c->W, c->st);
        if (derr != cudaSuccess) break;
        derr = cudaEventRecord(c->ev_used[b], c->st);
    }
    if (derr != cudaSuccess) {   // let the workers drain the list before reporting
        { std::lock_guard<std::mutex> g(mu); lo = hi + 1; }
        join_all();
        return fail(c, FEHT_E_CUDA, "chars upload: %s", cudaGetErrorString(derr));
    }
    join_all();
    if (worker_err.load()) return fail(c, FEHT_E_CUDA, "chars upload: a host packing worker failed: %s",
                                       cudaGetErrorString(cudaGetLastError()));
    CU(c, cudaStreamSynchronize(c->pack_stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->st));
    if (snp) {
        unsigned int dev_digits = 0;
        CU(c, cudaMemcpy(&dev_digits, c->snp_digit_flag.p, sizeof(dev_digits), cudaMemcpyDeviceToHost));
        if (dev_digits || host_digits.load())   // nothing is appended (n_units unchanged)
            return fail(c, FEHT_E_INVALID, "the SNP rows hold literal '0' / '1' calls, which count for all four alleles "
                        "(Table.hs:194-205): upload the expanded rows with feht_add_rows_chars");
    }
    c->last_upload_host_rows = host_rows.load();
    c->n_units += n;
    c->has_results = false;
    c->counts_ready = false;
    c->rank_is_perm = -1;
    return FEHT_OK;
}

int add_packed_common(feht_ctx *c, const uint64_t *const *src_planes, const int *dst_index,
                      int n_planes, int64_t n, int64_t stride, const int32_t *name_rank, int snp) {
    if (!c) return FEHT_E_INVALID;
    if (int rc = use_device(c)) return rc;
    if (int rc = check_mode(c, snp)) return rc;
    const int64_t need_words = (c->S + 63) / 64;
    if (n < 0 || stride < need_words) return fail(c, FEHT_E_INVALID, "row_stride_words < ceil(S/64)");
    if (n == 0) return FEHT_OK;
    for (int i = 0; i < n_planes; ++i)
        if (!src_planes[i]) return fail(c, FEHT_E_INVALID, "null plane");
    if (int rc = grow_rows(c, c->n_units + n, n_planes)) return rc;
    const int mult = snp ? 4 : 1;
    if (int rc = append_rank(c, name_rank, n * mult, c->n_units * mult)) return rc;
    c->min_row_len = std::min(c->min_row_len, c->S);

    const int64_t rows_per_piece = std::max<int64_t>(1, int64_t(kStageBytes / (size_t(stride) * 8)));
    if (size_t(stride) * 8 > kStageBytes) return fail(c, FEHT_E_INVALID, "row too long for staging");
    const int64_t pieces_per_plane = (n + rows_per_piece - 1) / rows_per_piece;
    const int64_t first_unit = c->n_units;
    // plane order: valid first so that value can be ANDed with it (binary mode)
    auto next_piece = [&](int i, const void **src, size_t *bytes) -> bool {
        const int64_t pl = i / pieces_per_plane, pc = i % pieces_per_plane;
        if (pl >= n_planes) return false;
        const int64_t r0 = pc * rows_per_piece, r1 = std::min(n, r0 + rows_per_piece);
        *src = src_planes[pl] + r0 * stride;
        *bytes = size_t(r1 - r0) * size_t(stride) * 8;
        return true;
    };
    auto consume = [&](unsigned char *dev, int i) -> int {
        const int64_t pl = i / pieces_per_plane, pc = i % pieces_per_plane;
        const int64_t r0 = pc * rows_per_piece, r1 = std::min(n, r0 + rows_per_piece);
        uint64_t *dst = c->plane[dst_index[pl]] + (first_unit + r0) * c->W;
        const uint64_t *and_with = nullptr;
        // value / hi / lo bits only count where the call is valid (plane 1, uploaded first): K1 relies on it
        if (dst_index[pl] != 1) and_with = c->plane[1] + (first_unit + r0) * c->W;
        CU(c, launch_repack_rows(reinterpret_cast<const uint64_t *>(dev), stride, dst, c->W, c->S,
                                 r1 - r0, and_with, c->st));
        return FEHT_OK;
    };
    if (int rc = staged_upload(c, next_piece, consume)) return rc;
    c->n_units += n;
    c->has_results = false;
    c->counts_ready = false;
    c->rank_is_perm = -1;
    return FEHT_OK;
}

}  // namespace

// =================================================================================== context ==
extern "C" const char *feht_version(void) { return "feht_b200 0.3 (sm_100a; K0 pack, K1 popc, K2 tcgen05 gemm i8/mxf4, K3 fet, K4 sort, K5b merge; multi-device)"; }

extern "C" int feht_create(feht_ctx **out, int device) {
    if (!out) return fail(nullptr, FEHT_E_INVALID, "out is null");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, FEHT_E_NODEVICE, "no CUDA device (%s); libfeht_cuda has no CPU fallback",
                    e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n_dev) return fail(nullptr, FEHT_E_INVALID, "device %d of %d", device, n_dev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, FEHT_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, FEHT_E_NODEVICE,
                    "device %d is sm_%d%d; libfeht_cuda is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    feht_ctx *c = new (std::nothrow) feht_ctx();
    if (!c) return fail(nullptr, FEHT_E_NOMEM, "out of host memory");
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    auto bail = [&](const char *what, cudaError_t err) {
        fail(nullptr, FEHT_E_CUDA, "%s: %s", what, cudaGetErrorString(err));
        feht_destroy(c);
        return FEHT_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    c->st = c->own_stream;
    for (int i = 0; i < 5; ++i)
        for (int k = 0; k < 2; ++k)
            if ((e = cudaEventCreate(&c->evr[k][i])) != cudaSuccess) return bail("event", e);
    if ((e = cudaStreamCreateWithFlags(&c->pack_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    for (int i = 0; i < 4; ++i) {
        if ((e = cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
        if ((e = cudaEventCreateWithFlags(&c->ev_used[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    }
    const std::vector<double> &tab = host_choose_table();
    if ((e = cudaMalloc(&c->d_choose, tab.size() * 8)) != cudaSuccess) return bail("cudaMalloc(choose)", e);
    if ((e = cudaMemcpy(c->d_choose, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail("cudaMemcpy(choose)", e);
    *out = c;
    return FEHT_OK;
}

extern "C" void feht_destroy(feht_ctx *c) {
    if (!c) return;
    if (c->multi) { multi::destroy(c->multi); c->multi = nullptr; }
    cudaSetDevice(c->device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    for (int i = 0; i < 3; ++i) if (c->plane[i]) cudaFree(c->plane[i]);
    release(c->rank); release(c->stage[0]); release(c->stage[1]);
    release(c->gemm_b); release(c->gemm_passes);
    release(c->masks); release(c->counts); release(c->groups); release(c->fracs); release(c->gcomps); release(c->sdefs); release(c->seg_count);
    release(c->r_comp); release(c->r_rec); release(c->o_row); release(c->o_rank);
    release(c->o_goa); release(c->o_gob); release(c->o_gta); release(c->o_gtb);
    release(c->r_p); release(c->o_p); release(c->o_ratio);
    release(c->sw_key_a); release(c->sw_key_b); release(c->sw_val_a); release(c->sw_val_b);
    release(c->sw_control); release(c->sw_oa); release(c->sw_coop);
    release(c->live_rows); release(c->live_count); release(c->snp_digit_flag);
    release(c->dd_keys); release(c->dd_owner); release(c->dd_count); release(c->dd_list); release(c->merge_dest);
    release(c->txt_counts); release(c->txt_starts); release(c->txt_name_len); release(c->txt_flags);
    if (c->d_choose) cudaFree(c->d_choose);
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 5; ++i) if (c->evr[k][i]) cudaEventDestroy(c->evr[k][i]);
    for (int i = 0; i < 4; ++i) {
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_used[i]) cudaEventDestroy(c->ev_used[i]);
    }
    for (cudaEvent_t ev : c->pack_ev) cudaEventDestroy(ev);
    if (c->pack_pinned) cudaFreeHost(c->pack_pinned);
    if (c->pack_stream) cudaStreamDestroy(c->pack_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

extern "C" const char *feht_last_error(const feht_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

// One handle for several GPUs of the box (feht_create(ctx, device_ids, n_devices)): rows shard over the
// listed devices, the merge of the ordered shards happens on device_ids[0] (multi.cu).
extern "C" int feht_create_multi(feht_ctx **out, const int *device_ids, int n_devices) {
    if (!out) return fail(nullptr, FEHT_E_INVALID, "out is null");
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 32)
        return fail(nullptr, FEHT_E_INVALID, "feht_create_multi: between 1 and 32 device ids");
    feht_ctx *h = nullptr;
    if (int rc = feht_create(&h, device_ids[0])) return rc;
    std::string err;
    MultiState *m = multi::create(device_ids, n_devices, &err);
    if (!m) {
        feht_destroy(h);
        // (a device that is missing or not sm_100-class is the usual reason)
        return fail(nullptr, FEHT_E_NODEVICE, "%s", err.c_str());
    }
    h->multi = m;
    *out = h;
    return FEHT_OK;
}

extern "C" int feht_num_devices(const feht_ctx *c) { return !c ? 0 : (c->multi ? multi::n_devices(c->multi) : 1); }

// (internal, multi.cu) the ordered result columns of a single-device context's last run
extern "C" int feht_internal_results(feht_ctx *c, ResultSoA *out, const int64_t **seg_off_host, int64_t *n_total, int *device) {
    if (!c || c->multi) return FEHT_E_INVALID;
    if (!c->has_results) return fail(c, FEHT_E_INVALID, "no results: call feht_run first");
    if (out) *out = ResultSoA{c->o_row.p, c->o_rank.p, c->o_goa.p, c->o_gob.p, c->o_gta.p, c->o_gtb.p, c->o_p.p, c->o_ratio.p};
    if (seg_off_host) *seg_off_host = c->seg_off.data();
    if (n_total) *n_total = c->n_surv;
    if (device) *device = c->device;
    return FEHT_OK;
}

extern "C" int feht_set_fet_mode(feht_ctx *c, int mode) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_fet_mode(c->multi, mode));
    if (mode != FEHT_FET_TWO_TAIL && mode != FEHT_FET_ONE_TAIL) return fail(c, FEHT_E_INVALID, "unknown FET mode %d", mode);
    c->fet_mode = mode;
    c->has_results = false;
    return FEHT_OK;
}

// The wire format on its own: a host that keeps the packed copy (or wants the upload off its critical path) packs
// with its cores and calls feht_add_rows_packed / feht_add_snp_packed later.  No GPU involved.
extern "C" int feht_pack_chars_host(const uint8_t *chars, const int64_t *row_offsets, int64_t n_rows, int64_t n_samples,
                                    int snp_mode, int64_t row_stride_words, uint64_t *plane0, uint64_t *plane1,
                                    uint64_t *plane2, int n_threads) {
    if (n_rows < 0 || n_samples <= 0 || row_stride_words < (n_samples + 63) / 64) return FEHT_E_INVALID;
    if (n_rows == 0) return FEHT_OK;
    if (!chars || !row_offsets || !plane0 || !plane1 || (snp_mode && !plane2)) return FEHT_E_INVALID;
    for (int64_t r = 0; r < n_rows; ++r)
        if (row_offsets[r + 1] < row_offsets[r]) return FEHT_E_INVALID;
    int threads = n_threads > 0 ? n_threads : int(std::thread::hardware_concurrency());
    threads = int(std::max<int64_t>(1, std::min<int64_t>(std::min(threads, 256), n_rows)));
    const bool simd = host_pack_available();
    std::atomic<int> digits{0};
    auto work = [&](int64_t r0, int64_t r1) {
        const size_t o = size_t(r0) * size_t(row_stride_words);
        if (snp_mode) {   // planes in the order feht_add_snp_packed takes them: hi, lo, valid
            const bool d = simd ? host_pack_snp(chars, row_offsets, r0, r1, n_samples, row_stride_words, plane0 + o, plane2 + o, plane1 + o)
                                : host_pack_snp_scalar(chars, row_offsets, r0, r1, n_samples, row_stride_words, plane0 + o, plane2 + o, plane1 + o);
            if (d) digits = 1;
        } else {
            if (simd) host_pack_binary(chars, row_offsets, r0, r1, n_samples, row_stride_words, plane0 + o, plane1 + o);
            else host_pack_binary_scalar(chars, row_offsets, r0, r1, n_samples, row_stride_words, plane0 + o, plane1 + o);
        }
    };
    if (threads == 1) {
        work(0, n_rows);
    } else {
        std::vector<std::thread> pool;
        const int64_t per = (n_rows + threads - 1) / threads;
        for (int t = 0; t < threads; ++t) {
            const int64_t r0 = t * per, r1 = std::min(n_rows, r0 + per);
            if (r0 < r1) pool.emplace_back(work, r0, r1);
        }
        for (std::thread &th : pool) th.join();
    }
    // SNP rows with literal '0' / '1' calls cannot be expressed in the 2-bit planes (Table.hs:194-205)
    return digits.load() ? FEHT_E_INVALID : FEHT_OK;
}

extern "C" int feht_set_host_threads(feht_ctx *c, int n_threads) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_host_threads(c->multi, n_threads));
    c->host_threads = n_threads < 0 ? -1 : n_threads;
    return FEHT_OK;
}

extern "C" int64_t feht_last_upload_host_rows(const feht_ctx *c) {
    if (c && c->multi) return multi::last_upload_host_rows(c->multi);
    return c ? c->last_upload_host_rows : FEHT_E_INVALID;
}

extern "C" int feht_set_stream(feht_ctx *c, void *cuda_stream) {
    if (!c) return FEHT_E_INVALID;
    if (c->multi) return fail(c, FEHT_E_INVALID, "a multi-device context runs every shard on its own stream");
    c->st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return FEHT_OK;
}

// ============================================================================== data matrix ==
extern "C" int feht_set_samples(feht_ctx *c, int64_t n_samples) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_samples(c->multi, n_samples));
    if (n_samples <= 0 || n_samples > (int64_t(1) << 24))
        return fail(c, FEHT_E_INVALID, "n_samples must be in [1, 2^24]");
    if (int rc = use_device(c)) return rc;
    CU(c, cudaStreamSynchronize(c->st));
    for (int i = 0; i < 3; ++i) { if (c->plane[i]) cudaFree(c->plane[i]); c->plane[i] = nullptr; }
    c->gemm_ready = false;
    c->plan_ready = false;
    // comparisons and categories are column lists / per-sample value vectors of the OLD sample count: drop them
    // (a category's value_of_sample has exactly S entries; keeping it across a larger S would be read out of bounds)
    c->comps.clear();
    c->cats.clear();
    c->counts_ready = false;
    c->S = n_samples;
    c->W = ((n_samples + 63) / 64 + 1) & ~int64_t(1);  // even number of 64-bit words: 16-byte rows
    c->snp = -1;
    c->n_units = c->cap_units = 0;
    c->rank_is_perm = -1;
    c->have_rank = -1;
    c->min_row_len = std::numeric_limits<int64_t>::max();
    c->has_results = false;
    return FEHT_OK;
}

extern "C" int feht_reserve_rows(feht_ctx *c, int64_t n_rows) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::reserve_rows(c->multi, n_rows));
    if (c->S <= 0) return fail(c, FEHT_E_INVALID, "feht_set_samples must be called first");
    if (int rc = use_device(c)) return rc;
    // planes are allocated lazily per mode; remember the wish
    if (c->snp == -1) { c->cap_units = 0; }
    const int n_planes = c->snp == 1 ? 3 : 2;
    if (c->snp == -1) {
        // allocate the two planes every mode needs; the SNP lo plane follows on first use
        return grow_rows(c, n_rows, 2);
    }
    return grow_rows(c, n_rows, n_planes);
}

extern "C" int feht_clear_rows(feht_ctx *c) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::clear_rows(c->multi));
    c->n_units = 0;
    c->snp = -1;
    c->rank_is_perm = -1;
    c->have_rank = -1;
    c->min_row_len = std::numeric_limits<int64_t>::max();
    c->has_results = false;
    c->counts_ready = false;
    return FEHT_OK;
}

extern "C" int feht_add_rows_chars(feht_ctx *c, const uint8_t *chars, const int64_t *row_offsets,
                                   int64_t n_rows, const int32_t *name_rank) {
    MULTI(c, multi::add_chars(c->multi, chars, row_offsets, n_rows, name_rank, 0));
    return add_chars_common(c, chars, row_offsets, n_rows, name_rank, 0);
}

extern "C" int feht_add_snp_chars(feht_ctx *c, const uint8_t *chars, const int64_t *row_offsets,
                                  int64_t n_sites, const int32_t *name_rank) {
    MULTI(c, multi::add_chars(c->multi, chars, row_offsets, n_sites, name_rank, 1));
    return add_chars_common(c, chars, row_offsets, n_sites, name_rank, 1);
}

// ------------------------------------------------------------------------- text upload (N2) --
// Pieces of whole lines (<= 64 MB) go to the device as they are; text_load.cu finds the lines and packs them.
// The copy of piece i + 1 runs while the kernels of piece i do.
extern "C" int feht_add_rows_text(feht_ctx *c, const char *text, int64_t len, char delimiter, int snp_mode,
                                  int64_t *n_lines_out) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::add_text(c->multi, text, len, delimiter, snp_mode, n_lines_out));
    if (int rc = use_device(c)) return rc;
    const int snp = snp_mode ? 1 : 0;
    if (int rc = check_mode(c, snp)) return rc;
    if (len < 0 || (len > 0 && !text)) return fail(c, FEHT_E_INVALID, "null text");
    if (delimiter == '\n' || delimiter == '\r') return fail(c, FEHT_E_INVALID, "the delimiter cannot be a line break");
    if (c->have_rank == 1)
        return fail(c, FEHT_E_INVALID, "text rows get their name ranks afterwards (feht_set_name_rank); the context "
                    "already holds rows with ranks");
    c->text_line_begin.clear();
    c->text_name_len.clear();
    if (n_lines_out) *n_lines_out = 0;
    if (len == 0) return FEHT_OK;
    constexpr int64_t kTextPiece = int64_t(64) << 20;
    static_assert(size_t(kTextPiece) <= kStageBytes, "a text piece must fit a staging buffer");
    for (int b = 0; b < 2; ++b) CU(c, ensure(c->stage[b], kStageBytes));
    CU(c, ensure(c->txt_counts, text_scratch_words(size_t(kTextPiece))));
    CU(c, ensure(c->txt_flags, 3));
    const unsigned long long flags0[3] = {~0ull, 0ull, 0ull};
    CU(c, cudaMemcpyAsync(c->txt_flags.p, flags0, sizeof(flags0), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));   // (also: nothing of an earlier call still reads the staging buffers)

    // piece boundaries: back up from every 64 MB mark to the last newline
    std::vector<int64_t> cut{0};
    while (cut.back() < len) {
        int64_t end = std::min(len, cut.back() + kTextPiece);
        if (end < len) {
            const void *nl = memrchr(text + cut.back(), '\n', size_t(end - cut.back()));
            if (!nl) return fail(c, FEHT_E_INVALID, "a line of the data file is longer than 64 MB");
            end = static_cast<const char *>(nl) - text + 1;
        }
        cut.push_back(end);
    }
    const int n_pieces = int(cut.size()) - 1;
    const int n_planes = snp ? 3 : 2;
    const int64_t first_unit = c->n_units;
    int64_t units = first_unit;
    auto issue_copy = [&](int i) -> cudaError_t {
        const int b = i & 1;
        cudaError_t e = cudaMemcpyAsync(c->stage[b].p, text + cut[size_t(i)], size_t(cut[size_t(i) + 1] - cut[size_t(i)]),
                                        cudaMemcpyHostToDevice, c->copy_stream);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(c->ev_copied[b], c->copy_stream);
    };
    CU(c, issue_copy(0));
    std::vector<int32_t> h_starts, h_names;
    for (int i = 0; i < n_pieces; ++i) {
        const int b = i & 1;
        const int64_t bytes = cut[size_t(i) + 1] - cut[size_t(i)];
        // (the kernels of piece i - 1, which read the other staging buffer, were waited for at the end of the last turn)
        if (i + 1 < n_pieces) CU(c, issue_copy(i + 1));
        CU(c, cudaStreamWaitEvent(c->st, c->ev_copied[b], 0));
        const uint8_t *d_text = c->stage[b].p;
        CU(c, launch_text_count(d_text, bytes, c->txt_counts.p, c->st));
        const int nb = int((bytes + 4095) / 4096);
        uint32_t newlines = 0;
        CU(c, cudaMemcpyAsync(&newlines, c->txt_counts.p + nb, 4, cudaMemcpyDeviceToHost, c->st));
        CU(c, cudaStreamSynchronize(c->st));
        const bool open_end = text[cut[size_t(i) + 1] - 1] != '\n';   // only the last piece can end without a newline
        const int64_t n_lines = int64_t(newlines) + (open_end ? 1 : 0);
        if (size_t(n_lines) + 2 > c->txt_starts.cap || !c->txt_starts.p) CU(c, ensure(c->txt_starts, size_t(n_lines) * 5 / 4 + 64));
        if (size_t(n_lines) > c->txt_name_len.cap || !c->txt_name_len.p) CU(c, ensure(c->txt_name_len, size_t(n_lines) * 5 / 4 + 64));
        // rows: one allocation sized from the first piece's line density, then grow_rows' own policy
        int64_t want = units + n_lines;
        if (i == 0 && n_pieces > 1) want = first_unit + int64_t(double(n_lines) * double(len) / double(bytes) * 1.01) + 64;
        if (int rc = grow_rows(c, want, n_planes)) return rc;
        CU(c, launch_text_starts(d_text, bytes, c->txt_counts.p, c->txt_starts.p, c->st));
        if (open_end) {
            const int32_t sentinel = int32_t(bytes + 1);
            CU(c, cudaMemcpyAsync(c->txt_starts.p + n_lines, &sentinel, 4, cudaMemcpyHostToDevice, c->st));
        }
        CU(c, launch_text_pack(d_text, c->txt_starts.p, n_lines, uint8_t(delimiter), snp, c->S, c->W,
                               c->plane[0] + units * c->W, c->plane[1] + units * c->W,
                               snp ? c->plane[2] + units * c->W : nullptr, c->txt_name_len.p, c->txt_flags.p,
                               c->num_sms, c->st));
        h_starts.resize(size_t(n_lines));
        h_names.resize(size_t(n_lines));
        if (n_lines > 0) {
            CU(c, cudaMemcpyAsync(h_starts.data(), c->txt_starts.p, size_t(n_lines) * 4, cudaMemcpyDeviceToHost, c->st));
            CU(c, cudaMemcpyAsync(h_names.data(), c->txt_name_len.p, size_t(n_lines) * 4, cudaMemcpyDeviceToHost, c->st));
        }
        CU(c, cudaStreamSynchronize(c->st));
        for (int64_t k = 0; k < n_lines; ++k) c->text_line_begin.push_back(cut[size_t(i)] + h_starts[size_t(k)]);
        c->text_name_len.insert(c->text_name_len.end(), h_names.begin(), h_names.end());
        units += n_lines;
    }
    unsigned long long flags[3] = {0, 0, 0};
    CU(c, cudaMemcpyAsync(flags, c->txt_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    if (flags[1]) {
        c->text_line_begin.clear(); c->text_name_len.clear();
        // `g:gnxs = xs` on the [] that BS.split gives for an empty line (Table.hs:168)
        return fail(c, FEHT_E_INVALID, "Irrefutable pattern failed for pattern g : gnxs (an empty line in the data file)");
    }
    if (snp && flags[2]) {
        c->text_line_begin.clear(); c->text_name_len.clear();
        return fail(c, FEHT_E_INVALID, "the SNP text holds literal '0' / '1' calls, which count for all four alleles "
                    "(Table.hs:194-205): upload the expanded rows with feht_add_rows_chars");
    }
    if (units > first_unit) {
        c->min_row_len = std::min<int64_t>(c->min_row_len, int64_t(flags[0]));
        if (c->have_rank == -1) c->have_rank = 0;
    }
    c->n_units = units;
    c->rank_is_perm = -1;
    c->has_results = false;
    c->counts_ready = false;
    if (n_lines_out) *n_lines_out = units - first_unit;
    return FEHT_OK;
}

extern "C" int feht_text_lines(feht_ctx *c, int64_t *line_begin, int32_t *name_len, int64_t cap) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::text_lines(c->multi, line_begin, name_len, cap));
    const int64_t n = int64_t(c->text_line_begin.size());
    if (cap < n) return fail(c, FEHT_E_INVALID, "feht_text_lines: %lld lines, room for %lld", (long long)n, (long long)cap);
    if (n > 0 && line_begin) std::memcpy(line_begin, c->text_line_begin.data(), size_t(n) * 8);
    if (n > 0 && name_len) std::memcpy(name_len, c->text_name_len.data(), size_t(n) * 4);
    return FEHT_OK;
}

extern "C" int feht_set_name_rank(feht_ctx *c, const int32_t *name_rank, int64_t n_marker_rows) {
    if (!c || !name_rank) return FEHT_E_INVALID;
    MULTI(c, multi::set_name_rank(c->multi, name_rank, n_marker_rows));
    if (int rc = use_device(c)) return rc;
    if (n_marker_rows != marker_rows(c))
        return fail(c, FEHT_E_INVALID, "feht_set_name_rank: %lld ranks for %lld marker rows", (long long)n_marker_rows,
                    (long long)marker_rows(c));
    if (n_marker_rows == 0) return FEHT_OK;
    c->have_rank = 1;
    c->rank_is_perm = -1;
    c->has_results = false;
    return append_rank(c, name_rank, n_marker_rows, 0);
}

extern "C" int feht_add_rows_packed(feht_ctx *c, const uint64_t *value_plane, const uint64_t *valid_plane,
                                    int64_t n_rows, int64_t row_stride_words, const int32_t *name_rank) {
    MULTI(c, multi::add_packed(c->multi, value_plane, valid_plane, nullptr, n_rows, row_stride_words, name_rank, 0));
    const uint64_t *src[2] = {valid_plane, value_plane};
    const int dst[2] = {1, 0};
    return add_packed_common(c, src, dst, 2, n_rows, row_stride_words, name_rank, 0);
}

extern "C" int feht_add_snp_packed(feht_ctx *c, const uint64_t *hi_plane, const uint64_t *lo_plane,
                                   const uint64_t *valid_plane, int64_t n_sites, int64_t row_stride_words,
                                   const int32_t *name_rank) {
    MULTI(c, multi::add_packed(c->multi, hi_plane, lo_plane, valid_plane, n_sites, row_stride_words, name_rank, 1));
    const uint64_t *src[3] = {valid_plane, hi_plane, lo_plane};
    const int dst[3] = {1, 0, 2};
    return add_packed_common(c, src, dst, 3, n_sites, row_stride_words, name_rank, 1);
}

extern "C" int feht_generate_synthetic(feht_ctx *c, const feht_synth_spec *spec, int snp_mode,
                                       int64_t first_row, int64_t n_rows) {
    if (!c || !spec) return FEHT_E_INVALID;
    MULTI(c, multi::generate_synthetic(c->multi, spec, snp_mode, first_row, n_rows));
    if (int rc = use_device(c)) return rc;
    if (int rc = check_mode(c, snp_mode ? 1 : 0)) return rc;
    if (spec->n_samples != c->S) return fail(c, FEHT_E_INVALID, "spec.n_samples != feht_set_samples");
    if (n_rows < 0) return fail(c, FEHT_E_INVALID, "n_rows < 0");
    if (c->have_rank == 1) return fail(c, FEHT_E_INVALID, "synthetic rows carry no name_rank");
    c->have_rank = 0;
    if (int rc = grow_rows(c, c->n_units + n_rows, snp_mode ? 3 : 2)) return rc;
    feht_synth_spec *d_spec = nullptr;
    CU(c, cudaMalloc(&d_spec, sizeof(feht_synth_spec)));
    CU(c, cudaMemcpyAsync(d_spec, spec, sizeof(feht_synth_spec), cudaMemcpyHostToDevice, c->st));
    const int64_t u0 = c->n_units;
    cudaError_t e = launch_synth(d_spec, snp_mode ? 1 : 0, first_row, n_rows, c->S, c->W,
                                 c->plane[0] + u0 * c->W, c->plane[1] + u0 * c->W,
                                 snp_mode ? c->plane[2] + u0 * c->W : nullptr, c->st);
    cudaError_t e2 = cudaStreamSynchronize(c->st);
    cudaFree(d_spec);
    CU(c, e);
    CU(c, e2);
    c->min_row_len = std::min(c->min_row_len, c->S);
    c->n_units += n_rows;
    c->has_results = false;
    c->counts_ready = false;
    return FEHT_OK;
}

extern "C" int feht_read_plane(feht_ctx *c, int plane, int64_t row0, int64_t n, uint64_t *out) {
    if (!c || !out) return FEHT_E_INVALID;
    MULTI(c, multi::read_plane(c->multi, plane, row0, n, out));
    if (int rc = use_device(c)) return rc;
    if (plane < 0 || plane > 2 || !c->plane[plane] || (plane == 2 && c->snp != 1))
        return fail(c, FEHT_E_INVALID, "no such plane");
    if (row0 < 0 || n < 0 || row0 + n > c->n_units) return fail(c, FEHT_E_INVALID, "row range");
    CU(c, cudaMemcpyAsync(out, c->plane[plane] + row0 * c->W, size_t(n) * c->W * 8, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}

extern "C" int64_t feht_row_stride_words(const feht_ctx *c) {
    if (c && c->multi) return multi::row_stride_words(c->multi);
    return c ? c->W : 0;
}
extern "C" int64_t feht_num_rows(const feht_ctx *c) {
    if (c && c->multi) return multi::num_rows(c->multi);
    return c ? marker_rows(c) : 0;
}

// ============================================================================== comparisons ==
extern "C" int64_t feht_add_comparison(feht_ctx *c, const int32_t *g1, int64_t n1, const int32_t *g2, int64_t n2) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::add_comparison(c->multi, g1, n1, g2, n2));
    if (n1 < 0 || n2 < 0 || (n1 && !g1) || (n2 && !g2)) return fail(c, FEHT_E_INVALID, "null group");
    Comp cp;
    cp.kind = 0;
    cp.g1.assign(g1, g1 + n1);
    cp.g2.assign(g2, g2 + n2);
    for (int32_t v : cp.g1) { if (v < 0) return fail(c, FEHT_E_BOUNDS, "negative column index"); cp.max_col = std::max<int64_t>(cp.max_col, v); }
    for (int32_t v : cp.g2) { if (v < 0) return fail(c, FEHT_E_BOUNDS, "negative column index"); cp.max_col = std::max<int64_t>(cp.max_col, v); }
    c->comps.push_back(std::move(cp));
    c->gemm_ready = false;
    c->plan_ready = false;  // explicit slots come first: the categories' slot base moves
    c->counts_ready = false;
    c->has_results = false;
    return int64_t(c->comps.size()) - 1;
}

extern "C" int64_t feht_add_category(feht_ctx *c, const int32_t *value_of_sample, int32_t n_values) {
    if (!c || !value_of_sample) return FEHT_E_INVALID;
    MULTI(c, multi::add_category(c->multi, value_of_sample, n_values));
    if (c->S <= 0) return fail(c, FEHT_E_INVALID, "feht_set_samples must be called first");
    if (n_values < 2) return fail(c, FEHT_E_INVALID, "a category needs at least 2 values");
    Cat cat;
    cat.n_values = n_values;
    cat.value_of_sample.assign(value_of_sample, value_of_sample + c->S);
    for (int64_t s = 0; s < c->S; ++s) {
        const int32_t v = cat.value_of_sample[size_t(s)];
        if (v >= n_values || v < -1) return fail(c, FEHT_E_INVALID, "value_of_sample[%lld] = %d out of range", (long long)s, v);
        if (v >= 0) cat.max_col = s;
    }
    const int ci = int(c->cats.size());
    c->cats.push_back(std::move(cat));
    c->gemm_ready = false;
    c->plan_ready = false;
    c->counts_ready = false;
    const int64_t first = int64_t(c->comps.size());
    for (int i = 0; i < n_values; ++i)
        for (int j = i + 1; j < n_values; ++j) {
            Comp cp; cp.kind = 1; cp.cat = ci; cp.v1 = i; cp.v2 = j; cp.max_col = c->cats[ci].max_col;
            c->comps.push_back(std::move(cp));
        }
    if (n_values > 2)  // Comparison.hs:329
        for (int v = 0; v < n_values; ++v) {
            Comp cp; cp.kind = 2; cp.cat = ci; cp.v1 = v; cp.v2 = -1; cp.max_col = c->cats[ci].max_col;
            c->comps.push_back(std::move(cp));
        }
    c->has_results = false;
    return first;
}

extern "C" int feht_clear_comparisons(feht_ctx *c) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::clear_comparisons(c->multi));
    c->comps.clear();
    c->cats.clear();
    c->gemm_ready = false;
    c->plan_ready = false;
    c->counts_ready = false;
    c->has_results = false;
    return FEHT_OK;
}

extern "C" int64_t feht_num_comparisons(const feht_ctx *c) {
    if (c && c->multi) return multi::num_comparisons(c->multi);
    return c ? int64_t(c->comps.size()) : 0;
}

extern "C" int feht_comparison_info(const feht_ctx *c, int64_t i, int32_t *kind, int32_t *category, int32_t *v1, int32_t *v2) {
    if (c && c->multi) return multi::comparison_info(c->multi, i, kind, category, v1, v2);
    if (!c || i < 0 || i >= int64_t(c->comps.size())) return FEHT_E_INVALID;
    const Comp &cp = c->comps[size_t(i)];
    if (kind) *kind = cp.kind;
    if (category) *category = cp.cat;
    if (v1) *v1 = cp.v1;
    if (v2) *v2 = cp.v2;
    return FEHT_OK;
}

extern "C" int feht_set_count_engine(feht_ctx *c, int engine) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_count_engine(c->multi, engine));
    if (engine < FEHT_ENGINE_AUTO || engine > FEHT_ENGINE_GEMM_F8) return fail(c, FEHT_E_INVALID, "unknown engine");
    c->engine = engine;
    c->counts_ready = false;
    return FEHT_OK;
}

// Emit only the comparisons [first, first + count) in the following runs (count < 0: all again).  For outputs too
// large for one run -- the reference's default ratioFilter 0 keeps every test (Comparison.hs:148), so a category of
// V values over R rows is (V(V-1)/2 + V) * R result rows -- the host walks the comparisons window by window:
// run, fetch, next window.  The contingency counts are computed by the first window's run and kept while a window
// stays set (they do not depend on the window); rows, comparisons, the engine or feht_set_samples reset them.
extern "C" int feht_set_comparison_window(feht_ctx *c, int64_t first, int64_t count) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_comparison_window(c->multi, first, count));
    if (first < 0) return fail(c, FEHT_E_INVALID, "window start < 0");
    if (count < 0) { first = 0; count = -1; }
    if (c->win_count < 0 || count < 0) c->counts_ready = false;   // entering / leaving windowed mode: count afresh
    c->win_first = first;
    c->win_count = count;
    c->plan_ready = false;     // the screen groups list the emitted comparisons
    c->has_results = false;
    return FEHT_OK;
}

extern "C" int feht_set_row_base(feht_ctx *c, int64_t row_base) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::set_row_base(c->multi, row_base));
    c->row_base = row_base;
    return FEHT_OK;
}

// ====================================================================================== run ==
namespace {
// which GEMM flavour FEHT_ENGINE_AUTO uses for categories; FEHT_GEMM=i8 / fp4 overrides
bool gemm_auto_fp4() {
    static const bool fp4 = [] {
        const char *e = getenv("FEHT_GEMM");
        if (e && std::string(e) == "i8") return false;
        if (e && std::string(e) == "fp4") return true;
        return kGemmAutoFp4;
    }();
    return fp4;
}
// The e5m2 flavour (one GEMM for both planes) issues half the MMAs but needs ~6 instructions per 4 calls in the
// expanders and ends up bound by them (config 4: 4.87 ms against 3.85 ms, profiles/README.md r1q): AUTO only takes
// it with FEHT_GEMM=f8
bool gemm_auto_f8() {
    static const bool f8 = [] {
        const char *e = getenv("FEHT_GEMM");
        return e && std::string(e) == "f8";
    }();
    return f8;
}

// GEMM columns of the e5m2 flavour, pass by pass: every value of every category gets ceil(members / 4095) columns
// (at least one), a category never straddles two passes.  Empty result = the categories do not fit.
std::vector<std::vector<GemmF8Column>> f8_columns(const feht_ctx *c, const std::vector<int> &cat_slot0,
                                                  const std::vector<int> &cat_tot) {
    std::vector<std::vector<GemmF8Column>> passes;
    const int max_cols = gemm_f8_max_cols(), max_members = gemm_f8_max_members();
    for (size_t k = 0; k < c->cats.size(); ++k) {
        const Cat &cat = c->cats[k];
        std::vector<std::vector<int32_t>> members(static_cast<size_t>(cat.n_values));
        for (int64_t s = 0; s < c->S; ++s) {
            const int v = cat.value_of_sample[size_t(s)];
            if (v >= 0) members[size_t(v)].push_back(int32_t(s));
        }
        std::vector<GemmF8Column> cols;
        for (int v = 0; v < cat.n_values; ++v) {
            const std::vector<int32_t> &m = members[size_t(v)];
            const size_t chunks = std::max<size_t>(1, (m.size() + size_t(max_members) - 1) / size_t(max_members));
            for (size_t q = 0; q < chunks; ++q) {
                GemmF8Column col;
                const size_t lo = q * size_t(max_members), hi = std::min(m.size(), lo + size_t(max_members));
                if (lo < hi) col.samples.assign(m.begin() + long(lo), m.begin() + long(hi));
                col.slot = cat_slot0[k] + v / 2;
                col.flags = uint8_t(v & 1);
                if (q + 1 == chunks) {
                    col.flags |= 2;
                    if ((v & 1) || v == cat.n_values - 1) col.flags |= 4;
                    if (v == cat.n_values - 1) { col.flags |= 8; col.tot_slot = cat_tot[k]; }
                }
                cols.push_back(std::move(col));
            }
        }
        if (int(cols.size()) > max_cols) return {};
        if (passes.empty() || int(passes.back().size() + cols.size()) > max_cols) passes.emplace_back();
        for (GemmF8Column &col : cols) passes.back().push_back(std::move(col));
    }
    return passes;
}
// FEHT_DEDUP=0: every exact test evaluates its own hypergeometric sum (A/B measurements)
bool dedup_enabled() {
    static const bool on = [] { const char *e = getenv("FEHT_DEDUP"); return !(e && atoi(e) == 0); }();
    return on;
}
// FEHT_SORT=onesweep keeps the tile-per-CTA sort for every size (A/B measurements and tests of that path)
bool sort_engine_allows_coop() {
    static const bool allow = [] {
        const char *e = getenv("FEHT_SORT");
        return !(e && std::string(e) == "onesweep");
    }();
    return allow;
}
}  // namespace

namespace {
// stage times of the run recorded in event set `slot` (waits for that run to finish)
int settle(feht_ctx *c, int slot) {
    feht_ctx::PendingRun &pr = c->pend[slot];
    if (!pr.active) return FEHT_OK;
    cudaEvent_t *ev = c->evr[slot];
    CU(c, cudaEventSynchronize(ev[4]));
    // ev: 0 | count | 1 | screen | 2 | A | 3 | B | 4 with (A, B) = (p-values, order + gather) for the p-value
    // order and (order, p-values fused with the gather) for the reference's |ratio| order
    float ms = 0, ms_screen = 0, ms_a = 0, ms_b = 0;
    CU(c, cudaEventElapsedTime(&ms, ev[0], ev[1])); c->tm.count_ms = ms;
    CU(c, cudaEventElapsedTime(&ms_screen, ev[1], ev[2]));
    CU(c, cudaEventElapsedTime(&ms_a, ev[2], ev[3]));
    CU(c, cudaEventElapsedTime(&ms_b, ev[3], ev[4]));
    c->tm.test_ms = ms_screen + (pr.fused ? ms_b : ms_a);
    c->tm.order_ms = pr.fused ? ms_a : ms_b;
    CU(c, cudaEventElapsedTime(&ms, ev[0], ev[4])); c->tm.total_ms = ms;
    c->tm.launches = pr.launches;
    c->tm.count_launches = pr.count_launches;
    c->tm.count_bytes = pr.count_bytes;
    c->acc[0] += 1; c->acc[1] += c->tm.count_ms; c->acc[2] += c->tm.test_ms; c->acc[3] += c->tm.order_ms;
    c->acc[4] += c->tm.total_ms; c->acc[5] += pr.launches; c->acc[6] += pr.count_launches;
    pr.active = false;
    return FEHT_OK;
}
// older run first, so that c->tm ends up describing the last one
int settle_all(feht_ctx *c) {
    if (int rc = settle(c, c->ev_slot ^ 1)) return rc;
    return settle(c, c->ev_slot);
}

int run_impl(feht_ctx *c, int correction, double ratio_filter, int sort_key, int64_t bonferroni_n, bool async) {
    if (!c) return FEHT_E_INVALID;
    if (int rc = use_device(c)) return rc;
    c->has_results = false;
    c->ev_slot ^= 1;
    if (int rc = settle(c, c->ev_slot)) return rc;   // the run before last: finished long ago
    cudaEvent_t *const ev = c->evr[c->ev_slot];
    if (c->S <= 0) return fail(c, FEHT_E_INVALID, "feht_set_samples must be called first");
    if (correction != FEHT_CORRECTION_NONE && correction != FEHT_CORRECTION_BONFERRONI)
        return fail(c, FEHT_E_INVALID, "unknown correction %d", correction);
    if (sort_key != FEHT_SORT_ABS_RATIO_DESC && sort_key != FEHT_SORT_PVALUE_ASC)
        return fail(c, FEHT_E_INVALID, "unknown sort key %d", sort_key);
    if (ratio_filter != ratio_filter) return fail(c, FEHT_E_INVALID, "ratio_filter is NaN");
    const int n_all = int(c->comps.size());
    // the comparisons this run emits: all of them, or the window of feht_set_comparison_window (local index i <->
    // comparison w0 + i; the pair-slots are counted for every comparison either way)
    int w0 = 0, n_comps = n_all;
    if (c->win_count >= 0) {
        w0 = int(std::min<int64_t>(c->win_first, n_all));
        n_comps = int(std::min<int64_t>(c->win_count, n_all - w0));
    }
    const int64_t R = marker_rows(c);
    if (n_comps > 65535) return fail(c, FEHT_E_INVALID, "more than 65535 comparisons in one run (use feht_set_comparison_window)");
    if (R > int64_t(0x7fffffff)) return fail(c, FEHT_E_INVALID, "more than 2^31-1 marker rows in one context");
    if (c->have_rank != 1 && c->row_base + R > int64_t(0x7fffffff))
        return fail(c, FEHT_E_INVALID, "row_base + rows exceeds 2^31-1: pass name_rank (the default rank is the global row index)");
    if (bonferroni_n <= 0) bonferroni_n = R;
    c->seg_off.assign(size_t(n_all) + 1, 0);
    c->n_surv = 0;
    if (n_comps == 0 || R == 0) {
        if (int rc = settle_all(c)) return rc;
        c->tm = Timings();
        c->has_results = true;
        return FEHT_OK;
    }
    c->seg_off.assign(size_t(n_comps) + 1, 0);   // offsets of the emitted comparisons; widened to all at the end

    // V.! bounds (Comparison.hs:123): every referenced column must exist in every row
    const int64_t row_len = std::min(c->min_row_len, c->S);
    for (const Comp &cp : c->comps)
        if (cp.max_col >= row_len)
            return fail(c, FEHT_E_BOUNDS, "index out of bounds: column %lld referenced, shortest row has %lld characters",
                        (long long)cp.max_col, (long long)row_len);

    // ---- pair-slots (what K1/K2 count) and screen groups (what K3a tests)
    const PopcPlan plan = make_popc_plan(c->W);
    const size_t mrow = size_t(plan.w4pad) * 2;  // 64-bit words per padded mask row
    std::vector<int> cat_slot0(c->cats.size(), -1), cat_tot(c->cats.size(), -1);
    int n_slots = 0;
    for (int i = 0; i < n_all; ++i) if (c->comps[size_t(i)].kind == 0) ++n_slots;
    const int n_explicit = n_slots;
    for (size_t k = 0; k < c->cats.size(); ++k) {
        cat_slot0[k] = n_slots;
        n_slots += (c->cats[k].n_values + 1) / 2;
    }
    // "whole category" slots (one-vs-rest = total - value) go last: the popcount engine does not scan them, it
    // adds up the category's value slots (the values partition the category's genomes)
    const int n_scanned = n_slots;
    for (size_t k = 0; k < c->cats.size(); ++k)
        if (c->cats[k].n_values > 2) cat_tot[k] = n_slots++;
    if (size_t(n_slots) * 2 * plan.w4pad * 16 > (size_t(1) << 31))
        return fail(c, FEHT_E_INVALID, "mask storage too large");
    // The masks and the screen plan only change with the comparisons (and S): a context that runs the same
    // comparisons again re-uses what is on the device (five small uploads less in front of every run).
    if (!c->plan_ready) {
    std::vector<uint64_t> hmask(size_t(n_slots) * 2 * mrow, 0);
    auto mask_row = [&](int slot, int which) { return hmask.data() + (size_t(slot) * 2 + which) * mrow; };
    std::vector<ScreenGroup> groups;
    std::vector<FracDef> fracs;
    std::vector<GroupComp> gcomps;
    std::vector<SlotDef> sdefs;
    int max_frac = 0, max_comp = 0, max_slot = 0;
    {
        // explicit comparisons: one pair-slot each, groups of up to 64
        constexpr int kExplicitPerGroup = 64;
        std::vector<int> cat_first(c->cats.size(), -1);  // first comparison index of each category
        int e = 0;
        for (int i = 0; i < n_all; ++i) {
            const Comp &cp = c->comps[size_t(i)];
            if (cp.kind != 0) {
                if (cat_first[size_t(cp.cat)] < 0) cat_first[size_t(cp.cat)] = i;
                continue;
            }
            uint64_t *m1 = mask_row(e, 0), *m2 = mask_row(e, 1);
            for (int32_t col : cp.g1) m1[col >> 6] |= 1ULL << (col & 63);
            for (int32_t col : cp.g2) m2[col >> 6] |= 1ULL << (col & 63);
            if (i < w0 || i >= w0 + n_comps) { ++e; continue; }   // counted, not emitted by this run
            if (groups.empty() || groups.back().n_comp >= kExplicitPerGroup || e % kExplicitPerGroup == 0)
                groups.push_back(ScreenGroup{int32_t(fracs.size()), 0, int32_t(gcomps.size()), 0,
                                             int32_t(sdefs.size()), 0, -1, 0});
            ScreenGroup &g = groups.back();
            fracs.push_back(FracDef{e, 0, -1, 0});
            fracs.push_back(FracDef{e, 1, -1, 0});
            sdefs.push_back(SlotDef{e, int16_t(g.n_frac), int16_t(g.n_frac + 1), -1, -1, 0});
            gcomps.push_back(GroupComp{i - w0, int16_t(g.n_frac), int16_t(g.n_frac + 1)});
            g.n_frac += 2;
            g.n_comp += 1;
            g.n_slot += 1;
            ++e;
        }
        // categories: value blocks of 64; group (a,b) holds the pairs between blocks a <= b, the
        // diagonal groups also hold the one-vs-rest comparisons of their block
        constexpr int kValueBlock = 64;
        for (size_t k = 0; k < c->cats.size(); ++k) {
            const Cat &cat = c->cats[k];
            const int V = cat.n_values;
            for (int64_t s = 0; s < c->S; ++s) {
                const int v = cat.value_of_sample[size_t(s)];
                if (v < 0) continue;
                mask_row(cat_slot0[k] + v / 2, v & 1)[s >> 6] |= 1ULL << (s & 63);
                if (cat_tot[k] >= 0) mask_row(cat_tot[k], 0)[s >> 6] |= 1ULL << (s & 63);
            }
            const int first = cat_first[k];
            // (local indices: comparison - w0; outside [0, n_comps) = not emitted by this run)
            auto pair_index = [&](int i, int j) { return first - w0 + i * V - i * (i + 1) / 2 + (j - i - 1); };
            const int ova_first = first - w0 + V * (V - 1) / 2;
            auto emitted = [&](int local) { return local >= 0 && local < n_comps; };
            const int nb = (V + kValueBlock - 1) / kValueBlock;
            for (int a = 0; a < nb; ++a)
                for (int b = a; b < nb; ++b) {
                    ScreenGroup g{int32_t(fracs.size()), 0, int32_t(gcomps.size()), 0, int32_t(sdefs.size()), 0,
                                  -1, 0};
                    const int a0 = a * kValueBlock, a1 = std::min(V, a0 + kValueBlock);
                    const int b0 = b * kValueBlock, b1 = std::min(V, b0 + kValueBlock);
                    auto plain = [&](int v) { return FracDef{cat_slot0[k] + v / 2, v & 1, -1, 0}; };
                    for (int v = a0; v < a1; ++v) fracs.push_back(plain(v));
                    g.n_frac = a1 - a0;
                    const int boff = (b == a) ? 0 : g.n_frac;   // fraction id of value b0
                    if (b != a) {
                        for (int v = b0; v < b1; ++v) fracs.push_back(plain(v));
                        g.n_frac += b1 - b0;
                    }
                    for (int i = a0; i < a1; ++i)
                        for (int j = std::max(b0, i + 1); j < b1; ++j) {
                            if (!emitted(pair_index(i, j))) continue;
                            gcomps.push_back(GroupComp{pair_index(i, j), int16_t(i - a0), int16_t(boff + j - b0)});
                            ++g.n_comp;
                        }
                    if (b == a && V > 2) {   // Comparison.hs:329: one-vs-rest only with more than two values
                        const int roff = g.n_frac;
                        for (int v = a0; v < a1; ++v) {
                            fracs.push_back(FracDef{cat_slot0[k] + v / 2, v & 1, cat_tot[k], 0});
                            if (!emitted(ova_first + v)) continue;
                            gcomps.push_back(GroupComp{ova_first + v, int16_t(v - a0), int16_t(roff + v - a0)});
                            ++g.n_comp;
                        }
                        g.n_frac += a1 - a0;
                    }
                    if (g.n_comp > 0) {
                        // slot view of the same fractions: one record load per pair-slot
                        g.tot_slot = (b == a && V > 2) ? cat_tot[k] : -1;
                        for (int f = 0; f < g.n_frac; ++f) {
                            const FracDef &fd = fracs[size_t(g.frac_off + f)];
                            SlotDef *sd = nullptr;
                            for (int q = g.slot_off; q < int(sdefs.size()); ++q)
                                if (sdefs[size_t(q)].slot == fd.slot) { sd = &sdefs[size_t(q)]; break; }
                            if (!sd) {
                                sdefs.push_back(SlotDef{fd.slot, -1, -1, -1, -1, 0});
                                sd = &sdefs.back();
                            }
                            int16_t &dst = fd.tot_slot >= 0 ? (fd.sel ? sd->r_hi : sd->r_lo) : (fd.sel ? sd->f_hi : sd->f_lo);
                            dst = int16_t(f);
                        }
                        g.n_slot = int(sdefs.size()) - g.slot_off;
                        groups.push_back(g);
                    } else {
                        fracs.resize(size_t(g.frac_off));
                    }
                }
        }
        for (const ScreenGroup &g : groups) {
            max_frac = std::max(max_frac, g.n_frac);
            max_comp = std::max(max_comp, g.n_comp);
            max_slot = std::max(max_slot, g.n_slot);
        }
    }
    CU(c, ensure(c->masks, size_t(n_slots) * 2 * plan.w4pad));
    CU(c, ensure(c->groups, groups.size()));
    CU(c, ensure(c->fracs, fracs.size()));
    CU(c, ensure(c->gcomps, gcomps.size()));
    CU(c, ensure(c->sdefs, sdefs.size()));
    CU(c, cudaMemcpyAsync(c->masks.p, hmask.data(), hmask.size() * 8, cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->groups.p, groups.data(), groups.size() * sizeof(ScreenGroup), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->fracs.p, fracs.data(), fracs.size() * sizeof(FracDef), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->gcomps.p, gcomps.data(), gcomps.size() * sizeof(GroupComp), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaMemcpyAsync(c->sdefs.p, sdefs.data(), sdefs.size() * sizeof(SlotDef), cudaMemcpyHostToDevice, c->st));
    CU(c, cudaStreamSynchronize(c->st));   // the host vectors die here (once per set of comparisons)
    // K3b flavour: do the hypergeometric batches keep their parameters in shared memory?  Yes when some comparison's
    // two groups hold fewer than 1000 samples together (every one of its tests is then an exact test, FET.hs:105-111)
    c->hyp_heavy = false;
    for (const Comp &cp : c->comps)
        if (cp.kind == 0 && cp.g1.size() + cp.g2.size() < 1000) c->hyp_heavy = true;
    for (const Cat &cat : c->cats) {
        std::vector<int64_t> cnt(static_cast<size_t>(cat.n_values), 0);
        for (int32_t v : cat.value_of_sample)
            if (v >= 0) ++cnt[size_t(v)];
        std::sort(cnt.begin(), cnt.end());
        if (cnt.size() >= 2 && cnt[0] + cnt[1] < 1000) c->hyp_heavy = true;   // the two smallest values make a pair
    }
    c->plan_groups = int(groups.size());
    c->plan_max_frac = max_frac; c->plan_max_comp = max_comp; c->plan_max_slot = max_slot;
    c->plan_ready = true;
    }
    CU(c, ensure(c->counts, size_t(n_slots) * size_t(R)));
    CU(c, ensure(c->seg_count, size_t(n_comps) * 2));
    if (ratio_filter > 0.0)   // survivor counters + cursor of the filtered screen (dense runs have neither)
        CU(c, cudaMemsetAsync(c->seg_count.p, 0, size_t(n_comps) * 2 * 8, c->st));

    int launches = 0, count_launches = 0;
    // ---- counting engine: explicit comparisons always go to K1; whole categories go to the int8
    // tcgen05 GEMM (K2) when asked for, or by default once there are enough value columns for the
    // tensor pipe to beat one masked popcount stream per value pair.
    const int n_cat_slots = n_slots - n_explicit;
    bool use_gemm = false;
    // FEHT_ENGINE_GEMM is the kind::i8 GEMM, FEHT_ENGINE_GEMM_FP4 the kind::mxf4 one; AUTO takes the faster (fp4)
    const bool fp4 = c->engine == FEHT_ENGINE_GEMM_FP4 || (c->engine == FEHT_ENGINE_AUTO && gemm_auto_fp4());
    bool f8 = c->engine == FEHT_ENGINE_GEMM_F8 || (c->engine == FEHT_ENGINE_AUTO && gemm_auto_f8());
    if (c->engine == FEHT_ENGINE_GEMM || c->engine == FEHT_ENGINE_GEMM_FP4 || c->engine == FEHT_ENGINE_GEMM_F8) {
        if (c->snp == 1) return fail(c, FEHT_E_INVALID, "the GEMM counting engine handles binary rows only");
        if (n_cat_slots == 0) return fail(c, FEHT_E_INVALID, "the GEMM counting engine needs a category (feht_add_category)");
        use_gemm = true;
    } else if (c->engine == FEHT_ENGINE_AUTO) {
        use_gemm = c->snp != 1 && (n_cat_slots >= kGemmAutoMinSlots ||
                                   (n_cat_slots >= kGemmAutoMinSlotsLarge && R >= kGemmAutoLargeRows));
    }
    if (use_gemm && f8 && !(c->gemm_ready && c->gemm_f8)) {
        // ---- e5m2 flavour: columns, pass descriptors and baked B blocks
        const std::vector<std::vector<GemmF8Column>> fp = f8_columns(c, cat_slot0, cat_tot);
        if (fp.empty()) {
            if (c->engine == FEHT_ENGINE_GEMM_F8)
                return fail(c, FEHT_E_INVALID, "a category needs more than %d GEMM columns (one per value and 4,095 of its "
                            "samples): use FEHT_ENGINE_GEMM_FP4", gemm_f8_max_cols());
            f8 = false;
        } else {
            const int n_passes = int(fp.size());
            const int n_stages = gemm_stage_count(c->S);
            std::vector<unsigned char> hpass(size_t(n_passes) * gemm_f8_pass_desc_bytes());
            std::vector<int> pad(static_cast<size_t>(n_passes));
            size_t total = 0;
            for (int p = 0; p < n_passes; ++p) {
                pad[size_t(p)] = ((int(fp[size_t(p)].size()) + 15) / 16) * 16;
                gemm_f8_fill_pass(hpass.data() + size_t(p) * gemm_f8_pass_desc_bytes(), fp[size_t(p)], pad[size_t(p)],
                                  int64_t(total));
                total += size_t(n_stages) * size_t(pad[size_t(p)]) * 128;
            }
            std::vector<uint8_t> hb(total);
            size_t off = 0;
            for (int p = 0; p < n_passes; ++p) {
                gemm_f8_bake_b(fp[size_t(p)], pad[size_t(p)], c->S, hb.data() + off);
                off += size_t(n_stages) * size_t(pad[size_t(p)]) * 128;
            }
            CU(c, ensure(c->gemm_b, total));
            CU(c, ensure(c->gemm_passes, hpass.size()));
            CU(c, cudaMemcpyAsync(c->gemm_b.p, hb.data(), total, cudaMemcpyHostToDevice, c->st));
            CU(c, cudaMemcpyAsync(c->gemm_passes.p, hpass.data(), hpass.size(), cudaMemcpyHostToDevice, c->st));
            CU(c, cudaStreamSynchronize(c->st));  // the host vectors die here
            c->gemm_n_passes = n_passes;
            c->gemm_f8 = true;
            c->gemm_ready = true;
        }
    }
    if (use_gemm && !f8 && (!c->gemm_ready || c->gemm_f8 || c->gemm_fp4 != fp4)) {
        const int per_pass = gemm_max_pairs_per_pass();
        const int n_passes = (n_cat_slots + per_pass - 1) / per_pass;
        const int n_stages = gemm_bblock_count(c->S, fp4);   // B blocks per pass
        // GEMM column gc = 2 * (slot - n_explicit) + sel  ->  list of samples with a 1
        std::vector<std::vector<int32_t>> cols(size_t(n_cat_slots) * 2);
        for (size_t k = 0; k < c->cats.size(); ++k) {
            const Cat &cat = c->cats[k];
            for (int64_t s = 0; s < c->S; ++s) {
                const int v = cat.value_of_sample[size_t(s)];
                if (v < 0) continue;
                cols[size_t(cat_slot0[k] - n_explicit + v / 2) * 2 + size_t(v & 1)].push_back(int32_t(s));
                if (cat_tot[k] >= 0) cols[size_t(cat_tot[k] - n_explicit) * 2].push_back(int32_t(s));
            }
        }
        std::vector<unsigned char> hpass(size_t(n_passes) * gemm_pass_desc_bytes());
        std::vector<int> pad(static_cast<size_t>(n_passes));
        size_t total = 0;
        for (int p = 0; p < n_passes; ++p) {
            const int pairs = std::min(per_pass, n_cat_slots - p * per_pass);
            pad[size_t(p)] = ((2 * pairs + 15) / 16) * 16;
            gemm_fill_pass(hpass.data() + size_t(p) * gemm_pass_desc_bytes(), n_explicit + p * per_pass, pairs,
                           pad[size_t(p)], int64_t(total));
            total += size_t(n_stages) * size_t(pad[size_t(p)]) * 128;
        }
        std::vector<uint8_t> hb(total);
        size_t off = 0;
        for (int p = 0; p < n_passes; ++p) {
            const int pairs = std::min(per_pass, n_cat_slots - p * per_pass);
            std::vector<std::vector<int32_t>> pc(cols.begin() + size_t(p) * per_pass * 2,
                                                 cols.begin() + size_t(p) * per_pass * 2 + size_t(pairs) * 2);
            gemm_bake_b(pc, pad[size_t(p)], c->S, fp4, hb.data() + off);
            off += size_t(n_stages) * size_t(pad[size_t(p)]) * 128;
        }
        CU(c, ensure(c->gemm_b, total));
        CU(c, ensure(c->gemm_passes, hpass.size()));
        CU(c, cudaMemcpyAsync(c->gemm_b.p, hb.data(), total, cudaMemcpyHostToDevice, c->st));
        CU(c, cudaMemcpyAsync(c->gemm_passes.p, hpass.data(), hpass.size(), cudaMemcpyHostToDevice, c->st));
        CU(c, cudaStreamSynchronize(c->st));  // the host vectors die here
        c->gemm_n_passes = n_passes;
        c->gemm_fp4 = fp4;
        c->gemm_f8 = false;
        c->gemm_ready = true;
    }
    c->last_engine = use_gemm ? (f8 ? FEHT_ENGINE_GEMM_F8 : (fp4 ? FEHT_ENGINE_GEMM_FP4 : FEHT_ENGINE_GEMM)) : FEHT_ENGINE_POPC;

    CU(c, cudaEventRecord(ev[0], c->st));
    // windows of one job (feht_set_comparison_window): the pair-slot counts do not depend on which comparisons are
    // emitted, so only the first window counts.  Never taken without a window: every ordinary run counts.
    const bool reuse_counts = c->win_count >= 0 && c->counts_ready && c->counts_engine == c->last_engine;
    c->counts_ready = false;
    const int n_popc_slots = use_gemm ? n_explicit : n_scanned;
    if (n_popc_slots > 0 && size_t(2) * plan.w4pad * 16 > size_t(200) * 1024)
        return fail(c, FEHT_E_INVALID, "%lld samples are too many for the popcount engine (a comparison's two masks must fit "
                    "in shared memory, about 800k samples); use feht_add_category with the GEMM engine", (long long)c->S);
    if (reuse_counts) {
        // nothing to launch
    } else if (c->snp == 1) {
        // one category of four values that covers every genome (config 3): its four masks partition the samples, and
        // the scan derives the last mask's counts from the row totals (count_popc.cu, PART)
        bool part = n_explicit == 0 && c->cats.size() == 1 && c->cats[0].n_values == 4 && n_popc_slots == 2;
        if (part)
            for (int32_t v : c->cats[0].value_of_sample)
                if (v < 0) { part = false; break; }
        CU(c, launch_count_popc_snp(c->plane[0], c->plane[1], c->plane[2], c->n_units, c->W, plan,
                                    c->masks.p, n_popc_slots, c->counts.p, c->num_sms, c->st, &count_launches, part));
    }
    else
        CU(c, launch_count_popc(c->plane[0], c->plane[1], c->n_units, c->W, plan, c->masks.p, n_popc_slots,
                                c->counts.p, c->num_sms, c->st, &count_launches));
    if (reuse_counts) {
    } else if (use_gemm && f8) {
        CU(c, launch_count_gemm_f8(c->plane[0], c->plane[1], c->n_units, c->W, c->S, c->gemm_b.p, c->gemm_passes.p,
                                   c->gemm_n_passes, c->counts.p, c->num_sms, c->st, &count_launches));
    } else if (use_gemm) {
        CU(c, launch_count_gemm(c->plane[0], c->plane[1], c->n_units, c->W, c->S, c->gemm_b.p,
                                c->gemm_passes.p, c->gemm_n_passes, c->gemm_fp4, c->counts.p, c->num_sms, c->st,
                                &count_launches));
    } else {
        for (size_t k = 0; k < c->cats.size(); ++k)
            if (cat_tot[k] >= 0) {
                CU(c, launch_sum_slots(c->counts.p, R, cat_slot0[k], (c->cats[k].n_values + 1) / 2, cat_tot[k], c->st));
                ++count_launches;
            }
    }
    launches += count_launches;
    c->counts_ready = c->win_count >= 0;
    c->counts_engine = c->last_engine;
    CU(c, cudaEventRecord(ev[1], c->st));

    // ---- K3 test + correct + filter
    TestParams tp{};
    tp.counts = c->counts.p;
    tp.groups = c->groups.p;
    tp.fracs = c->fracs.p;
    tp.gcomps = c->gcomps.p;
    tp.slots = c->sdefs.p;
    tp.max_slot = c->plan_max_slot;
    tp.n_groups = c->plan_groups;
    tp.max_frac = c->plan_max_frac;
    tp.max_comp = c->plan_max_comp;
    tp.n_rows = R;
    tp.n_comps = n_comps;
    tp.ratio_filter = ratio_filter;
    tp.name_rank = c->have_rank == 1 ? c->rank.p : nullptr;
    tp.row_base = c->row_base;
    tp.dense = ratio_filter <= 0.0 ? 1 : 0;
    tp.live_rows = nullptr;
    tp.live_count = nullptr;
    if (!tp.dense && R < (int64_t(1) << 31)) {
        // row lists of the prescreen kernel (FEHT_PRESCREEN=0: the screen kernel walks every tile and decides itself)
        static const bool off = [] { const char *e = getenv("FEHT_PRESCREEN"); return e && atoi(e) == 0; }();
        const size_t words = size_t(tp.n_groups) * size_t(R);
        if (!off && words <= (size_t(1) << 28)) {   // at most 1 GB of row indices
            CU(c, ensure(c->live_rows, words));
            CU(c, ensure(c->live_count, size_t(tp.n_groups)));
            tp.live_rows = c->live_rows.p;
            tp.live_count = c->live_count.p;
        }
    }
    if (tp.dense && c->have_rank == 1) {
        if (c->rank_is_perm == -1) {  // does name_rank enumerate [0, rows)?  (checked once per upload)
            DevBuf<uint32_t> flags;
            CU(c, ensure(flags, size_t((R + 31) / 32) + 2));
            unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(c->seg_count.p);
            CU(c, launch_count_distinct_ranks(c->rank.p, R, flags.p, d_cnt, c->st));
            unsigned long long distinct = 0;
            CU(c, cudaMemcpyAsync(&distinct, d_cnt, 8, cudaMemcpyDeviceToHost, c->st));
            CU(c, cudaStreamSynchronize(c->st));
            CU(c, cudaMemsetAsync(d_cnt, 0, 8, c->st));
            release(flags);
            c->rank_is_perm = distinct == (unsigned long long)R ? 1 : 0;
        }
        tp.dense_by_rank = c->rank_is_perm;
    }
    const bool rank_presorted = tp.dense && (c->have_rank != 1 || tp.dense_by_rank);
    // the reference's |ratio| order: K3a builds the sort keys, K4 sorts, K3b computes p row by row of the final
    // order and writes every output column (no separate gather); p-value order: K3b first, then K4 + gather
    const bool fused = sort_key == FEHT_SORT_ABS_RATIO_DESC;
    CU(c, ensure(c->sw_oa, 8));
    c->sw.or_and = c->sw_oa.p;
    auto ensure_survivors = [&](size_t nn) -> cudaError_t {
        cudaError_t e;
        if ((e = ensure(c->r_comp, nn)) != cudaSuccess || (e = ensure(c->r_rec, nn)) != cudaSuccess ||
            (e = ensure(c->sw_key_a, nn)) != cudaSuccess || (e = ensure(c->sw_key_b, nn)) != cudaSuccess ||
            (e = ensure(c->sw_val_a, nn)) != cudaSuccess || (e = ensure(c->sw_val_b, nn)) != cudaSuccess)
            return e;
        if (!fused && (e = ensure(c->r_p, nn)) != cudaSuccess) return e;
        tp.r_comp = c->r_comp.p; tp.rec = c->r_rec.p;
        c->sw.key_a = c->sw_key_a.p; c->sw.key_b = c->sw_key_b.p;
        c->sw.val_a = c->sw_val_a.p; c->sw.val_b = c->sw_val_b.p;
        tp.key = fused ? c->sw.key_a : nullptr;
        {   // FEHT_SORT_MERGE_COMP=0: the comparison index keeps its own radix pass
            static const bool off = [] { const char *e = getenv("FEHT_SORT_MERGE_COMP"); return e && atoi(e) == 0; }();
            tp.merge_comp = (fused && !off && n_comps >= 2 && n_comps <= 51) ? 1 : 0;
        }
        tp.val = c->sw.val_a;
        tp.or_and = c->sw.or_and;
        return fused ? launch_init_or_and(c->sw.or_and, c->st) : cudaSuccess;
    };
    if (tp.dense) {
        for (int i = 0; i <= n_comps; ++i) c->seg_off[size_t(i)] = int64_t(i) * R;
        if (c->seg_off[size_t(n_comps)] >= (int64_t(1) << 30))
            return fail(c, FEHT_E_INVALID, "2^30 or more surviving rows in one run");
        CU(c, ensure_survivors(size_t(c->seg_off[size_t(n_comps)])));
    } else {
        // K3a, filtered: one screen pass writes the survivors through a global cursor into arrays sized from the
        // last run (first run: 1/1024 of the tests, at least 1M slots); only if they did not fit is the pass
        // repeated with the exact size.  The host needs the counts anyway (result offsets, grid sizes).
        const double n_tests = double(R) * double(n_comps);
        int64_t cap = std::max<int64_t>(int64_t(1) << 20, std::min<int64_t>(int64_t(n_tests / 1024.0), int64_t(1) << 24));
        cap = std::max(cap, c->survivor_hint + c->survivor_hint / 4);
        std::vector<unsigned long long> cnt(static_cast<size_t>(n_comps) + 1);
        for (int attempt = 0;; ++attempt) {
            CU(c, ensure_survivors(size_t(cap)));
            tp.seg_count = c->seg_count.p;
            tp.cursor = c->seg_count.p + n_comps;
            tp.capacity = cap;
            CU(c, launch_test_emit(tp, c->st));
            ++launches;
            CU(c, cudaMemcpyAsync(cnt.data(), c->seg_count.p, (size_t(n_comps) + 1) * 8, cudaMemcpyDeviceToHost, c->st));
            CU(c, cudaStreamSynchronize(c->st));
            const unsigned long long total = cnt[size_t(n_comps)];
            if (total >= (1ull << 30)) return fail(c, FEHT_E_INVALID, "2^30 or more surviving rows in one run");
            if (int64_t(total) <= cap || attempt == 1) break;
            cap = int64_t(total);
            CU(c, cudaMemsetAsync(c->seg_count.p, 0, (size_t(n_comps) + 1) * 8, c->st));
        }
        for (int i = 0; i < n_comps; ++i) c->seg_off[size_t(i) + 1] = c->seg_off[size_t(i)] + int64_t(cnt[size_t(i)]);
        c->survivor_hint = c->seg_off[size_t(n_comps)];
    }
    const int64_t n = c->seg_off[size_t(n_comps)];
    c->n_surv = n;
    const size_t nn = size_t(n);
    CU(c, ensure(c->o_row, nn)); CU(c, ensure(c->o_rank, nn));
    CU(c, ensure(c->o_goa, nn)); CU(c, ensure(c->o_gob, nn)); CU(c, ensure(c->o_gta, nn)); CU(c, ensure(c->o_gtb, nn));
    CU(c, ensure(c->o_p, nn)); CU(c, ensure(c->o_ratio, nn));
    bool use_coop = false;
    const ResultSoA out{c->o_row.p, c->o_rank.p, c->o_goa.p, c->o_gob.p, c->o_gta.p, c->o_gtb.p, c->o_p.p, c->o_ratio.p};
    if (n > 0) {
        use_coop = sort_engine_allows_coop() && c->num_sms <= coop_sort_max_grid();
        if (use_coop) {
            if (!c->sw_coop.p) CU(c, ensure(c->sw_coop, coop_sort_control_words(c->num_sms)));
            c->sw.coop = c->sw_coop.p;
        } else {
            CU(c, ensure(c->sw_control, size_t(16) * 256 + 16 + size_t(16) * sort_status_words(n)));
        }
        c->sw.control = c->sw_control.p;
        if (tp.dense) {
            if (n_comps == 1 && c->comps[size_t(w0)].kind == 0) {
                int slot = 0;   // explicit comparisons take the pair-slots 0, 1, ... in the order they were added
                for (int i = 0; i < w0; ++i) if (c->comps[size_t(i)].kind == 0) ++slot;
                CU(c, launch_test_emit_single(tp, slot, c->st));
            }
            else
                CU(c, launch_test_emit(tp, c->st));
            ++launches;
        }
    }
    CU(c, cudaEventRecord(ev[2], c->st));
    uint32_t *perm = nullptr;
    auto order = [&]() -> int {
        if (use_coop) {
            // every pass in a single cooperative launch, nothing read back by the host
            CU(c, coop_sort_perm(c->sw, n, c->r_rec.p, c->r_comp.p, rank_presorted, c->num_sms, c->st, &launches));
        } else {
            CU(c, radix_sort_perm(c->sw, n, c->r_rec.p, c->r_comp.p, rank_presorted, c->st, &perm, &launches));
        }
        return FEHT_OK;
    };
    const bool hyp_heavy = c->hyp_heavy;   // decided with the comparison plan
    PvalueIo io{};
    io.rec = c->r_rec.p;
    io.r_comp = c->r_comp.p;
    if (n > 0 && hyp_heavy && dedup_enabled()) {
        // N4: one evaluation per distinct 2x2 table (FEHT_DEDUP=0 switches it off).  Table of 2n slots (a power of two,
        // at most 2^25 = 256 MB of keys); when it runs full a test simply computes its own p.
        size_t slots = 4096;
        while (slots < size_t(2) * size_t(n) && slots < (size_t(1) << 25)) slots <<= 1;
        CU(c, ensure(c->dd_keys, slots));
        CU(c, ensure(c->dd_owner, slots));
        CU(c, ensure(c->dd_list, size_t(n)));
        CU(c, ensure(c->dd_count, 1));
        CU(c, cudaMemsetAsync(c->dd_keys.p, 0, slots * 8, c->st));
        CU(c, cudaMemsetAsync(c->dd_count.p, 0, 4, c->st));
        io.dd_keys = c->dd_keys.p; io.dd_owner = c->dd_owner.p; io.dd_mask = uint32_t(slots - 1);
        io.dd_list = c->dd_list.p; io.dd_count = c->dd_count.p;
        ++launches;   // pvalue_fixup_kernel
    }
    if (n > 0 && fused) {
        if (int rc = order()) return rc;
        CU(c, cudaEventRecord(ev[3], c->st));
        // K3b fused with the gather: p (+ Bonferroni) of the survivors in their final order, all columns written
        io.perm_a = use_coop ? c->sw.val_a : perm;
        io.perm_b = use_coop ? c->sw.val_b : nullptr;
        io.d_sel = use_coop ? coop_sort_selector(c->sw, c->num_sms) : nullptr;
        io.out = out;
        io.row_base = c->row_base;
        CU(c, launch_pvalues(io, n, c->d_choose, correction, double(bonferroni_n), c->fet_mode, hyp_heavy, c->st));
        ++launches;
    } else if (n > 0) {
        // p-value order: K3b computes p and the keys, then K4 and the gather
        io.d_p = c->r_p.p;
        io.d_key = c->sw.key_a;
        io.d_val = c->sw.val_a;
        io.d_or_and = c->sw.or_and;
        CU(c, launch_pvalues(io, n, c->d_choose, correction, double(bonferroni_n), c->fet_mode, hyp_heavy, c->st));
        ++launches;
        CU(c, cudaEventRecord(ev[3], c->st));
        if (int rc = order()) return rc;
        if (use_coop)
            CU(c, launch_gather_results(c->sw.val_a, c->sw.val_b, coop_sort_selector(c->sw, c->num_sms), n, c->row_base,
                                        c->r_rec.p, c->r_p.p, out, c->st));
        else
            CU(c, launch_gather_results(perm, nullptr, nullptr, n, c->row_base, c->r_rec.p, c->r_p.p, out, c->st));
        ++launches;
    } else {
        CU(c, cudaEventRecord(ev[3], c->st));
    }
    CU(c, cudaEventRecord(ev[4], c->st));
    {
        feht_ctx::PendingRun &pr = c->pend[c->ev_slot];
        pr.active = true;
        pr.fused = fused;
        pr.launches = launches;
        pr.count_launches = count_launches;
        const int n_planes = c->snp == 1 ? 3 : 2;
        pr.count_bytes = double(c->n_units) * double((c->S + 63) / 64) * 8.0 * n_planes;
    }
    if (n_comps != n_all) {
        // result offsets for every comparison of the context: nothing before / after the window
        std::vector<int64_t> all(size_t(n_all) + 1, 0);
        for (int i = 0; i < n_all; ++i) {
            const int64_t rows_i = (i >= w0 && i < w0 + n_comps) ? c->seg_off[size_t(i - w0) + 1] - c->seg_off[size_t(i - w0)] : 0;
            all[size_t(i) + 1] = all[size_t(i)] + rows_i;
        }
        c->seg_off.swap(all);
    }
    if (!async)
        if (int rc = settle_all(c)) return rc;
    c->has_results = true;
    return FEHT_OK;
}
}  // namespace

extern "C" int feht_run(feht_ctx *c, int correction, double ratio_filter, int sort_key, int64_t bonferroni_n) {
    MULTI(c, multi::run(c->multi, correction, ratio_filter, sort_key, bonferroni_n));
    return run_impl(c, correction, ratio_filter, sort_key, bonferroni_n, false);
}

extern "C" int feht_run_async(feht_ctx *c, int correction, double ratio_filter, int sort_key, int64_t bonferroni_n) {
    MULTI(c, multi::run(c->multi, correction, ratio_filter, sort_key, bonferroni_n));   // (the merge needs every shard: blocking)
    return run_impl(c, correction, ratio_filter, sort_key, bonferroni_n, true);
}

extern "C" int64_t feht_result_count(const feht_ctx *c, int64_t comparison) {
    if (c && c->multi) return multi::result_count(c->multi, comparison);
    if (!c || !c->has_results || comparison < 0 || comparison + 1 >= int64_t(c->seg_off.size())) return FEHT_E_INVALID;
    return c->seg_off[size_t(comparison) + 1] - c->seg_off[size_t(comparison)];
}

extern "C" int64_t feht_result_total(const feht_ctx *c) {
    if (c && c->multi) return multi::result_total(c->multi);
    return (c && c->has_results) ? c->n_surv : FEHT_E_INVALID;
}

namespace {
// copy comparison c's ordered rows out of the result arrays (host or device destination)
int fetch_common(feht_ctx *c, int64_t comparison, int64_t *row, int32_t *goa, int32_t *gob, int32_t *gta,
                 int32_t *gtb, double *p, double *ratio, int32_t *name_rank, cudaMemcpyKind kind) {
    if (!c) return FEHT_E_INVALID;
    if (!c->has_results) return fail(c, FEHT_E_INVALID, "no results: call feht_run first");
    if (comparison < -1 || comparison + 1 >= int64_t(c->seg_off.size())) return fail(c, FEHT_E_INVALID, "comparison index");
    if (int rc = use_device(c)) return rc;
    // comparison == -1: the rows of every comparison, back to back in comparison order
    const int64_t o = comparison < 0 ? 0 : c->seg_off[size_t(comparison)];
    const size_t nn = comparison < 0 ? size_t(c->n_surv) : size_t(c->seg_off[size_t(comparison) + 1] - o);
    if (nn == 0) return FEHT_OK;
    if (row) CU(c, cudaMemcpyAsync(row, c->o_row.p + o, nn * 8, kind, c->st));
    if (goa) CU(c, cudaMemcpyAsync(goa, c->o_goa.p + o, nn * 4, kind, c->st));
    if (gob) CU(c, cudaMemcpyAsync(gob, c->o_gob.p + o, nn * 4, kind, c->st));
    if (gta) CU(c, cudaMemcpyAsync(gta, c->o_gta.p + o, nn * 4, kind, c->st));
    if (gtb) CU(c, cudaMemcpyAsync(gtb, c->o_gtb.p + o, nn * 4, kind, c->st));
    if (p) CU(c, cudaMemcpyAsync(p, c->o_p.p + o, nn * 8, kind, c->st));
    if (ratio) CU(c, cudaMemcpyAsync(ratio, c->o_ratio.p + o, nn * 8, kind, c->st));
    if (name_rank) CU(c, cudaMemcpyAsync(name_rank, c->o_rank.p + o, nn * 4, kind, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}
}  // namespace

extern "C" int feht_result_fetch(feht_ctx *c, int64_t comparison, int64_t *row, int32_t *goa, int32_t *gob,
                                 int32_t *gta, int32_t *gtb, double *p, double *ratio, int32_t *name_rank) {
    MULTI(c, multi::fetch(c->multi, comparison, row, goa, gob, gta, gtb, p, ratio, name_rank, false));
    return fetch_common(c, comparison, row, goa, gob, gta, gtb, p, ratio, name_rank, cudaMemcpyDeviceToHost);
}

extern "C" int feht_result_fetch_device(feht_ctx *c, int64_t comparison, int64_t *d_row, int32_t *d_goa,
                                        int32_t *d_gob, int32_t *d_gta, int32_t *d_gtb, double *d_p,
                                        double *d_ratio, int32_t *d_name_rank) {
    MULTI(c, multi::fetch(c->multi, comparison, d_row, d_goa, d_gob, d_gta, d_gtb, d_p, d_ratio, d_name_rank, true));
    return fetch_common(c, comparison, d_row, d_goa, d_gob, d_gta, d_gtb, d_p, d_ratio, d_name_rank,
                        cudaMemcpyDeviceToDevice);
}

extern "C" int feht_merge_sorted_device(feht_ctx *c, int n_shards, const int64_t *n, const double *const *d_key,
                                        const int32_t *const *d_name_rank, int sort_key, int64_t *const *d_dest) {
    if (!c) return FEHT_E_INVALID;
    if (n_shards < 0 || n_shards > merge_max_shards())
        return fail(c, FEHT_E_INVALID, "n_shards must be in [0, %d]", merge_max_shards());
    if (n_shards && (!n || !d_key || !d_name_rank || !d_dest)) return fail(c, FEHT_E_INVALID, "null shard arrays");
    if (sort_key != FEHT_SORT_ABS_RATIO_DESC && sort_key != FEHT_SORT_PVALUE_ASC)
        return fail(c, FEHT_E_INVALID, "unknown sort key %d", sort_key);
    for (int i = 0; i < n_shards; ++i)
        if (n[i] < 0 || (n[i] > 0 && (!d_key[i] || !d_name_rank[i] || !d_dest[i])))
            return fail(c, FEHT_E_INVALID, "shard %d: bad size or null pointer", i);
    if (int rc = use_device(c)) return rc;
    CU(c, launch_merge_rank(n_shards, n, d_key, d_name_rank, sort_key, d_dest, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}

extern "C" int feht_scatter_rows_device(feht_ctx *c, const void *d_src, void *d_dst, const int64_t *d_dest,
                                        int64_t n, int32_t elem_bytes) {
    if (!c) return FEHT_E_INVALID;
    if (n < 0 || elem_bytes <= 0 || (n && (!d_src || !d_dst || !d_dest))) return fail(c, FEHT_E_INVALID, "bad scatter arguments");
    if (int rc = use_device(c)) return rc;
    CU(c, launch_scatter_rows(d_src, d_dst, d_dest, n, elem_bytes, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    return FEHT_OK;
}

// K5b through the C ABI: the ordered shards of EVERY comparison of a row-sharded job, already on this context's GPU
// (one process per GPU: moved there by the host's transport), merged and scattered by one kernel launch.
extern "C" int feht_merge_shards_device(feht_ctx *c, int n_shards, int64_t n_comparisons, const feht_shard_rows *shards,
                                        int sort_key, const feht_shard_rows *out) {
    if (!c) return FEHT_E_INVALID;
    if (n_shards < 0 || n_shards > merge_max_shards())
        return fail(c, FEHT_E_INVALID, "n_shards must be in [0, %d]", merge_max_shards());
    if (n_comparisons < 0 || n_comparisons > 65535) return fail(c, FEHT_E_INVALID, "n_comparisons out of range");
    if (sort_key != FEHT_SORT_ABS_RATIO_DESC && sort_key != FEHT_SORT_PVALUE_ASC)
        return fail(c, FEHT_E_INVALID, "unknown sort key %d", sort_key);
    if (!out || !out->seg_off || (n_shards && !shards)) return fail(c, FEHT_E_INVALID, "null shard / output description");
    const size_t per = size_t(n_comparisons) + 1;
    std::vector<int64_t> hseg(per * size_t(n_shards + 1), 0);
    std::vector<MergeShardView> views(static_cast<size_t>(n_shards));
    int64_t total = 0;
    for (int i = 0; i < n_shards; ++i) {
        const feht_shard_rows &s = shards[i];
        if (!s.seg_off || s.seg_off[0] != 0) return fail(c, FEHT_E_INVALID, "shard %d: seg_off must start at 0", i);
        for (int64_t k = 0; k < n_comparisons; ++k)
            if (s.seg_off[k + 1] < s.seg_off[k]) return fail(c, FEHT_E_INVALID, "shard %d: seg_off must be non-decreasing", i);
        const int64_t n = s.seg_off[n_comparisons];
        if (n > 0 && (!s.row || !s.goa || !s.gob || !s.gta || !s.gtb || !s.p || !s.ratio || !s.name_rank))
            return fail(c, FEHT_E_INVALID, "shard %d: null column", i);
        std::memcpy(hseg.data() + per * size_t(i), s.seg_off, per * 8);
        for (int64_t k = 0; k < n_comparisons; ++k) hseg[per * size_t(n_shards) + size_t(k) + 1] += s.seg_off[k + 1] - s.seg_off[k];
        views[size_t(i)].res = ResultSoA{const_cast<int64_t *>(s.row), const_cast<int32_t *>(s.name_rank), const_cast<int32_t *>(s.goa),
                                         const_cast<int32_t *>(s.gob), const_cast<int32_t *>(s.gta), const_cast<int32_t *>(s.gtb),
                                         const_cast<double *>(s.p), const_cast<double *>(s.ratio)};
        views[size_t(i)].n = n;
        total += n;
    }
    for (int64_t k = 0; k < n_comparisons; ++k) hseg[per * size_t(n_shards) + size_t(k) + 1] += hseg[per * size_t(n_shards) + size_t(k)];
    std::memcpy(const_cast<int64_t *>(out->seg_off), hseg.data() + per * size_t(n_shards), per * 8);
    if (total == 0) return FEHT_OK;
    if (!out->row || !out->goa || !out->gob || !out->gta || !out->gtb || !out->p || !out->ratio || !out->name_rank)
        return fail(c, FEHT_E_INVALID, "null output column");
    if (int rc = use_device(c)) return rc;
    DevBuf<int64_t> dseg;
    CU(c, ensure(dseg, hseg.size()));
    CU(c, ensure(c->merge_dest, size_t(total) + merge_bounds_words(total, n_shards)));   // positions, then block bounds
    CU(c, cudaMemcpyAsync(dseg.p, hseg.data(), hseg.size() * 8, cudaMemcpyHostToDevice, c->st));
    for (int i = 0; i < n_shards; ++i) views[size_t(i)].d_seg_off = dseg.p + per * size_t(i);
    const ResultSoA o{const_cast<int64_t *>(out->row), const_cast<int32_t *>(out->name_rank), const_cast<int32_t *>(out->goa),
                      const_cast<int32_t *>(out->gob), const_cast<int32_t *>(out->gta), const_cast<int32_t *>(out->gtb),
                      const_cast<double *>(out->p), const_cast<double *>(out->ratio)};
    cudaError_t e = launch_merge_all(n_shards, int(n_comparisons), sort_key, views.data(), dseg.p + per * size_t(n_shards), o, 0,
                                     c->merge_dest.p, c->merge_dest.p + total, c->st);
    cudaError_t e2 = cudaStreamSynchronize(c->st);
    release(dseg);
    CU(c, e);
    CU(c, e2);
    return FEHT_OK;
}

extern "C" int feht_last_timings(const feht_ctx *cc, double *out, int n) {
    if (!cc || !out) return FEHT_E_INVALID;
    feht_ctx *c = const_cast<feht_ctx *>(cc);
    MULTI(c, multi::last_timings(c->multi, out, n));
    if (int rc = use_device(c)) return rc;
    if (int rc = settle_all(c)) return rc;
    const double v[8] = {c->tm.count_ms, c->tm.test_ms, c->tm.order_ms, c->tm.total_ms,
                         c->tm.launches, c->tm.count_launches, c->tm.count_bytes, double(c->last_engine)};
    for (int i = 0; i < n && i < 8; ++i) out[i] = v[i];
    return FEHT_OK;
}

extern "C" int feht_timings_accumulated(feht_ctx *c, double *out, int n, int reset) {
    if (!c || !out) return FEHT_E_INVALID;
    MULTI(c, multi::timings_accumulated(c->multi, out, n, reset));
    if (int rc = use_device(c)) return rc;
    if (int rc = settle_all(c)) return rc;
    for (int i = 0; i < n && i < 7; ++i) out[i] = c->acc[i];
    if (reset)
        for (double &a : c->acc) a = 0;
    return FEHT_OK;
}

// ==================================================================================== merge ==
extern "C" int feht_merge_sorted(int n_shards, const int64_t *n, const double *const *key,
                                 const int32_t *const *name_rank, int sort_key, int32_t *out_shard,
                                 int64_t *out_index) {
    if (n_shards < 0 || (n_shards && (!n || !key || !name_rank))) return FEHT_E_INVALID;
    struct Item { uint64_t k; int32_t rank; int32_t shard; int64_t idx; };
    auto key_bits = [&](double v) -> uint64_t {
        uint64_t b;
        if (sort_key == FEHT_SORT_PVALUE_ASC) { std::memcpy(&b, &v, 8); return b; }
        const double a = v < 0 ? -v : v;
        std::memcpy(&b, &a, 8);
        b &= 0x7fffffffffffffffULL;
        return ~b;
    };
    auto later = [](const Item &x, const Item &y) {
        if (x.k != y.k) return x.k > y.k;
        if (x.rank != y.rank) return uint32_t(x.rank) > uint32_t(y.rank);
        return x.shard > y.shard;
    };
    std::priority_queue<Item, std::vector<Item>, decltype(later)> heap(later);
    for (int s = 0; s < n_shards; ++s)
        if (n[s] > 0) heap.push(Item{key_bits(key[s][0]), name_rank[s][0], s, 0});
    int64_t o = 0;
    while (!heap.empty()) {
        Item it = heap.top();
        heap.pop();
        if (out_shard) out_shard[o] = it.shard;
        if (out_index) out_index[o] = it.idx;
        ++o;
        const int64_t nx = it.idx + 1;
        if (nx < n[it.shard])
            heap.push(Item{key_bits(key[it.shard][nx]), name_rank[it.shard][nx], it.shard, nx});
    }
    return FEHT_OK;
}

// ============================================================================== test hooks ==
extern "C" int64_t feht_choose_table(double *out, int64_t cap) {
    const std::vector<double> &t = host_choose_table();
    if (out && cap > 0) std::memcpy(out, t.data(), size_t(std::min<int64_t>(cap, int64_t(t.size()))) * 8);
    return int64_t(t.size());
}

extern "C" int feht_eval_tables(feht_ctx *c, const int32_t *goa, const int32_t *gob, const int32_t *gta,
                                const int32_t *gtb, int64_t n, double *p_out, double *ratio_out) {
    if (!c) return FEHT_E_INVALID;
    MULTI(c, multi::eval_tables(c->multi, goa, gob, gta, gtb, n, p_out, ratio_out));
    if (n < 0 || (n && (!goa || !gob || !gta || !gtb))) return fail(c, FEHT_E_INVALID, "null table arrays");
    if (n == 0) return FEHT_OK;
    if (int rc = use_device(c)) return rc;
    std::vector<int4> h(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        if (goa[i] < 0 || gob[i] < 0 || gta[i] < 0 || gtb[i] < 0)
            return fail(c, FEHT_E_INVALID, "negative count (the device path only sees popcounts)");
        h[size_t(i)] = make_int4(goa[i], gob[i], gta[i], gtb[i]);
    }
    DevBuf<int4> dt; DevBuf<double> dp, dr; DevBuf<SurvivorRec> drec;
    CU(c, ensure(dt, size_t(n))); CU(c, ensure(dp, size_t(n))); CU(c, ensure(dr, size_t(n))); CU(c, ensure(drec, size_t(n)));
    CU(c, cudaMemcpyAsync(dt.p, h.data(), size_t(n) * 16, cudaMemcpyHostToDevice, c->st));
    CU(c, launch_eval_tables(dt.p, n, c->d_choose, c->fet_mode, dp.p, dr.p, drec.p, c->st));
    if (p_out) CU(c, cudaMemcpyAsync(p_out, dp.p, size_t(n) * 8, cudaMemcpyDeviceToHost, c->st));
    if (ratio_out) CU(c, cudaMemcpyAsync(ratio_out, dr.p, size_t(n) * 8, cudaMemcpyDeviceToHost, c->st));
    CU(c, cudaStreamSynchronize(c->st));
    release(dt); release(dp); release(dr); release(drec);
    return FEHT_OK;
}
