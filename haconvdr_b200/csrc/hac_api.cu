// C ABI of the engine (include/hac_index.h): shard storage, add / reset, the search driver that
// chains scan -> refresh -> rescore -> select, and the merge / gather entry points.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hac_index.h"
#include "hac_common.cuh"
#include "hac_kernels.cuh"

using namespace hac;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return (e == cudaErrorMemoryAllocation) ? HAC_E_NOMEM : HAC_E_CUDA;
}
#define CU(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

constexpr int64_t kRowAlign = 256;          // scan tile width: segment capacities and chunk edges
constexpr int kMaxQueryBatch = 16384;
constexpr int kMaxChunks = 64;                  // corpus chunks (threshold refresh points) of one search
constexpr int kMaxEvents = 2 + 2 * kMaxChunks;  // timing events: search begin / end + one pair per scan launch

int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct Segment {
    float* rows = nullptr;        // [cap_rows][d] fp32
    uint8_t* shadow = nullptr;    // f16 tiled image of the same rows (allocated on first use when lazy)
    int64_t f16_rows = 0;         // rows [0, f16_rows) of the segment are present in the f16 image
    uint8_t* shadow8 = nullptr;   // int8 tiled image (d % 128 == 0 only)
    TileQ8* tiles8 = nullptr;     // per-128-row-tile constants of the int8 image
    OperandStats* stats = nullptr;
    int64_t cap_rows = 0;
    int64_t n_rows = 0;
    uint32_t base = 0;            // index-local row of rows[0]
    bool closed = false;          // no further appends (a larger segment follows)
};

struct Workspace {
    int nq_pad = 0;
    uint32_t cap = 0;
    int d = 0;
    float* q = nullptr;           // device copy of host queries
    uint8_t* q_shadow = nullptr;
    uint8_t* q_shadow8 = nullptr;
    QueryQ8* q_consts = nullptr;
    // int8 path: two log shortlists the scans append to (chunk i -> lg[i & 1]); the side-stream worker rescores a
    // log exactly and moves what can still reach the top-k into `cb`
    CandBuf lg[2] = {};
    int log_nq_pad = 0;
    uint32_t log_cap = 0;
    OperandStats* q_stats = nullptr;
    float *q_norm = nullptr, *q_err = nullptr, *margin = nullptr, *tau = nullptr, *thr = nullptr;
    float* q_shift = nullptr;     // q . centre per query (f16 path)
    float* scalars = nullptr;     // [0] absmax scratch, [1] margin_max, [2] screen_err_max
    unsigned long long* counters = nullptr;  // [0] emitted, [1] rescored
    CandBuf cb{};
    float* D = nullptr;
    int64_t* I = nullptr;
    int64_t out_elems = 0;
    void* host_pinned = nullptr;  // 64 B: overflow flag + statistics read-back
};

}  // namespace

namespace hac {
// hac_shards.cu reports its failures through the same per-thread message as every other entry point
int set_last_error(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace hac

struct hac_index {
    int d = 0, device = 0, sm_count = 148;
    // stream: the handle's own stream (host API) and the stream every int8 scan runs on - highest priority, so that
    // the persistent scan CTAs are placed before the side stream's worker CTAs.  side: lowest priority; rescore /
    // refresh / final select of the pipelined int8 search, co-resident with the scan of the following chunk.
    cudaStream_t stream = nullptr, side = nullptr;
    cudaEvent_t wdone[kMaxChunks] = {};     // worker (rescore + refresh) of chunk i finished; no timing
    cudaEvent_t ev_in = nullptr, ev_out = nullptr, ev_join = nullptr;
    std::vector<Segment> segs;
    int64_t ntotal = 0;
    int64_t id_base = 0;
    int64_t* id_table = nullptr;
    int64_t id_table_n = 0;
    OperandStats* corpus_stats = nullptr;   // index-wide maxima (device)
    float* add_scratch = nullptr;           // absmax scratch for add
    // screen centre (hac_prep.cu): [d] column means of the first rows added + [1] its norm; fixed until reset
    float* center = nullptr;
    double* center_accum = nullptr;
    bool center_enabled = true, center_valid = false;
    // unit order of the tensor-core scans (hac_scan_mma.cu UnitSchedule): -1 = per path (tile-major for the int8
    // screen, striped for the f16 one: measured 57.2 vs 61.1 ms and 83.1 vs 75.2 ms per search), 0 / 1 = forced
    int scan_tile_major = -1;
    int i8_cta_group = 2;                   // int8 screen: 2 = CTA pairs (a third less operand traffic from L2): 50.3 vs 56.9 ms
    // cross-shard threshold exchange (hac_set_threshold_exchange): buffers, and the epoch of the next searches (0 = off)
    ThrExchange exchange{};
    int64_t exchange_capacity = 0;
    int64_t exchange_epoch = 0;
    int cur_batch = 0;                      // index of the query batch inside the current search call
    int sticky_level = 0;                   // searches start in careful mode once the fast mode overflowed (until reset)
    Workspace ws;
    hac_stats stats{};
    cudaEvent_t ev[kMaxEvents] = {};
    bool events_ready = false;
    bool mma_configured = false;
    int mma_cta_group = 1;                  // 2 = CTA-pair variant (measured equal within noise; HAC_MMA_CTA_GROUP / hac_set_option)
    double chunk_growth = 4.0;              // chunk i+1 = growth * rows seen so far
    // f16 image: low mantissa bits forced to zero (corpus / queries).  The scan is power-limited and the tensor
    // core draws less with sparser mantissas: 3 dropped corpus bits (an 8-bit significand) run the scan 5 %
    // faster (76.3 -> 72.0 ms measured); the margin grows 0.76 -> 2.3 and 1.6x more rows are rescored (0.2 ms).
    int drop_bits_x = 3, drop_bits_q = 0;
    // keep an int8 image of the corpus too (rows*d bytes of HBM; HAC_PATH_I8).  On by default when d % 128 == 0: its
    // screen is the fastest path for k <= 128 at every batch size (50.3 vs 75.1 ms per search at 25.7M x 2514, k=100;
    // 3.0 vs 5.5 ms at one query); "build_i8" = 0 / HAC_BUILD_I8=0 saves the memory.
    bool build_i8 = true;
    bool i8_overflowed = false;             // the int8 screen overflowed on this corpus: AUTO stops choosing it (until reset)
    double i8_chunk_growth = 0.0;           // int8 chunk schedule, synchronous chunks: chunk = growth * rows seen so far (0 = by batch size and k)
    // Pipelined int8 search (option "i8_pipeline").  After a short synchronous prelude (thresholds must exist before chunks can be
    // large) the scans run back to back on `stream` while chunk i's rescore + refresh run on `side` beside the scan of
    // chunk i+1: scan i only waits for the worker of chunk i - i8_pipe_dist.  A stale threshold is a valid lower
    // bound, so nothing but the number of emitted rows depends on the overlap.
    // measured on one GPU: 47.1 ms pipelined vs 46.2 ms synchronous (the scan is power-bound: co-running rescores slow
    // it by what they save).  Giving the workers SMs of their own does not help either: the scan slows in proportion
    // to the SMs it gives up (132 of 148 SMs: 45.4 vs 40.8 ms of scan, profiles/r02_ab_warm_start_and_scan_sms.jsonl).
    bool i8_pipeline = false;
    int i8_pipe_dist = 2;                   // 1 = every scan waits for the previous chunk's worker (no overlap)
    // Warm start of the int8 search: an f16 image of the FIRST rows of the first segment (a few hundred MB).  A search
    // first runs the f16 screen over those rows - its margin is ~20x tighter, so it finds their exact top-k with ~10x
    // fewer rescored pairs than the int8 screen's loose early chunks (which emit k * e^(m8*z/sigma) pairs per query for
    // every e-fold of rows seen, whatever the chunk size) - and the int8 scan of the remaining rows starts with an
    // exact threshold.  -1 = automatic size, 0 = off, > 0 = rows.
    int64_t i8_warm_rows = -1;
    struct WarmSlab {
        uint8_t* shadow = nullptr;
        OperandStats* stats = nullptr;
        const float* src = nullptr;         // rows pointer of the segment the image was built from
        int64_t cap_rows = 0, rows = 0;     // rows [0, rows) are present
    } warm;
    double i8_pipe_growth = 0.125;          // pipelined chunks: max(i8_pipe_min_rows, growth * rows seen so far)
    int64_t i8_pipe_min_rows = 0;           // 0 = by batch size (about 75 us of scan per chunk)
    // the f16 image (rows*d*2 bytes) is only read by the f16 screen (k > i8_auto_max_k, int8 overflow fallback, forced
    // HAC_PATH_MMA): with the int8 image present it is built on first use instead of on add (25.7M rows: 138 -> 99 GB)
    int lazy_f16 = -1;                      // -1 = lazy exactly when the int8 image is built; 0 / 1 = forced
    // HAC_PATH_AUTO takes the int8 screen up to this k.  Its shortlist grows with k * e^(m8*z/sigma): with a warm slab
    // sized by k it still wins at k = 1000 on a large shard (70.7 vs 75.5 ms at 25.7M x 2514, 49.1 vs 72.4 ms at k = 250)
    // and the 39 GB f16 image of the corpus is never built ...
    int i8_auto_max_k = HAC_MAX_K;
    // ... for k > 128 only on shards of at least this many rows per k.  The int8 scan saves (rows - slab) x one int8
    // row of tensor time, the price is ~k * A * ln(rows / slab) more rescored rows per query: break-even ~10^4 rows per
    // k (k = 1000: 25.7M rows 70.7 vs 75.5 ms, 12.9M rows 37.7 vs ~37.6 ms, 3.2M rows 18.0 vs 17.1 ms).  Batches below
    // 128 queries run no warm slab and stream at the HBM rate: they need 256 rows per k and query.
    int64_t i8_large_k_rows_per_k = 12288;
    int default_path = HAC_PATH_MMA;        // what HAC_PATH_AUTO resolves to
    int i8_auto_max_queries = kMaxQueryBatch;          // HAC_PATH_AUTO takes the int8 screen up to this batch size when the image exists
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

__global__ void merge_stats_kernel(OperandStats* dst, const OperandStats* src) {
    dst->absmax = fmaxf(dst->absmax, src->absmax);
    dst->norm_max = fmaxf(dst->norm_max, src->norm_max);
    dst->hat_norm_max = fmaxf(dst->hat_norm_max, src->hat_norm_max);
    dst->err_norm_max = fmaxf(dst->err_norm_max, src->err_norm_max);
    dst->i8_beta_max = fmaxf(dst->i8_beta_max, src->i8_beta_max);
    dst->i8_gamma_max = fmaxf(dst->i8_gamma_max, src->i8_gamma_max);
}

void free_segment(Segment& s) {
    if (s.rows) cudaFree(s.rows);
    if (s.shadow) cudaFree(s.shadow);
    if (s.shadow8) cudaFree(s.shadow8);
    if (s.tiles8) cudaFree(s.tiles8);
    if (s.stats) cudaFree(s.stats);
    s = Segment{};
}

bool f16_is_lazy(const hac_index* idx) {
    return idx->lazy_f16 < 0 ? (idx->build_i8 && idx->d % kBlockK8 == 0) : idx->lazy_f16 != 0;
}

int alloc_segment(hac_index* idx, int64_t cap_rows, Segment* out) {
    Segment s;
    s.cap_rows = round_up(std::max<int64_t>(cap_rows, kRowAlign), kRowAlign);
    cudaError_t e = cudaMalloc(&s.rows, (size_t)s.cap_rows * idx->d * sizeof(float));
    if (e == cudaSuccess && !f16_is_lazy(idx)) e = cudaMalloc(&s.shadow, (size_t)shadow_bytes(s.cap_rows, idx->d));
    if (e == cudaSuccess) e = cudaMalloc(&s.stats, sizeof(OperandStats));
    if (e == cudaSuccess && idx->d % kBlockK8 == 0 && idx->build_i8) {
        e = cudaMalloc(&s.shadow8, (size_t)shadow8_bytes(s.cap_rows, idx->d));
        if (e == cudaSuccess) e = cudaMalloc(&s.tiles8, (size_t)shadow_tiles(s.cap_rows) * sizeof(TileQ8));
    }
    if (e != cudaSuccess) {
        free_segment(s);
        return fail_cuda(e, "segment allocation");
    }
    if (s.shadow) cudaMemsetAsync(s.shadow, 0, (size_t)shadow_bytes(s.cap_rows, idx->d), idx->stream);
    cudaMemsetAsync(s.stats, 0, sizeof(OperandStats), idx->stream);
    *out = s;
    return HAC_OK;
}

// returns the segment the next rows go to, allocating when the last one is full
int writable_segment(hac_index* idx, int64_t want_rows, Segment** out) {
    if (!idx->segs.empty()) {
        Segment& last = idx->segs.back();
        if (!last.closed && last.n_rows < last.cap_rows) {
            *out = &last;
            return HAC_OK;
        }
    }
    if ((int)idx->segs.size() >= kMaxSegments) return fail(HAC_E_STATE, "too many segments; call hac_reserve first");
    Segment s;
    const int64_t cap = std::max<int64_t>(want_rows, idx->ntotal);   // at least doubles the shard
    int rc = alloc_segment(idx, cap, &s);
    if (rc != HAC_OK) return rc;
    s.base = (uint32_t)idx->ntotal;
    idx->segs.push_back(s);
    *out = &idx->segs.back();
    return HAC_OK;
}

// rows [seg->f16_rows, end) of the segment -> f16 image (+ the norms of the f16 screen margin); seg->shadow exists
void convert_f16_rows(hac_index* idx, Segment* seg, int64_t end, cudaStream_t s) {
    const int d = idx->d;
    const int64_t r0 = seg->f16_rows, m = end - r0;
    if (m <= 0) return;
    const float* src = seg->rows + (size_t)r0 * d;
    launch_absmax(src, m * d, idx->add_scratch, s);
    launch_pick_scale(seg->stats, idx->add_scratch, /*keep_scale=*/r0 > 0 ? 1 : 0, s);
    const int64_t n_pad = std::min(round_up(end, kRowAlign), seg->cap_rows) - r0;
    launch_convert_rows(src, m, n_pad, d, seg->shadow, r0, seg->stats, nullptr, nullptr, idx->drop_bits_x,
                        idx->center_valid ? idx->center : nullptr, s);
    seg->f16_rows = end;
}

// rows the int8 search scans with the f16 screen first (multiple of kRowAlign; 0 = no warm start)
int64_t pick_warm_rows(const hac_index* idx, int nq, int k) {
    if (idx->segs.empty() || idx->i8_warm_rows == 0) return 0;
    const int64_t have = idx->segs[0].n_rows / kRowAlign * kRowAlign;
    int64_t want = idx->i8_warm_rows;
    if (want < 0) {
        // tensor-bound batches only: an f16 row costs twice an int8 row, every e-fold of warm rows saves one e-fold of
        // loosely filtered int8 emission (~k * A rescored rows per query, A = e^(m8*z/sigma) ~ 6-9).  The two marginal
        // costs meet at ~5000 * k rows whatever the corpus or batch size, and the optimum is flat around it.
        if (nq < 128) return 0;                               // Q = 128: 4.04 vs 4.15 ms with it, Q <= 32: slower (HBM-bound, an f16 row is twice the bytes)
        if (k <= 128) {
            // measured at 25.7M x 2514 (profiles/r02_ab_warm_start_and_scan_sms.jsonl): 128k rows 46.2 ms, 256k 44.4, 384k 44.8,
            // 768k 44.0 against 48.0 without; rescored pairs 19.9M -> 9.6M
            want = std::min<int64_t>(786432, std::max<int64_t>(32768, idx->ntotal / 32));
            if (idx->ntotal < 8 * want) return 0;
        } else {
            // k = 250: 49.1 ms with 768k rows, 50.8 with 1.5M, 51.1 with 3M (f16 screen alone: 72.4); k = 1000: 77.8 / 73.9 /
            // 70.7 (f16: 75.5; without a warm start 117.5) - profiles/r02_ab_k_large.jsonl
            want = std::min<int64_t>((int64_t)6144 * k, idx->ntotal / 4);
            if (want < 32768) return 0;
        }
    }
    want = std::min(want / kRowAlign * kRowAlign, have);
    return want >= 4096 ? want : 0;
}

void free_warm(hac_index* idx) {
    if (idx->warm.shadow) cudaFree(idx->warm.shadow);
    if (idx->warm.stats) cudaFree(idx->warm.stats);
    idx->warm = hac_index::WarmSlab{};
}

// f16 image of rows [0, want) of the first segment (built once per corpus, extended when `want` grows)
int ensure_warm(hac_index* idx, int64_t want, cudaStream_t s) {
    auto& wm = idx->warm;
    Segment& s0 = idx->segs[0];
    if (wm.src != s0.rows) wm.rows = 0;                      // another corpus (reset / reload): rebuild
    if (wm.cap_rows < want) {
        if (wm.shadow) cudaFree(wm.shadow);
        wm.shadow = nullptr;
        wm.cap_rows = 0;
        wm.rows = 0;
        const int64_t cap = round_up(want, kRowAlign);
        cudaError_t e = cudaMalloc(&wm.shadow, (size_t)shadow_bytes(cap, idx->d));
        if (e != cudaSuccess) return fail_cuda(e, "warm-start image allocation");
        wm.cap_rows = cap;
    }
    if (wm.stats == nullptr) CU(cudaMalloc(&wm.stats, sizeof(OperandStats)));
    if (wm.rows == 0) CU(cudaMemsetAsync(wm.stats, 0, sizeof(OperandStats), s));
    if (wm.rows < want) {
        Segment view;                                        // the slab seen as a segment of its own
        view.rows = s0.rows;
        view.shadow = wm.shadow;
        view.stats = wm.stats;
        view.f16_rows = wm.rows;
        view.cap_rows = wm.cap_rows;
        view.n_rows = want;
        convert_f16_rows(idx, &view, want, s);
        wm.rows = want;
        wm.src = s0.rows;
        CU(cudaGetLastError());
    }
    return HAC_OK;
}

// the f16 screen is about to run: allocate / complete the f16 image of every segment (no-op when it is up to date)
int ensure_f16_image(hac_index* idx, cudaStream_t s) {
    bool touched = false;
    for (auto& seg : idx->segs) {
        if (seg.shadow != nullptr && seg.f16_rows == seg.n_rows) continue;
        if (seg.shadow == nullptr) {
            cudaError_t e = cudaMalloc(&seg.shadow, (size_t)shadow_bytes(seg.cap_rows, idx->d));
            if (e != cudaSuccess) return fail_cuda(e, "f16 image allocation");
            cudaMemsetAsync(seg.shadow, 0, (size_t)shadow_bytes(seg.cap_rows, idx->d), s);
            seg.f16_rows = 0;
        }
        convert_f16_rows(idx, &seg, seg.n_rows, s);
        merge_stats_kernel<<<1, 1, 0, s>>>(idx->corpus_stats, seg.stats);
        touched = true;
    }
    if (touched) CU(cudaGetLastError());
    return HAC_OK;
}

enum class RowSource { Host, Device, Synthetic };

int add_rows(hac_index* idx, int64_t n, const float* src, RowSource kind, cudaStream_t user_stream, uint64_t seed,
             int64_t row0_global, int dist) {
    if (n < 0) return fail(HAC_E_INVALID, "add: negative row count");
    if (n == 0) return HAC_OK;
    if (kind != RowSource::Synthetic && src == nullptr) return fail(HAC_E_INVALID, "add: null row pointer");
    if (idx->ntotal + n > 0xFFFFFF00ll) return fail(HAC_E_INVALID, "add: shard would exceed 2^32 rows");
    if (idx->id_table != nullptr && idx->ntotal + n > idx->id_table_n) {
        // a stale table from a previous block must not silently translate new rows
        cudaFree(idx->id_table);
        idx->id_table = nullptr;
        idx->id_table_n = 0;
    }
    DeviceGuard guard(idx->device);
    cudaStream_t s = idx->stream;
    if (kind == RowSource::Device && user_stream != s) {
        // rows were produced on the caller's stream: order our stream after it
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        cudaEventRecord(e, user_stream);
        cudaStreamWaitEvent(s, e, 0);
        cudaEventDestroy(e);
    }
    const int d = idx->d;
    int64_t done = 0;
    while (done < n) {
        Segment* seg = nullptr;
        int rc = writable_segment(idx, n - done, &seg);
        if (rc != HAC_OK) return rc;
        const int64_t m = std::min(n - done, seg->cap_rows - seg->n_rows);
        float* dst = seg->rows + (size_t)seg->n_rows * d;
        if (kind == RowSource::Host) {
            // chunked so that the convert kernel of one chunk overlaps the copy of the next
            const int64_t chunk_rows = std::max<int64_t>(1, (64ll << 20) / (d * (int64_t)sizeof(float)));
            for (int64_t o = 0; o < m; o += chunk_rows) {
                const int64_t c = std::min(chunk_rows, m - o);
                CU(cudaMemcpyAsync(dst + (size_t)o * d, src + (size_t)(done + o) * d, (size_t)c * d * sizeof(float),
                                   cudaMemcpyHostToDevice, s));
            }
        } else if (kind == RowSource::Device) {
            CU(cudaMemcpyAsync(dst, src + (size_t)done * d, (size_t)m * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
        } else {
            launch_synth(dst, m, d, seed, row0_global + done, dist, s);
        }
        if (idx->center_enabled && !idx->center_valid && idx->ntotal == 0) {
            launch_column_mean(dst, std::min<int64_t>(m, 1 << 20), d, idx->center_accum, idx->center, s);
            idx->center_valid = true;
        }
        const float* center = idx->center_valid ? idx->center : nullptr;
        const int64_t end = seg->n_rows + m;
        if (seg->shadow != nullptr && seg->f16_rows == seg->n_rows) {
            // the f16 image exists and is complete: keep it so (otherwise it is (re)built on first use, ensure_f16_image)
            convert_f16_rows(idx, seg, end, s);
        }
        if (seg->shadow8 != nullptr) {
            // whole tiles are rebuilt from the fp32 rows (an append into a partly filled tile changes its scale)
            launch_convert_tiles_i8(seg->rows, end, d, seg->n_rows / kTileRows,
                                    std::min(round_up(end, kRowAlign), seg->cap_rows) / kTileRows, seg->shadow8,
                                    seg->tiles8, seg->stats, center, s);
        }
        merge_stats_kernel<<<1, 1, 0, s>>>(idx->corpus_stats, seg->stats);
        CU(cudaGetLastError());
        seg->n_rows = end;
        idx->ntotal += m;
        done += m;
    }
    // faiss copies on add: the caller may free / overwrite its buffer as soon as we return
    CU(cudaStreamSynchronize(s));
    return HAC_OK;
}

uint32_t cap_for_k(int k, int level) {
    uint32_t cap = k <= 128 ? 4096u : (k <= 512 ? 8192u : 16384u);
    if (level >= 1 && cap < 16384u) cap *= 2;
    return cap;
}

void free_workspace(Workspace& w) {
    void* ptrs[] = {w.q, w.q_shadow, w.q_shadow8, w.q_consts, w.q_stats, w.q_norm, w.q_err, w.margin, w.tau, w.thr, w.scalars, w.counters,
                    w.cb.score, w.cb.row, w.cb.exact, w.cb.count, w.cb.sorted, w.cb.overflow, w.D, w.I, w.q_shift};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (auto& l : w.lg) {
        if (l.score) cudaFree(l.score);
        if (l.row) cudaFree(l.row);
        if (l.count) cudaFree(l.count);
    }
    if (w.host_pinned) cudaFreeHost(w.host_pinned);
    w = Workspace{};
}

int ensure_workspace(hac_index* idx, int nq_pad, uint32_t cap, int64_t out_elems) {
    Workspace& w = idx->ws;
    if (w.host_pinned == nullptr) CU(cudaMallocHost(&w.host_pinned, 64));
    if (w.scalars == nullptr) {
        CU(cudaMalloc(&w.scalars, 8 * sizeof(float)));
        CU(cudaMalloc(&w.counters, 2 * sizeof(unsigned long long)));
        CU(cudaMalloc(&w.q_stats, sizeof(OperandStats)));
        CU(cudaMalloc(&w.cb.overflow, sizeof(uint32_t)));
    }
    if (nq_pad > w.nq_pad || idx->d != w.d) {
        // per-query arrays
        void* ptrs[] = {w.q, w.q_shadow, w.q_shadow8, w.q_consts, w.q_norm, w.q_err, w.margin, w.tau, w.thr, w.cb.count,
                        w.cb.sorted, w.cb.score, w.cb.row, w.cb.exact, w.q_shift};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        w.q = nullptr; w.q_shadow = nullptr; w.q_shadow8 = nullptr; w.q_consts = nullptr;
        w.q_norm = w.q_err = w.margin = w.tau = w.thr = w.q_shift = nullptr;
        w.cb.score = w.cb.exact = nullptr; w.cb.row = w.cb.count = w.cb.sorted = nullptr;
        const int np = std::max(nq_pad, w.nq_pad);
        w.nq_pad = 0; w.cap = 0;
        CU(cudaMalloc(&w.q, (size_t)np * idx->d * sizeof(float)));
        CU(cudaMalloc(&w.q_shadow, (size_t)shadow_bytes(np, idx->d)));
        CU(cudaMalloc(&w.q_shadow8, (size_t)shadow_bytes(np, idx->d)));     // (sized like the f16 image: enough)
        CU(cudaMalloc(&w.q_consts, np * sizeof(QueryQ8)));
        CU(cudaMalloc(&w.q_norm, np * sizeof(float)));
        CU(cudaMalloc(&w.q_err, np * sizeof(float)));
        CU(cudaMalloc(&w.margin, np * sizeof(float)));
        CU(cudaMalloc(&w.tau, np * sizeof(float)));
        CU(cudaMalloc(&w.thr, np * sizeof(float)));
        CU(cudaMalloc(&w.q_shift, np * sizeof(float)));
        CU(cudaMalloc(&w.cb.count, np * sizeof(uint32_t)));
        CU(cudaMalloc(&w.cb.sorted, np * sizeof(uint32_t)));
        w.nq_pad = np; w.d = idx->d;
    }
    if (cap > w.cap) {
        // per-(query, slot) arrays; the per-query arrays above are left alone
        void* ptrs[] = {w.cb.score, w.cb.row, w.cb.exact};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        w.cb.score = w.cb.exact = nullptr; w.cb.row = nullptr;
        w.cap = 0;
        CU(cudaMalloc(&w.cb.score, (size_t)w.nq_pad * cap * sizeof(float)));
        CU(cudaMalloc(&w.cb.row, (size_t)w.nq_pad * cap * sizeof(uint32_t)));
        CU(cudaMalloc(&w.cb.exact, (size_t)w.nq_pad * cap * sizeof(float)));
        w.cap = cap;
    }
    if (out_elems > w.out_elems) {
        if (w.D) cudaFree(w.D);
        if (w.I) cudaFree(w.I);
        w.D = nullptr; w.I = nullptr; w.out_elems = 0;
        CU(cudaMalloc(&w.D, out_elems * sizeof(float)));
        CU(cudaMalloc(&w.I, out_elems * sizeof(int64_t)));
        w.out_elems = out_elems;
    }
    w.cb.emitted = w.counters;
    return HAC_OK;
}

// log shortlists of the int8 path: [2][nq_pad][log_cap] (screen score, row) + [2][nq_pad] counters
int ensure_logs(hac_index* idx, int nq_pad, uint32_t log_cap) {
    Workspace& w = idx->ws;
    if (nq_pad <= w.log_nq_pad && log_cap <= w.log_cap) return HAC_OK;
    const int np = std::max(nq_pad, w.log_nq_pad);
    const uint32_t lc = std::max(log_cap, w.log_cap);
    for (auto& l : w.lg) {
        if (l.score) cudaFree(l.score);
        if (l.row) cudaFree(l.row);
        if (l.count) cudaFree(l.count);
        l = CandBuf{};
    }
    w.log_nq_pad = 0; w.log_cap = 0;
    for (auto& l : w.lg) {
        CU(cudaMalloc(&l.score, (size_t)np * lc * sizeof(float)));
        CU(cudaMalloc(&l.row, (size_t)np * lc * sizeof(uint32_t)));
        CU(cudaMalloc(&l.count, (size_t)np * sizeof(uint32_t)));
    }
    w.log_nq_pad = np; w.log_cap = lc;
    return HAC_OK;
}

struct HostReadback {
    uint32_t overflow;
    uint32_t pad;
    unsigned long long emitted;
    unsigned long long rescored;
    float margin_max;
    float screen_err_max;
};

// int8 screen path.  The corpus is cut in chunks; per chunk: int8 tensor-core scan with integer thresholds, its
// survivors appended to a log shortlist -> exact fp32 rescore of the log, survivors of THAT (exact score >= the k-th
// best exact score so far) moved to the main shortlist -> refresh (tau = k-th best exact score, everything below
// dropped, emission threshold raised).  A row is emitted when its upper bound  a8 + m8(q, tile)  reaches tau, so only
// ONE margin separates the screen from the exact threshold, and after the last chunk the shortlist IS the exact top-k.
//
// Schedule.  Thresholds must exist before a chunk may be large, so the first chunks (the prelude) are small and
// synchronous: scan i waits for the worker (rescore + refresh) of chunk i-1.  Once `min_rows` rows have been seen the
// search is pipelined: the scans run back to back on the handle's high-priority stream, the worker of chunk i runs on
// the low-priority side stream beside the scan of chunk i+1, and scan i only waits for the worker of chunk i-2 (which
// also frees the log buffer it writes, lg[i & 1]).  A stale threshold is a valid lower bound on the final one - the
// overlap changes how many rows are emitted, never the result.
struct ChunkPlan {
    int seg;
    int64_t r0, r1;
    int dist;          // the scan waits for the worker of chunk (index - dist)
};

// start_row: rows [0, start_row) of the first segment have been searched already (warm start); max_chunks: event budget
void plan_chunks_i8(const hac_index* idx, int nq, int nq_pad, int k, uint32_t log_cap, int64_t start_row, int max_chunks,
                    std::vector<ChunkPlan>& out, int* n_sync) {
    const bool few = nq <= 4 && k <= 128;
    const int n_qtiles = nq_pad / kTileRows;
    const int cg = (idx->i8_cta_group == 2 && n_qtiles % 2 == 0) ? 2 : 1;
    // one "round" of the tile-major scan: every CTA group takes one 256-row tile
    const int64_t quantum = kRowAlign * std::max(1, idx->sm_count / cg);
    // synchronous chunks: the first is emitted unfiltered and must fit the log; later ones grow with the rows seen so
    // far - a chunk is expected to emit about growth * k * exp(m8 * z / sigma) rows per query
    const int64_t first = few ? log_cap / 2
                              : std::min<int64_t>(log_cap / 2, std::max<int64_t>(512, round_up(2 * (int64_t)k, kRowAlign)));
    double sync_growth, pipe_growth = idx->i8_pipe_growth;
    int64_t min_rows;
    const bool pipeline = idx->i8_pipeline && idx->i8_pipe_dist >= 2;
    if (pipeline) {
        // prelude: large batches pay ~1 us per emitted row per query and ~40 us per synchronous chunk: growth 0.6
        sync_growth = idx->i8_chunk_growth > 0.0 ? idx->i8_chunk_growth
                      : few ? 4.0 : (nq >= 512 && k <= 128) ? 0.6 : (k <= 128 ? 2.0 : 1.0);
        // about 75 us of scan per pipelined chunk: 4 rounds at 2560 queries (a unit of 256 queries x 256 rows takes
        // ~2 us), more rows for smaller batches, whose scan is bound by HBM (~9M rows per ms)
        min_rows = idx->i8_pipe_min_rows > 0 ? idx->i8_pipe_min_rows
                                             : std::min<int64_t>(1 << 20, std::max<int64_t>(75776, 75776ll * 2560 / nq_pad));
        min_rows = round_up(min_rows, quantum);
    } else {
        // no overlap: every chunk boundary costs a full scan ramp-down + rescore + refresh, and fresher thresholds
        // emit fewer rows (measured at 25.7M x 2514, k=100: growth 2.0 -> 26.4M pairs, 53.5 ms; 1.0 -> 21.9M, 48.8 ms;
        // 0.6 -> 19.9M, 47.2 ms; 0.35 -> 18.5M, 46.9 ms); small batches keep few chunks (each costs ~40 us)
        sync_growth = idx->i8_chunk_growth > 0.0 ? idx->i8_chunk_growth
                      : few ? 4.0
                      : (nq >= 512 && k <= 128 && start_row >= (1 << 17)) ? 0.35      // behind a warm start: few chunks, all long
                      : (nq >= 512 && k > 128 && start_row > 0) ? 0.5                 // k = 250: 0.35 / 0.5 / 1.0 -> 49.1 / 49.2 / 51.7 ms; k = 1000: 77.8 / 78.1 / 82.4
                      : (nq >= 512 && k <= 128 && start_row > 0) ? 0.5
                      : (nq >= 512 && k <= 128 && idx->ntotal >= (8ll << 20)) ? 0.6
                      : (nq >= 512 && k <= 128 && idx->ntotal >= (1ll << 20)) ? 1.0
                      : (k <= 128 ? 2.0 : 1.0);
        min_rows = INT64_MAX;
    }
    for (int attempt = 0; attempt < 8; ++attempt) {
        out.clear();
        *n_sync = 0;
        int64_t rows_done = start_row;
        for (size_t si = 0; si < idx->segs.size(); ++si) {
            const Segment& seg = idx->segs[si];
            int64_t r = si == 0 ? start_row : 0;
            while (r < seg.n_rows) {
                const bool piped = pipeline && rows_done >= min_rows;
                int64_t size;
                if (piped) {
                    size = std::max<int64_t>(min_rows, (int64_t)(pipe_growth * (double)rows_done));
                    size = size / quantum * quantum;
                    if (seg.n_rows - (r + size) < min_rows / 2) size = seg.n_rows - r;     // no runt at the segment end
                } else {
                    size = rows_done == 0 ? first : (int64_t)(sync_growth * (double)rows_done);
                    if (pipeline) size = std::min(size, min_rows);
                    size = std::max<int64_t>(kRowAlign, size / kRowAlign * kRowAlign);
                }
                const int64_t r1 = std::min(seg.n_rows, r + size);
                out.push_back(ChunkPlan{(int)si, r, r1, piped ? idx->i8_pipe_dist : 1});
                if (!piped) ++*n_sync;
                rows_done += r1 - r;
                r = r1;
            }
        }
        if ((int)out.size() <= max_chunks) return;
        pipe_growth *= 1.5;                  // too many chunks for the event pool: coarser schedule
        sync_growth *= 1.5;
    }
}

// Returns 1 when a shortlist overflowed (caller falls back to the f16 path), 0 on success, < 0 on error.
int search_batch_i8(hac_index* idx, int nq, int nq_pad, const float* q_dev, int k, float* D_dev, int64_t* I_dev,
                    cudaStream_t s, const SegTable& segs) {
    const int d = idx->d;
    hac_stats& st = idx->stats;
    // a handful of queries: rescoring is cheap and every synchronous chunk costs ~40 us of launch / refresh latency,
    // so the log is doubled and the prelude grows twice as fast
    const bool few = nq <= 4 && k <= 128;
    const uint32_t cap = cap_for_k(k, 0);
    const uint32_t log_cap = few ? 2 * cap : cap;
    int rc = ensure_workspace(idx, nq_pad, cap, 0);
    if (rc == HAC_OK) rc = ensure_logs(idx, nq_pad, log_cap);
    if (rc != HAC_OK) return rc;
    Workspace& w = idx->ws;
    CandBuf cb = w.cb;
    cb.cap = cap;
    CandBuf lg[2];
    for (int i = 0; i < 2; ++i) {
        lg[i] = w.lg[i];
        lg[i].cap = log_cap;
        lg[i].exact = nullptr;
        lg[i].sorted = nullptr;
        lg[i].overflow = cb.overflow;
        lg[i].emitted = cb.emitted;
    }
    HostReadback* hr = static_cast<HostReadback*>(w.host_pinned);
    // warm start: the first rows are searched with the f16 screen (tight margin), the int8 scan continues behind them
    constexpr int kMaxWarmChunks = 8;
    const int64_t warm_rows = pick_warm_rows(idx, nq, k);
    std::vector<ChunkPlan> plan;
    int n_sync = 0;
    plan_chunks_i8(idx, nq, nq_pad, k, log_cap, warm_rows, kMaxChunks - kMaxWarmChunks, plan, &n_sync);
    if ((int)plan.size() > kMaxChunks - kMaxWarmChunks) return fail(HAC_E_STATE, "int8 search: chunk plan exceeds the event pool");
    const float* center = idx->center_valid ? idx->center : nullptr;
    // threshold exchange with the other shards: on when the caller armed an epoch for this search
    ThrExchange ex = idx->exchange;
    const bool use_ex = ex.n_peers > 0 && idx->exchange_epoch > 0 && (int64_t)nq * kExWords <= idx->exchange_capacity &&
                        idx->cur_batch < 16;
    ex.tag = (uint32_t)(((uint64_t)idx->exchange_epoch << 4) | (uint64_t)idx->cur_batch);
    if (use_ex) {
        // ranks at which every shard publishes its best scores: dense around a shard's fair share ks = ceil(k / G)
        // (where the global k-th best cuts an evenly mixed shard), sparser up to k (skewed shards)
        const int G = ex.n_peers + 1, ks = (k + G - 1) / G;
        const double mult[] = {0.3, 0.5, 0.7, 0.85, 1.0, 1.2, 1.5, 2.0, 3.0, 5.0};
        std::vector<int> r;
        for (double m : mult) r.push_back(std::min(k, std::max(1, (int)(m * ks + 0.5))));
        r.push_back(std::min(k, std::max(1, k / 2)));
        r.push_back(k);
        std::sort(r.begin(), r.end());
        r.erase(std::unique(r.begin(), r.end()), r.end());
        ex.n_ranks = (int)std::min<size_t>(r.size(), kExRanks);
        for (int i = 0; i < ex.n_ranks; ++i) ex.ranks[i] = r[r.size() - ex.n_ranks + i];   // keep the largest (k is last)
    }
    // A: scans (and everything before the first one); B: workers and the final select.  The caller's stream is
    // ordered before A at the start and after A at the end.
    // Without pipelined chunks everything is one dependency chain: it runs on A alone (no cross-stream event hops at the
    // chunk boundaries, two driver calls less per chunk).
    const bool two_streams = n_sync < (int)plan.size();
    cudaStream_t A = idx->stream, B = two_streams ? idx->side : idx->stream;
    if (s != A) {
        CU(cudaEventRecord(idx->ev_in, s));
        CU(cudaStreamWaitEvent(A, idx->ev_in, 0));
    }
    int launches = 0;
    cudaEventRecord(idx->ev[0], A);
    launch_init_search(cb, w.tau, w.thr, nq, nq_pad, A);
    cudaMemsetAsync(lg[0].count, 0, (size_t)nq_pad * sizeof(uint32_t), A);
    cudaMemsetAsync(lg[1].count, 0, (size_t)nq_pad * sizeof(uint32_t), A);
    launch_convert_queries_i8(q_dev, nq, nq_pad, d, w.q_shadow8, w.q_consts, A);
    if (center != nullptr) {
        launch_query_shift(q_dev, nq, nq_pad, d, center, w.q_shift, A);
        ++launches;
    }
    cudaMemsetAsync(w.scalars + 2, 0, sizeof(float), A);
    cudaMemsetAsync(w.counters + 1, 0, sizeof(unsigned long long), A);
    int n_warm = 0;                        // scan launches of the warm start (they take the first event pairs)
    if (warm_rows > 0) {
        rc = ensure_warm(idx, warm_rows, A);
        if (rc != HAC_OK) return rc;
        // f16 query image and the f16 screen margin against the slab's own statistics
        cudaMemsetAsync(w.q_stats, 0, sizeof(OperandStats), A);
        launch_absmax(q_dev, (int64_t)nq * d, w.scalars + 0, A);
        launch_pick_scale(w.q_stats, w.scalars + 0, 0, A);
        launch_convert_rows(q_dev, nq, nq_pad, d, w.q_shadow, 0, w.q_stats, w.q_norm, w.q_err, idx->drop_bits_q, nullptr, A);
        launch_margins(w.q_norm, w.q_err, idx->warm.stats, center ? center + d : nullptr, d, w.margin, w.scalars + 1, nq, A);
        launches += 4;
        const Segment& s0 = idx->segs[0];
        const double growth = std::min(idx->chunk_growth, std::max(1.0, (double)cap / (8.0 * k)));
        int64_t r = 0;
        while (r < warm_rows) {
            int64_t size = r == 0 ? cap / 2 : (int64_t)(growth * (double)r);
            size = std::max<int64_t>(kRowAlign, size / kRowAlign * kRowAlign);
            int64_t r1 = std::min(warm_rows, r + size);
            if (n_warm == kMaxWarmChunks - 1) r1 = warm_rows;
            cudaEventRecord(idx->ev[2 + 2 * n_warm], A);
            MmaScanArgs a;
            a.q_shadow = w.q_shadow;
            a.x_shadow = idx->warm.shadow;
            a.q_stats = w.q_stats;
            a.x_stats = idx->warm.stats;
            a.x_tiles = nullptr;
            a.q_consts = nullptr;
            a.q_shift = center != nullptr ? w.q_shift : nullptr;
            a.center_norm = nullptr;
            a.thr = w.thr;
            a.d = d;
            a.tile_major = idx->scan_tile_major < 0 ? 0 : idx->scan_tile_major;
            a.n_qtiles = nq_pad / kTileRows;
            a.ct0 = r / kRowAlign;
            a.ct1 = (r1 + kRowAlign - 1) / kRowAlign;
            a.seg_rows = r1;
            a.row_id_base = s0.base;
            a.cb = cb;
            CU(launch_scan_mma(a, idx->sm_count, idx->mma_cta_group, A));
            cudaEventRecord(idx->ev[3 + 2 * n_warm], A);
            launch_refresh(cb, k, w.margin, w.tau, w.thr, nq, A);
            launches += 2;
            ++n_warm;
            r = r1;
        }
        // the survivors' exact scores replace their screen scores; the exact phase starts from them
        cudaMemsetAsync(cb.sorted, 0, (size_t)nq_pad * sizeof(uint32_t), A);
        launch_rescore_new(cb, q_dev, d, segs, nq, w.scalars + 2, w.counters + 1, A);
        launch_fill_f32(w.tau, -INFINITY, nq, A);          // the k-th best SCREEN score is no bound on exact scores
        launches += 2;
    }
    launch_margins(nullptr, nullptr, nullptr, nullptr, d, w.margin, w.scalars + 1, nq, A);   // refresh margin = 0 (exact scores)
    launch_margins_i8(w.q_consts, idx->corpus_stats, d, w.q_norm /*scratch*/, w.scalars + 1, nq, A);   // statistics
    launches += 4;
    if (warm_rows > 0) {
        // tau = k-th best exact score of the warm rows = the first emission threshold of the int8 scan
        launch_refresh(cb, k, w.margin, w.tau, w.thr, nq, A, use_ex ? &ex : nullptr);
        ++launches;
    }
    const int n_chunks = (int)plan.size();
    int waited = -1;                       // A has already been ordered after the workers of chunks <= waited
    for (int i = 0; i < n_chunks; ++i) {
        const ChunkPlan& ch = plan[i];
        const Segment& seg = idx->segs[ch.seg];
        const int dep = i - ch.dist;
        if (two_streams && dep > waited) {
            CU(cudaStreamWaitEvent(A, idx->wdone[dep], 0));
            waited = dep;
        }
        cudaEvent_t ev_start = idx->ev[2 + 2 * (n_warm + i)], ev_stop = idx->ev[3 + 2 * (n_warm + i)];
        cudaEventRecord(ev_start, A);
        MmaScanArgs a;
        a.q_shadow = w.q_shadow8;
        a.x_shadow = seg.shadow8;
        a.q_stats = nullptr;
        a.x_stats = nullptr;
        a.q_shift = center != nullptr ? w.q_shift : nullptr;
        a.center_norm = center != nullptr ? center + d : nullptr;
        a.x_tiles = seg.tiles8;
        a.q_consts = w.q_consts;
        a.thr = w.thr;
        a.d = d;
        a.tile_major = idx->scan_tile_major < 0 ? 1 : idx->scan_tile_major;
        a.n_qtiles = nq_pad / kTileRows;
        a.ct0 = ch.r0 / kRowAlign;
        a.ct1 = (ch.r1 + kRowAlign - 1) / kRowAlign;
        a.seg_rows = std::min(seg.n_rows, ch.r1);
        a.row_id_base = seg.base;
        a.cb = lg[i & 1];
        CU(launch_scan_mma_i8(a, idx->sm_count, idx->i8_cta_group, A));
        CU(cudaEventRecord(ev_stop, A));
        // worker of the chunk (on the side stream when chunks are pipelined)
        if (two_streams) CU(cudaStreamWaitEvent(B, ev_stop, 0));
        if (!launch_rescore_log(lg[i & 1], cb, q_dev, d, segs, nq, w.tau, w.scalars + 2, w.counters + 1, B))
            return fail(HAC_E_INVALID, "int8 search: unsupported dimension");
        launch_refresh(cb, k, w.margin, w.tau, w.thr, nq, B, use_ex ? &ex : nullptr, lg[i & 1].count, /*small_cta=*/two_streams);
        if (two_streams) CU(cudaEventRecord(idx->wdone[i], B));
        launches += 3;
    }
    launch_final_select(cb, k, nq, idx->id_table, idx->id_base, D_dev, I_dev, /*use_score=*/true, B);
    ++launches;
    if (two_streams) {
        CU(cudaEventRecord(idx->ev_join, B));
        CU(cudaStreamWaitEvent(A, idx->ev_join, 0));
    }
    cudaEventRecord(idx->ev[1], A);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&hr->overflow, cb.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, A));
    CU(cudaMemcpyAsync(&hr->emitted, w.counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, A));
    CU(cudaMemcpyAsync(&hr->margin_max, w.scalars + 1, 2 * sizeof(float), cudaMemcpyDeviceToHost, A));
    if (s != A) {
        CU(cudaEventRecord(idx->ev_out, A));
        CU(cudaStreamWaitEvent(s, idx->ev_out, 0));
    }
    CU(cudaStreamSynchronize(A));
    st.n_chunks = n_chunks + n_warm;
    st.n_sync_chunks = n_sync + n_warm;
    st.pipelined = n_sync < n_chunks ? 1 : 0;
    st.warm_rows = (int32_t)warm_rows;
    st.kernel_launches = launches;
    st.candidates_emitted = (int64_t)hr->emitted;
    st.candidates_rescored = (int64_t)hr->rescored;
    st.margin_max = hr->margin_max;
    st.screen_err_max = hr->screen_err_max;
    float ms = 0.f, scan_ms = 0.f, tail_ms = 0.f;
    cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[1]);
    for (int i = 0; i < n_chunks + n_warm; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, idx->ev[2 + 2 * i], idx->ev[3 + 2 * i]);
        scan_ms += t;
    }
    if (n_chunks + n_warm > 0) cudaEventElapsedTime(&tail_ms, idx->ev[3 + 2 * (n_chunks + n_warm - 1)], idx->ev[1]);
    st.total_ms = ms;
    st.scan_ms = scan_ms;
    st.tail_ms = tail_ms;
    return hr->overflow ? 1 : 0;
}

// One query batch (nq <= kMaxQueryBatch) over the whole shard; all buffers on the device.
int search_batch(hac_index* idx, int nq, const float* q_dev, int k, float* D_dev, int64_t* I_dev, cudaStream_t s,
                 int path) {
    const int d = idx->d;
    const int nq_pad = (int)round_up(nq, kTileRows);
    // all paths end in the same exact fp32 scores; the tensor-core screens stream 1/4 (int8) or 1/2 (f16) of the
    // bytes of the fp32 rows, so they also win at the smallest batches (Q=1: 3.0 / 5.4 / 11.3 ms over 25.7M rows)
    bool have_i8 = !idx->segs.empty();
    for (const auto& sg : idx->segs) have_i8 = have_i8 && sg.shadow8 != nullptr;
    if (path == HAC_PATH_AUTO) {
        // with the int8 image present its screen wins at every k and batch size: large batches run the tensor pipe at
        // the int8 rate (38.5 vs 74.4 ms of scan at 25.7M x 2514), small ones stream half the bytes (Q=1: 2.8 vs 5.4
        // ms); rescoring its ~10^4 emitted rows per query (k = 100) runs at the HBM rate, and for larger k the f16 warm
        // slab grows with k.  Corpora that overflowed it go to the f16 screen.
        bool big_enough = true;
        if (k > 128) {
            const int64_t per_k = nq >= 128 ? idx->i8_large_k_rows_per_k
                                            : std::max<int64_t>(idx->i8_large_k_rows_per_k, 256 * (int64_t)nq);
            big_enough = idx->i8_large_k_rows_per_k == 0 || idx->ntotal >= per_k * k;
        }
        const bool take_i8 = have_i8 && idx->default_path == HAC_PATH_MMA && !idx->i8_overflowed &&
                             k <= idx->i8_auto_max_k && nq <= idx->i8_auto_max_queries && big_enough;
        path = take_i8 ? HAC_PATH_I8 : idx->default_path;
    }
    if (path == HAC_PATH_I8 && !have_i8) path = HAC_PATH_MMA;      // d % 128 != 0 or int8 image disabled
    if (path == HAC_PATH_GEMV && nq > 4) return fail(HAC_E_INVALID, "GEMV path takes at most 4 queries per batch");
    if (!idx->events_ready) {
        for (auto& e : idx->ev) CU(cudaEventCreate(&e));
        for (auto& e : idx->wdone) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&idx->ev_in, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&idx->ev_out, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&idx->ev_join, cudaEventDisableTiming));
        idx->events_ready = true;
    }
    if ((path == HAC_PATH_MMA || path == HAC_PATH_I8) && !idx->mma_configured) {
        CU(scan_mma_configure());
        idx->mma_configured = true;
    }
    hac_stats& st = idx->stats;
    st.path = path;
    st.retries = 0;
    st.n_sync_chunks = 0;
    st.pipelined = 0;
    st.warm_rows = 0;
    st.tail_ms = 0.f;

    SegTable segs;
    segs.n = (int)idx->segs.size();
    for (int i = 0; i < segs.n; ++i) {
        segs.base[i] = idx->segs[i].base;
        segs.rows[i] = idx->segs[i].rows;
    }
    if (path == HAC_PATH_I8) {
        // CTA pairs take two query tiles per unit: an odd tile count (8209 queries = 65 tiles) would fall back to the
        // single-CTA scan (measured 12 % slower), so one all-padding tile is added instead (it never emits)
        const int nq_pad8 = (idx->i8_cta_group == 2 && nq > kTileRows) ? (int)round_up(nq, 2 * kTileRows) : nq_pad;
        const int rc8 = search_batch_i8(idx, nq, nq_pad8, q_dev, k, D_dev, I_dev, s, segs);
        if (rc8 <= 0) return rc8;
        idx->i8_overflowed = true;
        path = HAC_PATH_MMA;            // shortlist overflow: redo with the f16 screen (and its careful mode)
        st.path = path;
        st.retries = 1;
    }
    if (path == HAC_PATH_MMA) {
        const int rcf = ensure_f16_image(idx, s);        // lazy f16 image: built on the first search that needs it
        if (rcf != HAC_OK) return rcf;
    }
    const int retries_before = st.retries;
    // level 0: fast mode - the whole search is enqueued without a host round trip; the overflow flag is
    //          read once at the end.
    // level 1: careful mode (only after level 0 overflowed: mass near-duplicates, adversarial data) -
    //          larger shortlist, the flag is read after every chunk; an overflowing chunk is rolled back and
    //          split in two; when a minimal chunk still overflows the carry-over is cut to its exact top-k
    //          (rescore + exact compaction), which bounds it by k whatever the data.  Always terminates.
    const float* center = idx->center_valid ? idx->center : nullptr;
    const int start_level = idx->sticky_level;
    for (int level = start_level; level < 2; ++level) {
        const bool careful = level == 1;
        const uint32_t cap = cap_for_k(k, level);
        int rc = ensure_workspace(idx, nq_pad, cap, 0);
        if (rc != HAC_OK) return rc;
        Workspace& w = idx->ws;
        CandBuf cb = w.cb;
        cb.cap = cap;   // rows of the candidate arrays are cap entries apart (may be below the allocation)
        HostReadback* hr = static_cast<HostReadback*>(w.host_pinned);
        int launches = 0, n_chunks = 0, n_ev = 2;
        cudaEventRecord(idx->ev[0], s);
        launch_init_search(cb, w.tau, w.thr, nq, nq_pad, s);
        ++launches;
        if (path == HAC_PATH_MMA) {
            cudaMemsetAsync(w.q_stats, 0, sizeof(OperandStats), s);
            launch_absmax(q_dev, (int64_t)nq * d, w.scalars + 0, s);
            launch_pick_scale(w.q_stats, w.scalars + 0, 0, s);
            launch_convert_rows(q_dev, nq, nq_pad, d, w.q_shadow, 0, w.q_stats, w.q_norm, w.q_err, idx->drop_bits_q,
                                nullptr, s);
            launch_margins(w.q_norm, w.q_err, idx->corpus_stats, center ? center + d : nullptr, d, w.margin,
                           w.scalars + 1, nq, s);
            launches += 4;
            if (center != nullptr) {
                launch_query_shift(q_dev, nq, nq_pad, d, center, w.q_shift, s);
                ++launches;
            }
        } else {
            launch_margins(nullptr, nullptr, nullptr, nullptr, d, w.margin, w.scalars + 1, nq, s);
            ++launches;
        }
        auto scan = [&](const Segment& seg, int64_t r, int64_t r1) -> int {
            const bool timed = n_ev + 2 <= kMaxEvents;
            if (timed) cudaEventRecord(idx->ev[n_ev], s);
            if (path == HAC_PATH_MMA) {
                MmaScanArgs a;
                a.q_shadow = w.q_shadow;
                a.x_shadow = seg.shadow;
                a.q_stats = w.q_stats;
                a.x_stats = seg.stats;
                a.q_shift = center != nullptr ? w.q_shift : nullptr;
                a.center_norm = nullptr;
                a.thr = w.thr;
                a.d = d;
                a.tile_major = idx->scan_tile_major < 0 ? 0 : idx->scan_tile_major;
                a.n_qtiles = nq_pad / kTileRows;
                a.ct0 = r / kRowAlign;
                a.ct1 = (r1 + kRowAlign - 1) / kRowAlign;
                a.seg_rows = std::min(seg.n_rows, r1);     // rows past r1 belong to a later chunk
                a.row_id_base = seg.base;
                a.cb = cb;
                CU(launch_scan_mma(a, idx->sm_count, idx->mma_cta_group, s));
            } else {
                launch_scan_gemv(seg.rows, r, r1, d, q_dev, nq, w.thr, cb, seg.base, idx->sm_count, s);
            }
            if (timed) {
                cudaEventRecord(idx->ev[n_ev + 1], s);
                n_ev += 2;
            }
            ++launches;
            ++n_chunks;
            return HAC_OK;
        };
        auto overflowed = [&](bool* flag) -> int {
            CU(cudaMemcpyAsync(&hr->overflow, cb.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            *flag = hr->overflow != 0;
            return HAC_OK;
        };
        // chunk schedule: the first chunk is emitted unfiltered (it must fit the shortlist), later
        // chunks grow geometrically with the rows already seen, so every chunk is expected to add
        // about growth * k * (margin factor) candidates per query.
        const double growth = std::min(idx->chunk_growth, std::max(1.0, (double)cap / (8.0 * k)));
        int64_t rows_done = 0;
        for (size_t si = 0; si < idx->segs.size(); ++si) {
            const Segment& seg = idx->segs[si];
            int64_t r = 0;
            while (r < seg.n_rows) {
                int64_t size = rows_done == 0 ? cap / 2 : (int64_t)(growth * (double)rows_done);
                size = std::max<int64_t>(kRowAlign, size / kRowAlign * kRowAlign);
                const int64_t r1 = std::min(seg.n_rows, r + size);
                if (!careful) {
                    rc = scan(seg, r, r1);
                    if (rc != HAC_OK) return rc;
                    launch_refresh(cb, k, w.margin, w.tau, w.thr, nq, s);
                    ++launches;
                } else {
                    std::vector<std::pair<int64_t, int64_t>> todo{{r, r1}};
                    while (!todo.empty()) {
                        const auto [a0, b0] = todo.back();
                        todo.pop_back();
                        bool ovf = false;
                        rc = scan(seg, a0, b0);
                        if (rc == HAC_OK) rc = overflowed(&ovf);
                        if (rc != HAC_OK) return rc;
                        if (ovf) {
                            launch_rollback(cb, nq, s);
                            ++launches;
                            if (b0 - a0 > 2 * kRowAlign) {
                                const int64_t mid = a0 + ((b0 - a0) / 2 + kRowAlign - 1) / kRowAlign * kRowAlign;
                                todo.push_back({mid, b0});     // LIFO: the lower half is scanned first
                                todo.push_back({a0, mid});
                                continue;
                            }
                            launch_rescore(cb, q_dev, d, segs, nq, w.scalars + 2, w.counters + 1, s);
                            launch_exact_compact(cb, k, w.margin, w.tau, w.thr, nq, s);
                            launches += 2;
                            rc = scan(seg, a0, b0);
                            if (rc == HAC_OK) rc = overflowed(&ovf);
                            if (rc != HAC_OK) return rc;
                            if (ovf) return fail(HAC_E_OVERFLOW, "shortlist overflow in a minimal chunk (k too close to the cap)");
                        }
                        launch_refresh(cb, k, w.margin, w.tau, w.thr, nq, s);
                        ++launches;
                    }
                }
                rows_done += r1 - r;
                r = r1;
            }
        }
        launch_rescore(cb, q_dev, d, segs, nq, w.scalars + 2, w.counters + 1, s);
        launch_final_select(cb, k, nq, idx->id_table, idx->id_base, D_dev, I_dev, /*use_score=*/false, s);
        launches += 2;
        cudaEventRecord(idx->ev[1], s);
        CU(cudaGetLastError());
        // read back the overflow flag and the statistics (56 bytes)
        CU(cudaMemcpyAsync(&hr->overflow, cb.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&hr->emitted, w.counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(&hr->margin_max, w.scalars + 1, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        st.n_chunks = n_chunks;
        st.n_sync_chunks = n_chunks;
        st.kernel_launches = launches;
        st.candidates_emitted = (int64_t)hr->emitted;
        st.candidates_rescored = (int64_t)hr->rescored;
        st.margin_max = hr->margin_max;
        st.screen_err_max = hr->screen_err_max;
        float ms = 0.f, scan_ms = 0.f;
        cudaEventElapsedTime(&ms, idx->ev[0], idx->ev[1]);
        for (int i = 2; i + 1 < n_ev; i += 2) {
            float t = 0.f;
            cudaEventElapsedTime(&t, idx->ev[i], idx->ev[i + 1]);
            scan_ms += t;
        }
        st.total_ms = ms;
        st.scan_ms = scan_ms;
        st.retries = retries_before + level - start_level;
        if (!hr->overflow) return HAC_OK;
        idx->sticky_level = 1;      // this corpus overflows the fast mode: later searches start in careful mode
    }
    return fail(HAC_E_OVERFLOW, "candidate shortlist overflowed even in careful mode");
}

int search_common(hac_index* idx, int64_t nq, const float* q, bool q_on_host, int k, float* D, int64_t* I,
                  bool out_on_host, cudaStream_t user_stream, int path) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "search: null index");
    if (nq < 0 || (nq > 0 && (q == nullptr || D == nullptr || I == nullptr)))
        return fail(HAC_E_INVALID, "search: null buffer");
    if (k <= 0 || k > HAC_MAX_K) {
        char buf[96];
        snprintf(buf, sizeof buf, "search: k=%d outside [1, %d]", k, HAC_MAX_K);
        return fail(HAC_E_INVALID, buf);
    }
    if (path != HAC_PATH_AUTO && path != HAC_PATH_GEMV && path != HAC_PATH_MMA && path != HAC_PATH_I8)
        return fail(HAC_E_INVALID, "search: unknown path");
    if (nq == 0) return HAC_OK;
    DeviceGuard guard(idx->device);
    // host API: the handle's own stream.  device API: exactly the caller's stream (NULL = the legacy
    // default stream, which is what torch's default stream is) so that the call is stream-ordered with
    // the producer of q and the consumer of D / I - and with the caller's caching allocator.
    cudaStream_t s = (q_on_host && out_on_host) ? idx->stream : user_stream;
    idx->stats.ntotal = idx->ntotal;
    // batches are sized by the path AUTO will resolve to: a "default_path" of GEMV takes 4 queries at a time
    const bool auto_gemv = path == HAC_PATH_AUTO && idx->default_path == HAC_PATH_GEMV;
    const int64_t max_batch = (path == HAC_PATH_GEMV || auto_gemv) ? 4 : kMaxQueryBatch;
    if (idx->ntotal == 0) {
        // faiss: empty index -> every slot unfilled
        if (out_on_host) {
            for (int64_t i = 0; i < nq * k; ++i) { D[i] = -FLT_MAX; I[i] = -1; }
        } else {
            launch_fill_empty(D, I, nq * k, s);
            CU(cudaStreamSynchronize(s));
        }
        return HAC_OK;
    }
    float total_ms = 0.f, scan_ms = 0.f, tail_ms = 0.f;
    int64_t emitted = 0, rescored = 0;
    int launches = 0, retries = 0;
    float margin_max = 0.f, err_max = 0.f;
    for (int64_t q0 = 0; q0 < nq; q0 += max_batch) {
        idx->cur_batch = (int)(q0 / max_batch);
        const int nb = (int)std::min<int64_t>(max_batch, nq - q0);
        const int nb_pad = (int)round_up(nb, kTileRows * std::max(idx->mma_cta_group, idx->i8_cta_group));
        int rc = ensure_workspace(idx, nb_pad, cap_for_k(k, 0), (q_on_host || out_on_host) ? (int64_t)nb * k : 0);
        if (rc != HAC_OK) return rc;
        const float* qd = q + (size_t)q0 * idx->d;
        if (q_on_host) {
            CU(cudaMemcpyAsync(idx->ws.q, qd, (size_t)nb * idx->d * sizeof(float), cudaMemcpyHostToDevice, s));
            qd = idx->ws.q;
        }
        float* Dd = out_on_host ? idx->ws.D : D + (size_t)q0 * k;
        int64_t* Id = out_on_host ? idx->ws.I : I + (size_t)q0 * k;
        rc = search_batch(idx, nb, qd, k, Dd, Id, s, path);
        if (rc != HAC_OK) return rc;
        if (out_on_host) {
            CU(cudaMemcpyAsync(D + (size_t)q0 * k, Dd, (size_t)nb * k * sizeof(float), cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyAsync(I + (size_t)q0 * k, Id, (size_t)nb * k * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
        }
        total_ms += idx->stats.total_ms; scan_ms += idx->stats.scan_ms;
        emitted += idx->stats.candidates_emitted; rescored += idx->stats.candidates_rescored;
        launches += idx->stats.kernel_launches; retries += idx->stats.retries;
        margin_max = std::max(margin_max, idx->stats.margin_max);
        err_max = std::max(err_max, idx->stats.screen_err_max);
        tail_ms += idx->stats.tail_ms;
    }
    idx->stats.tail_ms = tail_ms;
    idx->stats.total_ms = total_ms; idx->stats.scan_ms = scan_ms;
    idx->stats.candidates_emitted = emitted; idx->stats.candidates_rescored = rescored;
    idx->stats.kernel_launches = launches; idx->stats.retries = retries;
    idx->stats.margin_max = margin_max; idx->stats.screen_err_max = err_max;
    return HAC_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int hac_abi_version(void) { return HAC_ABI_VERSION; }
const char* hac_last_error(void) { return g_err.c_str(); }

int hac_create(int d, int device, hac_index** out) {
    if (out == nullptr) return fail(HAC_E_INVALID, "create: null out pointer");
    *out = nullptr;
    if (d <= 0 || d % 64 != 0 || d > 1024) return fail(HAC_E_INVALID, "create: d must be a multiple of 64 in [64, 1024]");
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return fail(HAC_E_INVALID, "create: no such CUDA device");
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(HAC_E_STATE, "this engine is built for sm_100a (B200) only");
    hac_index* idx = new hac_index();
    idx->d = d;
    idx->device = device;
    idx->sm_count = prop.multiProcessorCount;
    if (const char* cg = getenv("HAC_MMA_CTA_GROUP")) idx->mma_cta_group = atoi(cg) == 2 ? 2 : 1;
    // opt-in to the int8 image without touching the caller's code (the reference builds its index through faiss names)
    if (const char* b8 = getenv("HAC_BUILD_I8")) idx->build_i8 = atoi(b8) != 0;
    if (d % kBlockK8 != 0) idx->build_i8 = false;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    cudaError_t e = cudaStreamCreateWithPriority(&idx->stream, cudaStreamNonBlocking, prio_greatest);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&idx->side, cudaStreamNonBlocking, prio_least);
    if (const char* v = getenv("HAC_LAZY_F16")) idx->lazy_f16 = atoi(v) != 0;
    if (const char* v = getenv("HAC_I8_PIPELINE")) idx->i8_pipeline = atoi(v) != 0;
    if (e == cudaSuccess) e = cudaMalloc(&idx->corpus_stats, sizeof(OperandStats));
    if (e == cudaSuccess) e = cudaMalloc(&idx->add_scratch, 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&idx->center, (size_t)(d + 1) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&idx->center_accum, (size_t)d * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(idx->corpus_stats, 0, sizeof(OperandStats));
    if (e != cudaSuccess) {
        hac_destroy(idx);
        return fail_cuda(e, "create");
    }
    // HAC_OPTIONS="name=value,name=value": hac_set_option defaults for every new handle (A/B runs through callers
    // that only know the faiss names); an unknown name or bad value fails the create
    if (const char* opts = getenv("HAC_OPTIONS")) {
        std::string all(opts);
        size_t pos = 0;
        while (pos < all.size()) {
            size_t end = all.find(',', pos);
            if (end == std::string::npos) end = all.size();
            const std::string item = all.substr(pos, end - pos);
            pos = end + 1;
            if (item.empty()) continue;
            const size_t eq = item.find('=');
            if (eq == std::string::npos || hac_set_option(idx, item.substr(0, eq).c_str(), atoll(item.c_str() + eq + 1)) != HAC_OK) {
                hac_destroy(idx);
                return fail(HAC_E_INVALID, std::string("create: bad HAC_OPTIONS item '") + item + "'");
            }
        }
    }
    *out = idx;
    return HAC_OK;
}

int hac_destroy(hac_index* idx) {
    if (idx == nullptr) return HAC_OK;
    DeviceGuard guard(idx->device);
    cudaDeviceSynchronize();
    for (auto& s : idx->segs) free_segment(s);
    free_warm(idx);
    free_workspace(idx->ws);
    if (idx->id_table) cudaFree(idx->id_table);
    if (idx->corpus_stats) cudaFree(idx->corpus_stats);
    if (idx->add_scratch) cudaFree(idx->add_scratch);
    if (idx->center) cudaFree(idx->center);
    if (idx->center_accum) cudaFree(idx->center_accum);
    if (idx->events_ready) {
        for (auto& e : idx->ev) cudaEventDestroy(e);
        for (auto& e : idx->wdone) cudaEventDestroy(e);
        cudaEventDestroy(idx->ev_in);
        cudaEventDestroy(idx->ev_out);
        cudaEventDestroy(idx->ev_join);
    }
    if (idx->stream) cudaStreamDestroy(idx->stream);
    if (idx->side) cudaStreamDestroy(idx->side);
    delete idx;
    return HAC_OK;
}

int hac_reserve(hac_index* idx, int64_t n_rows) {
    if (idx == nullptr || n_rows < 0) return fail(HAC_E_INVALID, "reserve: bad argument");
    DeviceGuard guard(idx->device);
    const int64_t need = n_rows - idx->ntotal;
    if (!idx->segs.empty()) {
        Segment& last = idx->segs.back();
        if (!last.closed && last.cap_rows - last.n_rows >= need) return HAC_OK;
        if (last.n_rows == 0) {            // replace an empty tail segment
            free_segment(last);
            idx->segs.pop_back();
        } else {
            last.closed = true;            // its spare capacity stays unused
        }
    }
    if (need <= 0) return HAC_OK;
    if ((int)idx->segs.size() >= kMaxSegments) return fail(HAC_E_STATE, "reserve: too many segments");
    Segment s;
    int rc = alloc_segment(idx, need, &s);
    if (rc != HAC_OK) return rc;
    s.base = (uint32_t)idx->ntotal;
    idx->segs.push_back(s);
    CU(cudaStreamSynchronize(idx->stream));
    return HAC_OK;
}

int hac_add(hac_index* idx, int64_t n, const float* x_host) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "add: null index");
    return add_rows(idx, n, x_host, RowSource::Host, nullptr, 0, 0, 0);
}
int hac_add_device(hac_index* idx, int64_t n, const float* x_dev, void* stream) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "add: null index");
    return add_rows(idx, n, x_dev, RowSource::Device, static_cast<cudaStream_t>(stream), 0, 0, 0);
}
int hac_add_synthetic(hac_index* idx, int64_t n, uint64_t seed, int64_t row0, int dist) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "add: null index");
    if (dist != 0 && dist != 1) return fail(HAC_E_INVALID, "add_synthetic: dist must be 0 or 1");
    return add_rows(idx, n, nullptr, RowSource::Synthetic, nullptr, seed, row0, dist);
}
int hac_synth_fill_device(int device, float* out_dev, int64_t n, int d, uint64_t seed, int64_t row0, int dist,
                          void* stream) {
    if (out_dev == nullptr || n < 0 || d <= 0 || d % 2) return fail(HAC_E_INVALID, "synth_fill: bad argument");
    DeviceGuard guard(device);
    launch_synth(out_dev, n, d, seed, row0, dist, static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return HAC_OK;
}

int hac_reset(hac_index* idx) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "reset: null index");
    DeviceGuard guard(idx->device);
    CU(cudaStreamSynchronize(idx->stream));
    // keep the largest segment (capacity for the next block), drop the rest
    if (idx->segs.size() > 1) {
        size_t best = 0;
        for (size_t i = 1; i < idx->segs.size(); ++i)
            if (idx->segs[i].cap_rows > idx->segs[best].cap_rows) best = i;
        Segment keep = idx->segs[best];
        for (size_t i = 0; i < idx->segs.size(); ++i)
            if (i != best) free_segment(idx->segs[i]);
        idx->segs.assign(1, keep);
    }
    for (auto& s : idx->segs) {
        s.n_rows = 0;
        s.f16_rows = 0;
        s.base = 0;
        s.closed = false;
        CU(cudaMemsetAsync(s.stats, 0, sizeof(OperandStats), idx->stream));
    }
    CU(cudaMemsetAsync(idx->corpus_stats, 0, sizeof(OperandStats), idx->stream));
    idx->warm.rows = 0;                      // the warm-start image belongs to the rows that are gone
    if (idx->id_table) {
        cudaFree(idx->id_table);
        idx->id_table = nullptr;
        idx->id_table_n = 0;
    }
    idx->id_base = 0;
    idx->ntotal = 0;
    idx->center_valid = false;      // the next block gets its own centre
    idx->sticky_level = 0;
    idx->i8_overflowed = false;
    CU(cudaStreamSynchronize(idx->stream));
    return HAC_OK;
}

int hac_set_id_base(hac_index* idx, int64_t id_base) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "set_id_base: null index");
    idx->id_base = id_base;
    return HAC_OK;
}

int hac_set_id_table(hac_index* idx, const int64_t* ids_host, int64_t n) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "set_id_table: null index");
    DeviceGuard guard(idx->device);
    if (idx->id_table) {
        CU(cudaStreamSynchronize(idx->stream));
        cudaFree(idx->id_table);
        idx->id_table = nullptr;
        idx->id_table_n = 0;
    }
    if (ids_host == nullptr || n == 0) return HAC_OK;
    if (n < idx->ntotal) return fail(HAC_E_INVALID, "set_id_table: table shorter than ntotal");
    CU(cudaMalloc(&idx->id_table, n * sizeof(int64_t)));
    CU(cudaMemcpy(idx->id_table, ids_host, n * sizeof(int64_t), cudaMemcpyHostToDevice));
    idx->id_table_n = n;
    return HAC_OK;
}

int hac_search(hac_index* idx, int64_t nq, const float* q_host, int k, float* D_host, int64_t* I_host) {
    return search_common(idx, nq, q_host, true, k, D_host, I_host, true, nullptr, HAC_PATH_AUTO);
}
int hac_search_ex(hac_index* idx, int64_t nq, const float* q_host, int k, float* D_host, int64_t* I_host,
                  int path) {
    return search_common(idx, nq, q_host, true, k, D_host, I_host, true, nullptr, path);
}
int hac_search_device(hac_index* idx, int64_t nq, const float* q_dev, int k, float* D_dev, int64_t* I_dev,
                      void* stream) {
    return search_common(idx, nq, q_dev, false, k, D_dev, I_dev, false, static_cast<cudaStream_t>(stream),
                         HAC_PATH_AUTO);
}
int hac_search_device_ex(hac_index* idx, int64_t nq, const float* q_dev, int k, float* D_dev, int64_t* I_dev,
                         void* stream, int path) {
    return search_common(idx, nq, q_dev, false, k, D_dev, I_dev, false, static_cast<cudaStream_t>(stream), path);
}

int hac_merge_topk_device(int device, int n_lists, int64_t nq, int k, const float* D_lists_dev,
                          const int64_t* I_lists_dev, int k_out, float* D_out_dev, int64_t* I_out_dev,
                          void* stream) {
    if (n_lists <= 0 || nq < 0 || k <= 0 || k_out <= 0 || !D_lists_dev || !I_lists_dev || !D_out_dev || !I_out_dev)
        return fail(HAC_E_INVALID, "merge: bad argument");
    if ((int64_t)n_lists * k > 16384) return fail(HAC_E_INVALID, "merge: n_lists * k exceeds 16384");
    if (nq == 0) return HAC_OK;
    DeviceGuard guard(device);
    CU(launch_merge_topk(n_lists, nq, k, D_lists_dev, I_lists_dev, k_out, D_out_dev, I_out_dev,
                         static_cast<cudaStream_t>(stream)));
    return HAC_OK;
}

int hac_merge_topk_peers_device(int device, int n_lists, int64_t nq, int k, const float* const* D_list_ptrs,
                                const int64_t* const* I_list_ptrs, int k_out, float* D_out_dev, int64_t* I_out_dev,
                                void* stream) {
    if (n_lists <= 0 || n_lists > kMaxPeerLists || nq < 0 || k <= 0 || k_out <= 0 || !D_list_ptrs || !I_list_ptrs ||
        !D_out_dev || !I_out_dev)
        return fail(HAC_E_INVALID, "merge_peers: bad argument");
    if ((int64_t)n_lists * k > 16384) return fail(HAC_E_INVALID, "merge_peers: n_lists * k exceeds 16384");
    for (int i = 0; i < n_lists; ++i)
        if (!D_list_ptrs[i] || !I_list_ptrs[i]) return fail(HAC_E_INVALID, "merge_peers: null list pointer");
    if (nq == 0) return HAC_OK;
    DeviceGuard guard(device);
    CU(launch_merge_topk_peers(n_lists, nq, k, D_list_ptrs, I_list_ptrs, k_out, D_out_dev, I_out_dev,
                               static_cast<cudaStream_t>(stream)));
    return HAC_OK;
}

int hac_enable_peer_access(int device, int peer) {
    int n_dev = 0;
    CU(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev || peer < 0 || peer >= n_dev) return fail(HAC_E_INVALID, "enable_peer_access: no such device");
    if (device == peer) return HAC_OK;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return fail(HAC_E_STATE, "enable_peer_access: the devices are not peer-capable");
    DeviceGuard guard(device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return HAC_OK;
    }
    if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceEnablePeerAccess");
    return HAC_OK;
}

int hac_gather_ids_device(int device, const int64_t* table_dev, int64_t table_n, const int64_t* ids_dev, int64_t n,
                          int64_t* out_dev, void* stream) {
    if (!table_dev || !ids_dev || !out_dev || n < 0) return fail(HAC_E_INVALID, "gather: bad argument");
    DeviceGuard guard(device);
    launch_gather_ids(table_dev, table_n, ids_dev, n, out_dev, static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return HAC_OK;
}

int hac_set_threshold_exchange(hac_index* idx, uint64_t* mine_dev, const uint64_t* const* peers_dev, int n_peers,
                               int64_t capacity) {
    if (idx == nullptr) return fail(HAC_E_INVALID, "set_threshold_exchange: null index");
    if (n_peers < 0 || n_peers > kMaxPeerLists - 1 || capacity < 0 || (n_peers > 0 && (!mine_dev || !peers_dev)))
        return fail(HAC_E_INVALID, "set_threshold_exchange: bad argument");
    idx->exchange = ThrExchange{};
    idx->exchange_capacity = 0;
    if (n_peers == 0) return HAC_OK;
    for (int i = 0; i < n_peers; ++i) {
        if (!peers_dev[i]) return fail(HAC_E_INVALID, "set_threshold_exchange: null peer pointer");
        idx->exchange.peers[i] = reinterpret_cast<const unsigned long long*>(peers_dev[i]);
    }
    idx->exchange.mine = reinterpret_cast<unsigned long long*>(mine_dev);
    idx->exchange.n_peers = n_peers;
    idx->exchange_capacity = capacity;
    return HAC_OK;
}

int hac_reciprocal_rank_device(int device, const int64_t* pids_dev, int64_t nq, int k, const int64_t* rel_ptr_dev,
                               const int64_t* rel_pids_dev, float* rr_out_dev, int32_t* rank_out_dev, void* stream) {
    if (!pids_dev || !rel_ptr_dev || !rr_out_dev || !rank_out_dev || nq < 0 || k <= 0 || k > 4096)
        return fail(HAC_E_INVALID, "reciprocal_rank: bad argument");
    DeviceGuard guard(device);
    launch_reciprocal_rank(pids_dev, nq, k, rel_ptr_dev, rel_pids_dev, rr_out_dev, rank_out_dev,
                           static_cast<cudaStream_t>(stream));
    CU(cudaGetLastError());
    return HAC_OK;
}

// ---- shard files --------------------------------------------------------------------------------------------------
namespace {
constexpr uint64_t kShardAlign = 4096;
constexpr size_t kShardChunk = 64ull << 20;
struct ShardSection {
    uint64_t offset, bytes;
};
struct ShardHeader {
    char magic[8];
    uint32_t version, d;
    uint64_t n_rows, cap_rows;
    uint32_t flags, pad;
    ShardSection center, rows, i8, tiles;
    OperandStats stats;
};
static_assert(sizeof(ShardHeader) <= kShardAlign, "shard header must fit its page");
uint64_t shard_pad(uint64_t v) { return (v + kShardAlign - 1) / kShardAlign * kShardAlign; }

struct FileCloser {
    FILE* f;
    ~FileCloser() { if (f) fclose(f); }
};
struct PinnedPair {
    void* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~PinnedPair() {
        for (int i = 0; i < 2; ++i) {
            if (ev[i]) cudaEventDestroy(ev[i]);
            if (p[i]) cudaFreeHost(p[i]);
        }
    }
};

// device memory -> file at `offset`, staged through one pinned buffer
int write_section(FILE* f, uint64_t offset, const void* dev, uint64_t bytes, void* pinned, cudaStream_t s) {
    if (fseeko(f, (off_t)offset, SEEK_SET) != 0) return fail(HAC_E_STATE, "save_shard: seek failed");
    for (uint64_t o = 0; o < bytes; o += kShardChunk) {
        const size_t n = (size_t)std::min<uint64_t>(kShardChunk, bytes - o);
        CU(cudaMemcpyAsync(pinned, static_cast<const uint8_t*>(dev) + o, n, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (fwrite(pinned, 1, n, f) != n) return fail(HAC_E_STATE, "save_shard: short write (disk full?)");
    }
    return HAC_OK;
}
// file section -> device memory: the read of chunk i+1 overlaps the DMA of chunk i
int read_section(FILE* f, uint64_t offset, void* dev, uint64_t bytes, PinnedPair& pp, cudaStream_t s) {
    if (fseeko(f, (off_t)offset, SEEK_SET) != 0) return fail(HAC_E_STATE, "load_shard: seek failed");
    int b = 0;
    for (uint64_t o = 0; o < bytes; o += kShardChunk, b ^= 1) {
        const size_t n = (size_t)std::min<uint64_t>(kShardChunk, bytes - o);
        CU(cudaEventSynchronize(pp.ev[b]));                      // the DMA that last used this buffer has finished
        if (fread(pp.p[b], 1, n, f) != n) return fail(HAC_E_STATE, "load_shard: short read (truncated file)");
        CU(cudaMemcpyAsync(static_cast<uint8_t*>(dev) + o, pp.p[b], n, cudaMemcpyHostToDevice, s));
        CU(cudaEventRecord(pp.ev[b], s));
    }
    return HAC_OK;
}
}  // namespace

int hac_save_shard(hac_index* idx, const char* path) {
    if (idx == nullptr || path == nullptr) return fail(HAC_E_INVALID, "save_shard: null argument");
    if (idx->ntotal == 0) return fail(HAC_E_STATE, "save_shard: the index is empty");
    DeviceGuard guard(idx->device);
    cudaStream_t s = idx->stream;
    CU(cudaStreamSynchronize(s));
    const int d = idx->d;
    std::vector<const Segment*> segs;
    for (const auto& sg : idx->segs)
        if (sg.n_rows > 0) segs.push_back(&sg);
    const bool with_i8 = segs.size() == 1 && segs[0]->shadow8 != nullptr;
    ShardHeader h{};
    memcpy(h.magic, "HACSHD01", 8);
    h.version = 1;
    h.d = (uint32_t)d;
    h.n_rows = (uint64_t)idx->ntotal;
    h.cap_rows = (uint64_t)round_up(idx->ntotal, kRowAlign);
    h.flags = (with_i8 ? 1u : 0u) | (idx->center_valid ? 2u : 0u);
    uint64_t at = kShardAlign;
    h.center = {at, idx->center_valid ? (uint64_t)(d + 1) * sizeof(float) : 0};
    at = shard_pad(at + h.center.bytes);
    h.rows = {at, h.n_rows * (uint64_t)d * sizeof(float)};
    at = shard_pad(at + h.rows.bytes);
    h.i8 = {at, with_i8 ? (uint64_t)shadow8_bytes((int64_t)h.cap_rows, d) : 0};
    at = shard_pad(at + h.i8.bytes);
    h.tiles = {at, with_i8 ? (uint64_t)shadow_tiles((int64_t)h.cap_rows) * sizeof(TileQ8) : 0};
    const uint64_t total = shard_pad(at + h.tiles.bytes);
    CU(cudaMemcpy(&h.stats, idx->corpus_stats, sizeof(OperandStats), cudaMemcpyDeviceToHost));
    const std::string tmp = std::string(path) + ".tmp";
    void* pinned = nullptr;
    CU(cudaMallocHost(&pinned, kShardChunk));
    int rc = HAC_OK;
    {
        FileCloser fc{fopen(tmp.c_str(), "wb")};
        if (fc.f == nullptr) {
            cudaFreeHost(pinned);
            return fail(HAC_E_STATE, std::string("save_shard: cannot create ") + tmp);
        }
        std::vector<uint8_t> page(kShardAlign, 0);
        memcpy(page.data(), &h, sizeof h);
        if (fwrite(page.data(), 1, page.size(), fc.f) != page.size()) rc = fail(HAC_E_STATE, "save_shard: short write");
        if (rc == HAC_OK && h.center.bytes) rc = write_section(fc.f, h.center.offset, idx->center, h.center.bytes, pinned, s);
        uint64_t row_at = h.rows.offset;
        for (size_t i = 0; rc == HAC_OK && i < segs.size(); ++i) {
            const uint64_t nb = (uint64_t)segs[i]->n_rows * d * sizeof(float);
            rc = write_section(fc.f, row_at, segs[i]->rows, nb, pinned, s);
            row_at += nb;
        }
        if (rc == HAC_OK && with_i8) rc = write_section(fc.f, h.i8.offset, segs[0]->shadow8, h.i8.bytes, pinned, s);
        if (rc == HAC_OK && with_i8) rc = write_section(fc.f, h.tiles.offset, segs[0]->tiles8, h.tiles.bytes, pinned, s);
        if (rc == HAC_OK && (fflush(fc.f) != 0 || ftruncate(fileno(fc.f), (off_t)total) != 0))
            rc = fail(HAC_E_STATE, "save_shard: could not finish the file");
    }
    cudaFreeHost(pinned);
    if (rc != HAC_OK) {
        remove(tmp.c_str());
        return rc;
    }
    if (rename(tmp.c_str(), path) != 0) return fail(HAC_E_STATE, std::string("save_shard: cannot rename to ") + path);
    return HAC_OK;
}

int hac_load_shard(hac_index* idx, const char* path) {
    if (idx == nullptr || path == nullptr) return fail(HAC_E_INVALID, "load_shard: null argument");
    if (idx->ntotal != 0) return fail(HAC_E_STATE, "load_shard: the index must be empty (reset it first)");
    DeviceGuard guard(idx->device);
    cudaStream_t s = idx->stream;
    FileCloser fc{fopen(path, "rb")};
    if (fc.f == nullptr) return fail(HAC_E_INVALID, std::string("load_shard: cannot open ") + path);
    ShardHeader h{};
    if (fread(&h, 1, sizeof h, fc.f) != sizeof h || memcmp(h.magic, "HACSHD01", 8) != 0 || h.version != 1)
        return fail(HAC_E_INVALID, std::string("load_shard: not a version-1 shard file: ") + path);
    if ((int)h.d != idx->d) return fail(HAC_E_INVALID, "load_shard: dimension of the file differs from the index");
    if (h.n_rows == 0 || h.n_rows > 0xFFFFFF00ull || h.cap_rows != (uint64_t)round_up((int64_t)h.n_rows, kRowAlign) ||
        h.rows.bytes != h.n_rows * (uint64_t)h.d * sizeof(float))
        return fail(HAC_E_INVALID, "load_shard: corrupt header");
    if (fseeko(fc.f, 0, SEEK_END) != 0 || (uint64_t)ftello(fc.f) < h.rows.offset + h.rows.bytes ||
        (uint64_t)ftello(fc.f) < h.tiles.offset + h.tiles.bytes)
        return fail(HAC_E_INVALID, "load_shard: truncated file");
    const int d = idx->d;
    const bool file_i8 = (h.flags & 1u) != 0;
    if (file_i8 && (h.i8.bytes != (uint64_t)shadow8_bytes((int64_t)h.cap_rows, d) ||
                    h.tiles.bytes != (uint64_t)shadow_tiles((int64_t)h.cap_rows) * sizeof(TileQ8)))
        return fail(HAC_E_INVALID, "load_shard: int8 sections do not match the row count");
    // one segment of exactly the file's capacity (an existing larger empty one is kept)
    for (auto& sg : idx->segs) free_segment(sg);
    idx->segs.clear();
    idx->warm.rows = 0;
    Segment seg;
    int rc = alloc_segment(idx, (int64_t)h.cap_rows, &seg);
    if (rc != HAC_OK) return rc;
    seg.base = 0;
    idx->segs.push_back(seg);
    Segment& sg = idx->segs.back();
    PinnedPair pp;
    for (int i = 0; i < 2; ++i) {
        CU(cudaMallocHost(&pp.p[i], kShardChunk));
        CU(cudaEventCreateWithFlags(&pp.ev[i], cudaEventDisableTiming));
    }
    idx->center_valid = false;
    if ((h.flags & 2u) && idx->center_enabled) {
        if (h.center.bytes != (uint64_t)(d + 1) * sizeof(float)) return fail(HAC_E_INVALID, "load_shard: corrupt centre section");
        rc = read_section(fc.f, h.center.offset, idx->center, h.center.bytes, pp, s);
        if (rc != HAC_OK) return rc;
        idx->center_valid = true;
    } else if (h.flags & 2u) {
        return fail(HAC_E_STATE, "load_shard: the file holds a centred image but center_screen is off on this index");
    }
    rc = read_section(fc.f, h.rows.offset, sg.rows, h.rows.bytes, pp, s);
    if (rc != HAC_OK) return rc;
    sg.n_rows = (int64_t)h.n_rows;
    sg.f16_rows = 0;
    const float* center = idx->center_valid ? idx->center : nullptr;
    if (sg.shadow8 != nullptr) {
        if (file_i8) {
            rc = read_section(fc.f, h.i8.offset, sg.shadow8, h.i8.bytes, pp, s);
            if (rc == HAC_OK) rc = read_section(fc.f, h.tiles.offset, sg.tiles8, h.tiles.bytes, pp, s);
            if (rc != HAC_OK) return rc;
            CU(cudaMemcpyAsync(sg.stats, &h.stats, sizeof(OperandStats), cudaMemcpyHostToDevice, s));
            // the f16 statistics of the file belong to an image this index has not built: the lazy build recomputes them
            CU(cudaMemcpyAsync(idx->corpus_stats, &h.stats, sizeof(OperandStats), cudaMemcpyHostToDevice, s));
        } else {
            // written from a multi-segment shard: the image is rebuilt from the rows (one conversion pass)
            launch_convert_tiles_i8(sg.rows, sg.n_rows, d, 0, (int64_t)h.cap_rows / kTileRows, sg.shadow8, sg.tiles8, sg.stats,
                                    center, s);
            merge_stats_kernel<<<1, 1, 0, s>>>(idx->corpus_stats, sg.stats);
        }
    }
    if (sg.shadow != nullptr) {                                  // eager f16 image requested ("lazy_f16" = 0)
        sg.f16_rows = 0;
        convert_f16_rows(idx, &sg, sg.n_rows, s);
        merge_stats_kernel<<<1, 1, 0, s>>>(idx->corpus_stats, sg.stats);
    }
    CU(cudaGetLastError());
    idx->ntotal = (int64_t)h.n_rows;
    CU(cudaStreamSynchronize(s));
    return HAC_OK;
}

int hac_pinned_alloc(size_t bytes, void** out_host) {
    if (out_host == nullptr) return fail(HAC_E_INVALID, "pinned_alloc: null out pointer");
    *out_host = nullptr;
    cudaError_t e = cudaMallocHost(out_host, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocHost");
    return HAC_OK;
}
int hac_pinned_free(void* host) {
    if (host) cudaFreeHost(host);
    return HAC_OK;
}

int hac_set_option(hac_index* idx, const char* name, int64_t value) {
    if (idx == nullptr || name == nullptr) return fail(HAC_E_INVALID, "set_option: null argument");
    if (strcmp(name, "mma_cta_group") == 0) {
        if (value != 1 && value != 2) return fail(HAC_E_INVALID, "mma_cta_group must be 1 or 2");
        idx->mma_cta_group = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "default_path") == 0) {
        if (value != HAC_PATH_GEMV && value != HAC_PATH_MMA && value != HAC_PATH_I8)
            return fail(HAC_E_INVALID, "default_path must be a HAC_PATH_* scan path");
        idx->default_path = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "f16_drop_bits_corpus") == 0 || strcmp(name, "f16_drop_bits_queries") == 0) {
        if (value < 0 || value > 8) return fail(HAC_E_INVALID, "drop bits must be in [0, 8]");
        if (name[14] == 'c') {
            if (idx->ntotal != 0) return fail(HAC_E_STATE, "f16_drop_bits_corpus must be set on an empty index");
            idx->drop_bits_x = (int)value;
        } else {
            idx->drop_bits_q = (int)value;
        }
        return HAC_OK;
    }
    if (strcmp(name, "scan_tile_major") == 0) {
        if (value > 2) return fail(HAC_E_INVALID, "scan_tile_major must be -1, 0, 1 or 2");
        idx->scan_tile_major = value < 0 ? -1 : (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_cta_group") == 0) {
        if (value != 1 && value != 2) return fail(HAC_E_INVALID, "i8_cta_group must be 1 or 2");
        idx->i8_cta_group = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "center_screen") == 0) {
        if (idx->ntotal != 0) return fail(HAC_E_STATE, "center_screen must be set on an empty index");
        idx->center_enabled = value != 0;
        idx->center_valid = false;
        return HAC_OK;
    }
    if (strcmp(name, "i8_pipeline") == 0) { idx->i8_pipeline = value != 0; return HAC_OK; }
    if (strcmp(name, "i8_warm_rows") == 0) {
        if (value < -1 || value > (int64_t)1 << 26) return fail(HAC_E_INVALID, "i8_warm_rows out of range");
        idx->i8_warm_rows = value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_pipe_dist") == 0) {
        if (value != 1 && value != 2) return fail(HAC_E_INVALID, "i8_pipe_dist must be 1 or 2");
        idx->i8_pipe_dist = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_pipe_growth_x1000") == 0) {
        if (value < 10 || value > 4000) return fail(HAC_E_INVALID, "i8_pipe_growth_x1000 must be in [10, 4000]");
        idx->i8_pipe_growth = (double)value / 1000.0;
        return HAC_OK;
    }
    if (strcmp(name, "i8_pipe_min_rows") == 0) {
        if (value != 0 && (value < 4096 || value > (1ll << 26))) return fail(HAC_E_INVALID, "i8_pipe_min_rows must be 0 or in [4096, 2^26]");
        idx->i8_pipe_min_rows = value;
        return HAC_OK;
    }
    if (strcmp(name, "lazy_f16") == 0) {
        if (idx->ntotal != 0 || !idx->segs.empty()) return fail(HAC_E_STATE, "lazy_f16 must be set on an empty index");
        idx->lazy_f16 = value < 0 ? -1 : (value != 0);
        return HAC_OK;
    }
    if (strcmp(name, "exchange_epoch") == 0) {
        if (value < 0 || value > 0x0FFFFFFF) return fail(HAC_E_INVALID, "exchange_epoch out of range");
        idx->exchange_epoch = value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_chunk_growth_x100") == 0) {
        if (value != 0 && (value < 10 || value > 1600)) return fail(HAC_E_INVALID, "i8_chunk_growth_x100 must be 0 or in [10, 1600]");
        idx->i8_chunk_growth = (double)value / 100.0;
        return HAC_OK;
    }
    if (strcmp(name, "i8_large_k_rows_per_k") == 0) {
        if (value < 0) return fail(HAC_E_INVALID, "i8_large_k_rows_per_k must be >= 0");
        idx->i8_large_k_rows_per_k = value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_auto_max_k") == 0) {
        if (value < 0 || value > HAC_MAX_K) return fail(HAC_E_INVALID, "i8_auto_max_k out of range");
        idx->i8_auto_max_k = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "i8_auto_max_queries") == 0) {
        if (value < 0 || value > kMaxQueryBatch) return fail(HAC_E_INVALID, "i8_auto_max_queries out of range");
        idx->i8_auto_max_queries = (int)value;
        return HAC_OK;
    }
    if (strcmp(name, "build_i8") == 0) {
        if (idx->ntotal != 0 || !idx->segs.empty()) return fail(HAC_E_STATE, "build_i8 must be set on an empty index");
        idx->build_i8 = value != 0 && idx->d % kBlockK8 == 0;
        return HAC_OK;
    }
    if (strcmp(name, "chunk_growth_x100") == 0) {
        if (value < 110 || value > 1600) return fail(HAC_E_INVALID, "chunk_growth_x100 must be in [110, 1600]");
        idx->chunk_growth = (double)value / 100.0;
        return HAC_OK;
    }
    return fail(HAC_E_INVALID, std::string("set_option: unknown option ") + name);
}

int64_t hac_ntotal(const hac_index* idx) { return idx ? idx->ntotal : -1; }
int hac_dim(const hac_index* idx) { return idx ? idx->d : -1; }
int hac_device(const hac_index* idx) { return idx ? idx->device : -1; }
int hac_get_stats(const hac_index* idx, hac_stats* out) {
    if (idx == nullptr || out == nullptr) return fail(HAC_E_INVALID, "get_stats: null argument");
    *out = idx->stats;
    out->ntotal = idx->ntotal;
    int64_t b32 = 0, bsh = 0, b8 = 0;
    for (const auto& s : idx->segs) {
        b32 += s.cap_rows * (int64_t)idx->d * 4;
        if (s.shadow) bsh += shadow_bytes(s.cap_rows, idx->d);
        if (s.shadow8) b8 += shadow8_bytes(s.cap_rows, idx->d) + shadow_tiles(s.cap_rows) * (int64_t)sizeof(TileQ8);
    }
    out->bytes_fp32 = b32;
    out->bytes_shadow = bsh;
    out->bytes_i8 = b8;
    return HAC_OK;
}

}  // extern "C"
