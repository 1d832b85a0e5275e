// Declarations shared by the kernel translation units and the C-ABI layer.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hac {

constexpr int kMaxSegments = 32;
constexpr int kMaxPeerLists = 16;

// Per-operand statistics kept on the device (written by the convert kernels, read by the
// margin / scan kernels), so a search never has to synchronise with the host.
struct OperandStats {
    float absmax;        // max |x| over the rows converted so far
    float scale;         // power-of-two factor applied before rounding to f16
    float inv_scale;     // 1 / scale
    float norm_max;      // max_j ||x_j||_2            (fp32 rows)
    float hat_norm_max;  // max_j ||f16(x_j*scale)/scale||_2
    float err_norm_max;  // max_j ||x_j - f16(x_j*scale)/scale||_2
    float i8_beta_max;   // int8 image: max over tiles of beta  (max ||x_j||)
    float i8_gamma_max;  // int8 image: max over tiles of gamma (max ||x_j - alpha*xi_j||)
};

// int8 shadow: constants of one 128-row tile (written by the int8 convert kernel).
//   x_j ~= alpha * xi_j (xi int8);  beta = max_j ||x_j||;  gamma = max_j ||x_j - alpha*xi_j||  over the tile's rows
struct TileQ8 {
    float alpha, beta, gamma, pad;
};
// per-query constants of the int8 query image: q ~= t * qi;  eps = ||q - t*qi||;  nhat = ||t*qi||;  norm = ||q||
struct QueryQ8 {
    float t, eps, nhat, norm;
};

// Candidate shortlist: per query a list of (screen score, index-local row) pairs.
struct CandBuf {
    float* score;        // [nq_pad][cap]  screen (approximate) score, unscaled units
    uint32_t* row;       // [nq_pad][cap]  index-local row
    float* exact;        // [nq_pad][cap]  exact fp32 score (filled by the rescore kernel)
    uint32_t* count;     // [nq_pad]       entries appended (may exceed cap -> overflow)
    uint32_t* sorted;    // [nq_pad]       entries already sorted/compacted by the last refresh
    uint32_t* overflow;  // [1]
    unsigned long long* emitted;  // [1] total pairs that passed the filter (statistics)
    uint32_t cap;
};

// Cross-shard threshold exchange (one process per GPU, corpus sharded): every shard publishes, per query, its best
// exact scores so far at n_ranks fixed ranks into a buffer its peers can read over NVLink.  A word (shard, rank c, L)
// claims "this shard holds >= c rows scoring >= L"; the largest T at which the claims of all shards add up to k rows
// is a lower bound on the GLOBAL k-th best (refresh_kernel), to which every shard raises its threshold.  Words are
// (tag << 32 | float bits), kExWords per query; a reader ignores words whose tag is not the current search's, so no
// barrier or reset is needed between searches.
constexpr int kExWords = 16;      // = HAC_EXCHANGE_WORDS_PER_QUERY
constexpr int kExRanks = 12;      // published ranks per query (<= kExWords)
struct ThrExchange {
    unsigned long long* mine;                          // [capacity] this shard's published claims (peer-readable)
    const unsigned long long* peers[kMaxPeerLists];    // the other shards' buffers (peer-mapped device pointers)
    int n_peers;                                       // 0 = no exchange
    uint32_t tag;
    int n_ranks;
    int ranks[kExRanks];                               // ascending, 1-based, last = k
};

struct SegTable {
    int n;
    uint32_t base[kMaxSegments];     // index-local row of the segment's first row
    const float* rows[kMaxSegments]; // fp32 rows of the segment
};

// ---- prep ----------------------------------------------------------------------------------
void launch_absmax(const float* x, int64_t n_elems, float* absmax_out, cudaStream_t s);
// scale <- 2^(12 - ceil(log2(absmax))) unless *stats already holds a scale and keep_scale != 0
void launch_pick_scale(OperandStats* stats, const float* absmax_in, int keep_scale, cudaStream_t s);
// rows [0,n) of x -> shadow rows [row0, row0+n); rows [n, n_pad) are written as zeros.
// drop_bits: low mantissa bits of the f16 image forced to zero (0 = full 11-bit significand)
// center: nullptr, or d floats subtracted from every row before rounding (the image then holds x - c)
void launch_convert_rows(const float* x, int64_t n, int64_t n_pad, int d, uint8_t* shadow, int64_t row0,
                         OperandStats* stats, float* row_norm, float* row_err, int drop_bits, const float* center,
                         cudaStream_t s);
// center[0..d) <- column means of rows [0, n) of x, center[d] <- ||center|| (upper bound); accum: d doubles of scratch
void launch_column_mean(const float* x, int64_t n, int d, double* accum, float* center, cudaStream_t s);
// shift[q] = q . center for q < nq, 0 for the padding up to nq_pad
void launch_query_shift(const float* q, int nq, int nq_pad, int d, const float* center, float* shift, cudaStream_t s);
void launch_synth(float* out, int64_t n, int d, uint64_t seed, int64_t row0, int dist, cudaStream_t s);
// int8 images: whole 128-row tiles [tile0, tile1) of a segment are (re)built from its fp32 rows (n_rows valid)
// center: nullptr or the screen centre (the image and the tile norms are then those of x - c)
void launch_convert_tiles_i8(const float* rows, int64_t n_rows, int d, int64_t tile0, int64_t tile1, uint8_t* shadow8,
                             TileQ8* tiles, OperandStats* stats, const float* center, cudaStream_t s);
// diagnostic: margin[q] = eps_q*beta_max + nhat_q*gamma_max + slack, the loosest per-tile int8 margin of query q
void launch_margins_i8(const QueryQ8* q_consts, const OperandStats* corpus, int d, float* margin, float* margin_max,
                       int nq, cudaStream_t s);
void launch_convert_queries_i8(const float* q, int nq, int nq_pad, int d, uint8_t* q_shadow8, QueryQ8* consts,
                               cudaStream_t s);

// ---- search state --------------------------------------------------------------------------
void launch_init_search(CandBuf cb, float* tau, float* thr, int nq, int nq_pad, cudaStream_t s);
// margin[q] = ||e_q|| * Xhat_max + ||q|| * Ex_max + slack   (0 when corpus==nullptr: exact scan)
// center_norm: nullptr, or a pointer to ||c|| when the corpus image (and its statistics) is centred
void launch_margins(const float* q_norm, const float* q_err, const OperandStats* corpus, const float* center_norm,
                    int d, float* margin, float* margin_max, int nq, cudaStream_t s);
// per query: sort the shortlist, raise tau to the k-th best screen score, drop entries below tau - 2m
// ex: optional threshold exchange (exact-score shortlists only, i.e. the int8 path); nullptr = none
// clear_count: optional [nq] counters zeroed by the kernel (the log shortlist the preceding rescore consumed)
// small_cta: 256 threads per query whatever the batch size (co-resident with a running scan kernel)
void launch_refresh(CandBuf cb, int k, const float* margin, float* tau, float* thr, int nq, cudaStream_t s,
                    const ThrExchange* ex = nullptr, uint32_t* clear_count = nullptr, bool small_cta = false);
// exact fp32 scores of every shortlisted pair
void launch_rescore(CandBuf cb, const float* q, int d, SegTable segs, int nq, float* screen_err_max,
                    unsigned long long* rescored, cudaStream_t s);
// same for the entries appended since the last refresh only ([sorted, count)); their screen score is
// replaced by the exact one, so the refresh that follows works on exact scores (int8 path)
void launch_rescore_new(CandBuf cb, const float* q, int d, SegTable segs, int nq, float* screen_err_max,
                        unsigned long long* rescored, cudaStream_t s);
// pipelined int8 search: the scan of chunk i appends its survivors to a LOG shortlist `lg`; this kernel rescores every
// log entry exactly and appends those that can still reach the top-k (exact score >= tau[q], the k-th best exact score
// so far) to the main shortlist `cb`.  Small register / shared-memory footprint so that its CTAs are co-resident with
// the persistent scan kernel of the NEXT chunk (side stream).  d must be a multiple of 128.
bool launch_rescore_log(CandBuf lg, CandBuf cb, const float* q, int d, SegTable segs, int nq, const float* tau,
                        float* screen_err_max, unsigned long long* rescored, cudaStream_t s);
// final top-k by (exact score desc, id asc) with id translation
// use_score: the `score` array already holds exact scores (int8 path) - rank by it instead of `exact`
void launch_final_select(CandBuf cb, int k, int nq, const int64_t* id_table, int64_t id_base, float* D,
                         int64_t* I, bool use_score, cudaStream_t s);
void launch_fill_empty(float* D, int64_t* I, int64_t n, cudaStream_t s);
void launch_fill_f32(float* p, float value, int n, cudaStream_t s);
// careful mode: undo the appends since the last refresh / cut the rescored shortlist to its exact top-k
void launch_rollback(CandBuf cb, int nq, cudaStream_t s);
void launch_exact_compact(CandBuf cb, int k, const float* margin, float* tau, float* thr, int nq, cudaStream_t s);

// ---- scans ---------------------------------------------------------------------------------
// exact fp32 streaming scan of rows [r0, r1) of one segment, 1..4 queries
void launch_scan_gemv(const float* rows, int64_t r0, int64_t r1, int d, const float* q, int nq,
                      const float* thr, CandBuf cb, uint32_t row_id_base, int sm_count, cudaStream_t s);
// tcgen05 f16 screen of 256-row column tiles [ct0, ct1) of one segment against all query tiles
struct MmaScanArgs {
    const uint8_t* q_shadow;
    const uint8_t* x_shadow;
    const OperandStats* q_stats;
    const OperandStats* x_stats;
    const TileQ8* x_tiles;      // int8 path: per-128-row-tile constants of the segment (else nullptr)
    const QueryQ8* q_consts;    // int8 path: per-query constants
    const float* thr;       // [n_qtiles*128] unscaled emission thresholds
    const float* q_shift;   // [n_qtiles*128] q . c of a centred corpus image, or nullptr
    const float* center_norm;   // int8 path: pointer to ||c|| (device), or nullptr
    int d;
    int n_qtiles;
    int tile_major;         // 1: every CTA walks whole corpus tiles (all query tiles back to back); 0: units striped;
                            // 2: query-stationary CTA pairs (int8 screen with CTA pairs only, else as 1)
    int64_t ct0, ct1;       // 256-row tiles of the segment
    int64_t seg_rows;       // valid rows of the segment
    uint32_t row_id_base;
    CandBuf cb;
};
// cta_group = 1: one CTA per tile; 2: CTA pairs (cluster of 2) sharing each MMA; needs an even n_qtiles
cudaError_t launch_scan_mma(const MmaScanArgs& a, int sm_count, int cta_group, cudaStream_t s);
// int8 screen (kind::i8, s32 accumulators, integer thresholds); x_tiles / q_consts must be set
cudaError_t launch_scan_mma_i8(const MmaScanArgs& a, int sm_count, int cta_group, cudaStream_t s);
cudaError_t scan_mma_configure();

// ---- merge / gather ---------------------------------------------------------------------------
cudaError_t launch_merge_topk(int n_lists, int64_t nq, int k, const float* D_lists, const int64_t* I_lists,
                              int k_out, float* D_out, int64_t* I_out, cudaStream_t s);
cudaError_t launch_merge_topk_peers(int n_lists, int64_t nq, int k, const float* const* D_ptrs,
                                    const int64_t* const* I_ptrs, int k_out, float* D_out, int64_t* I_out,
                                    cudaStream_t s);
void launch_gather_ids(const int64_t* table, int64_t table_n, const int64_t* ids, int64_t n, int64_t* out,
                       cudaStream_t s);
// per query: rank (1-based, 0 = none) and reciprocal rank of the first relevant pid in the deduplicated ranking
void launch_reciprocal_rank(const int64_t* pids, int64_t nq, int k, const int64_t* rel_ptr, const int64_t* rel_pids,
                            float* rr_out, int32_t* rank_out, cudaStream_t s);

}  // namespace hac
