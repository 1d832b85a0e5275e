// Tensor-core screens: int8 (default) or f16 scores on the 5th-gen tensor cores with a fused threshold filter.
//
//   f16 : D[queries x 256 rows] (fp32, TMEM) += Qhat[queries x 64]  * Xhat[256 x 64]^T   per K block, 12 blocks for d = 768
//   int8: D[queries x 256 rows] (s32,  TMEM) += Qi  [queries x 128] * Xi  [256 x 128]^T  per K block,  6 blocks for d = 768
//   operands staged global -> shared by 16 KiB bulk async copies of the pre-swizzled shadow tiles (hac_common.cuh),
//   tcgen05.mma issued by one thread, accumulators double-buffered in TMEM (2 x 256 columns), epilogue warps read
//   them back with tcgen05.ld and compare every score against the owning query's emission threshold (int8: one
//   integer threshold per query and 128-row tile) - only the (rare) survivors are appended to the shortlist in HBM,
//   the score tile itself never leaves the SM.
//
// Template kCG (CTA group) x kI8 (operand kind):
//   kCG = 1  one CTA per SM works alone: tile 128 queries x 256 rows, 4 stages of 48 KiB.
//   kCG = 2  a CTA pair (cluster of 2, one TPC) shares each MMA (cta_group::2, M = 256): every CTA
//            stages its own 128 queries and HALF of the 256 corpus rows, so shared-memory fills and
//            operand reads per FLOP drop by a third; 6 stages of 32 KiB.  The pair's leader issues
//            the MMAs; the peer forwards "my half has landed" to the leader through a remote
//            mbarrier arrive; tcgen05.commit multicasts slot-free / accumulator-ready to both CTAs.
//   Defaults: int8 -> kCG = 2 (single-tile batches run kCG = 1), f16 -> kCG = 1 (power-bound either way).
// Warp roles (320 threads, persistent): warp 0 = copy producer, warp 1 = MMA issuer (leader) or
// forwarder (peer) and owner of the TMEM allocation, warps 2..9 = epilogue (TMEM lane quarter =
// warp % 4, one query per thread; warps 2..5 drain columns 0..127 of a tile, warps 6..9 columns 128..255).
// Unit order (UnitSchedule): tile-major for int8 (every CTA group takes whole 256-row corpus tiles and runs all
// query tiles of a tile back to back: one HBM fetch per tile, L2 re-reads from the same SMs), striped for f16
// (consecutive CTAs take the query tiles of one corpus tile) - each measured faster for its kind.
// Query-stationary form (kQS, int8 CTA pairs, "scan_tile_major" = 2): a CTA pair keeps ONE query tile pair for the whole
// launch - its int8 image (d/128 pieces of 16 KiB per CTA) is loaded once and stays in shared memory - and only the
// corpus pieces stream through an 8-stage ring.  The pairs form lanes of n_qgroups pairs that walk the same corpus
// tiles at the same time (one HBM fetch, L2 hits for the others), so the L2 -> SM operand traffic per MMA is half of
// the tile-major form's, which is what the power-capped clock pays for.
// Roofline: tensor pipe (HBM for batches of one query tile); algorithmic work = 2 * queries * rows * d per launch.
#include <limits.h>

#include <algorithm>

#include "hac_common.cuh"
#include "hac_kernels.cuh"

namespace hac {

namespace {

constexpr int kTileM = 128;                                // queries per CTA tile (TMEM lanes)
constexpr int kTileN = 256;                                // corpus rows per tile (TMEM columns)
constexpr int kThreads = 320;                              // producer warp, MMA warp, 8 epilogue warps
constexpr int kStash = 16;                                 // per-thread survivors kept until the TMEM buffer is released
constexpr int kMaxStages = 8;                              // barrier slots (pipeline stages)
constexpr int kSmemQS = 232448;                            // query-stationary form: all 227 KiB a CTA may have

template <int kCG>
struct Cfg {
    static constexpr int kStages = kCG == 1 ? 4 : 6;
    static constexpr int kABytes = kPieceBytes;                              // 128 queries x 64 k
    static constexpr int kBBytes = kCG == 1 ? 2 * kPieceBytes : kPieceBytes; // 256 or 128 rows x 64 k
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

struct Barriers {
    uint64_t full[kMaxStages];       // this CTA's operand bytes of the stage have landed
    uint64_t peer_full[kMaxStages];  // (leader only) the peer CTA's bytes have landed
    uint64_t empty[kMaxStages];      // the MMAs reading the stage have retired
    uint64_t tmem_full[2];           // accumulator buffer complete
    uint64_t tmem_empty[2];          // (leader only when kCG = 2) accumulator buffer drained
    uint64_t a_full;                 // query-stationary form: this CTA's resident query pieces have landed
    uint64_t peer_a_full;            // (leader only) the peer CTA's resident query pieces have landed
    uint32_t tmem_base;
};

// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 21); ++spin) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x2000;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// integer emission threshold of the int8 screen for one (query, 128-row tile).  The image holds v = fl(x - c)
// (c = 0 without a centre), q ~= t*qi, v ~= alpha*xi, A = qi.xi (exact s32):
//   fl32(q.x) <= t*alpha*A + shift + eps*beta + nhat*gamma + E      [Cauchy-Schwarz on the two quantisation errors]
//   E <= ((d/32 + 8) * 2^-23 + 2^-23) * ||q|| * (beta + ||c||)      [rounding of the exact fp32 dot, of x - c, of q.c]
// so a row can only matter if  A >= (thr - shift - eps*beta - nhat*gamma - E) / (t*alpha).
__device__ __forceinline__ int i8_threshold(float thr, float shift, float cnorm, const QueryQ8& qc, const TileQ8& tile,
                                            int d) {
    if (thr == INFINITY) return INT_MAX;                 // padded query: never emits
    if (thr == -INFINITY) return INT_MIN;                // no threshold yet: everything is emitted
    const float slack = (float)(d / 32 + 16) * 1.2e-7f * qc.norm * (tile.beta + cnorm) + 1e-6f * (fabsf(thr) + fabsf(shift));
    const float num = thr - shift - (qc.eps * tile.beta + qc.nhat * tile.gamma) - slack;
    const float den = qc.t * tile.alpha;
    if (!(den > 0.f)) return num <= 0.f ? INT_MIN : INT_MAX;
    float tf = num / den;
    tf -= fabsf(tf) * 1e-5f + 1.f;                       // conservative under fp32 rounding: may only lower it
    if (tf <= -2.0e9f) return INT_MIN;
    if (tf >= 2.0e9f) return INT_MAX;
    return (int)floorf(tf);
}

// 3-input float max (FMNMX3 on sm_100); NaNs are ignored like fmaxf, so a NaN score never passes a threshold
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Order in which one CTA group walks its share of the (corpus tile, query group) units of a launch.
// Body: the corpus tiles are dealt out whole - group b takes tiles b, b+G, b+2G, ... and runs ALL query groups of a
// tile back to back, so a tile is pulled from HBM once and re-read from L2 by the same SM a few microseconds later,
// whatever the other CTAs are doing.  (Striping the units of one tile over 20 neighbouring CTAs relies on those CTAs
// staying in step; the int8 scan, whose units are half as long, drifted apart and re-fetched every tile ~7 times:
// ncu DRAM read 93.6 GB for 13.2 GB of rows, L2 hit rate 64 %.)
// Tail: the last (n_ct mod G) tiles, fewer than one per group, are striped unit by unit for load balance.
// Query-stationary (mode 2): group b keeps query group b % n_qg and belongs to lane b / n_qg of G / n_qg lanes; lane l takes
// the corpus tiles l, l + lanes, l + 2 lanes, ... - the n_qg groups of a lane read the same tile at the same time.
struct UnitSchedule {
    int64_t G, b, body_rounds, body_units, n_mine;
    int n_qg;
    int g_div, g_mod;                // G = g_div * n_qg + g_mod: one striped step advances g_div tiles and g_mod query groups
    bool qs;
    int64_t qs_lanes, qs_lane;
    int qs_qg;
    __device__ UnitSchedule(int64_t n_ct, int n_qgroups, int64_t n_groups, int64_t group, int mode) {
        G = n_groups; b = group; n_qg = n_qgroups;
        qs = mode == 2;
        const bool tile_major = mode != 0;
        if (qs) {
            qs_lanes = G / n_qg;
            qs_lane = b / n_qg;
            qs_qg = (int)(b - qs_lane * n_qg);
            body_rounds = 0;
            n_mine = (qs_lane < qs_lanes && qs_lane < n_ct) ? (n_ct - qs_lane + qs_lanes - 1) / qs_lanes : 0;
            body_units = n_mine;
            g_div = g_mod = 0;
            return;
        }
        qs_lanes = qs_lane = 0;
        qs_qg = 0;
        body_rounds = tile_major ? n_ct / G : 0;
        body_units = body_rounds * n_qg;
        const int64_t tail_units = (n_ct - body_rounds * G) * n_qg;
        n_mine = body_units + (tail_units > b ? (tail_units - b + G - 1) / G : 0);
        g_div = (int)(G / n_qg);
        g_mod = (int)(G - (int64_t)g_div * n_qg);
    }
    // i-th unit of this group -> (corpus tile relative to ct0, query group)
    __device__ __forceinline__ void get(int64_t i, int64_t& ct_rel, int& qg) const {
        if (qs) {
            ct_rel = qs_lane + i * qs_lanes;
            qg = qs_qg;
        } else if (i < body_units) {
            const int64_t r = i / n_qg;
            ct_rel = r * G + b;
            qg = (int)(i - r * n_qg);
        } else {
            const int64_t u = b + (i - body_units) * G;
            const int64_t t = u / n_qg;
            ct_rel = body_rounds * G + t;
            qg = (int)(u - t * n_qg);
        }
    }
};

// Walks a CTA group's units in order.  UnitSchedule::get costs two 64-bit divisions (~200 instructions); every one of
// the 256 epilogue threads needs the current and the next unit, which made the bookkeeping of a unit three times
// as long as the drain of its accumulator columns.  The iterator steps incrementally instead (additions only; one
// get() at the start and one where the striped tail begins).
struct UnitIter {
    const UnitSchedule& s;
    int64_t i, ct_rel;
    int qg;
    __device__ explicit UnitIter(const UnitSchedule& sched) : s(sched), i(0), ct_rel(0), qg(0) {
        if (s.n_mine > 0) s.get(0, ct_rel, qg);
    }
    __device__ __forceinline__ bool valid() const { return i < s.n_mine; }
    __device__ __forceinline__ void next() {
        ++i;
        if (i >= s.n_mine) return;
        if (s.qs) {
            ct_rel += s.qs_lanes;
        } else if (i == s.body_units) {
            s.get(i, ct_rel, qg);
        } else if (i < s.body_units) {           // tile-major body: all query groups of a tile, then the tile G further
            if (++qg == s.n_qg) { qg = 0; ct_rel += s.G; }
        } else {                                 // striped tail: unit index advances by G
            ct_rel += s.g_div;
            qg += s.g_mod;
            if (qg >= s.n_qg) { qg -= s.n_qg; ++ct_rel; }
        }
    }
};

template <int kCG, bool kI8, bool kQS = false>
__global__ void __launch_bounds__(kThreads, 1) scan_mma_kernel(const MmaScanArgs a) {
    using C = Cfg<kCG>;
    static_assert(!kQS || (kCG == 2 && kI8), "the query-stationary form exists for int8 CTA pairs");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_count = a.d / (kI8 ? kBlockK8 : kBlockK);
    // query-stationary: [resident query pieces][ring of corpus pieces]; else [ring of (query, corpus) stages]
    const int resident_bytes = kQS ? kb_count * kPieceBytes : 0;
    const int stage_bytes = kQS ? kPieceBytes : C::kStageBytes;
    const int n_stages = kQS ? min(kMaxStages, (kSmemQS - 1024 - 512 - resident_bytes) / kPieceBytes) : C::kStages;
    uint8_t* ring = smem + resident_bytes;
    Barriers* bars = reinterpret_cast<Barriers*>(ring + n_stages * stage_bytes);
    const int n_bar_slots = n_stages;

    const uint32_t cta_rank = kCG == 2 ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    // work unit: (corpus tile of 256 rows) x (query tile of 128*kCG queries); one unit per CTA group
    const int n_qgroups = a.n_qtiles / kCG;
    const UnitSchedule sched(a.ct1 - a.ct0, n_qgroups, gridDim.x / kCG, blockIdx.x / kCG, kQS ? 2 : (a.tile_major != 0 ? 1 : 0));

    if constexpr (kCG == 2) cluster_sync_all();          // both CTAs resident before any cross-CTA traffic
    if (threadIdx.x == 0) {
        for (int i = 0; i < n_bar_slots; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->peer_full[i], 1);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tmem_full[i], 1);
            mbar_init(&bars->tmem_empty[i], 256 * kCG);      // every epilogue thread of the CTA group arrives
        }
        mbar_init(&bars->a_full, 1);
        mbar_init(&bars->peer_a_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<kCG>(&bars->tmem_base, 512);
    tc_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ---------------- producer: shadow pieces -> shared memory ----------------
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            if constexpr (kQS) {
                if (sched.n_mine > 0) {                  // the pair's query tile: loaded once, resident for the launch
                    const int qt = sched.qs_qg * kCG + (int)cta_rank;
                    const uint8_t* srcA = a.q_shadow + (size_t)qt * kb_count * kPieceBytes;
                    mbar_arrive_expect_tx(&bars->a_full, (uint32_t)resident_bytes);
                    for (int kb = 0; kb < kb_count; ++kb)
                        bulk_g2s(smem + kb * kPieceBytes, srcA + (size_t)kb * kPieceBytes, kPieceBytes, &bars->a_full);
                }
            }
            for (UnitIter u(sched); u.valid(); u.next()) {
                const int64_t ct = a.ct0 + u.ct_rel;
                const int qt = u.qg * kCG + (int)cta_rank;
                const uint8_t* srcA = a.q_shadow + (size_t)qt * kb_count * kPieceBytes;
                // kCG = 1: both 128-row halves of the corpus tile; kCG = 2: this CTA's half only
                const uint8_t* srcB = a.x_shadow + (size_t)(2 * ct + (kCG == 2 ? cta_rank : 0)) * kb_count * kPieceBytes;
                for (int kb = 0; kb < kb_count; ++kb) {
                    wait_or_trap(&bars->empty[stage], phase ^ 1);
                    if constexpr (kQS) {
                        mbar_arrive_expect_tx(&bars->full[stage], kPieceBytes);
                        bulk_g2s(ring + stage * kPieceBytes, srcB + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                    } else {
                        uint8_t* sA = ring + stage * C::kStageBytes;
                        uint8_t* sB = sA + C::kABytes;
                        mbar_arrive_expect_tx(&bars->full[stage], C::kStageBytes);
                        bulk_g2s(sA, srcA + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                        bulk_g2s(sB, srcB + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                        if constexpr (kCG == 1)
                            bulk_g2s(sB + kPieceBytes, srcB + (size_t)(kb_count + kb) * kPieceBytes, kPieceBytes,
                                     &bars->full[stage]);
                    }
                    if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ---------------- MMA issuer ----------------
            if (elect_one()) {
                constexpr uint32_t idesc = kI8 ? umma_idesc_i8(kTileM * kCG, kTileN) : umma_idesc_f16(kTileM * kCG, kTileN);
                uint32_t stage = 0, phase = 0, it = 0;
                if constexpr (kQS) {
                    if (sched.n_mine > 0) {
                        wait_or_trap(&bars->a_full, 0);
                        wait_or_trap(&bars->peer_a_full, 0);
                    }
                }
                for (int64_t i = 0; i < sched.n_mine; ++i, ++it) {
                    const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                    wait_or_trap(&bars->tmem_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * kTileN;
                    for (int kb = 0; kb < kb_count; ++kb) {
                        wait_or_trap(&bars->full[stage], phase);
                        if constexpr (kCG == 2) wait_or_trap(&bars->peer_full[stage], phase);
                        tc_fence_after();
                        const uint32_t sA = kQS ? smem_u32(smem + kb * kPieceBytes) : smem_u32(ring + stage * C::kStageBytes);
                        const uint32_t sB = kQS ? smem_u32(ring + stage * kPieceBytes) : sA + C::kABytes;
                        const uint64_t descA = umma_desc_k128(sA);
                        const uint64_t descB = umma_desc_k128(sB);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // advance 32 B along K inside the swizzle atom (16 f16 or 32 int8 elements): +2 in the
                            // >>4 address field
                            if constexpr (kI8) umma_i8<kCG>(tmem_d, descA + 2 * k, descB + 2 * k, idesc, (kb | k) != 0);
                            else umma_f16<kCG>(tmem_d, descA + 2 * k, descB + 2 * k, idesc, (kb | k) != 0);
                        }
                        // slot free / accumulator ready, signalled when the MMAs above retire
                        if constexpr (kCG == 1) {
                            umma_commit(&bars->empty[stage]);
                            if (kb == kb_count - 1) umma_commit(&bars->tmem_full[acc]);
                        } else {
                            umma_commit_2cta(&bars->empty[stage], 0b11);
                            if (kb == kb_count - 1) umma_commit_2cta(&bars->tmem_full[acc], 0b11);
                        }
                        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else {
            // ---------------- peer forwarder: my half of the stage has landed ----------------
            if (elect_one()) {
                uint32_t stage = 0, phase = 0;
                if constexpr (kQS) {
                    if (sched.n_mine > 0) {
                        wait_or_trap(&bars->a_full, 0);
                        mbar_arrive_cluster(&bars->peer_a_full, 0);
                    }
                }
                for (int64_t i = 0; i < sched.n_mine; ++i) {
                    for (int kb = 0; kb < kb_count; ++kb) {
                        wait_or_trap(&bars->full[stage], phase);
                        mbar_arrive_cluster(&bars->peer_full[stage], 0);
                        if (++stage == (uint32_t)n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ---------------- epilogue: fused threshold filter ----------------
        const int quarter = warp & 3;                    // TMEM lanes [32*quarter, 32*quarter+32)
        const int half = (warp - 2) >> 2;                // columns [128*half, 128*half+128) of every tile
        float scale = 1.f, inv_scale = 1.f;
        if constexpr (!kI8) {
            scale = a.q_stats->scale * a.x_stats->scale;
            inv_scale = a.q_stats->inv_scale * a.x_stats->inv_scale;
        }
        float cnorm = 0.f;
        if constexpr (kI8) {
            if (a.center_norm != nullptr) cnorm = *a.center_norm;
        }
        unsigned long long emitted = 0;
        uint32_t it = 0;
        // Survivors are first stashed per thread (local memory) and appended to the shortlist only
        // AFTER the accumulator buffer has been handed back to the MMA warp, so the round trip of the
        // global atomic overlaps the next tile's MMAs instead of stalling the TMEM pipeline.
        // The append itself is deferred by one more tile: the atomic that reserves the slots is issued when a
        // tile ends, its result is first used when the NEXT tile ends, so no warp ever waits for it.
        float stash_v[2][kStash];
        uint32_t stash_r[2][kStash];
        uint32_t pend_n = 0, pend_base = 0;
        int pend_q = 0;
        auto flush_pending = [&](int buf) {
            if (pend_n != 0) {
                if (pend_base + pend_n > a.cb.cap) *a.cb.overflow = 1u;
                for (uint32_t i = 0; i < pend_n; ++i) {
                    if (pend_base + i < a.cb.cap) {
                        a.cb.score[(size_t)pend_q * a.cb.cap + pend_base + i] = stash_v[buf][i];
                        a.cb.row[(size_t)pend_q * a.cb.cap + pend_base + i] = stash_r[buf][i];
                    }
                }
                pend_n = 0;
            }
        };
        // Per-unit constants (thresholds) are fetched one unit ahead, so their global-load latency is
        // hidden behind the TMEM drain of the current tile.
        struct UnitConsts {
            float thr;
            float shift;             // q . c of a centred corpus image, else 0
            QueryQ8 qc;
            TileQ8 t0, t1;
        };
        auto fetch = [&](int64_t ct_rel, int qg) {
            UnitConsts c;
            const int64_t ct = a.ct0 + ct_rel;
            const int q = (qg * kCG + (int)cta_rank) * kTileM + quarter * 32 + lane;
            // the thresholds rise WHILE this launch runs (the previous chunks' rescore / refresh work on a side
            // stream): read them past the non-coherent L1; a stale value is a valid (lower) threshold
            c.thr = __ldcg(a.thr + q);
            c.shift = a.q_shift != nullptr ? a.q_shift[q] : 0.f;
            if constexpr (kI8) {
                c.qc = a.q_consts[q];
                c.t0 = a.x_tiles[2 * ct];
                c.t1 = a.x_tiles[2 * ct + 1];
            }
            return c;
        };
        UnitConsts cur{};
        UnitIter u(sched), ahead(sched);
        if (u.valid()) cur = fetch(u.ct_rel, u.qg);
        ahead.next();
        for (; u.valid(); u.next(), ahead.next(), ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int64_t ct = a.ct0 + u.ct_rel;
            const int qt = u.qg * kCG + (int)cta_rank;
            const int q = qt * kTileM + quarter * 32 + lane;
            // threshold in accumulator units (power-of-two scale); the accumulator lacks the per-query constant
            // q.c of a centred image, so it is taken off the threshold (rounded down: may only lower it) and added
            // back to the scores that are stored
            const float shift = cur.shift;
            float thr_c = cur.thr - shift;
            if (shift != 0.f) thr_c -= (fabsf(cur.thr) + fabsf(shift)) * 2.4e-7f;
            const float thr_s = thr_c * scale;
            const int64_t row0 = ct * kTileN;
            // int8: one integer threshold and one de-quantisation factor per 128-row half of the tile
            int thr_i[2] = {0, 0};
            float deq[2] = {0.f, 0.f};
            if constexpr (kI8) {
                thr_i[0] = i8_threshold(cur.thr, shift, cnorm, cur.qc, cur.t0, a.d);
                thr_i[1] = i8_threshold(cur.thr, shift, cnorm, cur.qc, cur.t1, a.d);
                deq[0] = cur.qc.t * cur.t0.alpha;
                deq[1] = cur.qc.t * cur.t1.alpha;
            }
            UnitConsts nxt{};
            if (ahead.valid()) nxt = fetch(ahead.ct_rel, ahead.qg);
            uint32_t stash_n = 0;
            // one 32-column group of this thread's query: compare, stash or append the survivors
            auto process = [&](const uint32_t (&v)[32], int c) {
                const int thr_c = thr_i[c >> 2];
                const float out_scale = kI8 ? deq[c >> 2] : inv_scale;
                auto passes = [&](uint32_t bits) {
                    if constexpr (kI8) return (int)bits >= thr_c;
                    else return __uint_as_float(bits) >= thr_s;
                };
                auto value = [&](uint32_t bits) {
                    if constexpr (kI8) return fmaf((float)(int)bits, out_scale, shift);
                    else return fmaf(__uint_as_float(bits), out_scale, shift);
                };
                // The common case is "no survivor in these 32 columns".  Four independent chains of eight columns, one
                // 3-input max per two columns, and a single compare: 18 instructions, dependency depth 6.  (One chain
                // of 32 is one instruction shorter and measurably slower - the drain is latency-bound: 37.8 vs 35.4 ms
                // of scan, profiles/r02_ab_scan_epilogue_hybrid.jsonl.)  The octet maxima also tell the survivor
                // search below where to look.
                bool any;
                bool hit[4] = {true, true, true, true};        // which column octets hold a survivor
                if constexpr (kI8) {
                    int m[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        m[g] = __vimax3_s32((int)v[8 * g], (int)v[8 * g + 1], (int)v[8 * g + 2]);
                        m[g] = __vimax3_s32(m[g], (int)v[8 * g + 3], (int)v[8 * g + 4]);
                        m[g] = __vimax3_s32(m[g], (int)v[8 * g + 5], (int)v[8 * g + 6]);
                        m[g] = max(m[g], (int)v[8 * g + 7]);
                    }
                    any = max(__vimax3_s32(m[0], m[1], m[2]), m[3]) >= thr_c;
                    if (any) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) hit[g] = m[g] >= thr_c;
                    }
                } else {
                    float m[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        m[g] = fmax3(__uint_as_float(v[8 * g]), __uint_as_float(v[8 * g + 1]), __uint_as_float(v[8 * g + 2]));
                        m[g] = fmax3(m[g], __uint_as_float(v[8 * g + 3]), __uint_as_float(v[8 * g + 4]));
                        m[g] = fmax3(m[g], __uint_as_float(v[8 * g + 5]), __uint_as_float(v[8 * g + 6]));
                        m[g] = fmaxf(m[g], __uint_as_float(v[8 * g + 7]));
                    }
                    any = fmaxf(fmax3(m[0], m[1], m[2]), m[3]) >= thr_s;
                    if (any) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) hit[g] = m[g] >= thr_s;
                    }
                }
                // Survivors are rare once a threshold exists (a few per 100 000 columns), but a warp takes this branch
                // whenever ONE of its 32 queries has one - every seventh 32-column group at k = 100 - so its length
                // decides how often a tile's drain outlasts the next tile's MMAs (the round-1 code built a 32-bit mask and
                // walked 32 predicated stores, ~300 instructions: 40.4 vs 35.8 ms of scan, k = 1 vs k = 100 34.4 vs
                // 38.6 ms).  Sparse case: only the octets that hold a survivor are expanded, straight into the stash.
                // (Narrowing further to quads is slower again - more branches: 36.6 vs 35.8 ms.)
                const int n_hit = (int)hit[0] + (int)hit[1] + (int)hit[2] + (int)hit[3];
                if (any && n_hit < 4 && stash_n + 8u * (uint32_t)n_hit <= (uint32_t)kStash &&
                    a.seg_rows - (row0 + c * 32) >= 32) {
                    const uint32_t row_id = a.row_id_base + (uint32_t)(row0 + c * 32);
                    const uint32_t before = stash_n;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (hit[g]) {
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                const int j = 8 * g + jj;
                                if (passes(v[j])) {
                                    stash_v[it & 1][stash_n] = value(v[j]);
                                    stash_r[it & 1][stash_n] = row_id + j;
                                    ++stash_n;
                                }
                            }
                        }
                    }
                    emitted += stash_n - before;
                } else if (any) {
                    // dense case (no threshold yet, ragged segment end, stash full): mask of the 32 columns
                    uint32_t mask = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (passes(v[j]) ? 1u : 0u) << j;
                    const int64_t rem = a.seg_rows - (row0 + c * 32);      // rows past the segment end are padding
                    if (rem < 32) mask &= rem <= 0 ? 0u : ((1u << rem) - 1u);
                    const uint32_t n = __popc(mask);
                    const uint32_t row_id = a.row_id_base + (uint32_t)(row0 + c * 32);
                    if (n != 0 && stash_n + n <= (uint32_t)kStash) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if ((mask >> j) & 1u) {
                                stash_v[it & 1][stash_n] = value(v[j]);
                                stash_r[it & 1][stash_n] = row_id + j;
                                ++stash_n;
                            }
                        }
                    } else if (n != 0) {
                        // loose-threshold phase (first chunks): one atomic per 32-column group
                        const uint32_t base = atomicAdd(a.cb.count + q, n);
                        if (base + n > a.cb.cap) *a.cb.overflow = 1u;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t pos = base + __popc(mask & ((1u << j) - 1u));
                            if (((mask >> j) & 1u) && pos < a.cb.cap) {
                                a.cb.score[(size_t)q * a.cb.cap + pos] = value(v[j]);
                                a.cb.row[(size_t)q * a.cb.cap + pos] = row_id + j;
                            }
                        }
                    }
                    emitted += n;
                }
                __syncwarp();
            };
            wait_or_trap(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kTileN;
            // Two epilogue warps share each TMEM lane quarter, one per 128-column half of the tile: the drain of a tile
            // is a chain of TMEM-load latencies (ncu, int8 scan with one warp per quarter: tensor pipe 72 % active in
            // both CTA-group variants - the MMAs of a tile take 3072 cycles, its drain ~4300), so halving the chain per
            // warp is what takes the epilogue off the critical path.
            // two register buffers: the next 32 columns are in flight while the current ones are compared
            constexpr int kGroupsPerWarp = kTileN / 32 / 2;
            const int c0 = half * kGroupsPerWarp;
            uint32_t va[32], vb[32];
            tmem_ld_32x32(taddr + c0 * 32, va);
            tmem_ld_wait();
#pragma unroll 1
            for (int c = c0; c < c0 + kGroupsPerWarp; c += 2) {
                tmem_ld_32x32(taddr + (c + 1) * 32, vb);
                process(va, c);
                tmem_ld_wait();
                if (c + 2 < c0 + kGroupsPerWarp) tmem_ld_32x32(taddr + (c + 2) * 32, va);
                process(vb, c + 1);
                tmem_ld_wait();
            }
            tc_fence_before();
            if constexpr (kCG == 1) mbar_arrive(&bars->tmem_empty[acc]);
            else mbar_arrive_cluster(&bars->tmem_empty[acc], 0);       // the leader's barrier counts both CTAs
            flush_pending((it & 1) ^ 1);                  // the previous tile's survivors: their slots have arrived
            if (stash_n != 0) {
                pend_base = atomicAdd(a.cb.count + q, stash_n);   // result consumed one tile later
                pend_n = stash_n;
                pend_q = q;
            }
            __syncwarp();
            cur = nxt;
        }
        flush_pending((it & 1) ^ 1);
        if (emitted) atomicAdd(a.cb.emitted, emitted);
    }

    tc_fence_before();
    if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) tmem_dealloc<kCG>(tmem_base, 512);
}

}  // namespace

cudaError_t scan_mma_configure() {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kernel, int bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set(scan_mma_kernel<1, false>, Cfg<1>::kSmemBytes);
    set(scan_mma_kernel<1, true>, Cfg<1>::kSmemBytes);
    set(scan_mma_kernel<2, false>, Cfg<2>::kSmemBytes);
    set(scan_mma_kernel<2, true>, Cfg<2>::kSmemBytes);
    set(scan_mma_kernel<2, true, true>, kSmemQS);
    return e;
}

namespace {
// cluster-of-2 launch: one CTA pair per unit, pairs <= sm_count / 2
template <bool kI8>
cudaError_t launch_pairs(const MmaScanArgs& a, int sm_count, cudaStream_t s) {
    const int64_t n_units = (a.ct1 - a.ct0) * (a.n_qtiles / 2);
    const int pairs = (int)(n_units < sm_count / 2 ? n_units : sm_count / 2);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg<2>::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if constexpr (kI8) {
        // query-stationary form: lanes of n_qgroups pairs; needs at least one full lane and a ring of >= 4 stages
        const int n_qg = a.n_qtiles / 2;
        const int64_t lanes = std::min<int64_t>((sm_count / 2) / n_qg, a.ct1 - a.ct0);
        const int ring = (kSmemQS - 1024 - 512 - (a.d / kBlockK8) * kPieceBytes) / kPieceBytes;
        if (a.tile_major == 2 && lanes >= 1 && ring >= 4) {
            cfg.gridDim = dim3((unsigned)(2 * lanes * n_qg));
            cfg.dynamicSmemBytes = kSmemQS;
            return cudaLaunchKernelEx(&cfg, scan_mma_kernel<2, true, true>, a);
        }
    }
    return cudaLaunchKernelEx(&cfg, scan_mma_kernel<2, kI8>, a);
}
}  // namespace

cudaError_t launch_scan_mma_i8(const MmaScanArgs& a, int sm_count, int cta_group, cudaStream_t s) {
    const int64_t n_units = (a.ct1 - a.ct0) * a.n_qtiles;
    if (n_units <= 0) return cudaSuccess;
    if (a.x_tiles == nullptr || a.q_consts == nullptr || a.d % kBlockK8 != 0) return cudaErrorInvalidValue;
    if (cta_group == 2 && (a.n_qtiles % 2) == 0) return launch_pairs<true>(a, sm_count, s);
    const int grid = (int)(n_units < sm_count ? n_units : sm_count);
    scan_mma_kernel<1, true><<<grid, kThreads, Cfg<1>::kSmemBytes, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scan_mma(const MmaScanArgs& a, int sm_count, int cta_group, cudaStream_t s) {
    const int64_t n_ctiles = a.ct1 - a.ct0;
    if (n_ctiles <= 0 || a.n_qtiles <= 0) return cudaSuccess;
    if (a.tile_major == 2) {                 // the query-stationary order is an int8 form: the f16 scan keeps its default
        MmaScanArgs b = a;
        b.tile_major = 0;
        return launch_scan_mma(b, sm_count, cta_group, s);
    }
    if (cta_group == 2 && (a.n_qtiles % 2) == 0) return launch_pairs<false>(a, sm_count, s);
    const int64_t n_units = n_ctiles * a.n_qtiles;
    const int grid = (int)(n_units < sm_count ? n_units : sm_count);
    scan_mma_kernel<1, false><<<grid, kThreads, Cfg<1>::kSmemBytes, s>>>(a);
    return cudaGetLastError();
}

}  // namespace hac
