// Selection side of the search: per-query shortlist refresh (k-th best screen score -> new
// emission threshold), exact fp32 rescore of the shortlist, final top-k, cross-shard merge.
// The full score matrix never exists: these kernels only ever see the few hundred candidates
// per query that passed the fused filter of the scan kernels.
#include <float.h>
#include <limits.h>

#include <algorithm>

#include "hac_common.cuh"
#include "hac_kernels.cuh"

namespace hac {

// ---------------------------------------------------------------------------------------------
// shared-memory bitonic sort, descending, P a power of two
template <typename K>
__device__ __forceinline__ void bitonic_sort_desc(K* a, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const K x = a[lo], y = a[hi];
                if ((x < y) == desc) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

__device__ __forceinline__ uint64_t cand_key(float score, uint32_t row) {
    return ((uint64_t)float_key(score) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}

// ---------------------------------------------------------------------------------------------
__global__ void init_search_kernel(CandBuf cb, float* tau, float* thr, int nq, int nq_pad) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) {
        *cb.overflow = 0u;
        *cb.emitted = 0ull;
    }
    if (q >= nq_pad) return;
    cb.count[q] = 0u;
    cb.sorted[q] = 0u;
    tau[q] = -INFINITY;
    thr[q] = q < nq ? -INFINITY : INFINITY;   // padded queries never emit
}

void launch_init_search(CandBuf cb, float* tau, float* thr, int nq, int nq_pad, cudaStream_t s) {
    init_search_kernel<<<(nq_pad + 255) / 256, 256, 0, s>>>(cb, tau, thr, nq, nq_pad);
}

// m_q bounds |screen score - exact fp32 score| for every row of the index.  With the corpus image centred on c
// (hac_prep.cu) the screen score is  sum qhat*xhat + fl(q.c)  and xhat approximates v = fl(x - c):
//   |sum qhat*xhat - q.(x-c)| <= ||q - qhat||*||xhat|| + ||q||*||v - xhat||              (Cauchy-Schwarz)
//                                + 2^-24 * ||q|| * ||x - c||                              (rounding of x - c)
//   tensor-core fp32 accumulation (truncating, any order):      <= d * 2^-23 * ||qhat|| * ||xhat||
//   rounding of the exact fp32 dot (per-lane chain of d/32 FMAs + 5 butterfly adds, hac_common.cuh; the generic
//   kernel's chain is no longer):                               <= (d/32 + 5) * 2^-24 * ||q|| * ||x||
//   fl(q.c), the subtraction thr - q.c and the addition a + q.c: <= 3 * 2^-24 * ||q|| * (||c|| + ||x||)
// All norms of the statistics are those of the centred rows; ||x|| <= ||x - c|| + ||c||.  Every term is taken
// with a factor >= 2 of safety.
__global__ void margins_kernel(const float* q_norm, const float* q_err, const OperandStats* corpus,
                               const float* center_norm, int d, float* margin, float* margin_max, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    float m = 0.f;
    if (corpus != nullptr) {
        const float qn = q_norm[q], qe = q_err[q];
        const float cn = center_norm != nullptr ? *center_norm : 0.f;
        const float xc = fmaxf(corpus->norm_max, corpus->hat_norm_max);      // centred rows
        const float xu = xc + cn;                                            // uncentred rows
        m = qe * corpus->hat_norm_max + qn * corpus->err_norm_max
            + (float)d * 2.38418579e-7f * (qn + qe) * xc                     // d * 2^-22: tensor-core accumulation
            + (float)(d / 32 + 8) * 1.1920929e-7f * qn * xu                  // (d/32 + 8) * 2^-23: exact fp32 dot
            + 9.5367432e-7f * qn * (xu + cn);                                // 2^-20: centring round-offs
        m *= 1.001f;
    }
    margin[q] = m;
    atomicMax(reinterpret_cast<int*>(margin_max), __float_as_int(m));
}

__global__ void margins_i8_kernel(const QueryQ8* q_consts, const OperandStats* corpus, int d, float* margin,
                                  float* margin_max, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const QueryQ8 c = q_consts[q];
    const float m = c.eps * corpus->i8_beta_max + c.nhat * corpus->i8_gamma_max +
                    (float)d * 2.4e-7f * c.norm * corpus->i8_beta_max;
    margin[q] = m;
    atomicMax(reinterpret_cast<int*>(margin_max), __float_as_int(m));
}
void launch_margins_i8(const QueryQ8* q_consts, const OperandStats* corpus, int d, float* margin, float* margin_max,
                       int nq, cudaStream_t s) {
    cudaMemsetAsync(margin_max, 0, sizeof(float), s);
    margins_i8_kernel<<<(nq + 127) / 128, 128, 0, s>>>(q_consts, corpus, d, margin, margin_max, nq);
}

void launch_margins(const float* q_norm, const float* q_err, const OperandStats* corpus, const float* center_norm,
                    int d, float* margin, float* margin_max, int nq, cudaStream_t s) {
    cudaMemsetAsync(margin_max, 0, sizeof(float), s);
    margins_kernel<<<(nq + 127) / 128, 128, 0, s>>>(q_norm, q_err, corpus, center_norm, d, margin, margin_max, nq);
}

// ---------------------------------------------------------------------------------------------
// One CTA per query.  tau <- the k-th best screen score seen so far (a valid lower bound on the final
// k-th best screen score), found by an 8-bit MSB radix select over the order-preserving keys; every
// true top-k row then has screen score >= tau - 2m, so entries below thr = tau - 2m are dropped
// (block-wide stream compaction, in place) and later chunks only emit rows with score >= thr.
// No sort: the shortlist stays unordered until the final select.
constexpr int kRefreshMaxThreads = 1024;   // 256 per query for large batches, 1024 when a few queries must finish fast

__global__ void __launch_bounds__(kRefreshMaxThreads) refresh_kernel(CandBuf cb, int k,
                                                                  const float* __restrict__ margin,
                                                                  float* __restrict__ tau,
                                                                  float* __restrict__ thr, const ThrExchange ex,
                                                                  uint32_t* __restrict__ clear_count) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem_raw);          // [cap]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_want;
    __shared__ uint32_t warp_cnt[kRefreshMaxThreads / 32];
    __shared__ uint32_t s_base;
    __shared__ uint32_t s_top_n, s_best_key;
    __shared__ uint32_t top[1024];                                  // exchange: keys of this shard's best entries, sorted
    __shared__ uint32_t claim[kMaxPeerLists * kExRanks];            // exchange: claim[shard * n_ranks + j] = key or 0
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_threads = blockDim.x;
    // exchange: thread t < (n_peers + 1) * n_ranks owns the word (shard t / n_ranks, rank slot t % n_ranks); shard n_peers
    // is this shard itself.  The loads cross NVLink (~2 us), so they are issued first and consumed last.
    const int n_claims = ex.n_peers > 0 ? (ex.n_peers + 1) * ex.n_ranks : 0;
    unsigned long long word = 0ull;
    if (tid < n_claims) {
        const int sh = tid / ex.n_ranks, j = tid - sh * ex.n_ranks;
        const unsigned long long* src = (sh < ex.n_peers ? ex.peers[sh] : ex.mine) + (size_t)q * kExWords + j;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
    }
    if (tid == 0) {
        s_top_n = 0u; s_best_key = 0u;
        if (clear_count != nullptr) clear_count[q] = 0u;     // the log the preceding rescore consumed is free again
    }
    const uint32_t raw = cb.count[q];
    const int cnt = (int)min(raw, cb.cap);
    if (raw > cb.cap && tid == 0) *cb.overflow = 1u;
    if ((uint32_t)cnt == cb.sorted[q] && ex.n_peers == 0) return;   // nothing new since the last refresh (uniform per block)
    float* sc = cb.score + (size_t)q * cb.cap;
    uint32_t* rw = cb.row + (size_t)q * cb.cap;
    for (int i = tid; i < cnt; i += n_threads) keys[i] = float_key(sc[i]);
    // key of the want-th best entry (1-based, want <= cnt): 8-bit MSB radix select; every thread of the block calls it
    auto select_kth = [&](uint32_t want_rank) -> uint32_t {
        __syncthreads();
        if (tid == 0) { s_prefix = 0u; s_want = want_rank; }
        uint32_t mask = 0u;
#pragma unroll 1
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (tid < 256) hist[tid] = 0u;
            __syncthreads();
            const uint32_t prefix = s_prefix;
            for (int i = tid; i < cnt; i += n_threads) {
                const uint32_t key = keys[i];
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                // lane l owns bins [255-8l-7, 255-8l] (descending order across lanes)
                uint32_t local[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { local[j] = hist[255 - (lane * 8 + j)]; sum += local[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const uint32_t excl = incl - sum, want = s_want;
                if (excl < want && want <= incl) {                    // the wanted entry falls into this lane's bins
                    uint32_t cum = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (cum < want && want <= cum + local[j]) {
                            s_prefix = prefix | ((uint32_t)(255 - (lane * 8 + j)) << shift);
                            s_want = want - cum;
                        }
                        cum += local[j];
                    }
                }
            }
            mask |= 0xFFu << shift;
            __syncthreads();
        }
        return s_prefix;
    };
    float tau_new = tau[q];
    uint32_t kth_key = 0u;                       // key of this shard's k-th best entry (0: fewer than k entries)
    if (cnt >= k) {
        kth_key = select_kth((uint32_t)k);
        tau_new = fmaxf(tau_new, key_float(kth_key));
    }
    if (ex.n_peers > 0) {
        // Claims.  Shard r publishes L_r(c) = its c-th best exact score so far for a few fixed ranks c (ex.ranks, the
        // same on every shard): "shard r holds >= c rows scoring >= L_r(c)".  For any T, the shards together hold at
        // least  sum_r max{ c : L_r(c) >= T }  rows scoring >= T; the largest T for which that sum reaches k is a lower
        // bound on the global k-th best - about the global k-th best itself when the ranks resolve each shard's share,
        // where a shard's own k-th best only knows 1/G of the rows.  Claims only strengthen during a search and a
        // missing or stale word is simply no claim, so no ordering between the shards is needed.
        __syncthreads();
        for (int i = tid; i < cnt; i += n_threads) {
            const uint32_t key = keys[i];
            if (key >= kth_key) {
                const uint32_t pos = atomicAdd(&s_top_n, 1u);
                if (pos < 1024u) top[pos] = key;
            }
        }
        __syncthreads();
        const uint32_t m = s_top_n;                 // this shard's entries that can matter: its top-k (ties included)
        const bool have_top = m <= 1024u;
        if (have_top) {
            const int P = next_pow2((int)max(m, 2u));
            for (int i = (int)m + tid; i < P; i += n_threads) top[i] = 0u;
            bitonic_sort_desc(top, P);              // (block-wide; starts and ends with a barrier)
        }
        if (tid < n_claims) {
            const int sh = tid / ex.n_ranks, j = tid - sh * ex.n_ranks;
            uint32_t key = 0u;
            if ((uint32_t)(word >> 32) == ex.tag) key = float_key(__uint_as_float((uint32_t)word));
            if (sh == ex.n_peers && have_top && m >= (uint32_t)ex.ranks[j]) {
                const uint32_t now = top[ex.ranks[j] - 1];
                if (now > key) {                    // an earlier claim stays true; publish only a stronger one
                    key = now;
                    const unsigned long long w = ((unsigned long long)ex.tag << 32) |
                                                 (unsigned long long)__float_as_uint(key_float(key));
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(ex.mine + (size_t)q * kExWords + j), "l"(w)
                                 : "memory");
                }
            }
            claim[tid] = key;
        }
        __syncthreads();
        if (tid < n_claims && claim[tid] != 0u) {
            const uint32_t T = claim[tid];          // candidate threshold: does the sum of the claims reach k at T?
            int total = 0;
            for (int sh = 0; sh <= ex.n_peers; ++sh) {
                int best = 0;
                for (int j = 0; j < ex.n_ranks; ++j)
                    if (claim[sh * ex.n_ranks + j] >= T) best = max(best, ex.ranks[j]);
                total += best;
            }
            if (total >= k) atomicMax(&s_best_key, T);
        }
        __syncthreads();
        if (s_best_key != 0u) tau_new = fmaxf(tau_new, key_float(s_best_key));
    }
    float thr_new = -INFINITY;
    if (tau_new > -INFINITY) {
        thr_new = tau_new - 2.f * margin[q];
        thr_new -= fabsf(thr_new) * 1e-6f;       // keep the cut conservative under fp32 rounding
    }
    // in-place compaction of entries with score >= thr_new, one round of blockDim.x entries at a time
    const uint32_t thr_key = float_key(thr_new);
    if (tid == 0) s_base = 0u;
    __syncthreads();
    for (int r0 = 0; r0 < cnt; r0 += n_threads) {
        const int i = r0 + tid;
        uint32_t key = 0u, row = 0u;
        bool keep = false;
        if (i < cnt) {
            key = keys[i];
            row = rw[i];
            keep = thr_new == -INFINITY || key >= thr_key;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(ballot);
        __syncthreads();                                   // all reads of this round done, counts visible
        uint32_t off = s_base;
        for (int w = 0; w < warp; ++w) off += warp_cnt[w];
        if (keep) {
            const uint32_t pos = off + __popc(ballot & ((1u << lane) - 1u));
            sc[pos] = key_float(key);
            rw[pos] = row;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < (n_threads >> 5); ++w) t += warp_cnt[w];
            s_base += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        cb.count[q] = s_base;
        cb.sorted[q] = s_base;
        tau[q] = tau_new;
        thr[q] = thr_new;
    }
}

void launch_refresh(CandBuf cb, int k, const float* margin, float* tau, float* thr, int nq, cudaStream_t s,
                    const ThrExchange* ex, uint32_t* clear_count, bool small_cta) {
    const size_t smem = (size_t)cb.cap * sizeof(uint32_t);
    // the attribute is per device (several devices per process are possible), so it is set per launch
    if (smem > 40 * 1024) cudaFuncSetAttribute(refresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ThrExchange none{};
    const int threads = (nq <= 64 && !small_cta) ? kRefreshMaxThreads : 256;
    refresh_kernel<<<nq, threads, smem, s>>>(cb, k, margin, tau, thr, ex ? *ex : none, clear_count);
}

// ---------------------------------------------------------------------------------------------
// Exact fp32 score of every shortlisted (query,row) pair: one warp per pair, the same per-lane
// FMA order and butterfly as the GEMV scan (hac_common.cuh), so both paths agree bitwise.
__device__ __forceinline__ const float* seg_row_ptr(const SegTable& segs, uint32_t row, int d) {
    int s = 0;
#pragma unroll 1
    for (int i = 1; i < segs.n; ++i)
        if (row >= segs.base[i]) s = i;
    return segs.rows[s] + (size_t)(row - segs.base[s]) * d;
}

template <bool kNewOnly>
__global__ void __launch_bounds__(256) rescore_kernel(CandBuf cb, const float* __restrict__ qmat, int d,
                                                      SegTable segs, float* __restrict__ screen_err_max,
                                                      unsigned long long* __restrict__ rescored) {
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int cnt = (int)min(cb.count[q], cb.cap);
    const int first = kNewOnly ? (int)min(cb.sorted[q], (uint32_t)cnt) : 0;   // entries before it are exact already
    const int d4 = d >> 2;
    const float4* qrow = reinterpret_cast<const float4*>(qmat + (size_t)q * d);
    float worst = 0.f;
    for (int slot = first + warp; slot < cnt; slot += n_warps) {
        const uint32_t row = cb.row[(size_t)q * cb.cap + slot];
        const float4* xrow = reinterpret_cast<const float4*>(seg_row_ptr(segs, row, d));
        float acc = 0.f;
        for (int i = lane; i < d4; i += 32) acc = lane_fma4(acc, __ldg(qrow + i), __ldg(xrow + i));
        acc = warp_sum(acc);
        if (lane == 0) {
            cb.exact[(size_t)q * cb.cap + slot] = acc;
            worst = fmaxf(worst, fabsf(acc - cb.score[(size_t)q * cb.cap + slot]));
            if (kNewOnly) cb.score[(size_t)q * cb.cap + slot] = acc;   // the refresh then ranks exact scores
        }
    }
    if (lane == 0) {
        if (worst > 0.f) atomicMax(reinterpret_cast<int*>(screen_err_max), __float_as_int(worst));
        if (warp == 0) atomicAdd(rescored, (unsigned long long)(cnt - first));
    }
}

// d = 128 * VPL: the query slice lives in registers, two shortlisted rows are fetched per warp and iteration with
// all 2*VPL 16-byte loads issued up front (the generic kernel above serialises one load per loop trip and is
// latency-bound at ~2 TB/s; this one streams the random 3 KB rows at the HBM rate).  Same per-lane FMA order.
template <bool kNewOnly, int VPL>
__global__ void __launch_bounds__(128) rescore_vec_kernel(CandBuf cb, const float* __restrict__ qmat, SegTable segs,
                                                          float* __restrict__ screen_err_max,
                                                          unsigned long long* __restrict__ rescored) {
    constexpr int d = VPL * 128;
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int cnt = (int)min(cb.count[q], cb.cap);
    const int first = kNewOnly ? (int)min(cb.sorted[q], (uint32_t)cnt) : 0;
    float4 qv[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) qv[i] = __ldg(reinterpret_cast<const float4*>(qmat + (size_t)q * d) + i * 32 + lane);
    const uint32_t* rows = cb.row + (size_t)q * cb.cap;
    float worst = 0.f;
    // gridDim.y CTAs share one query's shortlist (small batches: the pairs, not the queries, fill the GPU).
    // The row indices and the screen scores of the NEXT iteration are fetched while the current rows are in flight,
    // so one global-memory latency per iteration is exposed, not two.
    const int step = 2 * n_warps * (int)gridDim.y;
    int slot = first + 2 * (warp + n_warps * (int)blockIdx.y);
    uint32_t ra = 0, rb = 0;
    float old_a = 0.f, old_b = 0.f;
    auto fetch_meta = [&](int sl) {
        if (sl < cnt) {
            const bool hb = sl + 1 < cnt;
            ra = rows[sl];
            rb = hb ? rows[sl + 1] : ra;
            if (lane == 0) {
                old_a = cb.score[(size_t)q * cb.cap + sl];
                old_b = hb ? cb.score[(size_t)q * cb.cap + sl + 1] : 0.f;
            }
        }
    };
    fetch_meta(slot);
    for (; slot < cnt; slot += step) {
        const bool has_b = slot + 1 < cnt;
        const size_t o = (size_t)q * cb.cap + slot;
        const float4* pa = reinterpret_cast<const float4*>(seg_row_ptr(segs, ra, d));
        const float4* pb = reinterpret_cast<const float4*>(seg_row_ptr(segs, rb, d));
        const float cur_a = old_a, cur_b = old_b;
        float4 xa[VPL], xb[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            xa[i] = __ldg(pa + i * 32 + lane);
            xb[i] = __ldg(pb + i * 32 + lane);
        }
        fetch_meta(slot + step);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            a = lane_fma4(a, qv[i], xa[i]);
            b = lane_fma4(b, qv[i], xb[i]);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            cb.exact[o] = a;
            worst = fmaxf(worst, fabsf(a - cur_a));
            if (kNewOnly) cb.score[o] = a;
            if (has_b) {
                cb.exact[o + 1] = b;
                worst = fmaxf(worst, fabsf(b - cur_b));
                if (kNewOnly) cb.score[o + 1] = b;
            }
        }
    }
    if (lane == 0) {
        if (worst > 0.f) atomicMax(reinterpret_cast<int*>(screen_err_max), __float_as_int(worst));
        if (warp == 0 && blockIdx.y == 0) atomicAdd(rescored, (unsigned long long)(cnt - first));
    }
}

template <bool kNewOnly>
static void launch_rescore_any(CandBuf cb, const float* q, int d, SegTable segs, int nq, float* screen_err_max,
                               unsigned long long* rescored, cudaStream_t s) {
    // about 32 CTAs of 4 warps per SM in total (all resident at once): with few queries each shortlist is split over
    // many CTAs, so a chunk's few thousand new pairs are fetched in one or two latencies instead of a serial loop
    // (measured at Q=1: 45-86 us per chunk with 64 CTAs, the largest non-scan item of the search)
    const int per_query = (int)std::min<int64_t>(4736 / std::max(nq, 1), ((int64_t)cb.cap + 7) / 8);
    const dim3 grid((unsigned)nq, (unsigned)std::max(1, per_query));
    switch (d) {
        case 128: rescore_vec_kernel<kNewOnly, 1><<<grid, 128, 0, s>>>(cb, q, segs, screen_err_max, rescored); break;
        case 256: rescore_vec_kernel<kNewOnly, 2><<<grid, 128, 0, s>>>(cb, q, segs, screen_err_max, rescored); break;
        case 512: rescore_vec_kernel<kNewOnly, 4><<<grid, 128, 0, s>>>(cb, q, segs, screen_err_max, rescored); break;
        case 768: rescore_vec_kernel<kNewOnly, 6><<<grid, 128, 0, s>>>(cb, q, segs, screen_err_max, rescored); break;
        case 1024: rescore_vec_kernel<kNewOnly, 8><<<grid, 128, 0, s>>>(cb, q, segs, screen_err_max, rescored); break;
        default: rescore_kernel<kNewOnly><<<nq, 256, 0, s>>>(cb, q, d, segs, screen_err_max, rescored); break;
    }
}

void launch_rescore(CandBuf cb, const float* q, int d, SegTable segs, int nq, float* screen_err_max,
                    unsigned long long* rescored, cudaStream_t s) {
    cudaMemsetAsync(screen_err_max, 0, sizeof(float), s);
    cudaMemsetAsync(rescored, 0, sizeof(unsigned long long), s);
    launch_rescore_any<false>(cb, q, d, segs, nq, screen_err_max, rescored, s);
}

// accumulates into screen_err_max / rescored (the caller zeroes them once per search)
void launch_rescore_new(CandBuf cb, const float* q, int d, SegTable segs, int nq, float* screen_err_max,
                        unsigned long long* rescored, cudaStream_t s) {
    launch_rescore_any<true>(cb, q, d, segs, nq, screen_err_max, rescored, s);
}

// ---------------------------------------------------------------------------------------------
// Pipelined int8 search, worker side.  The scan of a chunk appends its survivors (screen score, row) to a log shortlist;
// this kernel computes the exact fp32 score of every log entry (same per-lane FMA order as everywhere else) and moves
// the entries that can still reach the top-k - exact score >= tau[q], the k-th best exact score found so far - into the
// main shortlist.  It runs on a side stream WHILE the scan kernel of the next chunk occupies every SM, so its footprint
// is what fits beside that kernel (168 regs x 320 threads, 198 KB of shared memory per SM): 128 threads, <= 80
// registers (the query slice sits in 3 KB of shared memory instead of registers), two rows per warp in flight.
template <int VPL>
__global__ void __launch_bounds__(128, 6) rescore_log_kernel(CandBuf lg, CandBuf cb, const float* __restrict__ qmat,
                                                             SegTable segs, const float* __restrict__ tau,
                                                             float* __restrict__ screen_err_max,
                                                             unsigned long long* __restrict__ rescored) {
    constexpr int d = VPL * 128;
    __shared__ float4 qs[VPL * 32];
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const uint32_t raw = lg.count[q];
    const int cnt = (int)min(raw, lg.cap);
    if (cnt == 0) return;                                   // uniform per block
    if (raw > lg.cap && threadIdx.x == 0) *cb.overflow = 1u;
    for (int i = threadIdx.x; i < VPL * 32; i += blockDim.x)
        qs[i] = __ldg(reinterpret_cast<const float4*>(qmat + (size_t)q * d) + i);
    __syncthreads();
    const float tau_q = tau[q];
    const uint32_t* rows = lg.row + (size_t)q * lg.cap;
    const float* old = lg.score + (size_t)q * lg.cap;
    float worst = 0.f;
    const int step = 2 * n_warps * (int)gridDim.y;
    int slot = 2 * (warp + n_warps * (int)blockIdx.y);
    uint32_t ra = 0, rb = 0;
    float old_a = 0.f, old_b = 0.f;
    auto fetch_meta = [&](int sl) {
        if (sl < cnt) {
            const bool hb = sl + 1 < cnt;
            ra = rows[sl];
            rb = hb ? rows[sl + 1] : ra;
            if (lane == 0) {
                old_a = old[sl];
                old_b = hb ? old[sl + 1] : 0.f;
            }
        }
    };
    auto keep = [&](float score, uint32_t row) {
        if (score >= tau_q) {
            const uint32_t pos = atomicAdd(cb.count + q, 1u);
            if (pos < cb.cap) {
                cb.score[(size_t)q * cb.cap + pos] = score;
                cb.row[(size_t)q * cb.cap + pos] = row;
            } else {
                *cb.overflow = 1u;
            }
        }
    };
    fetch_meta(slot);
    for (; slot < cnt; slot += step) {
        const bool has_b = slot + 1 < cnt;
        const uint32_t row_a = ra, row_b = rb;
        const float4* pa = reinterpret_cast<const float4*>(seg_row_ptr(segs, row_a, d));
        const float4* pb = reinterpret_cast<const float4*>(seg_row_ptr(segs, row_b, d));
        const float cur_a = old_a, cur_b = old_b;
        float4 xa[VPL], xb[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            xa[i] = __ldg(pa + i * 32 + lane);
            xb[i] = __ldg(pb + i * 32 + lane);
        }
        fetch_meta(slot + step);
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 qv = qs[i * 32 + lane];
            a = lane_fma4(a, qv, xa[i]);
            b = lane_fma4(b, qv, xb[i]);
        }
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            worst = fmaxf(worst, fabsf(a - cur_a));
            keep(a, row_a);
            if (has_b) {
                worst = fmaxf(worst, fabsf(b - cur_b));
                keep(b, row_b);
            }
        }
    }
    if (lane == 0) {
        if (worst > 0.f) atomicMax(reinterpret_cast<int*>(screen_err_max), __float_as_int(worst));
        if (warp == 0 && blockIdx.y == 0) atomicAdd(rescored, (unsigned long long)cnt);
    }
}

bool launch_rescore_log(CandBuf lg, CandBuf cb, const float* q, int d, SegTable segs, int nq, const float* tau,
                        float* screen_err_max, unsigned long long* rescored, cudaStream_t s) {
    // large batches: one CTA per query (a chunk's log holds a few hundred entries per query); a handful of queries:
    // the log is split over many CTAs so that its rows are fetched in one or two latencies
    const int per_query = (int)std::min<int64_t>(std::max(1, 1184 / std::max(nq, 1)), ((int64_t)lg.cap + 7) / 8);
    const dim3 grid((unsigned)nq, (unsigned)std::max(1, per_query));
#define HAC_RESCORE_LOG(V) rescore_log_kernel<V><<<grid, 128, 0, s>>>(lg, cb, q, segs, tau, screen_err_max, rescored)
    switch (d) {
        case 128: HAC_RESCORE_LOG(1); break;
        case 256: HAC_RESCORE_LOG(2); break;
        case 384: HAC_RESCORE_LOG(3); break;
        case 512: HAC_RESCORE_LOG(4); break;
        case 640: HAC_RESCORE_LOG(5); break;
        case 768: HAC_RESCORE_LOG(6); break;
        case 896: HAC_RESCORE_LOG(7); break;
        case 1024: HAC_RESCORE_LOG(8); break;
        default: return false;
    }
#undef HAC_RESCORE_LOG
    return true;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) final_select_kernel(CandBuf cb, int k, const int64_t* __restrict__ id_table,
                                                           int64_t id_base, float* __restrict__ D,
                                                           int64_t* __restrict__ I, bool use_score) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    const int q = blockIdx.x;
    const int cnt = (int)min(cb.count[q], cb.cap);
    const int P = next_pow2(cnt);
    const float* ex = (use_score ? cb.score : cb.exact) + (size_t)q * cb.cap;
    const uint32_t* rw = cb.row + (size_t)q * cb.cap;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = i < cnt ? cand_key(ex[i], rw[i]) : 0ull;
    bitonic_sort_desc(keys, P);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        float score = -FLT_MAX;
        int64_t id = -1;
        if (j < cnt) {
            const uint64_t kk = keys[j];
            const uint32_t row = 0xFFFFFFFFu - (uint32_t)kk;
            score = key_float((uint32_t)(kk >> 32));
            id = id_table != nullptr ? id_table[row] : id_base + (int64_t)row;
        }
        D[(size_t)q * k + j] = score;
        I[(size_t)q * k + j] = id;
    }
}

// ---------------------------------------------------------------------------------------------
// Careful mode (shortlist overflow recovery).
// rollback: forget the appends made since the last refresh (entries [0, sorted) are intact).
__global__ void rollback_kernel(CandBuf cb, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *cb.overflow = 0u;
    if (q < nq) cb.count[q] = cb.sorted[q];
}
void launch_rollback(CandBuf cb, int nq, cudaStream_t s) {
    rollback_kernel<<<(nq + 255) / 256, 256, 0, s>>>(cb, nq);
}

// exact compaction: the shortlist (already rescored) is cut to its k best entries by
// (exact score desc, row asc); those entries now carry their exact score.  A row dropped here is
// ranked behind k rows with smaller ids or larger scores and can never re-enter the top-k, so the
// carry-over between chunks is bounded by k whatever the data looks like (mass duplicates).
__global__ void __launch_bounds__(512) exact_compact_kernel(CandBuf cb, int k, const float* __restrict__ margin,
                                                            float* __restrict__ tau, float* __restrict__ thr) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    const int q = blockIdx.x;
    const int cnt = (int)min(cb.count[q], cb.cap);
    const int P = next_pow2(cnt);
    float* sc = cb.score + (size_t)q * cb.cap;
    uint32_t* rw = cb.row + (size_t)q * cb.cap;
    const float* ex = cb.exact + (size_t)q * cb.cap;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = i < cnt ? cand_key(ex[i], rw[i]) : 0ull;
    bitonic_sort_desc(keys, P);
    const int keep = min(cnt, k);
    for (int i = threadIdx.x; i < keep; i += blockDim.x) {
        sc[i] = key_float((uint32_t)(keys[i] >> 32));
        rw[i] = 0xFFFFFFFFu - (uint32_t)keys[i];
    }
    if (threadIdx.x == 0) {
        cb.count[q] = (uint32_t)keep;
        cb.sorted[q] = (uint32_t)keep;
        if (cnt >= k) {
            const float t = fmaxf(tau[q], key_float((uint32_t)(keys[k - 1] >> 32)));
            float th = t - 2.f * margin[q];
            th -= fabsf(th) * 1e-6f;
            tau[q] = t;
            thr[q] = th;
        }
    }
}
void launch_exact_compact(CandBuf cb, int k, const float* margin, float* tau, float* thr, int nq, cudaStream_t s) {
    const size_t smem = (size_t)cb.cap * sizeof(uint64_t);
    // the attribute is per device (several devices per process are possible), so it is set per launch
    if (smem > 40 * 1024) cudaFuncSetAttribute(exact_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    exact_compact_kernel<<<nq, 512, smem, s>>>(cb, k, margin, tau, thr);
}

void launch_final_select(CandBuf cb, int k, int nq, const int64_t* id_table, int64_t id_base, float* D,
                         int64_t* I, bool use_score, cudaStream_t s) {
    const size_t smem = (size_t)cb.cap * sizeof(uint64_t);
    // the attribute is per device (several devices per process are possible), so it is set per launch
    if (smem > 48 * 1024) cudaFuncSetAttribute(final_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    final_select_kernel<<<nq, 512, smem, s>>>(cb, k, id_table, id_base, D, I, use_score);
}

__global__ void fill_empty_kernel(float* D, int64_t* I, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        D[i] = -FLT_MAX;
        I[i] = -1;
    }
}
void launch_fill_empty(float* D, int64_t* I, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    fill_empty_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, s>>>(D, I, n);
}

__global__ void fill_f32_kernel(float* p, float value, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}
void launch_fill_f32(float* p, float value, int n, cudaStream_t s) {
    if (n > 0) fill_f32_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, value, n);
}

// ---------------------------------------------------------------------------------------------
// k-way merge of n_lists SORTED lists per query (shards or corpus blocks) on the device, by rank
// computation instead of a sort: the final position of an item is its position in its own list plus,
// for every other list, the number of items there that rank before it (binary search) - one pass,
// no sorting network, no barriers between steps.  Order: (score desc, id asc, list asc); fillers
// (id -1, score -FLT_MAX) sort last.
struct MergeItem {
    uint32_t key;
    uint32_t list;
    uint64_t id;   // int64 id viewed unsigned: fillers (-1) sort last among equal scores
};
__device__ __forceinline__ bool ranks_before(const MergeItem& a, const MergeItem& b) {
    if (a.key != b.key) return a.key > b.key;
    if (a.id != b.id) return a.id < b.id;
    return a.list < b.list;
}

__device__ __forceinline__ void merge_by_rank(const MergeItem* items, int n_lists, int k, int k_out, float* D_out,
                                              int64_t* I_out) {
    const int total = n_lists * k;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const MergeItem me = items[i];
        const int l = i / k;
        int rank = i - l * k;
        for (int o = 0; o < n_lists; ++o) {
            if (o == l) continue;
            const MergeItem* other = items + o * k;
            int lo = 0, hi = k;                       // first position of `other` that does not rank before me
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ranks_before(other[mid], me)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k_out) {
            const int64_t id = (int64_t)me.id;
            D_out[rank] = id < 0 ? -FLT_MAX : key_float(me.key);
            I_out[rank] = id;
        }
    }
    for (int j = total + threadIdx.x; j < k_out; j += blockDim.x) {
        D_out[j] = -FLT_MAX;
        I_out[j] = -1;
    }
}

constexpr int kMergeThreads = 256;

__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(int n_lists, int64_t nq, int k,
                                                                   const float* __restrict__ D_lists,
                                                                   const int64_t* __restrict__ I_lists, int k_out,
                                                                   float* __restrict__ D_out,
                                                                   int64_t* __restrict__ I_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MergeItem* items = reinterpret_cast<MergeItem*>(smem_raw);
    const int64_t q = blockIdx.x;
    const int total = n_lists * k;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int l = i / k, j = i - l * k;
        const size_t src = ((size_t)l * nq + q) * k + j;
        MergeItem it;
        it.key = float_key(D_lists[src]);
        it.list = (uint32_t)l;
        it.id = (uint64_t)I_lists[src];
        items[i] = it;
    }
    __syncthreads();
    merge_by_rank(items, n_lists, k, k_out, D_out + q * k_out, I_out + q * k_out);
}

// Same merge, but every list lives in a different buffer - typically the symmetric-memory result
// buffers of the peer GPUs, read here directly over NVLink (ld.global on mapped peer pointers), so the
// cross-shard exchange and the merge are ONE kernel and no staging copy of the candidates exists.
struct PeerLists {
    const float* D[kMaxPeerLists];
    const int64_t* I[kMaxPeerLists];
};

__global__ void __launch_bounds__(kMergeThreads) merge_topk_peers_kernel(PeerLists lists, int n_lists, int64_t nq,
                                                                         int k, int k_out,
                                                                         float* __restrict__ D_out,
                                                                         int64_t* __restrict__ I_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MergeItem* items = reinterpret_cast<MergeItem*>(smem_raw);
    const int64_t q = blockIdx.x;
    const int total = n_lists * k;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int l = i / k, j = i - l * k;
        MergeItem it;
        // plain loads of peer memory: the producer ranks finished (cross-GPU barrier) before this launch
        it.key = float_key(lists.D[l][q * k + j]);
        it.list = (uint32_t)l;
        it.id = (uint64_t)lists.I[l][q * k + j];
        items[i] = it;
    }
    __syncthreads();
    merge_by_rank(items, n_lists, k, k_out, D_out + q * k_out, I_out + q * k_out);
}

cudaError_t launch_merge_topk_peers(int n_lists, int64_t nq, int k, const float* const* D_ptrs,
                                    const int64_t* const* I_ptrs, int k_out, float* D_out, int64_t* I_out,
                                    cudaStream_t s) {
    if (n_lists > kMaxPeerLists) return cudaErrorInvalidValue;
    PeerLists lists;
    for (int i = 0; i < kMaxPeerLists; ++i) {
        lists.D[i] = i < n_lists ? D_ptrs[i] : nullptr;
        lists.I[i] = i < n_lists ? I_ptrs[i] : nullptr;
    }
    const size_t smem = (size_t)n_lists * k * sizeof(MergeItem);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(merge_topk_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    merge_topk_peers_kernel<<<(unsigned)nq, kMergeThreads, smem, s>>>(lists, n_lists, nq, k, k_out, D_out, I_out);
    return cudaGetLastError();
}

cudaError_t launch_merge_topk(int n_lists, int64_t nq, int k, const float* D_lists, const int64_t* I_lists,
                              int k_out, float* D_out, int64_t* I_out, cudaStream_t s) {
    const size_t smem = (size_t)n_lists * k * sizeof(MergeItem);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    merge_topk_kernel<<<(unsigned)nq, kMergeThreads, smem, s>>>(n_lists, nq, k, D_lists, I_lists, k_out, D_out, I_out);
    return cudaGetLastError();
}

__global__ void gather_ids_kernel(const int64_t* __restrict__ table, int64_t table_n,
                                  const int64_t* __restrict__ ids, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = ids[i];
        out[i] = (v >= 0 && v < table_n) ? table[v] : -1;
    }
}
void launch_gather_ids(const int64_t* table, int64_t table_n, const int64_t* ids, int64_t n, int64_t* out,
                       cudaStream_t s) {
    if (n <= 0) return;
    gather_ids_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, s>>>(table, table_n, ids, n, out);
}

// ---------------------------------------------------------------------------------------------
// Reciprocal rank of the first relevant passage per query - what trec_eval's `recip_rank` gives for the run
// file the PRJ drivers write (test_PRJ_topiocqa.py:232-255 ranking, :290-299 run lines, :326-338 evaluation),
// without writing or parsing a run file.  One CTA per query.  The ranking is the pids in first-occurrence order
// (a pid seen before is skipped, :249-255); unfilled trailing slots are (0, 0) tuples, i.e. pid 0 at the last
// rank, and because the evaluator's run dict keeps the LAST line of a repeated passage, a real pid 0 ranked
// earlier moves to the end as well whenever padding exists.
constexpr int kRrThreads = 128;
__global__ void __launch_bounds__(kRrThreads) reciprocal_rank_kernel(const int64_t* __restrict__ pids, int k,
                                                                     const int64_t* __restrict__ rel_ptr,
                                                                     const int64_t* __restrict__ rel_pids,
                                                                     float* __restrict__ rr_out,
                                                                     int32_t* __restrict__ rank_out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    int64_t* pid = reinterpret_cast<int64_t*>(smem_raw);            // [k]
    uint8_t* first = reinterpret_cast<uint8_t*>(pid + k);           // [k]
    __shared__ int s_distinct, s_best;
    const int64_t q = blockIdx.x;
    const int64_t r0 = rel_ptr[q], r1 = rel_ptr[q + 1];
    if (threadIdx.x == 0) { s_distinct = 0; s_best = INT_MAX; }
    for (int j = threadIdx.x; j < k; j += kRrThreads) pid[j] = pids[q * k + j];
    __syncthreads();
    int mine = 0;
    for (int j = threadIdx.x; j < k; j += kRrThreads) {
        const int64_t p = pid[j];
        bool f = p >= 0;                                            // -1: unfilled search slot (k > ntotal)
        for (int i = 0; f && i < j; ++i) f = pid[i] != p;
        first[j] = f ? 1 : 0;
        mine += f ? 1 : 0;
    }
    atomicAdd(&s_distinct, mine);
    __syncthreads();
    const bool padded = s_distinct < k;
    int n_ranked = 0;                                               // distinct pids that keep their place
    if (padded) {
        for (int j = threadIdx.x; j < k; j += kRrThreads)
            if (first[j] && pid[j] == 0) first[j] = 0;              // pid 0 is re-ranked behind everything
        __syncthreads();
    }
    for (int j = threadIdx.x; j < k; j += kRrThreads) {
        if (!first[j]) continue;
        const int64_t p = pid[j];
        bool rel = false;
        for (int64_t r = r0; r < r1 && !rel; ++r) rel = rel_pids[r] == p;
        if (rel) {
            int rank = 1;
            for (int i = 0; i < j; ++i) rank += first[i];
            atomicMin(&s_best, rank);
        }
    }
    if (padded && threadIdx.x == 0) {
        bool rel0 = false;
        for (int64_t r = r0; r < r1 && !rel0; ++r) rel0 = rel_pids[r] == 0;
        if (rel0) {
            for (int i = 0; i < k; ++i) n_ranked += first[i];
            atomicMin(&s_best, n_ranked + 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int best = s_best;
        rank_out[q] = best == INT_MAX ? 0 : best;
        rr_out[q] = best == INT_MAX ? 0.f : 1.f / (float)best;
    }
}

void launch_reciprocal_rank(const int64_t* pids, int64_t nq, int k, const int64_t* rel_ptr, const int64_t* rel_pids,
                            float* rr_out, int32_t* rank_out, cudaStream_t s) {
    if (nq <= 0) return;
    const size_t smem = (size_t)k * (sizeof(int64_t) + 1);
    reciprocal_rank_kernel<<<(unsigned)nq, kRrThreads, smem, s>>>(pids, k, rel_ptr, rel_pids, rr_out, rank_out);
}

}  // namespace hac
