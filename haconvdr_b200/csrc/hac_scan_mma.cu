// Large-batch scan: f16 screen scores on the 5th-gen tensor cores with a fused threshold filter.
//
//   D[128 queries x 256 rows] (fp32, TMEM) += Qhat[128 x 64] * Xhat[256 x 64]^T   per K block,
//   12 K blocks for d = 768, operands staged global -> shared by 16 KiB bulk async copies of the
//   pre-swizzled shadow tiles (hac_common.cuh), tcgen05.mma issued by one thread, accumulators
//   double-buffered in TMEM (2 x 256 columns), epilogue warps read them back with tcgen05.ld and
//   compare every score against the owning query's emission threshold - only the (rare) survivors
//   are appended to the shortlist in HBM, the score tile itself never leaves the SM.
//
// Warp roles (192 threads, one persistent CTA per SM): warp 0 = copy producer, warp 1 = MMA issuer
// (owns the TMEM allocation), warps 2..5 = epilogue (TMEM lane quarter = warp % 4, one query per thread).
// Tile order: consecutive CTAs take the query tiles of the same 256-row corpus tile, so a corpus
// tile is pulled from HBM once and re-read from L2 by the other query tiles.
// Roofline: tensor pipe; algorithmic FLOPs = 2 * queries * rows * d per launch.
#include "hac_common.cuh"
#include "hac_kernels.cuh"

namespace hac {

namespace {

constexpr int kStages = 4;
constexpr int kTileM = 128;                                // queries per tile (TMEM lanes)
constexpr int kTileN = 256;                                // corpus rows per tile (TMEM columns)
constexpr int kABytes = kPieceBytes;                       // 16 KiB
constexpr int kBBytes = 2 * kPieceBytes;                   // 32 KiB
constexpr int kStageBytes = kABytes + kBBytes;             // 48 KiB
constexpr int kThreads = 192;
constexpr int kStash = 8;                                  // per-thread survivors kept until the TMEM buffer is released
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

struct Barriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
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

__global__ void __launch_bounds__(kThreads, 1) scan_mma_kernel(const MmaScanArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kb_count = a.d / kBlockK;
    const int64_t n_ctiles = a.ct1 - a.ct0;
    const int64_t n_tiles = n_ctiles * a.n_qtiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tmem_full[i], 1);
            mbar_init(&bars->tmem_empty[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<1>(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 0) {
        // ---------------- producer: shadow pieces -> shared memory ----------------
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int64_t ct = a.ct0 + t / a.n_qtiles;
                const int qt = (int)(t % a.n_qtiles);
                const uint8_t* srcA = a.q_shadow + (size_t)qt * kb_count * kPieceBytes;
                const uint8_t* srcB0 = a.x_shadow + (size_t)(2 * ct) * kb_count * kPieceBytes;
                const uint8_t* srcB1 = srcB0 + (size_t)kb_count * kPieceBytes;
                for (int kb = 0; kb < kb_count; ++kb) {
                    wait_or_trap(&bars->empty[stage], phase ^ 1);
                    uint8_t* sA = smem + stage * kStageBytes;
                    uint8_t* sB = sA + kABytes;
                    mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
                    bulk_g2s(sA, srcA + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                    bulk_g2s(sB, srcB0 + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                    bulk_g2s(sB + kPieceBytes, srcB1 + (size_t)kb * kPieceBytes, kPieceBytes, &bars->full[stage]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(kTileM, kTileN);
            uint32_t stage = 0, phase = 0;
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                wait_or_trap(&bars->tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kTileN;
                for (int kb = 0; kb < kb_count; ++kb) {
                    wait_or_trap(&bars->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sA = smem_u32(smem + stage * kStageBytes);
                    const uint64_t descA = umma_desc_k128(sA);
                    const uint64_t descB = umma_desc_k128(sA + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address field
                        umma_f16<1>(tmem_d, descA + 2 * k, descB + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&bars->empty[stage]);          // frees the smem stage when these MMAs retire
                    if (kb == kb_count - 1) umma_commit(&bars->tmem_full[acc]);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ---------------- epilogue: fused threshold filter ----------------
        const int quarter = warp & 3;                    // TMEM lanes [32*quarter, 32*quarter+32)
        const float scale = a.q_stats->scale * a.x_stats->scale;
        const float inv_scale = a.q_stats->inv_scale * a.x_stats->inv_scale;
        unsigned long long emitted = 0;
        uint32_t it = 0;
        // Survivors are first stashed per thread (local memory) and appended to the shortlist only
        // AFTER the accumulator buffer has been handed back to the MMA warp, so the round trip of the
        // global atomic overlaps the next tile's MMAs instead of stalling the TMEM pipeline.
        float stash_v[kStash];
        uint32_t stash_r[kStash];
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int64_t ct = a.ct0 + t / a.n_qtiles;
            const int qt = (int)(t % a.n_qtiles);
            const int q = qt * kTileM + quarter * 32 + lane;
            const float thr_s = a.thr[q] * scale;        // threshold in accumulator units (power-of-two scale)
            const int64_t row0 = ct * kTileN;
            uint32_t stash_n = 0;
            wait_or_trap(&bars->tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kTileN;
#pragma unroll 1
            for (int c = 0; c < kTileN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                tmem_ld_wait();
                bool any = false;
#pragma unroll
                for (int j = 0; j < 32; ++j) any |= (__uint_as_float(v[j]) >= thr_s);
                if (any) {
                    uint32_t mask = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) >= thr_s ? 1u : 0u) << j;
                    const int64_t rem = a.seg_rows - (row0 + c * 32);      // rows past the segment end are padding
                    if (rem < 32) mask &= rem <= 0 ? 0u : ((1u << rem) - 1u);
                    const uint32_t n = __popc(mask);
                    const uint32_t row_id = a.row_id_base + (uint32_t)(row0 + c * 32);
                    if (n != 0 && stash_n + n <= (uint32_t)kStash) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if ((mask >> j) & 1u) {
                                stash_v[stash_n] = __uint_as_float(v[j]) * inv_scale;
                                stash_r[stash_n] = row_id + j;
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
                                a.cb.score[(size_t)q * a.cb.cap + pos] = __uint_as_float(v[j]) * inv_scale;
                                a.cb.row[(size_t)q * a.cb.cap + pos] = row_id + j;
                            }
                        }
                    }
                    emitted += n;
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&bars->tmem_empty[acc]);
            if (stash_n != 0) {
                const uint32_t base = atomicAdd(a.cb.count + q, stash_n);
                if (base + stash_n > a.cb.cap) *a.cb.overflow = 1u;
                for (uint32_t i = 0; i < stash_n; ++i) {
                    if (base + i < a.cb.cap) {
                        a.cb.score[(size_t)q * a.cb.cap + base + i] = stash_v[i];
                        a.cb.row[(size_t)q * a.cb.cap + base + i] = stash_r[i];
                    }
                }
            }
            __syncwarp();
        }
        if (emitted) atomicAdd(a.cb.emitted, emitted);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

}  // namespace

cudaError_t scan_mma_configure() {
    return cudaFuncSetAttribute(scan_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

cudaError_t launch_scan_mma(const MmaScanArgs& a, int sm_count, cudaStream_t s) {
    const int64_t n_tiles = (a.ct1 - a.ct0) * a.n_qtiles;
    if (n_tiles <= 0) return cudaSuccess;
    const int grid = (int)(n_tiles < sm_count ? n_tiles : sm_count);
    scan_mma_kernel<<<grid, kThreads, kSmemBytes, s>>>(a);
    return cudaGetLastError();
}

}  // namespace hac
