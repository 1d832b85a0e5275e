// Small-batch scan (1..4 queries): exact fp32 scores by streaming the fp32 rows once from HBM.
// One warp per row, 128-bit coalesced loads, queries resident in registers, per-lane FMA chains and
// an xor-butterfly (the engine's exact score, see hac_common.cuh), fused threshold filter - only rows
// whose score reaches the query's current threshold are appended to the shortlist.
// Roofline: HBM; algorithmic bytes = rows * d * 4 per launch.
#include "hac_common.cuh"
#include "hac_kernels.cuh"

namespace hac {

__device__ __forceinline__ void emit_candidate(const CandBuf& cb, int q, uint32_t row, float score) {
    const uint32_t pos = atomicAdd(cb.count + q, 1u);
    if (pos < cb.cap) {
        cb.score[(size_t)q * cb.cap + pos] = score;
        cb.row[(size_t)q * cb.cap + pos] = row;
    } else {
        *cb.overflow = 1u;
    }
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

template <int QB, int VPL, bool kRagged>
__global__ void __launch_bounds__(256) scan_gemv_kernel(const float* __restrict__ rows, int64_t r0, int64_t r1,
                                                        int d, const float* __restrict__ qmat,
                                                        const float* __restrict__ thr, CandBuf cb,
                                                        uint32_t row_id_base) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = d >> 2;

    float4 qv[QB][VPL];
    float th[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        th[qi] = thr[qi];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = i * 32 + lane;
            qv[qi][i] = (!kRagged || c < d4) ? __ldg(reinterpret_cast<const float4*>(qmat + (size_t)qi * d) + c)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    unsigned long long emitted = 0;
    // two rows per warp per iteration: 2*VPL independent 16-byte loads in flight per lane
    for (int64_t r = r0 + 2 * gw; r < r1; r += 2 * n_warps) {
        const bool has_b = r + 1 < r1;
        const float4* pa = reinterpret_cast<const float4*>(rows + (size_t)r * d);
        const float4* pb = reinterpret_cast<const float4*>(rows + (size_t)(has_b ? r + 1 : r) * d);
        float4 xa[VPL], xb[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = i * 32 + lane;
            if (!kRagged || c < d4) {
                xa[i] = ld_stream(pa + c);
                xb[i] = ld_stream(pb + c);
            } else {
                xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                xb[i] = xa[i];
            }
        }
        float sa[QB], sb[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                if (!kRagged || i * 32 + lane < d4) {
                    a = lane_fma4(a, qv[qi][i], xa[i]);
                    b = lane_fma4(b, qv[qi][i], xb[i]);
                }
            }
            sa[qi] = warp_sum(a);
            sb[qi] = warp_sum(b);
        }
        if (lane == 0) {
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                if (sa[qi] >= th[qi]) {
                    emit_candidate(cb, qi, row_id_base + (uint32_t)r, sa[qi]);
                    ++emitted;
                }
                if (has_b && sb[qi] >= th[qi]) {
                    emit_candidate(cb, qi, row_id_base + (uint32_t)(r + 1), sb[qi]);
                    ++emitted;
                }
            }
        }
    }
    if (lane == 0 && emitted) atomicAdd(cb.emitted, emitted);
}

// ---------------------------------------------------------------------------------------------
// Staged variant (d a multiple of 128): a producer warp streams groups of 16 consecutive rows
// (16*d*4 bytes, contiguous in HBM) into a 4-deep shared-memory ring with one bulk async copy each,
// 8 consumer warps take two rows of the group each (two independent FMA chains per query).  Up to
// 192 KiB are in flight per SM independent of the register budget, so 2..4 queries (96 query registers
// per lane) still stream near the HBM rate.
// The per-lane FMA order and the butterfly are those of the direct kernel: scores are bitwise equal.
constexpr int kGemvStages = 4;
constexpr int kGemvRowsPerStage = 16;
constexpr int kGemvConsumerWarps = 8;
constexpr int kGemvThreads = 32 * (1 + kGemvConsumerWarps);

__device__ __forceinline__ void gemv_wait(uint64_t* bar, uint32_t parity) {
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

template <int QB, int VPL>
__global__ void __launch_bounds__(kGemvThreads, 1) scan_gemv_staged_kernel(const float* __restrict__ rows, int64_t r0,
                                                                           int64_t r1, const float* __restrict__ qmat,
                                                                           const float* __restrict__ thr, CandBuf cb,
                                                                           uint32_t row_id_base) {
    constexpr int d = VPL * 128;
    constexpr int kRowBytes = d * 4;
    constexpr int kStageBytes = kGemvRowsPerStage * kRowBytes;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kGemvStages * kStageBytes);
    uint64_t* empty = full + kGemvStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_groups = (r1 - r0 + kGemvRowsPerStage - 1) / kGemvRowsPerStage;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kGemvStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kGemvConsumerWarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
                const int64_t row = r0 + g * kGemvRowsPerStage;
                const int64_t left = r1 - row;
                const uint32_t bytes = (uint32_t)((left < kGemvRowsPerStage ? left : kGemvRowsPerStage) * kRowBytes);
                gemv_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], bytes);
                bulk_g2s(smem + stage * kStageBytes, rows + (size_t)row * d, bytes, &full[stage]);
                if (++stage == kGemvStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    const int slot = warp - 1;                       // this warp owns rows slot and slot + 8 of every group
    float4 qv[QB][VPL];
    float th[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        th[qi] = thr[qi];
#pragma unroll
        for (int i = 0; i < VPL; ++i) qv[qi][i] = __ldg(reinterpret_cast<const float4*>(qmat + (size_t)qi * d) + i * 32 + lane);
    }
    unsigned long long emitted = 0;
    uint32_t stage = 0, phase = 0;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int64_t row_a = r0 + g * kGemvRowsPerStage + slot;
        const int64_t row_b = row_a + kGemvConsumerWarps;
        gemv_wait(&full[stage], phase);
        float4 xa[VPL], xb[VPL];
        const float4* src = reinterpret_cast<const float4*>(smem + stage * kStageBytes + slot * kRowBytes);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            // rows past r1 were not copied: read them anyway (stale smem) and ignore the result below
            xa[i] = src[i * 32 + lane];
            xb[i] = src[kGemvConsumerWarps * (kRowBytes / 16) + i * 32 + lane];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);   // both rows are in registers: hand the slot back
        float sa[QB], sb[QB];
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                a = lane_fma4(a, qv[qi][i], xa[i]);
                b = lane_fma4(b, qv[qi][i], xb[i]);
            }
            sa[qi] = warp_sum(a);
            sb[qi] = warp_sum(b);
        }
        if (lane == 0) {
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                if (row_a < r1 && sa[qi] >= th[qi]) {
                    emit_candidate(cb, qi, row_id_base + (uint32_t)row_a, sa[qi]);
                    ++emitted;
                }
                if (row_b < r1 && sb[qi] >= th[qi]) {
                    emit_candidate(cb, qi, row_id_base + (uint32_t)row_b, sb[qi]);
                    ++emitted;
                }
            }
        }
        if (++stage == kGemvStages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0 && emitted) atomicAdd(cb.emitted, emitted);
}

template <int QB, int VPL>
static bool launch_staged(const float* rows, int64_t r0, int64_t r1, const float* q, const float* thr, CandBuf cb,
                          uint32_t row_id_base, int sm_count, cudaStream_t s) {
    constexpr int smem = kGemvStages * kGemvRowsPerStage * VPL * 512 + 128 + 2 * kGemvStages * 8;
    if (smem > 227 * 1024) return false;
    cudaFuncSetAttribute(scan_gemv_staged_kernel<QB, VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int64_t n_groups = (r1 - r0 + kGemvRowsPerStage - 1) / kGemvRowsPerStage;
    const int grid = (int)(n_groups < sm_count ? n_groups : sm_count);
    scan_gemv_staged_kernel<QB, VPL><<<grid, kGemvThreads, smem, s>>>(rows, r0, r1, q, thr, cb, row_id_base);
    return true;
}

template <int QB>
static void launch_qb(const float* rows, int64_t r0, int64_t r1, int d, const float* q, const float* thr,
                      CandBuf cb, uint32_t row_id_base, int sm_count, cudaStream_t s) {
    const int64_t n_rows = r1 - r0;
    int64_t blocks = (n_rows + 15) / 16;                  // 8 warps x 2 rows
    const int64_t max_blocks = (int64_t)sm_count * 8;     // a few persistent waves
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    const int vpl = (d + 127) / 128;
    const bool ragged = (d % 128) != 0;
    if (!ragged && r1 - r0 >= 4096) {               // staged variant for the long chunks
        bool done = false;
        switch (vpl) {
            case 1: done = launch_staged<QB, 1>(rows, r0, r1, q, thr, cb, row_id_base, sm_count, s); break;
            case 2: done = launch_staged<QB, 2>(rows, r0, r1, q, thr, cb, row_id_base, sm_count, s); break;
            case 4: done = launch_staged<QB, 4>(rows, r0, r1, q, thr, cb, row_id_base, sm_count, s); break;
            case 6: done = launch_staged<QB, 6>(rows, r0, r1, q, thr, cb, row_id_base, sm_count, s); break;
            default: break;
        }
        if (done) return;
    }
#define HAC_GEMV_CASE(V)                                                                                   \
    case V:                                                                                                \
        if (ragged)                                                                                        \
            scan_gemv_kernel<QB, V, true><<<(int)blocks, 256, 0, s>>>(rows, r0, r1, d, q, thr, cb, row_id_base); \
        else                                                                                               \
            scan_gemv_kernel<QB, V, false><<<(int)blocks, 256, 0, s>>>(rows, r0, r1, d, q, thr, cb, row_id_base); \
        break;
    switch (vpl) {
        HAC_GEMV_CASE(1)
        HAC_GEMV_CASE(2)
        HAC_GEMV_CASE(3)
        HAC_GEMV_CASE(4)
        HAC_GEMV_CASE(5)
        HAC_GEMV_CASE(6)
        HAC_GEMV_CASE(7)
        HAC_GEMV_CASE(8)
        default: break;
    }
#undef HAC_GEMV_CASE
}

void launch_scan_gemv(const float* rows, int64_t r0, int64_t r1, int d, const float* q, int nq, const float* thr,
                      CandBuf cb, uint32_t row_id_base, int sm_count, cudaStream_t s) {
    if (r1 <= r0) return;
    switch (nq) {
        case 1: launch_qb<1>(rows, r0, r1, d, q, thr, cb, row_id_base, sm_count, s); break;
        case 2: launch_qb<2>(rows, r0, r1, d, q, thr, cb, row_id_base, sm_count, s); break;
        case 3: launch_qb<3>(rows, r0, r1, d, q, thr, cb, row_id_base, sm_count, s); break;
        case 4: launch_qb<4>(rows, r0, r1, d, q, thr, cb, row_id_base, sm_count, s); break;
        default: break;
    }
}

}  // namespace hac
