// Shared device helpers for the sm_100a kernels: PTX wrappers (mbarrier, bulk async copy,
// tcgen05 / TMEM), shadow-tile geometry and the exact fp32 dot product that defines the
// engine's final scores.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hac {

// ------------------------------------------------------------------------------------------
// Geometry of the f16 "shadow" copy (the tensor-core operand image of corpus and queries).
// A row tile is 128 rows; K is cut in 64-element blocks.  One (row tile, k block) piece is a
// 16 KiB image of the shared-memory layout tcgen05.mma expects for a K-major operand with
// 128-byte swizzle: 16 swizzle atoms of 8 rows x 128 B, chunk c (16 B) of row r stored at
//   (r/8)*1024 + (r%8)*128 + ((c ^ (r%8)) * 16).
// Pieces are contiguous in HBM, [tile][kblock][16 KiB], so one cp.async.bulk of 16 KiB fills
// a pipeline stage with no tensor map and no shared-memory bank conflicts.
// ------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kBlockK = 64;
constexpr int kPieceBytes = kTileRows * kBlockK * 2;  // 16384

__host__ __device__ inline int64_t shadow_tiles(int64_t rows) { return (rows + kTileRows - 1) / kTileRows; }
__host__ __device__ inline int64_t shadow_bytes(int64_t rows, int d) {
    return shadow_tiles(rows) * (int64_t)(d / kBlockK) * kPieceBytes;
}
// byte offset of the 16-byte chunk holding elements [k8*8, k8*8+8) of row r
__device__ __forceinline__ int64_t shadow_chunk_offset(int64_t r, int k8, int d) {
    const int64_t tile = r / kTileRows;
    const int rr = (int)(r % kTileRows);
    const int kb = k8 >> 3, c = k8 & 7;
    return (tile * (d / kBlockK) + kb) * (int64_t)kPieceBytes + (rr >> 3) * 1024 + (rr & 7) * 128 +
           ((c ^ (rr & 7)) << 4);
}

// The int8 shadow uses the same piece image with one byte per element: a K block is 128 elements
// (one 128-byte swizzle row), chunk c = 16 consecutive int8 values.
constexpr int kBlockK8 = 128;
__host__ __device__ inline int64_t shadow8_bytes(int64_t rows, int d) {
    return shadow_tiles(rows) * (int64_t)(d / kBlockK8) * kPieceBytes;
}
__device__ __forceinline__ int64_t shadow8_chunk_offset(int64_t r, int k16, int d) {
    const int64_t tile = r / kTileRows;
    const int rr = (int)(r % kTileRows);
    const int kb = k16 >> 3, c = k16 & 7;
    return (tile * (d / kBlockK8) + kb) * (int64_t)kPieceBytes + (rr >> 3) * 1024 + (rr & 7) * 128 +
           ((c ^ (rr & 7)) << 4);
}

// ------------------------------------------------------------------------------------------
// The exact score.  One warp per (query, row): lane l accumulates float4 groups l, l+32, ...
// sequentially with FMA (x,y,z,w order), then an xor-butterfly sums the 32 partials.  Every
// kernel that produces a final score calls this one function, so the GEMV scan and the
// shortlist rescore are bitwise identical.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float lane_fma4(float acc, const float4 a, const float4 b) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}

// order-preserving float <-> uint key (larger float -> larger key)
__device__ __forceinline__ uint32_t float_key(float f) {
    uint32_t u = __float_as_uint(f + 0.0f);   // -0.0 and +0.0 are the same score: one key
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------------------------------
// PTX: mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// arrive on the same-offset barrier of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t}" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// PTX: bulk async copy global -> shared (TMA engine, non-tensor form: UBLKCP in SASS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ------------------------------------------------------------------------------------------
// PTX: tcgen05 (5th-gen tensor cores) and tensor memory
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    if constexpr (kCtaGroup == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                     "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
                     "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if constexpr (kCtaGroup == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout_type [61,64) with SWIZZLE_128B = 2)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)0 << 16;               // LBO unused: one swizzle atom along K
    d |= (uint64_t)(1024 >> 4) << 32;     // SBO
    d |= (uint64_t)1 << 46;               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;               // SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D=f32, A=B=f16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4)                       // c_format = F32
           | (0u << 7) | (0u << 10)        // a_format = b_format = F16
           | (0u << 15) | (0u << 16)       // K-major A and B
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::i8: D=s32, A=B=signed 8-bit, both K-major; K = 32 per instruction
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4)                       // c_format = S32
           | (1u << 7) | (1u << 10)        // a_format = b_format = signed int8
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <int kCtaGroup>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (kCtaGroup == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
template <int kCtaGroup>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    if constexpr (kCtaGroup == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
}
// make the mbarrier track completion of all tcgen05 ops issued so far by this thread
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(cta_mask)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (lane = thread)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace hac
