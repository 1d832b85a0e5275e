// Operand preparation: fp32 rows -> f16 tiled/swizzled shadow (+ error norms for the rigorous
// screen margin), synthetic corpus generator.  All HBM-bound, 128-bit vectorised, one warp per row.
#include <algorithm>

#include "hac_common.cuh"
#include "hac_kernels.cuh"

namespace hac {

// ---------------------------------------------------------------------------------------------
__global__ void absmax_kernel(const float4* __restrict__ x, int64_t n4, float* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative floats order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

void launch_absmax(const float* x, int64_t n_elems, float* absmax_out, cudaStream_t s) {
    cudaMemsetAsync(absmax_out, 0, sizeof(float), s);
    const int64_t n4 = n_elems / 4;
    if (n4 == 0) return;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    absmax_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(x), n4, absmax_out);
}

// scale = 2^(12 - ceil(log2(absmax))): |x*scale| <= 4096, three octaves of headroom below the
// f16 maximum for later appends to the same segment; values that still overflow saturate in the
// convert kernel and show up in err_norm_max (the margin stays rigorous).
__global__ void pick_scale_kernel(OperandStats* st, const float* absmax_in, int keep_scale) {
    const float am = *absmax_in;
    st->absmax = fmaxf(st->absmax, am);
    if (keep_scale && st->scale != 0.f) return;
    int e = 0;
    float scale = 1.f;
    if (am > 0.f && isfinite(am)) {
        frexpf(am, &e);                 // am = f * 2^e, f in [0.5, 1)  ->  am <= 2^e
        int p = 12 - e;
        p = max(-100, min(100, p));
        scale = ldexpf(1.f, p);
    }
    st->scale = scale;
    st->inv_scale = 1.f / scale;
}

void launch_pick_scale(OperandStats* stats, const float* absmax_in, int keep_scale, cudaStream_t s) {
    pick_scale_kernel<<<1, 1, 0, s>>>(stats, absmax_in, keep_scale);
}

// ---------------------------------------------------------------------------------------------
// One warp per row.  Each lane converts 16-byte output chunks (8 elements): two float4 loads,
// scale, round-to-nearest f16 (saturating), one 16 B store into the swizzled tile image.
// Per-row sums of x^2, xhat^2 and (x - xhat)^2 feed the screen margin.
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ x, int64_t n, int64_t n_pad,
                                                           int d, uint8_t* __restrict__ shadow, int64_t row0,
                                                           OperandStats* __restrict__ stats,
                                                           float* __restrict__ row_norm,
                                                           float* __restrict__ row_err, int drop_bits,
                                                           const float* __restrict__ center) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float scale = stats->scale, inv_scale = stats->inv_scale;
    const int chunks = d >> 3;
    float w_norm = 0.f, w_hat = 0.f, w_err = 0.f;
    for (int64_t r = warp; r < n_pad; r += n_warps) {
        float s_x = 0.f, s_h = 0.f, s_e = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            uint4 packed = make_uint4(0u, 0u, 0u, 0u);
            if (r < n) {
                const float4* src = reinterpret_cast<const float4*>(x + r * d + c * 8);
                float4 a = __ldg(src), b = __ldg(src + 1);
                if (center != nullptr) {
                    // the image holds x - c (see column_mean_kernel): norms and errors below are those of the
                    // centred row, which is what the screen margin is made of
                    const float4 ca = __ldg(reinterpret_cast<const float4*>(center + c * 8));
                    const float4 cb = __ldg(reinterpret_cast<const float4*>(center + c * 8) + 1);
                    a.x -= ca.x; a.y -= ca.y; a.z -= ca.z; a.w -= ca.w;
                    b.x -= cb.x; b.y -= cb.y; b.z -= cb.z; b.w -= cb.w;
                }
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                __half h[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t = v[i] * scale;
                    t = fminf(fmaxf(t, -65504.f), 65504.f);
                    h[i] = __float2half_rn(t);
                    if (drop_bits > 0) {
                        // keep fewer significant bits (round to nearest): a sparser mantissa toggles less in the
                        // tensor-core multipliers; the margin is computed from the actual error, so it stays rigorous
                        unsigned short u = __half_as_ushort(h[i]);
                        u = (unsigned short)((u + (1u << (drop_bits - 1))) & ~((1u << drop_bits) - 1u));
                        h[i] = __ushort_as_half(u);
                    }
                    const float back = __half2float(h[i]) * inv_scale;
                    const float e = v[i] - back;
                    s_x = fmaf(v[i], v[i], s_x);
                    s_h = fmaf(back, back, s_h);
                    s_e = fmaf(e, e, s_e);
                }
                packed = *reinterpret_cast<uint4*>(h);
            }
            *reinterpret_cast<uint4*>(shadow + shadow_chunk_offset(row0 + r, c, d)) = packed;
        }
        if (r < n) {
            s_x = warp_sum(s_x);
            s_h = warp_sum(s_h);
            s_e = warp_sum(s_e);
            // round the norms up a little: they are used as upper bounds
            const float nx = sqrtf(s_x) * 1.0001f, nh = sqrtf(s_h) * 1.0001f, ne = sqrtf(s_e) * 1.0001f;
            if (row_norm != nullptr && lane == 0) {
                row_norm[r] = nx;
                row_err[r] = ne;
            }
            w_norm = fmaxf(w_norm, nx);
            w_hat = fmaxf(w_hat, nh);
            w_err = fmaxf(w_err, ne);
        } else if (row_norm != nullptr && lane == 0) {
            row_norm[r] = 0.f;
            row_err[r] = 0.f;
        }
    }
    if (lane == 0 && w_norm > 0.f) {
        atomicMax(reinterpret_cast<int*>(&stats->norm_max), __float_as_int(w_norm));
        atomicMax(reinterpret_cast<int*>(&stats->hat_norm_max), __float_as_int(w_hat));
        atomicMax(reinterpret_cast<int*>(&stats->err_norm_max), __float_as_int(w_err));
    }
}

void launch_convert_rows(const float* x, int64_t n, int64_t n_pad, int d, uint8_t* shadow, int64_t row0,
                         OperandStats* stats, float* row_norm, float* row_err, int drop_bits, const float* center,
                         cudaStream_t s) {
    if (n_pad <= 0) return;
    int64_t blocks = (n_pad + 7) / 8;  // 8 warps per block, one row per warp per pass
    if (blocks > 148 * 8) blocks = 148 * 8;
    convert_rows_kernel<<<(int)blocks, 256, 0, s>>>(x, n, n_pad, d, shadow, row0, stats, row_norm, row_err, drop_bits,
                                                    center);
}

// ---------------------------------------------------------------------------------------------
// Screen centre.  ANCE-style embeddings share a large common component (scores crowd around |mu|^2 with a small
// spread), and the f16 rounding error of a row - hence the screen margin - scales with the row's norm.  The f16
// image therefore holds x - c for one fixed vector c (the column means of the first rows added), the scan
// adds the per-query constant q.c back, and the margin is made of the norms of the CENTRED rows.  Any c keeps the
// bound rigorous; a good one only makes it tighter (3.5x on the x = mu + 0.3*eps stress distribution).
__global__ void __launch_bounds__(256) column_sum_kernel(const float* __restrict__ x, int64_t n, int d,
                                                         double* __restrict__ accum) {
    // blockIdx.y: column group of 256; blockIdx.x: row stripe
    const int col = blockIdx.y * 256 + threadIdx.x;
    if (col >= d) return;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = blockIdx.x * per, r1 = min(n, r0 + per);
    float acc = 0.f;
    double total = 0.0;
    int run = 0;
    for (int64_t r = r0; r < r1; ++r) {
        acc += __ldg(x + r * d + col);
        if (++run == 256) { total += (double)acc; acc = 0.f; run = 0; }
    }
    total += (double)acc;
    if (r1 > r0) atomicAdd(accum + col, total);
}
// center[0..d) <- means (non-finite -> 0); center[d] <- ||c|| rounded up
__global__ void __launch_bounds__(1024) column_mean_finish_kernel(const double* __restrict__ accum, int64_t n, int d,
                                                                  float* __restrict__ center) {
    __shared__ float warp_part[32];
    float sq = 0.f;
    for (int j = threadIdx.x; j < d; j += blockDim.x) {
        float c = (float)(accum[j] / (double)n);
        if (!isfinite(c)) c = 0.f;
        center[j] = c;
        sq = fmaf(c, c, sq);
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? warp_part[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) center[d] = sqrtf(v) * 1.0001f;
    }
}

void launch_column_mean(const float* x, int64_t n, int d, double* accum, float* center, cudaStream_t s) {
    cudaMemsetAsync(accum, 0, (size_t)d * sizeof(double), s);
    if (n <= 0) {
        cudaMemsetAsync(center, 0, (size_t)(d + 1) * sizeof(float), s);
        return;
    }
    const int stripes = (int)std::min<int64_t>(296, (n + 63) / 64);
    column_sum_kernel<<<dim3(stripes, (d + 255) / 256), 256, 0, s>>>(x, n, d, accum);
    column_mean_finish_kernel<<<1, 1024, 0, s>>>(accum, n, d, center);
}

// shift[q] = q . c (double accumulation, rounded to fp32): the constant the centred screen score is missing.
// One warp per query; padded queries get 0.
__global__ void __launch_bounds__(256) query_shift_kernel(const float* __restrict__ q, int nq, int nq_pad, int d,
                                                          const float* __restrict__ center,
                                                          float* __restrict__ shift) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= nq_pad) return;
    double acc = 0.0;
    if (r < nq)
        for (int j = lane; j < d; j += 32) acc = fma((double)__ldg(q + (size_t)r * d + j), (double)__ldg(center + j), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        float v = (float)acc;
        shift[r] = isfinite(v) ? v : 0.f;
    }
}
void launch_query_shift(const float* q, int nq, int nq_pad, int d, const float* center, float* shift, cudaStream_t s) {
    if (nq_pad <= 0) return;
    query_shift_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, s>>>(q, nq, nq_pad, d, center, shift);
}

// ---------------------------------------------------------------------------------------------
// int8 images.  Corpus: one CTA per 128-row tile; the tile shares one scale alpha = absmax/127, so the
// scan epilogue compares raw s32 accumulators against one integer threshold per (query, tile).
// Per tile the kernel also records beta = max ||x_j|| and gamma = max ||x_j - alpha*xi_j|| (upper bounds),
// the two norms the rigorous screen margin needs.
__device__ __forceinline__ uint32_t pack4_i8(int a, int b, int c, int d) {
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}
__device__ __forceinline__ int quant_i8(float x, float inv, float alpha, float& sx, float& se) {
    int q = __float2int_rn(x * inv);
    q = max(-127, min(127, q));
    const float e = fmaf(-alpha, (float)q, x);
    sx = fmaf(x, x, sx);
    se = fmaf(e, e, se);
    return q;
}
// quantise the 16 elements of chunk c16 of one row and store them into the swizzled piece image
__device__ __forceinline__ float4 ld4_centred(const float4* src, const float4* center) {
    float4 v = __ldg(src);
    if (center != nullptr) {
        const float4 c = __ldg(center);
        v.x -= c.x; v.y -= c.y; v.z -= c.z; v.w -= c.w;
    }
    return v;
}
__device__ __forceinline__ void convert_chunk_i8(const float* __restrict__ row, const float* __restrict__ center,
                                                 int c16, float inv, float alpha, uint8_t* dst, float& sx,
                                                 float& se) {
    const float4* src = reinterpret_cast<const float4*>(row + c16 * 16);
    const float4* csrc = center != nullptr ? reinterpret_cast<const float4*>(center + c16 * 16) : nullptr;
    uint4 out;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = ld4_centred(src + i, csrc != nullptr ? csrc + i : nullptr);
        w[i] = pack4_i8(quant_i8(v.x, inv, alpha, sx, se), quant_i8(v.y, inv, alpha, sx, se),
                        quant_i8(v.z, inv, alpha, sx, se), quant_i8(v.w, inv, alpha, sx, se));
    }
    out = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4*>(dst) = out;
}

__global__ void __launch_bounds__(256) convert_tiles_i8_kernel(const float* __restrict__ rows, int64_t n_rows, int d,
                                                               int64_t tile0, uint8_t* __restrict__ shadow8,
                                                               TileQ8* __restrict__ tiles,
                                                               OperandStats* __restrict__ stats,
                                                               const float* __restrict__ center) {
    __shared__ int s_absmax, s_beta, s_gamma;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tile = tile0 + blockIdx.x;
    const int64_t row_base = tile * kTileRows;
    const int chunks = d >> 4;
    if (threadIdx.x == 0) { s_absmax = 0; s_beta = 0; s_gamma = 0; }
    __syncthreads();
    float m = 0.f;
    for (int rr = warp; rr < kTileRows; rr += 8) {
        const int64_t r = row_base + rr;
        if (r >= n_rows) break;
        const float4* src = reinterpret_cast<const float4*>(rows + (size_t)r * d);
        const float4* csrc = reinterpret_cast<const float4*>(center);
        for (int i = lane; i < (d >> 2); i += 32) {
            const float4 v = ld4_centred(src + i, center != nullptr ? csrc + i : nullptr);   // the image holds x - c
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) atomicMax(&s_absmax, __float_as_int(m));
    __syncthreads();
    const float absmax = __int_as_float(s_absmax);
    const bool usable = absmax > 0.f && isfinite(absmax);
    const float alpha = usable ? absmax / 127.f : 1.f;
    const float inv = usable ? 127.f / absmax : 0.f;
    for (int rr = warp; rr < kTileRows; rr += 8) {
        const int64_t r = row_base + rr;
        float sx = 0.f, se = 0.f;
        for (int c = lane; c < chunks; c += 32) {
            uint8_t* dst = shadow8 + shadow8_chunk_offset(r, c, d);
            if (r < n_rows) convert_chunk_i8(rows + (size_t)r * d, center, c, inv, alpha, dst, sx, se);
            else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (r < n_rows) {
            sx = warp_sum(sx);
            se = warp_sum(se);
            if (lane == 0) {
                atomicMax(&s_beta, __float_as_int(sqrtf(sx)));
                atomicMax(&s_gamma, __float_as_int(sqrtf(se)));
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileQ8 t;
        t.alpha = alpha;
        t.beta = __int_as_float(s_beta) * 1.0001f;                                   // upper bounds
        t.gamma = __int_as_float(s_gamma) * 1.0001f + 1e-6f * __int_as_float(s_beta);
        t.pad = 0.f;
        tiles[tile] = t;
        if (stats != nullptr) {
            atomicMax(reinterpret_cast<int*>(&stats->i8_beta_max), __float_as_int(t.beta));
            atomicMax(reinterpret_cast<int*>(&stats->i8_gamma_max), __float_as_int(t.gamma));
        }
    }
}

void launch_convert_tiles_i8(const float* rows, int64_t n_rows, int d, int64_t tile0, int64_t tile1, uint8_t* shadow8,
                             TileQ8* tiles, OperandStats* stats, const float* center, cudaStream_t s) {
    if (tile1 <= tile0) return;
    convert_tiles_i8_kernel<<<(unsigned)(tile1 - tile0), 256, 0, s>>>(rows, n_rows, d, tile0, shadow8, tiles, stats,
                                                                      center);
}

// queries: one warp per query, one scale per query
__global__ void __launch_bounds__(256) convert_queries_i8_kernel(const float* __restrict__ q, int nq, int nq_pad,
                                                                 int d, uint8_t* __restrict__ q_shadow8,
                                                                 QueryQ8* __restrict__ consts) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= nq_pad) return;
    const int chunks = d >> 4;
    if (r >= nq) {
        for (int c = lane; c < chunks; c += 32)
            *reinterpret_cast<uint4*>(q_shadow8 + shadow8_chunk_offset(r, c, d)) = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) consts[r] = QueryQ8{0.f, 0.f, 0.f, 0.f};
        return;
    }
    const float* row = q + (size_t)r * d;
    float m = 0.f;
    for (int i = lane; i < (d >> 2); i += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const bool usable = m > 0.f && isfinite(m);
    const float t = usable ? m / 127.f : 1.f, inv = usable ? 127.f / m : 0.f;
    float sx = 0.f, se = 0.f;
    for (int c = lane; c < chunks; c += 32)
        convert_chunk_i8(row, nullptr, c, inv, t, q_shadow8 + shadow8_chunk_offset(r, c, d), sx, se);
    sx = warp_sum(sx);
    se = warp_sum(se);
    if (lane == 0) {
        QueryQ8 c;
        c.t = t;
        c.norm = sqrtf(sx) * 1.0001f;
        c.eps = sqrtf(se) * 1.0001f + 1e-6f * c.norm;
        c.nhat = c.norm + c.eps;                      // ||t*qi|| <= ||q|| + ||q - t*qi||
        consts[r] = c;
    }
}

void launch_convert_queries_i8(const float* q, int nq, int nq_pad, int d, uint8_t* q_shadow8, QueryQ8* consts,
                               cudaStream_t s) {
    if (nq_pad <= 0) return;
    convert_queries_i8_kernel<<<(nq_pad * 32 + 255) / 256, 256, 0, s>>>(q, nq, nq_pad, d, q_shadow8, consts);
}

// ---------------------------------------------------------------------------------------------
// Synthetic rows: element (r, c) is a pure function of (seed, r, c) - counter-based, so shards
// of any size reproduce the same global corpus.  splitmix64 -> two 24-bit uniforms -> Box-Muller.
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float2 normal_pair(uint64_t seed, uint64_t counter) {
    const uint64_t h = splitmix64(splitmix64(seed) ^ (counter * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull));
    const float u1 = ((float)(uint32_t)(h >> 40) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
    const float u2 = (float)(uint32_t)((h >> 8) & 0xFFFFFFu) * (1.0f / 16777216.0f);  // [0, 1)
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

__global__ void synth_kernel(float2* __restrict__ out, int64_t n_pairs, int d2, uint64_t seed, int64_t row0,
                             int dist) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pairs;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d2;
        const int c2 = (int)(i - r * d2);
        float2 z = normal_pair(seed, (uint64_t)((row0 + r) * d2 + c2));
        if (dist == 1) {
            const float2 mu = normal_pair(seed ^ 0x5DEECE66Dull, 0xFFFFFFFF00000000ull + (uint64_t)c2);
            z.x = mu.x + 0.3f * z.x;
            z.y = mu.y + 0.3f * z.y;
        }
        out[i] = z;
    }
}

void launch_synth(float* out, int64_t n, int d, uint64_t seed, int64_t row0, int dist, cudaStream_t s) {
    const int64_t n_pairs = n * (d / 2);
    if (n_pairs <= 0) return;
    int64_t blocks = (n_pairs + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    synth_kernel<<<(int)blocks, 256, 0, s>>>(reinterpret_cast<float2*>(out), n_pairs, d / 2, seed, row0, dist);
}

}  // namespace hac
