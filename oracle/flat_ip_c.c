/*
 * C restatement of the reference retrieval arithmetic - TEST INFRASTRUCTURE ONLY.
 *
 * Restates what `index.search(query_embeddings, topN)` does at
 * /root/reference/src/test_HAConvDR_topiocqa.py:102 through faiss 1.7.x
 * IndexFlatIP (third-party, not vendored; README.md:12 pins "faiss-gpu 1.7.2"):
 * fp32 inner products, a per-query min-heap of size k whose root is replaced
 * only on a strictly larger score (rows visited in increasing index order), a
 * final reorder to descending score, -FLT_MAX / -1 fill for unfilled slots.
 * Ties (unspecified in faiss) are fixed to (score desc, index asc).
 *
 * Two entry points:
 *   oracle_heap_addn / oracle_heap_reorder : the selection half, fed with score
 *       blocks produced by a BLAS sgemm on the Python side (faiss' 4096x1024
 *       blocking) - this is the timed "port" CPU baseline;
 *   oracle_search_naive : self-contained scalar path (own dot products) used to
 *       validate the NumPy restatement without any BLAS.
 * Parity status: unpinned for the faiss arithmetic (see oracle/__init__.py).
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* heap order: a "less" b  <=>  a is evicted before b. */
static inline int evict_before(float va, int64_t ia, float vb, int64_t ib) {
    return (va < vb) || (va == vb && ia > ib);
}

static void sift_down(int k, float* val, int64_t* idx, int pos) {
    float v = val[pos];
    int64_t id = idx[pos];
    for (;;) {
        int l = 2 * pos + 1, r = l + 1, m;
        if (l >= k) break;
        m = (r < k && evict_before(val[r], idx[r], val[l], idx[l])) ? r : l;
        if (!evict_before(val[m], idx[m], v, id)) break;
        val[pos] = val[m];
        idx[pos] = idx[m];
        pos = m;
    }
    val[pos] = v;
    idx[pos] = id;
}

void oracle_heap_init(int64_t nq, int k, float* val, int64_t* idx) {
    for (int64_t i = 0; i < nq * (int64_t)k; ++i) {
        val[i] = -FLT_MAX;
        idx[i] = -1;
    }
}

/* ip: [nq, nb] row-major score block for corpus rows [j0, j0+nb). */
void oracle_heap_addn(int64_t nq, int k, float* val, int64_t* idx,
                      const float* ip, int64_t nb, int64_t ld, int64_t j0) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        float* hv = val + qi * k;
        int64_t* hi = idx + qi * k;
        const float* row = ip + qi * ld;
        for (int64_t j = 0; j < nb; ++j) {
            float s = row[j];
            /* strict '>' against the heap minimum; empty slots hold -FLT_MAX/-1 and
               are replaced by any real row (id -1 sorts as "evict first"). */
            if (s > hv[0] || hi[0] < 0) {
                hv[0] = s;
                hi[0] = j0 + j;
                sift_down(k, hv, hi, 0);
            }
        }
    }
}

static int cmp_desc(const void* a, const void* b) {
    const float va = *(const float*)a, vb = *(const float*)b;
    const int64_t ia = *(const int64_t*)((const char*)a + 8), ib = *(const int64_t*)((const char*)b + 8);
    if (va != vb) return va > vb ? -1 : 1;
    if (ia < 0 || ib < 0) return (ia < 0) - (ib < 0);
    return ia < ib ? -1 : (ia > ib);
}

void oracle_heap_reorder(int64_t nq, int k, float* val, int64_t* idx) {
#pragma omp parallel for schedule(static)
    for (int64_t qi = 0; qi < nq; ++qi) {
        struct { float v; int32_t pad; int64_t i; }* tmp = malloc((size_t)k * 16);
        for (int j = 0; j < k; ++j) { tmp[j].v = val[qi * k + j]; tmp[j].pad = 0; tmp[j].i = idx[qi * k + j]; }
        qsort(tmp, (size_t)k, 16, cmp_desc);
        for (int j = 0; j < k; ++j) { val[qi * k + j] = tmp[j].v; idx[qi * k + j] = tmp[j].i; }
        free(tmp);
    }
}

/* 16 interleaved partial sums, combined pairwise: a fixed fp32 summation order
   that gcc can vectorise without -ffast-math. */
static inline float dot_f32(const float* a, const float* b, int d) {
    float acc[16];
    int i, l;
    for (l = 0; l < 16; ++l) acc[l] = 0.0f;
    for (i = 0; i + 16 <= d; i += 16)
        for (l = 0; l < 16; ++l) acc[l] += a[i + l] * b[i + l];
    for (; i < d; ++i) acc[i & 15] += a[i] * b[i];
    for (l = 8; l > 0; l >>= 1)
        for (i = 0; i < l; ++i) acc[i] += acc[i + l];
    return acc[0];
}

void oracle_search_naive(int64_t nq, int64_t n, int d, const float* q, const float* x,
                         int k, float* D, int64_t* I) {
    oracle_heap_init(nq, k, D, I);
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t qi = 0; qi < nq; ++qi) {
        float* hv = D + qi * k;
        int64_t* hi = I + qi * k;
        for (int64_t j = 0; j < n; ++j) {
            float s = dot_f32(q + qi * d, x + j * d, d);
            if (s > hv[0] || hi[0] < 0) {
                hv[0] = s;
                hi[0] = j;
                sift_down(k, hv, hi, 0);
            }
        }
    }
    oracle_heap_reorder(nq, k, D, I);
}
