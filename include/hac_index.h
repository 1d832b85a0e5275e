/*
 * hac_index.h - C ABI of the B200-native exact inner-product (IndexFlatIP) engine.
 *
 * Drop-in boundary for the retrieval hot path of fengranMark/HAConvDR.  Every entry
 * point replaces one call the reference makes into faiss (SWIG Python binding) in
 *   src/test_HAConvDR_topiocqa.py / src/test_HAConvDR_qrecc.py (identical line numbers)
 *   src/test_PRJ_topiocqa.py / src/test_PRJ_qrecc.py (same code at :44-171)
 * Plain pointers and sizes only; no torch / numpy types.  One handle owns one
 * device shard (one process per GPU, or one handle per device inside a process).
 *
 * Conventions
 *   - every function returns 0 on success, a negative HAC_E_* code on failure;
 *     hac_last_error() returns the message of the calling thread's last failure.
 *   - a handle is not re-entrant; different handles may be used from different threads.
 *   - "host" pointers are ordinary or pinned host memory, "dev" pointers are device
 *     memory on the handle's device; `stream` is a cudaStream_t passed as void* and is used
 *     as given (NULL = the CUDA legacy default stream, i.e. torch's default stream): device
 *     entry points are stream-ordered with the caller's work and return after the stream
 *     reached the end of the call.  Host entry points run on the handle's own stream.
 *   - result order is the deterministic total order (score desc, id asc); unfilled
 *     slots (k > ntotal) carry score -FLT_MAX and id -1, as faiss does.
 */
#ifndef HAC_INDEX_H_
#define HAC_INDEX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HAC_ABI_VERSION 3

#define HAC_OK 0
#define HAC_E_INVALID (-1)   /* bad argument (dimension mismatch, k out of range, null pointer) */
#define HAC_E_CUDA (-2)      /* CUDA runtime error (message holds cudaGetErrorString) */
#define HAC_E_NOMEM (-3)     /* device or pinned-host allocation failed */
#define HAC_E_OVERFLOW (-4)  /* candidate shortlist overflowed in every retry mode */
#define HAC_E_STATE (-5)     /* call not valid in the current state */

#define HAC_MAX_K 1024       /* faiss-gpu caps k at 2048; BASELINE's sweep tops out at 1000 */

/* scan paths (hac_search*_ex `path` argument / hac_stats.path) */
#define HAC_PATH_AUTO 0      /* int8 screen when its image exists (d % 128 == 0; large k: on large shards), else the f16 screen */
#define HAC_PATH_GEMV 1      /* exact fp32 HBM-streaming scan, small query batches */
#define HAC_PATH_MMA 2       /* tcgen05 f16 screen + exact fp32 rescore of the shortlist */
#define HAC_PATH_I8 3        /* tcgen05 int8 screen (s32 accumulate) + exact fp32 rescore of every emitted row */

typedef struct hac_index hac_index;

typedef struct hac_stats {
    int32_t path;                 /* HAC_PATH_* actually used by the last search */
    int32_t retries;              /* shortlist-overflow retries taken by the last search */
    int32_t n_chunks;             /* corpus chunks (threshold refresh points) of the last search */
    int32_t kernel_launches;      /* kernels launched by the last search */
    int64_t candidates_emitted;   /* (query,row) pairs that passed the fused filter */
    int64_t candidates_rescored;  /* pairs whose exact fp32 score was recomputed */
    float   margin_max;           /* largest per-query screen margin m_q (score units) */
    float   screen_err_max;       /* largest |screen score - exact score| seen among rescored pairs */
    float   scan_ms;              /* device time of the scan kernels of the last search (CUDA events) */
    float   total_ms;             /* device time of the whole last search (CUDA events) */
    int64_t ntotal;
    int64_t bytes_fp32;           /* HBM held by the fp32 rows */
    int64_t bytes_shadow;         /* HBM held by the f16 tiled image (0 until a search needs it, see "lazy_f16") */
    int64_t bytes_i8;             /* HBM held by the int8 tiled image and its per-tile constants */
    int32_t n_sync_chunks;        /* chunks of the last search whose scan waited for the previous chunk's thresholds */
    int32_t pipelined;            /* 1 = the last search overlapped rescoring with the scans (int8 screen) */
    float   tail_ms;              /* device time between the end of the last scan and the end of the last search */
    int32_t warm_rows;            /* int8 screen: leading rows the last search scanned with the f16 screen first ("i8_warm_rows") */
} hac_stats;

/* ---- lifetime ------------------------------------------------------------------------------
 * hac_create  <- faiss.IndexFlatIP(768) + StandardGpuResources()/index_cpu_to_gpu_multiple
 *                (src/test_HAConvDR_topiocqa.py:46-66).  d must be a multiple of 64, <= 1024.
 * hac_destroy <- Python GC of the faiss index object. */
int hac_create(int d, int device, hac_index** out);
int hac_destroy(hac_index* idx);

/* Optional: pre-size the shard so later adds do not allocate (HBM-resident corpus). */
int hac_reserve(hac_index* idx, int64_t n_rows);

/* ---- add / reset ---------------------------------------------------------------------------
 * hac_add        <- index.add(passage_embedding) (src/test_HAConvDR_topiocqa.py:98):
 *                   copies n rows of d fp32 from host memory; ids are insertion order.
 * hac_add_device <- same, rows already on the device (loader / synthetic generator).
 * hac_reset      <- index.reset() (src/test_HAConvDR_topiocqa.py:122): ntotal -> 0, id base and
 *                   id table cleared, capacity is kept for the next block. */
int hac_add(hac_index* idx, int64_t n, const float* x_host);
int hac_add_device(hac_index* idx, int64_t n, const float* x_dev, void* stream);
int hac_reset(hac_index* idx);

/* Synthetic corpus rows generated on the device: row r (global index row0 + i) is a pure
 * function of (seed, r), so a shard is reproducible for any shard count.
 * dist 0: i.i.d. N(0,1); dist 1: shared mean + 0.3*N(0,1) ("anisotropic", ANCE-like). */
int hac_add_synthetic(hac_index* idx, int64_t n, uint64_t seed, int64_t row0, int dist);
/* Same generator into caller memory (device pointer), for queries and for test read-back. */
int hac_synth_fill_device(int device, float* out_dev, int64_t n, int d, uint64_t seed,
                          int64_t row0, int dist, void* stream);

/* ---- id translation ------------------------------------------------------------------------
 * Returned ids are  id_base + local_row  unless an id table is set, in which case they are
 * id_table[local_row].  Replaces  passage_embedding2id[I]  (src/test_HAConvDR_topiocqa.py:110)
 * and the shard-base translation faiss IndexShards does on the host.
 * The table is copied to the device; n must equal the rows it will cover. */
int hac_set_id_base(hac_index* idx, int64_t id_base);
int hac_set_id_table(hac_index* idx, const int64_t* ids_host, int64_t n);

/* ---- search --------------------------------------------------------------------------------
 * hac_search        <- D, I = index.search(query_embeddings, topN)
 *                      (src/test_HAConvDR_topiocqa.py:102): host fp32 queries [nq,d] in,
 *                      host D fp32 [nq,k] and I int64 [nq,k] out.
 * hac_search_device <- same with device buffers (torch tensor handoff via data_ptr()).
 * *_ex variants force a scan path (HAC_PATH_*); the plain ones pick by batch size. */
int hac_search(hac_index* idx, int64_t nq, const float* q_host, int k, float* D_host, int64_t* I_host);
int hac_search_device(hac_index* idx, int64_t nq, const float* q_dev, int k, float* D_dev,
                      int64_t* I_dev, void* stream);
int hac_search_ex(hac_index* idx, int64_t nq, const float* q_host, int k, float* D_host,
                  int64_t* I_host, int path);
int hac_search_device_ex(hac_index* idx, int64_t nq, const float* q_dev, int k, float* D_dev,
                         int64_t* I_dev, void* stream, int path);

/* ---- cross-shard / cross-block merge ----------------------------------------------------------
 * Merges `n_lists` top-k lists per query (laid out [n_lists][nq][k], device memory), each already
 * sorted by (score desc, id asc) as hac_search returns them, into one top-k_out list per query by
 * (score desc, id asc, list asc).  Replaces faiss IndexShards' host merge and the
 * reference's pure-Python pairwise merge (src/test_HAConvDR_topiocqa.py:126-149). */
int hac_merge_topk_device(int device, int n_lists, int64_t nq, int k, const float* D_lists_dev,
                          const int64_t* I_lists_dev, int k_out, float* D_out_dev,
                          int64_t* I_out_dev, void* stream);

/* Same merge with every list in its own buffer: `D_list_ptrs` / `I_list_ptrs` are HOST arrays of n_lists
 * (<= 16) DEVICE pointers, each to a [nq][k] list.  With symmetric-memory result buffers these are the
 * peer GPUs' buffers, read in-kernel over NVLink, so the cross-shard exchange and the merge are one
 * kernel (the caller orders it after a cross-GPU barrier).  Replaces NCCL all-gather + merge. */
int hac_merge_topk_peers_device(int device, int n_lists, int64_t nq, int k, const float* const* D_list_ptrs,
                                const int64_t* const* I_list_ptrs, int k_out, float* D_out_dev,
                                int64_t* I_out_dev, void* stream);

/* Cross-shard threshold exchange (one process per GPU, corpus sharded; optional).  A shard alone only knows the k-th
 * best score of its own 1/G of the rows, so it emits and rescores as many candidates as a whole-corpus search would.
 * With the exchange every shard publishes, per query, its best exact scores so far at a few fixed ranks c_1 < c_2 < ...
 * (around k/G, up to k) into `mine_dev` and reads what its peers published through `peers_dev` (HOST array of
 * n_peers = G-1 <= 15 peer-mapped DEVICE pointers, e.g. symmetric memory read over NVLink).  A word (r, c_j, L) claims
 * "shard r holds at least c_j rows scoring >= L"; any threshold T for which the claims add up to k rows is a lower
 * bound on the GLOBAL k-th best, and every shard raises its threshold to the largest such T.  Buffers hold `capacity`
 * u64 words (tag << 32 | float bits), HAC_EXCHANGE_WORDS_PER_QUERY per query, zero-initialised by the caller.  The
 * exchange is armed per search with hac_set_option(idx, "exchange_epoch", e): e > 0 must be the same on all shards for
 * the same search and different from the previous searches' (stale entries are ignored by tag, no barrier or reset
 * between searches; claims only ever strengthen within a search, so any mix of old and new words is valid), e = 0
 * switches it off.  With it armed a shard's result holds only its candidates for the GLOBAL top-k (possibly fewer
 * than k, the rest filled with -FLT_MAX / -1): it is meant to be followed by hac_merge_topk_*.  Applies to the int8
 * screen (exact-score shortlists).  n_peers = 0 clears the buffers.
 * Replaces nothing in the reference: faiss IndexShards searches its shards independently. */
#define HAC_EXCHANGE_WORDS_PER_QUERY 16
int hac_set_threshold_exchange(hac_index* idx, uint64_t* mine_dev, const uint64_t* const* peers_dev, int n_peers,
                               int64_t capacity);

/* Several shards inside ONE process (the faiss IndexShards form, src/test_HAConvDR_topiocqa.py:55-66 with n_gpu > 1):
 * lets kernels running on `device` load and store memory of `peer` (cudaDeviceEnablePeerAccess; already enabled or
 * device == peer is not an error), so that hac_merge_topk_peers_device and the threshold exchange can take plain
 * device pointers of the other GPUs.  With one process per GPU the same is achieved with symmetric memory. */
int hac_enable_peer_access(int device, int peer);

/* ---- shard group: the whole IndexShards search in one call --------------------------------------------------------
 * hac_shards_create  <- faiss.index_cpu_to_gpu_multiple(vres, vdev, cpu_index, co) with co.shard = True
 *                       (src/test_HAConvDR_topiocqa.py:55-66): groups n (<= 16) shard handles of the same dimension,
 *                       one per device (several on one device are allowed), enables peer access between their devices
 *                       and starts one host thread per shard - faiss' IndexShards keeps a C++ thread per sub-index too.
 *                       The handles are borrowed: add / reset / id bases stay per shard (hac_add, hac_set_id_base),
 *                       the group must be destroyed before them.  Like a handle, a group is not re-entrant.
 * hac_shards_search  <- D, I = index.search(query_embeddings, topN) on that sharded index (:102): every shard thread
 *                       copies the host queries to its device and runs the single-shard search with the cross-shard
 *                       threshold exchange armed (peer-mapped buffers owned by the group), then ONE merge kernel on the
 *                       first shard's device reads all lists in place over NVLink and the merged [nq,k] lists go to the
 *                       host.  Results are those of one index holding all rows (ids as the shards translate them).
 * hac_shards_search_device  same with q / D / I on the FIRST shard's device; `stream` orders the merge (the queries
 *                       must be complete on `stream`, which is synchronised first).
 * hac_shards_set_exchange   0 = shards search independently (what faiss does), 1 (default) = threshold exchange.
 * hac_shards_peer_access    1 when every device of the group can address every other one (else: gather copies, no
 *                       exchange).
 * hac_shards_last_phases    host-side milestones of the last search in ms since the call began: [0] slowest shard
 *                       done, [1] merge enqueued, [2] results on the host, [3] fastest shard done. */
typedef struct hac_shards hac_shards;
int hac_shards_create(hac_index* const* shards, int n, hac_shards** out);
int hac_shards_destroy(hac_shards* grp);
int hac_shards_search(hac_shards* grp, int64_t nq, const float* q_host, int k, float* D_host, int64_t* I_host);
int hac_shards_search_device(hac_shards* grp, int64_t nq, const float* q_dev, int k, float* D_dev, int64_t* I_dev,
                             void* stream);
int hac_shards_set_exchange(hac_shards* grp, int on);
int hac_shards_peer_access(const hac_shards* grp);
int hac_shards_last_phases(const hac_shards* grp, float* out_ms, int n);

/* offset -> pid gather on the device (src/test_HAConvDR_topiocqa.py:250): out[i] =
 * table[ids[i]] for ids >= 0, -1 otherwise.  table/ids/out are device pointers. */
int hac_gather_ids_device(int device, const int64_t* table_dev, int64_t table_n, const int64_t* ids_dev,
                          int64_t n, int64_t* out_dev, void* stream);

/* Reciprocal rank of the first relevant passage per query, fused on the device: what the PRJ drivers obtain by
 * writing the run file (src/test_PRJ_topiocqa.py:232-255, :290-299) and evaluating `recip_rank` on it with
 * pytrec_eval (:326-338), the per-query score `improve_judge` (:443-472) then compares.  pids [nq,k]: the
 * offset -> pid translated result (hac_gather_ids_device); a pid already seen for the query is skipped and slots
 * left unfilled by that count as pid 0 at the last rank, as in the run file.  rel_ptr [nq+1] / rel_pids: CSR list of
 * each query's relevant pids.  rr_out [nq] = 1/rank or 0, rank_out [nq] = 1-based rank or 0.  Device pointers. */
int hac_reciprocal_rank_device(int device, const int64_t* pids_dev, int64_t nq, int k, const int64_t* rel_ptr_dev,
                               const int64_t* rel_pids_dev, float* rr_out_dev, int32_t* rank_out_dev, void* stream);

/* ---- shard files (engine-native corpus format) -----------------------------------------------------
 * What `gen_doc_embeddings.py:127-155` would write per shard instead of a dozen pickles: ONE 4 KiB-aligned file with
 * the fp32 rows, the int8 tensor-core image of the same rows (tiled / swizzled exactly as it lies in HBM), its
 * per-tile constants, the screen centre and the corpus statistics.
 * hac_save_shard  writes the resident shard of `idx` (any number of segments; the int8 image is included when the
 *                 shard is a single segment, else it is rebuilt on load).
 * hac_load_shard  fills an EMPTY index from such a file by plain reads into page-locked staging + DMA: no pickle walk,
 *                 no fp32 -> int8 conversion, no statistics pass.  d must match; id base / table are not part of the file.
 * Layout (little endian; every section starts on a 4 KiB boundary):
 *   header  4096 B: "HACSHD01", u32 version, u32 d, u64 n_rows, u64 cap_rows, u32 flags (1 = int8 image, 2 = centre),
 *                   u32 pad, u64 offset / u64 bytes of the sections {centre, rows, int8 image, tile constants},
 *                   8 x f32 corpus statistics
 *   centre  (d + 1) f32;  rows  n_rows * d f32;  int8 image  ceil(cap_rows / 128) * (d / 128) * 16 KiB;
 *   tile constants  ceil(cap_rows / 128) * 4 f32 */
int hac_save_shard(hac_index* idx, const char* path);
int hac_load_shard(hac_index* idx, const char* path);

/* ---- pinned staging (loader) -----------------------------------------------------------------
 * Page-locked host buffers the block-pickle loader reads file payloads into, so that
 * hac_add runs H2D at full PCIe rate without an extra host copy. */
int hac_pinned_alloc(size_t bytes, void** out_host);
int hac_pinned_free(void* host);

/* ---- tuning knobs ---------------------------------------------------------------------------
 * Named integer options (unknown names -> HAC_E_INVALID).  The environment variable HAC_OPTIONS
 * ("name=value,name=value") applies them to every handle at hac_create - for callers that only know the faiss names:
 *   "mma_cta_group"  1 = one CTA per scan tile, 2 = CTA pairs sharing each MMA (cta_group::2)
 *   "chunk_growth_x100"  corpus-chunk growth factor of the threshold schedule, in percent (default 400)
 *   "default_path"   the scan path HAC_PATH_AUTO resolves to (HAC_PATH_GEMV / _MMA / _I8)
 *   "build_i8"       1 (default when d % 128 == 0) = also keep an int8 image of the corpus (rows*d bytes of HBM) for
 *                    HAC_PATH_I8, 0 = save the memory (HAC_PATH_I8 then falls back to HAC_PATH_MMA); empty index
 *                    only; the environment variable HAC_BUILD_I8 sets the default of new handles
 *   "scan_tile_major"  unit order of the tensor-core scans: 1 = every CTA group walks whole 256-row corpus tiles (all
 *                    query tiles of a tile back to back: one HBM fetch per tile, L2 re-reads from the same SMs), 0 = units
 *                    striped over the CTAs, -1 (default) = per path (1 for the int8 screen, 0 for the f16 one)
 *   "i8_cta_group"   2 (default) = CTA pairs share each int8 MMA (cta_group::2), 1 = one CTA per tile
 *   "center_screen"  1 (default) = the f16 / int8 images hold x - c, c = column means of the first rows added after a
 *                    reset, and the scan adds q.c back: embeddings with a large shared component (ANCE) get a
 *                    margin made of the centred norms; results are unaffected (exact rescore); empty index only
 *   "exchange_epoch"  arms the cross-shard threshold exchange for the next searches (see hac_set_threshold_exchange)
 *   "i8_auto_max_k" / "i8_auto_max_queries"  with the int8 image present, HAC_PATH_AUTO takes the int8 screen up to
 *                    this k / batch size (default: every k <= HAC_MAX_K, every batch); larger ones run the f16 screen,
 *                    0 = never
 *   "i8_large_k_rows_per_k"  ... and for k > 128 only on shards holding at least this many rows per k (default 12288;
 *                    batches below 128 queries: at least 256 rows per k and query); smaller shards run the f16 screen,
 *                    whose rescoring does not grow with k; 0 = no such limit
 *   "f16_drop_bits_corpus" / "f16_drop_bits_queries"  low mantissa bits of the f16 image forced to zero
 *                    (0..8, default 3 / 0): sparser operands draw less tensor-core power, the screen margin
 *                    is computed from the actual rounding error so exactness is unaffected; corpus: empty index only
 *   "lazy_f16"       -1 (default) = the f16 image is built on first use exactly when the int8 image exists, 0 / 1 = forced
 *   "i8_chunk_growth_x100"  int8 screen, synchronous chunks: chunk = growth * rows seen so far, in percent (0 = by batch
 *                    size and k)
 *   "i8_pipeline"    1 = pipelined int8 search: after a synchronous prelude the scans run back to back and the rescore +
 *                    refresh of chunk i run on a side stream beside the scan of chunk i+1 (HAC_I8_PIPELINE sets the
 *                    default of new handles); "i8_pipe_dist" (1 / 2), "i8_pipe_growth_x1000", "i8_pipe_min_rows" shape
 *                    its chunk schedule
 *   "i8_warm_rows"   int8 screen warm start: the first rows of the shard are searched with the f16 screen (an f16 image
 *                    of just those rows, ~0.6 GB at the default size), whose margin is ~20x tighter, so the int8 scan
 *                    of the rest starts from an exact threshold instead of emitting through many loosely filtered
 *                    early chunks; -1 (default) = automatic (tensor-bound batches; ~800 k rows for k <= 128 on shards
 *                    >= 8x the slab, 6144 * k rows up to a quarter of the shard for larger k), 0 = off, > 0 = that many rows
 */
int hac_set_option(hac_index* idx, const char* name, int64_t value);

/* ---- introspection --------------------------------------------------------------------------- */
int64_t hac_ntotal(const hac_index* idx);
int hac_dim(const hac_index* idx);
int hac_device(const hac_index* idx);
int hac_get_stats(const hac_index* idx, hac_stats* out);
int hac_abi_version(void);
const char* hac_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* HAC_INDEX_H_ */
