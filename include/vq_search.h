/*
 * vq_search.h — C-ABI of the B200-native frame-embedding search engine.
 *
 * This is the drop-in boundary for the query hot path of adhney/video-quierer.  The
 * reference is pure Python and has no FFI of its own, so each entry
 * point below cites the reference *code* it replaces; the ctypes binding a maintainer
 * adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain `extern "C"`, pointers + sizes only; no torch / C++ types cross the boundary;
 *   - every function returns 0 on success or a negative VQ_E* code; `vq_last_error()`
 *     returns a thread-local human-readable message for the last failure;
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - the caller owns all memory (torch allocates; the library borrows for the call);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises the device unless stated;
 *   - rows of a store are `ld` elements apart (`ld >= dim`, padding must be zero;
 *     `ld % 32 == 0` for fp32 stores, `ld % 64 == 0` for bf16 stores);
 *   - result order everywhere: score descending, ties by ascending row; missing results
 *     (k > n, NaN scores) have row = -1 and score = -inf.
 */
#ifndef VQ_SEARCH_H_
#define VQ_SEARCH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ_ABI_VERSION 1

/* element types of a frame-embedding store */
#define VQ_F32  0
#define VQ_BF16 1

/* error codes */
#define VQ_OK            0
#define VQ_EINVAL       -1   /* bad argument (message says which) */
#define VQ_ECUDA        -2   /* a CUDA call / launch failed */
#define VQ_EWORKSPACE   -3   /* workspace too small */
#define VQ_EUNSUPPORTED -4   /* shape not supported by the requested path */

/* query normalisation applied inside vq_scan_topk / vq_hnsw_search */
#define VQ_NORM_NONE   0     /* queries are used as given */
#define VQ_NORM_EPS    1     /* q / (|q| + 1e-10)  — video_search_overhaul.py:49-50 */
#define VQ_NORM_PLAIN  2     /* q / |q|            — src/indexes/hnsw.py:250,499 */

/* scan path selection */
#define VQ_SCAN_AUTO   0     /* pick by batch size and store dtype */
#define VQ_SCAN_FMA    1     /* fp32-FMA HBM-streaming path ("GEMV" path), any store dtype */
#define VQ_SCAN_MMA    2     /* tcgen05/TMEM tile-GEMM path (kind::f16 over a bf16 store; an fp32 store is refused
                                with VQ_EUNSUPPORTED: exact fp32 results at tensor-core speed come from
                                vq_search_exact over the bf16 + fp32 twins) */
#define VQ_SCAN_FMA32  3     /* the FMA path with its widest query tile (32 per pass; experiments) */

int         vq_abi_version(void);
const char* vq_last_error(void);
/* Name of the scan kernel the last vq_scan_topk call on this thread dispatched to and how
 * many kernels it launched (for bench.py's gpu_launches accounting). */
const char* vq_last_scan_path(void);
int         vq_last_launch_count(void);
/* Roofline instrumentation: when enabled, vq_scan_topk / vq_hnsw_search bracket their dominant
 * kernel with CUDA events on the launch stream; vq_profile_last_kernel_ms() waits for the end
 * event and returns that kernel's device time in ms (negative if none was recorded). */
int         vq_profile_enable(int on);
float       vq_profile_last_kernel_ms(void);

/* (a) L2-normalisation of embeddings, in place.                       [kernel: l2norm_rows]
 * Replaces: `embedding / embedding.norm()` video_search_overhaul.py:226,289;
 *           `vector / np.linalg.norm(vector)` src/indexes/hnsw.py:157.
 * x: [rows, ld] fp32; eps_mode is VQ_NORM_EPS or VQ_NORM_PLAIN. */
int vq_l2_normalize(float* x, int64_t rows, int dim, int ld, int eps_mode, void* stream);

/* Store ingest: copy `rows` fp32 embeddings (src: [rows, src_ld]) into a store of dtype
 * `dst_dtype` (dst: [rows, dst_ld], zero-padding columns dim..dst_ld), optionally
 * normalising each row first (norm_mode = VQ_NORM_*).               [kernel: ingest_rows]
 * Replaces: `self.embeddings.append(embedding.astype(np.float32))` video_search_overhaul.py:33
 * and the per-query `np.vstack(self.embeddings)` (:46) — the matrix is built once. */
int vq_ingest_rows(const float* src, int64_t rows, int dim, int src_ld,
                   void* dst, int dst_dtype, int dst_ld, int norm_mode, void* stream);

/* (b)+(c) exact batched query x frame inner-product scan with fused per-query top-k.
 * Replaces: SimpleVideoIndex.search, video_search_overhaul.py:40-64 (normalise :49-50,
 * np.dot :53, argsort top-k :56) and the sequential batch loop src/api/routes.py:627-634.
 *   store      [n, ld] of store_dtype, rows assumed unit-norm (never re-normalised, :53)
 *   queries    [b, dim] fp32 (dense, row stride = dim)
 *   out_scores [b, k] fp32, out_rows [b, k] int32  (best first)
 *   workspace  >= vq_scan_workspace_bytes(...) bytes, 256-byte aligned
 * Scores never round-trip to HBM: only per-CTA top-k candidate lists are written. */
size_t vq_scan_workspace_bytes(int64_t n, int dim, int ld, int store_dtype, int b, int k, int path);
int vq_scan_topk(const void* store, int64_t n, int dim, int ld, int store_dtype,
                 const float* queries, int b, int k, int query_norm,
                 float* out_scores, int32_t* out_rows,
                 void* workspace, size_t workspace_bytes, int path, void* stream);

/* Merge g candidate lists per query into one global top-k_out.        [kernel: topk_merge]
 * New (the reference is single-process); this is the shard/merge layer of SURVEY.md §8(e).
 *   scores/rows  [g, b, k_in] (fp32 / int32, local row numbers, row < 0 = empty slot);
 *                consecutive shards are g_stride elements apart (0 = dense, b*k_in), so one
 *                packed all-gather buffer [g][scores|rows] can be merged in place
 *   shard_offsets[g] int64 added to local rows (may be NULL = all zero)
 *   out_scores   [b, k_out] fp32, out_rows [b, k_out] int64 (global rows) */
int vq_topk_merge(const float* scores, const int32_t* rows, int g, int64_t g_stride, int b, int k_in,
                  const int64_t* shard_offsets, int k_out,
                  float* out_scores, int64_t* out_rows, void* stream);

/* Shard/merge over NVLink peer memory: ONE kernel per rank replaces ncclAllGather of the per-GPU
 * top-k + vq_topk_merge (the reference is single-process and has no counterpart).
 * [kernel: peer_exchange_merge]
 * Every rank owns an exchange window (header + double-buffered 16-byte candidate lines for `world`
 * ranks x b_max queries x k_max entries; a line = {score, epoch, row, epoch}, each 8-byte half validates
 * itself, so neither side needs a fence) that all peers map through CUDA IPC:
 *   vq_peer_window_create   cudaMalloc + zero + IPC handle (VQ_PEER_HANDLE_BYTES bytes, host memory) —
 *                           the handles are exchanged by the host (e.g. torch.distributed.all_gather_object)
 *   vq_peer_window_open     maps a PEER's window into this process (enables peer access lazily)
 *   vq_peer_window_close / vq_peer_window_destroy / vq_peer_window_status (host copy of epoch + error word)
 * vq_peer_exchange_merge is a collective over the ranks sharing the windows: all ranks call it with the
 * same b, k, k_out, in the same order, one stream per window set.  It pushes this rank's local candidates
 *   scores / rows [b, k] (best first per query; row < 0 = empty slot, local row numbers)
 * into every peer's window with 16-byte stores over NVLink, polls (one warp per query, no barrier) its own
 * window until the peers' lines of the same query carry this epoch and merges world*k -> k_out (score desc,
 * global row asc):
 *   windows_dev   device array of `world` window base pointers, index = rank (own window included)
 *   shard_offsets [world] int64 added to local rows (may be NULL)
 *   out_scores [b, k_out] fp32, out_rows [b, k_out] int64 — identical on every rank
 *   out_status    optional device int32, set to 1 if a peer did not arrive within VQ_PEER_TIMEOUT_MS
 *                 (default 10 s; the error is also sticky in the window header)
 * The epoch counter lives in the window, so the launch can be captured in a CUDA graph and replayed. */
#define VQ_PEER_HANDLE_BYTES 64
size_t vq_peer_window_bytes(int world, int b_max, int k_max);
int vq_peer_window_create(size_t bytes, void** local_ptr, unsigned char* handle_out_host);
int vq_peer_window_open(const unsigned char* handle_host, void** peer_ptr);
int vq_peer_window_close(void* peer_ptr);
int vq_peer_window_destroy(void* local_ptr);
int vq_peer_window_status(const void* local_ptr, uint32_t* epoch_host, uint32_t* error_host);
int vq_peer_exchange_merge(const void* windows_dev, int world, int rank, int b_max, int k_max,
                           const float* scores, const int32_t* rows, int b, int k,
                           const int64_t* shard_offsets, int k_out,
                           float* out_scores, int64_t* out_rows, int32_t* out_status, void* stream);

/* Ingest side of a sharded step: all-gather of the query batch over NVLink peer memory, so that every
 * rank copies only ITS slice of the host batch over PCIe (rows [rank*per, min(b, (rank+1)*per)), per =
 * ceil(b / world)).  Same windows (vq_peer_window_create/open with vq_peer_rows_window_bytes), same
 * self-validating line protocol and collective rules as vq_peer_exchange_merge, its own window set.
 *   slice  [rows of this rank, dim] fp32 dense (dim even)     out  [b, dim] fp32 dense, identical on all ranks
 * Replaces: nothing in the reference (single process); SURVEY.md 8(e) "queries are replicated to all
 * ranks".                                                                 [kernel: peer_allgather_rows] */
size_t vq_peer_rows_window_bytes(int world, int b_max, int ld_max);
int vq_peer_allgather_rows(const void* windows_dev, int world, int rank, int b_max, int ld_max,
                           const float* slice, int b, int dim, float* out, int32_t* out_status, void* stream);

/* Exact fp32 re-score of candidate rows (two-stage mode: bf16 scan selects k_cand, this
 * re-scores them from the fp32 shadow store and keeps the best k).     [kernel: rescore_rows]
 *   cand_rows [b, k_cand] int32 (row < 0 ignored); queries [b, ld] fp32 *already normalised*
 *   and zero padded to the store's ld (i.e. the output of vq_ingest_rows on the queries)
 *   out_scores [b, k] fp32, out_rows [b, k] int32; workspace >= b*k_cand*4 bytes */
int vq_rescore_topk(const float* store_f32, int64_t n, int dim, int ld,
                    const float* queries, int b, const int32_t* cand_rows, int k_cand, int k,
                    float* out_scores, int32_t* out_rows,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Exact search at tensor-core speed, ONE call, ONE pass over the store — the default search path of
 * B200FlatIndex and what bench.py measures.                  [kernels: scan_mma_bf16<exact>, exact_finish]
 * Replaces: SimpleVideoIndex.search, video_search_overhaul.py:40-64, with identical results: the ids
 * are those of the fp32 scan and the scores are fp32 FMA dot products of the fp32 rows.
 *   store_bf16 / store_f32  [n, ld] twins of the SAME rows (same ld, ld % 64 == 0, ld <= 768)
 *   store_bounds  device float[2] = {max |x^|, max |x^ - x|} over the rows (x^ = the bf16 row, x = the fp32
 *                 row), maintained with vq_store_bounds; any upper bounds are valid
 *   out_overflow  [b] int32: 1 = more rows than the gather buffer holds came within the error bound of the
 *                 k-th best (mass ties / duplicates) — re-run that query with vq_scan_topk on the fp32 store
 *   out_stats     optional [b, 2] int32: rows the scan gathered, rows re-scored in fp32 (roofline accounting)
 * How: the tensor-core scan of the bf16 copy keeps a running k-th best S and gathers every row whose
 * bf16-operand score reaches S - 2*eps_q, eps_q = |q^ - q| max|x^| + |q| max|x^ - x| + fp32 accumulation
 * slack (a Cauchy-Schwarz bound of |q^.x^ - q.x|, computed per query).  The exact k-th best s_k is
 * >= S_final - eps_q, so every row of the exact top-k has bf16 score >= s_k - eps_q >= S - 2*eps_q and is
 * gathered.  exact_finish re-scores the best candidates in fp32, which yields a lower bound s_lb <= s_k
 * reached by k real rows, then every other candidate with bf16 score >= s_lb - eps_q, and returns the best k
 * by (exact score desc, row asc).
 * k <= 64: the running k-th best comes from per-query register lists plus the k-th largest of the maxima all
 * CTAs publish (cooperative bound); 64 < k <= 128 (BASELINE config 4: k = 100): the cooperative bound alone.
 * vq_search_exact_supported says whether a shape is served (ld <= 768, ld % 64 == 0, k <= 128, and for k > 64 a
 * store large enough to bootstrap the bound); otherwise use vq_search_collect / vq_scan_topk. */
int    vq_search_exact_supported(int64_t n, int dim, int ld, int b, int k);
size_t vq_search_exact_workspace_bytes(int64_t n, int dim, int ld, int b, int k);
int vq_search_exact(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                    const float* queries, int b, int k, int query_norm, const float* store_bounds,
                    float* out_scores, int32_t* out_rows, int32_t* out_overflow, int32_t* out_stats,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Fold rows [0, rows) of the twins into store_bounds (atomic max; zero the two floats once, then call after
 * every append with the appended rows).                                        [kernel: store_bounds] */
int vq_store_bounds(const float* store_f32, const void* store_bf16, int64_t rows, int ld,
                    float* store_bounds, void* stream);

/* Two-stage exact search in ONE call (superseded by vq_search_exact; kept for k_cand experiments): the tensor-core scan of the bf16
 * copy selects k_cand candidates per query, they are re-scored exactly (fp32 FMA, same arithmetic as
 * vq_rescore_topk) from the fp32 copy of the SAME rows, and the best k by exact score are returned.
 * Replaces the same reference code as vq_scan_topk (video_search_overhaul.py:40-64).
 *   store_bf16 / store_f32  [n, ld] twins (same ld, ld % 64 == 0)
 *   queries    [b, dim] raw fp32; query_norm as in vq_scan_topk
 *   score_eps  bound on |bf16-operand score - fp32 score|: both operands are rounded to 8 significant bits
 *              (unit roundoff 2^-8 each), so (2^-7 + 2^-16) * |q| * max row norm + accumulation slack
 *   out_uncertified [b] int32: 0 = the result is provably the exact top-k (every row outside the
 *              candidate set has exact score <= k_cand-th candidate score + score_eps < k-th exact
 *              score); 1 = could not be certified (near-duplicate heavy data): re-run that query with
 *              vq_scan_topk on the fp32 store.                 [kernels: scan_mma_bf16, scan_finish] */
size_t vq_search_two_stage_workspace_bytes(int64_t n, int dim, int ld, int b, int k_cand);
int vq_search_two_stage(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                        const float* queries, int b, int k, int k_cand, int query_norm, float score_eps,
                        float* out_scores, int32_t* out_rows, int32_t* out_uncertified,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Collect pass — the exact answer for queries vq_search_two_stage could not certify, still at
 * tensor-core speed: every row whose bf16-operand score reaches thresholds[q] is gathered (at most cap
 * per query), ALL gathered rows are re-scored exactly from the fp32 copy and the best k by exact score
 * are returned.  With thresholds[q] = s_k - score_eps, s_k being any lower bound of the exact k-th best
 * score (e.g. the k-th score vq_search_two_stage returned), every row of the exact top-k is gathered:
 * its exact score is >= s_k, so its bf16-operand score is >= s_k - score_eps.
 *   thresholds   [b] fp32 (device), or NULL: derived from the store itself (the k-th largest per-tile
 *                maximum of up to 256 sample tiles, reached by k distinct rows, k <= 256) — at least k
 *                rows are gathered, and the k-th exact score returned is a lower bound of the k-th best of
 *                any store containing these rows (step 1 of the large-k search);
 *   cap          candidate slots per query (k <= cap <= 16384), k <= 1024
 *   store_bounds NULL: every gathered row is re-scored in fp32.  Else the store's device float[2] of vq_store_bounds:
 *                only the best candidates by bf16 score and those within the per-query rounding bound of the k-th
 *                exact score among them are re-scored (same argument and final stage as vq_search_exact) — what
 *                makes k = 100 over millions of rows cheap (BASELINE config 4)
 *   out_overflow [b] int32: 1 = more than cap rows reached the threshold, the result is incomplete
 *                (re-run that query with vq_scan_topk on the fp32 store). [kernels: scan_mma_bf16<collect>, scan_finish] */
size_t vq_search_collect_workspace_bytes(int64_t n, int dim, int ld, int b, int cap);
int vq_search_collect(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                      const float* queries, int b, int k, int query_norm, const float* thresholds, int cap,
                      const float* store_bounds, float* out_scores, int32_t* out_rows, int32_t* out_overflow,
                      void* workspace, size_t workspace_bytes, void* stream);

/* (d) HNSW greedy/beam search, one warp per query.                     [kernel: hnsw_search]
 * Replaces: HNSWIndex.search / OptimizedHNSWIndex.search, src/indexes/hnsw.py:238-280,
 * :488-528 and _search_layer :76-121 (same stop rule :103 and admit rule :113).
 *   store      [n, ld] of store_dtype (unit-norm rows)
 *   levels     [n] int32 top level of each node
 *   adj0       [n, m0] int32 layer-0 neighbours, -1 padded
 *   upper_off  [n] int32 first upper-layer slot of a node (-1 if level 0)
 *   upper_adj  [slots, m] int32: slot(node, lv) = upper_off[node] + lv - 1
 *   entry / max_level: entry point and its level;  ef = max(ef_search, k) like :264
 *   out_dist   [b, k] fp32 (1 - dot, ascending), out_rows [b, k] int32
 *   out_stats  optional [b, 4] uint32: distance evaluations, expanded nodes, visited-set
 *              overflow flag, reserved (feeds the gather-bandwidth roofline)
 *   visited_capacity  slots of the per-query visited set (0 = 16*ef); a query whose set fills up
 *              sets its overflow flag and should be re-run with a larger capacity */
size_t vq_hnsw_workspace_bytes(int b, int ld, int ef);
int vq_hnsw_search(const void* store, int64_t n, int dim, int ld, int store_dtype,
                   const int32_t* levels, const int32_t* adj0, int m0,
                   const int32_t* upper_off, const int32_t* upper_adj, int m,
                   int32_t entry, int max_level, int ef,
                   const float* queries, int b, int k, int query_norm,
                   float* out_dist, int32_t* out_rows, uint32_t* out_stats, int visited_capacity,
                   void* workspace, size_t workspace_bytes, void* stream);

/* HNSW construction on the GPU (the reference's per-insert Python build,
 * src/indexes/hnsw.py:150-229, is ~10 ms/insert and unusable at 1M).   [kernels: knn_*, link_*]
 * Builds one layer: for every member node, its `m_out` graph neighbours chosen from the
 * exact `k_cand` nearest members (brute-force scan + fused top-k), then reverse edges are
 * added and every list is pruned back to `m_out` closest (hnsw.py:197-223 semantics).
 *   members    [n_members] int32 node ids participating in this layer (NULL = all n rows)
 *   adj_out    [n_members, m_out] int32 (-1 padded), neighbours as node ids
 *   diversify  0 = closest-m of the exact k nearest (reference selection rule hnsw.py:123-148 on exact candidates),
 *              1 = HNSW diversity heuristic over the exact k_cand nearest,
 *              2 = incremental (k_cand == m_out): the reference's construction ORDER — node i links to its m_out nearest
 *                  among the nodes before it (hnsw.py:183-199), every node keeps the closest m_out of all links it ever
 *                  received (:202-223) — computed data-parallel (causal scan + one sort), not insert by insert,
 *              3 = sequential (k_cand == m_out): the reference's add() with exact candidates, batch by batch — links in both
 *                  directions, closest-m_out prune that removes the dropped link at BOTH ends (:217-223), so lists stay
 *                  short of m_out and the long links of early inserts survive */
size_t vq_hnsw_layer_workspace_bytes(int64_t n_members, int dim, int ld, int store_dtype, int k_cand, int m_out);
int vq_hnsw_build_layer(const void* store, int64_t n, int dim, int ld, int store_dtype,
                        const int32_t* members, int64_t n_members,
                        int k_cand, int m_out, int diversify,
                        int32_t* adj_out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQ_SEARCH_H_ */
