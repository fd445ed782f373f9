// C-ABI entry points (include/vq_search.h) for normalise / ingest / exact scan / merge / rescore.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "vq_common.cuh"

// ---- kernels implemented in the other translation units
int vq_scan_fma_tile_rows(int bt);
size_t vq_scan_fma_smem(int bt, int ld, int k);
int vq_scan_fma_grid(int n, int bt);
int vq_scan_fma_launch(const void* store, int n, int ld, int store_dtype, const float* q, int bt, int k,
                       float* part_scores, int* part_rows, int grid, cudaStream_t stream);
int vq_topk_merge_launch(const float* scores, const int* rows, int g, long long g_stride, int b_out, int k_in,
                         const long long* offsets, int k_out, float* out_scores, void* out_rows,
                         int rows64, int negate_out, cudaStream_t stream);
int vq_ingest_launch(const float* src, long long rows, int dim, int src_ld, void* dst, int dst_dtype,
                     int dst_ld, int mode, cudaStream_t stream);
int vq_rescore_launch(const float* store, int ld, const float* queries, int qld, const int* cand, int b,
                      int k_cand, float* out_scores, cudaStream_t stream);
bool vq_scan_finish_lists_supported(int g, int k_in, int k_out);
int vq_scan_finish_lists_launch(const float* scores, const int* rows, int g, long long g_stride, int b, int k_in, int k_out,
                                float* out_scores, int* out_rows, cudaStream_t stream);
// tcgen05 path (scan_mma.cu)
bool vq_scan_mma_supported(int64_t n, int dim, int ld, int store_dtype, int b, int k);
size_t vq_scan_mma_workspace(int64_t n, int ld, int store_dtype, int b, int k);
int vq_scan_mma_run(const void* store, int64_t n, int dim, int ld, int store_dtype, const float* queries, int query_norm,
                    int b, int k_sel, const float* store_f32, float eps, int k_out, float* out_scores, int32_t* out_rows,
                    int32_t* out_bad, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches);

size_t vq_scan_mma_collect_workspace(int64_t n, int ld, int b, int cap);
int vq_scan_mma_collect(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld, const float* queries,
                        int query_norm, int b, const float* thresholds, int cap, const float* bounds, int k, float* out_scores,
                        int32_t* out_rows, int32_t* out_overflow, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches);

size_t vq_scan_mma_exact_workspace(int64_t n, int ld, int b, int k);
int vq_scan_mma_exact(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld, const float* queries,
                      int query_norm, int b, int k, const float* bounds, float* out_scores, int32_t* out_rows,
                      int32_t* out_overflow, int32_t* out_stats, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches);
bool vq_scan_mma_exact_supported(int64_t n, int ld, int b, int k);
int vq_store_bounds_launch(const float* f32, const void* bf16, long long rows, int ld, float* bounds, cudaStream_t stream);
int vq_fill_empty_launch(float* scores, int* rows, long long count, cudaStream_t stream);

// ----------------------------------------------------------------------------- error state
static thread_local char g_err[512] = "";
static thread_local char g_path[64] = "";
static thread_local int g_launches = 0;

void vq_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void vq_note_launch(const char* path, int launches) {
    if (path) { strncpy(g_path, path, sizeof(g_path) - 1); g_path[sizeof(g_path) - 1] = 0; }
    g_launches = launches;
}

// ---- roofline instrumentation (events around the dominant kernel of the last call)
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static thread_local bool g_ev_valid = false;
void vq_prof_begin(cudaStream_t s) {
    if (!g_prof_on) return;
    if (!g_ev0) { cudaEventCreate(&g_ev0); cudaEventCreate(&g_ev1); }
    cudaEventRecord(g_ev0, s);
    g_ev_valid = false;
}
void vq_prof_end(cudaStream_t s) {
    if (!g_prof_on || !g_ev0) return;
    cudaEventRecord(g_ev1, s);
    g_ev_valid = true;
}

// PDL classes: 0 prep/reset, 1 boot scan, 2 boot_select, 3 main scan, 4 finish.  Default: off (measured
// on B200 with the whole chain enabled: +35 us per step — see DESIGN.md); VQ_PDL=<mask> to experiment.
int vq_pdl_mask() {
    static const int mask = getenv("VQ_PDL") ? atoi(getenv("VQ_PDL")) : 0;
    return mask;
}

// Launch classes 0-3 (query prep, threshold bootstrap, boot_select, main scan) form the critical chain of a
// search step; class 4 (final selection + re-score) and the shard exchange are its tail.  With several steps
// in flight on different streams (graphs.PipelinedSearch) the tail of step i and the head of step i+1 become
// ready at the same moment.  VQ_PRIO=1 launches the head at the highest stream priority so that its CTAs
// are placed first.  Measured and NOT adopted (default off): on one GPU scanning a 125k-row shard at batch
// 1024 the step went 0.189 -> 0.210 ms with two steps in flight and 0.184 -> 0.187 ms with three.
int vq_launch_priority(int launch_class) {
    static const int on = getenv("VQ_PRIO") ? atoi(getenv("VQ_PRIO")) : 0;
    static int hi = 0;
    static bool init = false;
    if (!on) return 0;
    if (!init) {
        int lo_p = 0, hi_p = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p) == cudaSuccess) hi = hi_p;   // numerically lowest = highest
        init = true;
    }
    return launch_class <= 3 ? hi : 0;
}

int vq_num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int pow2_at_least(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static int check_store(int64_t n, int dim, int ld, int store_dtype) {
    VQ_CHECK_ARG(store_dtype == VQ_F32 || store_dtype == VQ_BF16, "store_dtype must be VQ_F32 or VQ_BF16, got %d", store_dtype);
    VQ_CHECK_ARG(n >= 0 && n < (int64_t)INT_MAX, "n=%lld out of range (shard rows must fit int32)", (long long)n);
    VQ_CHECK_ARG(dim > 0 && ld >= dim, "need 0 < dim <= ld (dim=%d ld=%d)", dim, ld);
    const int mult = store_dtype == VQ_BF16 ? 64 : 32;
    VQ_CHECK_ARG(ld % mult == 0, "ld=%d must be a multiple of %d for this store dtype", ld, mult);
    return VQ_OK;
}

struct ScanPlan {
    int bt;            // FMA query tile
    int grid;          // FMA grid
    size_t q_bytes, part_bytes;
};
static ScanPlan fma_plan(int64_t n, int ld, int b, int k, int force_bt) {
    ScanPlan p;
    p.bt = force_bt ? force_bt : (b >= 16 ? 16 : pow2_at_least(b));
    // shrink the tile until its shared memory fits
    while (p.bt > 1 && vq_scan_fma_smem(p.bt, ld, k) > 200 * 1024) p.bt >>= 1;
    p.grid = vq_scan_fma_grid((int)n, p.bt);
    // the per-CTA lists are reduced by scan_finish (bisection select, one CTA per query) when they fit its
    // shared-memory key pool; for large k the grid is trimmed to make them fit (never below 128 CTAs)
    if ((long long)p.grid * k > 16384 && 16384 / k >= 128) p.grid = 16384 / k;
    const int b_pad = (int)align_up((size_t)b, (size_t)p.bt);
    p.q_bytes = align_up((size_t)b_pad * ld * 4, 256);
    p.part_bytes = align_up((size_t)p.grid * p.bt * k * 4, 256);
    return p;
}

extern "C" {

int vq_abi_version(void) { return VQ_ABI_VERSION; }
const char* vq_last_error(void) { return g_err; }
const char* vq_last_scan_path(void) { return g_path; }
int vq_last_launch_count(void) { return g_launches; }

int vq_profile_enable(int on) { g_prof_on = on != 0; return VQ_OK; }
float vq_profile_last_kernel_ms(void) {
    if (!g_ev_valid) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(g_ev1) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.f;
    return ms;
}

int vq_l2_normalize(float* x, int64_t rows, int dim, int ld, int eps_mode, void* stream) {
    VQ_CHECK_ARG(x != nullptr || rows == 0, "x is NULL");
    VQ_CHECK_ARG(rows >= 0 && dim > 0 && ld >= dim, "bad shape rows=%lld dim=%d ld=%d", (long long)rows, dim, ld);
    VQ_CHECK_ARG(eps_mode == VQ_NORM_EPS || eps_mode == VQ_NORM_PLAIN, "eps_mode must be VQ_NORM_EPS or VQ_NORM_PLAIN");
    const int rc = vq_ingest_launch(x, rows, dim, ld, x, VQ_F32, ld, eps_mode, (cudaStream_t)stream);
    vq_note_launch("l2norm_rows", rows ? 1 : 0);
    return rc;
}

int vq_ingest_rows(const float* src, int64_t rows, int dim, int src_ld, void* dst, int dst_dtype, int dst_ld,
                   int norm_mode, void* stream) {
    VQ_CHECK_ARG((src && dst) || rows == 0, "src/dst is NULL");
    VQ_CHECK_ARG(rows >= 0 && dim > 0 && src_ld >= dim && dst_ld >= dim, "bad shape rows=%lld dim=%d src_ld=%d dst_ld=%d",
                 (long long)rows, dim, src_ld, dst_ld);
    VQ_CHECK_ARG(dst_dtype == VQ_F32 || dst_dtype == VQ_BF16, "dst_dtype must be VQ_F32 or VQ_BF16");
    VQ_CHECK_ARG(norm_mode >= VQ_NORM_NONE && norm_mode <= VQ_NORM_PLAIN, "bad norm_mode %d", norm_mode);
    const int rc = vq_ingest_launch(src, rows, dim, src_ld, dst, dst_dtype, dst_ld, norm_mode, (cudaStream_t)stream);
    vq_note_launch("ingest_rows", rows ? 1 : 0);
    return rc;
}

size_t vq_scan_workspace_bytes(int64_t n, int dim, int ld, int store_dtype, int b, int k, int path) {
    (void)dim;
    if (n <= 0 || b <= 0 || k <= 0) return 256;
    const ScanPlan p = fma_plan(n, ld, b, k, path == 3 ? 32 : 0);
    size_t fma = p.q_bytes + 2 * p.part_bytes;
    size_t mma = 0;
    if (path != VQ_SCAN_FMA && path != 3) mma = vq_scan_mma_workspace(n, ld, store_dtype, b, k);
    return (fma > mma ? fma : mma) + 256;
}

int vq_scan_topk(const void* store, int64_t n, int dim, int ld, int store_dtype, const float* queries, int b, int k,
                 int query_norm, float* out_scores, int32_t* out_rows, void* workspace, size_t workspace_bytes,
                 int path, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = check_store(n, dim, ld, store_dtype);
    if (rc) return rc;
    VQ_CHECK_ARG(b >= 0 && k > 0 && k <= 1024, "need b >= 0 and 0 < k <= 1024 (b=%d k=%d)", b, k);
    VQ_CHECK_ARG(query_norm >= VQ_NORM_NONE && query_norm <= VQ_NORM_PLAIN, "bad query_norm %d", query_norm);
    VQ_CHECK_ARG(path >= 0 && path <= 3, "bad path %d", path);
    if (b == 0) { vq_note_launch("none", 0); return VQ_OK; }
    VQ_CHECK_ARG(store && queries && out_scores && out_rows && workspace, "NULL pointer argument");
    VQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    VQ_CHECK_ARG(((uintptr_t)store & 15) == 0, "store must be 16-byte aligned");
    const size_t need = vq_scan_workspace_bytes(n, dim, ld, store_dtype, b, k, path);
    if (workspace_bytes < need) {
        vq_set_error("workspace too small: %zu < %zu", workspace_bytes, need);
        return VQ_EWORKSPACE;
    }
    int launches = 0;
    if (n == 0) {   // empty store: all slots empty (the reference returns [] — video_search_overhaul.py:42-43)
        rc = vq_fill_empty_launch(out_scores, out_rows, (long long)b * k, stream);   // score -inf, row -1 (header contract)
        vq_note_launch("empty", 1);
        return rc;
    }

    bool use_mma = false;
    if (path == VQ_SCAN_MMA) {
        if (!vq_scan_mma_supported(n, dim, ld, store_dtype, b, k)) {
            vq_set_error("tcgen05 scan path does not support n=%lld dim=%d ld=%d dtype=%d b=%d k=%d", (long long)n, dim,
                         ld, store_dtype, b, k);
            return VQ_EUNSUPPORTED;
        }
        use_mma = true;
    } else if (path == VQ_SCAN_AUTO) {
        // bf16 store: the tensor path streams the store at ~HBM speed for every batch size
        // (measured 5.7 TB/s at b=1 vs 4.6 TB/s for the FMA path), so it is always preferred.
        // fp32 store: kind::tf32 would break the 1e-5 score parity, so AUTO never picks it.
        use_mma = store_dtype == VQ_BF16 && vq_scan_mma_supported(n, dim, ld, store_dtype, b, k);
    }

    unsigned char* ws = (unsigned char*)workspace;
    if (use_mma) {
        int l2 = 0;
        rc = vq_scan_mma_run(store, n, dim, ld, store_dtype, queries, query_norm, b, k, nullptr, 0.f, k, out_scores,
                             out_rows, nullptr, ws, workspace_bytes, stream, &l2);
        if (rc) return rc;
        vq_note_launch("scan_mma_bf16", l2);
        return VQ_OK;
    }

    const ScanPlan p = fma_plan(n, ld, b, k, path == 3 ? 32 : 0);
    float* qn = (float*)ws;                               // [b_pad, ld] normalised, zero padded
    const int b_pad = (int)align_up((size_t)b, (size_t)p.bt);
    VQ_CUDA(cudaMemsetAsync(qn, 0, (size_t)b_pad * ld * 4, stream));
    rc = vq_ingest_launch(queries, b, dim, dim, qn, VQ_F32, ld, query_norm, stream);
    if (rc) return rc;
    launches += 1;

    float* part_s = (float*)(ws + p.q_bytes);
    int* part_r = (int*)(ws + p.q_bytes + p.part_bytes);
    for (int q0 = 0; q0 < b; q0 += p.bt) {
        int bt = p.bt;
        const int left = b - q0;
        if (left < bt) bt = pow2_at_least(left);            // padded queries are all-zero rows of qn
        int grid = vq_scan_fma_grid((int)n, bt);
        if (grid > p.grid) grid = p.grid;
        if (q0 == 0) vq_prof_begin(stream);
        rc = vq_scan_fma_launch(store, (int)n, ld, store_dtype, qn + (size_t)q0 * ld, bt, k, part_s, part_r, grid, stream);
        if (q0 == 0) vq_prof_end(stream);
        if (rc) return rc;
        if (vq_scan_finish_lists_supported(grid, k, k))
            rc = vq_scan_finish_lists_launch(part_s, part_r, grid, (long long)bt * k, left < bt ? left : bt, k, k,
                                             out_scores + (size_t)q0 * k, out_rows + (size_t)q0 * k, stream);
        else
            rc = vq_topk_merge_launch(part_s, part_r, grid, (long long)bt * k, left < bt ? left : bt, k, nullptr, k,
                                      out_scores + (size_t)q0 * k, out_rows + (size_t)q0 * k, 0, 0, stream);
        if (rc) return rc;
        launches += 2;
    }
    vq_note_launch(store_dtype == VQ_BF16 ? "scan_fma_bf16" : "scan_fma_f32", launches);
    return VQ_OK;
}

int vq_topk_merge(const float* scores, const int32_t* rows, int g, int64_t g_stride, int b, int k_in,
                  const int64_t* shard_offsets, int k_out, float* out_scores, int64_t* out_rows, void* stream) {
    if (g_stride == 0) g_stride = (int64_t)b * k_in;
    VQ_CHECK_ARG(g_stride >= (int64_t)b * k_in, "g_stride %lld < b*k_in", (long long)g_stride);
    VQ_CHECK_ARG(g > 0 && b >= 0 && k_in > 0 && k_out > 0 && k_out <= 1024, "bad shape g=%d b=%d k_in=%d k_out=%d", g, b, k_in, k_out);
    if (b == 0) return VQ_OK;
    VQ_CHECK_ARG(scores && rows && out_scores && out_rows, "NULL pointer argument");
    const int rc = vq_topk_merge_launch(scores, rows, g, (long long)g_stride, b, k_in, (const long long*)shard_offsets, k_out, out_scores,
                                        out_rows, 1, 0, (cudaStream_t)stream);
    vq_note_launch("topk_merge", 1);
    return rc;
}

int vq_rescore_topk(const float* store_f32, int64_t n, int dim, int ld, const float* queries, int b,
                    const int32_t* cand_rows, int k_cand, int k, float* out_scores, int32_t* out_rows,
                    void* workspace, size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = check_store(n, dim, ld, VQ_F32);
    if (rc) return rc;
    VQ_CHECK_ARG(b >= 0 && k_cand > 0 && k > 0 && k <= k_cand && k <= 1024, "bad shape b=%d k_cand=%d k=%d", b, k_cand, k);
    if (b == 0) return VQ_OK;
    VQ_CHECK_ARG(store_f32 && queries && cand_rows && out_scores && out_rows && workspace, "NULL pointer argument");
    if (workspace_bytes < (size_t)b * k_cand * 4) {
        vq_set_error("rescore workspace too small: %zu < %zu", workspace_bytes, (size_t)b * k_cand * 4);
        return VQ_EWORKSPACE;
    }
    float* tmp = (float*)workspace;                       // the b*k_cand exact scores
    rc = vq_rescore_launch(store_f32, ld, queries, ld, cand_rows, b, k_cand, tmp, stream);
    if (rc == VQ_OK)
        rc = vq_topk_merge_launch(tmp, cand_rows, 1, (long long)b * k_cand, b, k_cand, nullptr, k, out_scores, out_rows, 0, 0, stream);
    vq_note_launch("rescore_rows", 2);
    return rc;
}

size_t vq_search_two_stage_workspace_bytes(int64_t n, int dim, int ld, int b, int k_cand) {
    (void)dim;
    if (n <= 0 || b <= 0 || k_cand <= 0) return 256;
    return vq_scan_mma_workspace(n, ld, VQ_BF16, b, k_cand) + 256;
}

int vq_search_two_stage(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                        const float* queries, int b, int k, int k_cand, int query_norm, float score_eps,
                        float* out_scores, int32_t* out_rows, int32_t* out_uncertified,
                        void* workspace, size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = check_store(n, dim, ld, VQ_BF16);
    if (rc) return rc;
    VQ_CHECK_ARG(b >= 0 && k > 0 && k_cand >= k, "need b >= 0 and 0 < k <= k_cand (b=%d k=%d k_cand=%d)", b, k, k_cand);
    VQ_CHECK_ARG(query_norm >= VQ_NORM_NONE && query_norm <= VQ_NORM_PLAIN, "bad query_norm %d", query_norm);
    VQ_CHECK_ARG(score_eps >= 0.f, "score_eps must be >= 0");
    if (b == 0) { vq_note_launch("none", 0); return VQ_OK; }
    VQ_CHECK_ARG(n > 0, "two-stage search needs a non-empty store");
    VQ_CHECK_ARG(store_bf16 && store_f32 && queries && out_scores && out_rows && out_uncertified && workspace,
                 "NULL pointer argument");
    VQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    VQ_CHECK_ARG(((uintptr_t)store_bf16 & 15) == 0 && ((uintptr_t)store_f32 & 15) == 0, "stores must be 16-byte aligned");
    if (!vq_scan_mma_supported(n, dim, ld, VQ_BF16, b, k_cand)) {
        vq_set_error("two-stage search does not support n=%lld dim=%d ld=%d b=%d k_cand=%d", (long long)n, dim, ld, b, k_cand);
        return VQ_EUNSUPPORTED;
    }
    int launches = 0;
    rc = vq_scan_mma_run(store_bf16, n, dim, ld, VQ_BF16, queries, query_norm, b, k_cand, store_f32, score_eps, k,
                         out_scores, out_rows, out_uncertified, workspace, workspace_bytes, stream, &launches);
    if (rc) return rc;
    vq_note_launch("scan_mma_bf16+rescore", launches);
    return VQ_OK;
}

int vq_store_bounds(const float* store_f32, const void* store_bf16, int64_t rows, int ld, float* bounds, void* stream) {
    VQ_CHECK_ARG(rows >= 0 && ld > 0 && ld % 64 == 0, "bad shape rows=%lld ld=%d", (long long)rows, ld);
    VQ_CHECK_ARG(bounds != nullptr, "bounds is NULL");
    if (rows == 0) return VQ_OK;
    VQ_CHECK_ARG(store_f32 && store_bf16, "NULL store pointer");
    const int rc = vq_store_bounds_launch(store_f32, store_bf16, rows, ld, bounds, (cudaStream_t)stream);
    vq_note_launch("store_bounds", 1);
    return rc;
}

size_t vq_search_exact_workspace_bytes(int64_t n, int dim, int ld, int b, int k) {
    (void)dim;
    if (n <= 0 || b <= 0 || k <= 0) return 256;
    return vq_scan_mma_exact_workspace(n, ld, b, k) + 256;
}

int vq_search_exact_supported(int64_t n, int dim, int ld, int b, int k) {
    (void)dim;
    return vq_scan_mma_exact_supported(n, ld, b, k) ? 1 : 0;
}

int vq_search_exact(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                    const float* queries, int b, int k, int query_norm, const float* store_bounds,
                    float* out_scores, int32_t* out_rows, int32_t* out_overflow, int32_t* out_stats,
                    void* workspace, size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = check_store(n, dim, ld, VQ_BF16);
    if (rc) return rc;
    VQ_CHECK_ARG(b >= 0 && k > 0 && k <= 128, "need b >= 0 and 0 < k <= 128 (b=%d k=%d)", b, k);
    VQ_CHECK_ARG(query_norm >= VQ_NORM_NONE && query_norm <= VQ_NORM_PLAIN, "bad query_norm %d", query_norm);
    if (b == 0) { vq_note_launch("none", 0); return VQ_OK; }
    VQ_CHECK_ARG(n > 0, "exact search needs a non-empty store");
    VQ_CHECK_ARG(store_bf16 && store_f32 && queries && store_bounds && out_scores && out_rows && out_overflow && workspace,
                 "NULL pointer argument");
    VQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    VQ_CHECK_ARG(((uintptr_t)store_bf16 & 15) == 0 && ((uintptr_t)store_f32 & 15) == 0, "stores must be 16-byte aligned");
    int launches = 0;
    rc = vq_scan_mma_exact(store_bf16, store_f32, n, dim, ld, queries, query_norm, b, k, store_bounds, out_scores, out_rows,
                           out_overflow, out_stats, workspace, workspace_bytes, stream, &launches);
    if (rc) return rc;
    vq_note_launch("scan_mma_bf16<exact>+finish", launches);
    return VQ_OK;
}

size_t vq_search_collect_workspace_bytes(int64_t n, int dim, int ld, int b, int cap) {
    (void)dim;
    if (n <= 0 || b <= 0 || cap <= 0) return 256;
    return vq_scan_mma_collect_workspace(n, ld, b, cap) + 256;
}

int vq_search_collect(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld,
                      const float* queries, int b, int k, int query_norm, const float* thresholds, int cap, const float* store_bounds,
                      float* out_scores, int32_t* out_rows, int32_t* out_overflow,
                      void* workspace, size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    int rc = check_store(n, dim, ld, VQ_BF16);
    if (rc) return rc;
    VQ_CHECK_ARG(b >= 0 && k > 0 && k <= 1024 && cap >= k && cap <= 16384, "need b >= 0, 0 < k <= 1024, k <= cap <= 16384 (b=%d k=%d cap=%d)", b, k, cap);
    VQ_CHECK_ARG(query_norm >= VQ_NORM_NONE && query_norm <= VQ_NORM_PLAIN, "bad query_norm %d", query_norm);
    if (b == 0) { vq_note_launch("none", 0); return VQ_OK; }
    VQ_CHECK_ARG(n > 0, "collect pass needs a non-empty store");
    VQ_CHECK_ARG(store_bf16 && store_f32 && queries && out_scores && out_rows && out_overflow && workspace,
                 "NULL pointer argument");                       // thresholds may be NULL (derived from the store)
    VQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    int launches = 0;
    rc = vq_scan_mma_collect(store_bf16, store_f32, n, dim, ld, queries, query_norm, b, thresholds, cap, store_bounds, k, out_scores, out_rows,
                             out_overflow, workspace, workspace_bytes, stream, &launches);
    if (rc) return rc;
    vq_note_launch("scan_mma_bf16<collect>+rescore", launches);
    return VQ_OK;
}

}  // extern "C"
