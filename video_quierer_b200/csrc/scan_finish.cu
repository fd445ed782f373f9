// (c) final per-query selection of the tensor-core scan, optionally fused with the exact re-score.
//
// The scan kernel (scan_mma.cu) appends, per query, only the list entries that can still belong to
// the global top-k (they reach the shared k-th-best bound): typically k..3k candidates instead of
// groups*k.  One CTA per query
//   1. sorts the candidates by (score desc, row asc) with a bitonic network in shared memory,
//   2. plain mode: writes the best k_out;
//      two-stage mode (store_f32 != NULL): normalises the query in fp32 exactly like vq_ingest_rows,
//      re-scores the best k_sel candidates from the fp32 copy with the same fp32 FMA chain as
//      rescore_rows_kernel, selects the best k_out by exact score and CERTIFIES the result: every row
//      outside the candidate set has bf16-operand score <= the k_sel-th candidate score, hence exact
//      score <= that + eps; if the exact k_out-th score is not below that bound nothing was missed.
// This replaces topk_merge (groups*k candidates per query, one CTA each: 118 us at batch 32) +
// ingest + rescore + merge + the torch ops of the certification by one launch.
//
// Reference sites: np.argsort(sim)[::-1][:k] video_search_overhaul.py:56 (selection),
// q / (|q| + 1e-10) :49-50 and np.dot :53 (the exact fp32 score).
#include <stdlib.h>

#include "vq_common.cuh"

namespace {

__device__ int g_finish_dbg = 0;
#define FDBG(i) do { if (dbg_on && tid == 0 && q == 0) tmark[i] = clock64(); } while (0)

constexpr int kThreads = 256;
constexpr int kMaxSel = 64;
constexpr int kSelMax = 1024;            // keys sorted after the bisection (more only with mass ties)
constexpr int kSelStop = 64;             // the bisection stops once this few keys (>= k_sel) are left
constexpr int kRankMax = 256;            // pools up to this size are sorted by rank counting

// monotone map: larger score -> smaller key (ascending sort = best first); -0 is folded into +0
__device__ __forceinline__ uint32_t score_key(float s) {
    if (s == 0.f) s = 0.f;
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~u;
}
__device__ __forceinline__ float key_score(uint32_t k) {
    uint32_t u = ~k;
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

__global__ void __launch_bounds__(kThreads)
scan_finish_kernel(int mode,      // 0 plain top-k of the scan scores, 1 two-stage + certificate, 2 re-score ALL candidates
                   const float* __restrict__ cand_s, const int* __restrict__ cand_r, const int* __restrict__ cand_cnt,
                   int k_in, long long g_stride,   // g_stride > 0: candidates are cap/k_in lists [list][query][k_in] (FMA scan)
                   int cap, int k_sel, const float* __restrict__ store_f32, int ld, int dim,
                   const float* __restrict__ queries, int query_norm, float eps, int k_out,
                   float* __restrict__ out_s, int* __restrict__ out_r, int* __restrict__ out_bad, int sort_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);        // [sort_cap]
    unsigned long long* sel = keys + sort_cap;                                           // [kSelMax]
    unsigned long long* ranked = sel + kSelMax;                                          // [kRankMax]
    float* qn = reinterpret_cast<float*>(ranked + kRankMax);                             // [ld]   (two-stage)
    float* ex_s = qn + ld;                                                                // [kMaxSel]
    int* ex_r = reinterpret_cast<int*>(ex_s + kMaxSel);                                   // [kMaxSel]
    __shared__ int part[2][kThreads / 32];
    __shared__ unsigned key_min, key_max;
    __shared__ int sel_cnt, n_valid_s;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool dbg_on = g_finish_dbg != 0;
    long long tmark[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    vq_pdl_wait();
    vq_pdl_trigger();
    FDBG(0);

    int n = cap;                               // list layout: every slot is read, empty ones (row < 0) sort last
    if (g_stride == 0) { n = cand_cnt[q]; n = n < cap ? n : cap; }
    if (tid == 0) { key_min = 0xffffffffu; key_max = 0u; sel_cnt = 0; n_valid_s = 0; if (out_bad && mode != 2) out_bad[q] = 0; }
    __syncthreads();
    // ---- load: candidates -> 64-bit keys in shared memory (+ their score-key range); meanwhile the
    // last warp normalises the query in fp32 (same arithmetic as ingest_rows_kernel)
    if (store_f32 != nullptr && warp == kThreads / 32 - 1) {
        const float* s = queries + (size_t)q * dim;
        float sum = 0.f;
        if (query_norm != VQ_NORM_NONE) {
            for (int c = lane; c < dim; c += 32) { const float v = s[c]; sum = fmaf(v, v, sum); }
            sum = vq_warp_sum(sum);
        }
        float d = 1.f;
        if (query_norm != VQ_NORM_NONE) { d = sqrtf(sum); if (query_norm == VQ_NORM_EPS) d += 1e-10f; }
        for (int c = lane; c < ld; c += 32) qn[c] = c < dim ? (query_norm == VQ_NORM_NONE ? s[c] : s[c] / d) : 0.f;
    }
    const size_t base = (size_t)q * cap;
    if (mode != 2) {
        unsigned mn = 0xffffffffu, mx = 0u;
        int valid = 0;
        for (int i = tid; i < n; i += kThreads) {
            const size_t at = g_stride > 0 ? (size_t)(i / k_in) * (size_t)g_stride + (size_t)q * k_in + (size_t)(i % k_in) : base + i;
            const int row = cand_r[at];
            unsigned long long key = ~0ull;
            if (row >= 0) {
                const unsigned kh = score_key(cand_s[at]);
                key = ((unsigned long long)kh << 32) | (unsigned)row;
                mn = kh < mn ? kh : mn;
                mx = kh > mx ? kh : mx;
                ++valid;
            }
            keys[i] = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
            valid += __shfl_xor_sync(0xffffffffu, valid, o);
        }
        if (lane == 0 && valid > 0) { atomicMin(&key_min, mn); atomicMax(&key_max, mx); atomicAdd(&n_valid_s, valid); }
    } else {
        // collect mode: the scan scores only selected the candidates; every one of them is re-scored
        // exactly (4 rows per warp in flight) and the keys are built from the EXACT scores
        __syncthreads();                                     // qn is complete
        const float4* y = reinterpret_cast<const float4*>(qn);
        unsigned mn = 0xffffffffu, mx = 0u;
        for (int c0 = warp; c0 < n; c0 += 4 * (kThreads / 32)) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int rr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * (kThreads / 32);
                rr[u] = c < n ? cand_r[base + c] : -1;
            }
            for (int j = lane; j < ld / 4; j += 32) {
                float4 a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    a[u] = rr[u] >= 0 ? reinterpret_cast<const float4*>(store_f32 + (size_t)rr[u] * ld)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 bq = y[j];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u] = fmaf(a[u].x, bq.x, acc[u]); acc[u] = fmaf(a[u].y, bq.y, acc[u]);
                    acc[u] = fmaf(a[u].z, bq.z, acc[u]); acc[u] = fmaf(a[u].w, bq.w, acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float t = vq_warp_sum(acc[u]);
                const int c = c0 + u * (kThreads / 32);
                if (c < n) {
                    const unsigned kh = score_key(t);
                    if (lane == 0) keys[c] = ((unsigned long long)kh << 32) | (unsigned)rr[u];
                    mn = kh < mn ? kh : mn;
                    mx = kh > mx ? kh : mx;
                }
            }
        }
        if (lane == 0 && n > 0) { atomicMin(&key_min, mn); atomicMax(&key_max, mx); }
        if (tid == 0) n_valid_s = n;
        if (tid == 0 && out_bad) out_bad[q] = cand_cnt[q] > cap ? 1 : 0;      // overflow: result incomplete
    }
    __syncthreads();
    FDBG(1);

    // ---- selection: the scan's shared bound is weak (a CTA only knows its own k-th best), so n is
    // typically 1-3 thousand.  Bisect the score-key range until at most kSelStop keys (but at least
    // k_sel) lie below the pivot: one block-wide count and one barrier per step.
    int m = n;                                   // number of keys that take part in the final sort
    unsigned long long* pool = keys;
    const int sel_stop = k_sel > kSelStop ? k_sel : kSelStop;
    if (n > sel_stop) {
        unsigned lo = key_min, hi = key_max;     // invariant: count(key_hi <= hi) = cnt_hi >= min(k_sel, valid)
        int cnt_hi = n_valid_s, it = 0;
        while (lo < hi && cnt_hi > sel_stop) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (int i = tid; i < n; i += kThreads) c += ((unsigned)(keys[i] >> 32) <= mid) ? 1 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) part[it & 1][warp] = c;
            __syncthreads();
            int total = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) total += part[it & 1][w];
            if (total >= k_sel) { hi = mid; cnt_hi = total; } else lo = mid + 1;
            ++it;
        }
        for (int i = tid; i < n; i += kThreads) {
            const unsigned long long key = keys[i];
            if ((unsigned)(key >> 32) <= hi) {
                const int at = atomicAdd(&sel_cnt, 1);
                if (at < kSelMax) sel[at] = key;
            }
        }
        __syncthreads();
        if (sel_cnt <= kSelMax) { m = sel_cnt; pool = sel; }      // else (mass ties): sort everything in place
    }
    FDBG(2);

    // ---- sort the pool (best first)
    unsigned long long* sorted = pool;
    if (m <= kRankMax) {            // small: every thread ranks one key against all (keys are unique)
        if (tid < m) {
            const unsigned long long mine = pool[tid];
            int rank = 0;
            for (int j = 0; j < m; ++j) rank += pool[j] < mine ? 1 : 0;
            ranked[rank] = mine;
        }
        sorted = ranked;
        __syncthreads();
    } else {
        int P = 2;
        while (P < m) P <<= 1;
        for (int i = m + tid; i < P; i += kThreads) pool[i] = ~0ull;
        __syncthreads();
        for (int size = 2; size <= P; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < (P >> 1); t += kThreads) {
                    const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
                    const int hi = lo | stride;
                    const bool asc = (lo & size) == 0;
                    const unsigned long long a = pool[lo], b = pool[hi];
                    if ((a > b) == asc) { pool[lo] = b; pool[hi] = a; }
                }
                __syncthreads();
            }
        }
    }
    FDBG(3);
    keys = sorted;                               // entries [0, min(n, k_sel)) are what follows needs

    if (mode != 1) {
        const int n_valid = n_valid_s;
        for (int i = tid; i < k_out; i += kThreads) {
            const bool ok = i < n_valid;
            out_s[(size_t)q * k_out + i] = ok ? key_score((unsigned)(keys[i] >> 32)) : VQ_NEG_INF;
            out_r[(size_t)q * k_out + i] = ok ? (int)(unsigned)keys[i] : -1;
        }
        return;
    }

    // ---- two-stage: exact fp32 re-score of the best k_sel candidates
    const int n_sel = n_valid_s < k_sel ? n_valid_s : k_sel;
    FDBG(4);
    // one warp per candidate row (rescore_rows_kernel's fp32 FMA chain); 4 rows of a warp are loaded
    // together so that their HBM latencies overlap
    for (int c0 = warp; c0 < n_sel; c0 += 4 * (kThreads / 32)) {
        const float4* y = reinterpret_cast<const float4*>(qn);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * (kThreads / 32);
            rr[u] = c < n_sel ? (int)(unsigned)keys[c] : -1;
        }
        for (int j = lane; j < ld / 4; j += 32) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                a[u] = rr[u] >= 0 ? reinterpret_cast<const float4*>(store_f32 + (size_t)rr[u] * ld)[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 bq = y[j];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[u] = fmaf(a[u].x, bq.x, acc[u]); acc[u] = fmaf(a[u].y, bq.y, acc[u]);
                acc[u] = fmaf(a[u].z, bq.z, acc[u]); acc[u] = fmaf(a[u].w, bq.w, acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float t = vq_warp_sum(acc[u]);
            const int c = c0 + u * (kThreads / 32);
            if (lane == 0 && c < n_sel) { ex_s[c] = t; ex_r[c] = rr[u]; }
        }
    }
    __syncthreads();
    FDBG(5);
    // ---- best k_out by exact score (score desc, row asc): every thread ranks one candidate
    for (int o = n_sel + tid; o < k_out; o += kThreads) {         // fewer candidates than k_out: empty slots
        out_s[(size_t)q * k_out + o] = VQ_NEG_INF;
        out_r[(size_t)q * k_out + o] = -1;
    }
    if (tid < n_sel) {
        const float ms = ex_s[tid];
        const int mr = ex_r[tid];
        int rank = 0;
        for (int j = 0; j < n_sel; ++j) rank += (j != tid && !vq_better(ms, mr, ex_s[j], ex_r[j])) ? 1 : 0;
        if (rank < k_out) {
            out_s[(size_t)q * k_out + rank] = ms;
            out_r[(size_t)q * k_out + rank] = mr;
        }
        if (rank == k_out - 1 && out_bad) {
            // rows were dropped only if at least k_sel candidates exist; the dropped ones score <= the k_sel-th
            out_bad[q] = (n_valid_s >= k_sel && ms < key_score((unsigned)(keys[k_sel - 1] >> 32)) + eps) ? 1 : 0;
        }
    }
    FDBG(6);
    if (dbg_on && tid == 0 && q == 0)
        printf("[finish dbg] n=%d m=%d load=%lld select=%lld sort=%lld rescore=%lld final=%lld cycles\n", n, m, tmark[1] - tmark[0],
               tmark[2] - tmark[1], tmark[3] - tmark[2], tmark[5] - tmark[4], tmark[6] - tmark[5]);
}

}  // namespace

static int finish_launch(int mode, const float* cand_s, const int* cand_r, const int* cand_cnt, int k_in, long long g_stride,
                         int cap, int b, int k_sel,
                          const float* store_f32, int ld, int dim, const float* queries, int query_norm, float eps,
                          int k_out, float* out_scores, int* out_rows, int* out_bad, cudaStream_t stream) {
    if (b <= 0) return VQ_OK;
    // modes 0 / 1 rank their k_sel exact scores with one thread each out of a 64-entry buffer; mode 2 (collect)
    // selects straight from the sorted keys and goes up to the sort pool
    if (k_sel > (mode != 1 ? kSelMax : kMaxSel) || k_out > k_sel) {
        vq_set_error("scan_finish: need k_out <= k_sel <= %d (k_sel=%d k_out=%d)", mode != 1 ? kSelMax : kMaxSel, k_sel, k_out);
        return VQ_EUNSUPPORTED;
    }
    int sort_cap = 2;
    while (sort_cap < cap) sort_cap <<= 1;
    const size_t smem = (size_t)sort_cap * 8 + kSelMax * 8 + kRankMax * 8 + (size_t)ld * 4 + kMaxSel * 8 + 64;
    if (smem > 220 * 1024) {
        vq_set_error("scan_finish: %d candidate slots per query do not fit shared memory", cap);
        return VQ_EUNSUPPORTED;
    }
    static std::atomic<unsigned long long> attr_done{0};
    if (vq_first_use_on_device(&attr_done)) {
        const int on = getenv("VQ_FINISH_DEBUG") ? 1 : 0;
        cudaMemcpyToSymbol(g_finish_dbg, &on, sizeof(int));
        VQ_CUDA(cudaFuncSetAttribute(scan_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        vq_mark_used(&attr_done);
    }
    const cudaError_t e = vq_launch(4, scan_finish_kernel, dim3(b), dim3(kThreads), smem, stream, mode, cand_s, cand_r, cand_cnt, k_in, g_stride, cap, k_sel,
                                    store_f32, ld, dim, queries, query_norm, eps, k_out, out_scores, out_rows, out_bad, sort_cap);
    if (e != cudaSuccess) {
        vq_set_error("launch of scan_finish_kernel failed: %s", cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    return VQ_OK;
}

int vq_scan_finish_launch(int mode, const float* cand_s, const int* cand_r, const int* cand_cnt, int cap, int b, int k_sel,
                          const float* store_f32, int ld, int dim, const float* queries, int query_norm, float eps,
                          int k_out, float* out_scores, int* out_rows, int* out_bad, cudaStream_t stream) {
    return finish_launch(mode, cand_s, cand_r, cand_cnt, 1, 0, cap, b, k_sel, store_f32, ld, dim, queries, query_norm, eps, k_out,
                         out_scores, out_rows, out_bad, stream);
}

// Best k_out of g sorted-or-not candidate lists per query laid out [list][query][k_in] (the per-CTA lists of
// the FMA scan): the same bisection select + rank sort, one CTA per query.  g * k_in <= 16384.
bool vq_scan_finish_lists_supported(int g, int k_in, int k_out) { return (long long)g * k_in <= 16384 && k_out <= kSelMax; }
int vq_scan_finish_lists_launch(const float* scores, const int* rows, int g, long long g_stride, int b, int k_in, int k_out,
                                float* out_scores, int* out_rows, cudaStream_t stream) {
    return finish_launch(0, scores, rows, nullptr, k_in, g_stride, g * k_in, b, k_out, nullptr, 32, 32, nullptr, VQ_NORM_NONE, 0.f,
                         k_out, out_scores, out_rows, nullptr, stream);
}
