// (c) final stage of the single-pass exact search: out of the rows the exact-mode scan gathered
// (scan_mma.cu, kModeExact: every row whose bf16-operand score came within 2*eps of the running k-th
// best), produce the exact fp32 top-k.  One CTA per query:
//   1. candidates -> 64-bit (bf16 score key, row) keys in shared memory;
//   2. bisection on the score key until the best A >= k_sel (and at most ~64) candidates are isolated;
//   3. A is re-scored from the fp32 copy (same fp32 FMA chain as rescore_rows_kernel / the FMA scan);
//      s_lb = the k-th best exact score inside A — reached by k real rows, hence a lower bound of the
//      exact k-th best s_k of the store;
//   4. every other candidate whose bf16 score reaches s_lb - eps is re-scored as well: a row of the exact
//      top-k has exact score >= s_k >= s_lb, i.e. bf16 score >= s_lb - eps, and it is among the
//      candidates (see filter_collect) — nothing can be missed;
//   5. the best k of all re-scored rows by (exact score desc, row asc).
// Exact by construction; out_overflow[q] = 1 only if the scan gathered more than `cap` rows or more than
// kXMax rows need the re-score (mass ties) — the caller re-runs those queries on the fp32 FMA scan.
//
// Reference sites: np.argsort(sim)[::-1][:k] video_search_overhaul.py:56 (selection),
// q / (|q| + 1e-10) :49-50 and np.dot :53 (the exact fp32 score).
#include "vq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kXMaxLarge = 2048;         // rows re-scored per query at most (k_sel > 64: the large-k route)
constexpr int kXMaxSmall = 1024;         // ... for k <= 64 (measured: <= 140 on clustered 1M-row stores): 12 KB less shared memory, 4 instead of 3 CTAs per SM
constexpr int kSelStop = 64;             // the bisection stops once this few keys (>= k_sel) are left
constexpr int kSortCapSmem = 1024;       // bf16-score keys kept in shared memory; a query that gathered more reads the rest from
                                         // its gather buffer (L2): 22 KB per CTA = 8 CTAs per SM, a batch of 1024 in ONE wave
                                         // (with all `cap` = 4096 keys in shared memory: 46 KB, 4 CTAs per SM, two waves)

// exact fp32 scores of rows[0..cnt) -> 64-bit (exact score key, row) keys; one warp per row, 4 rows of a
// warp in flight so that their HBM latencies overlap
__device__ __forceinline__ void rescore_to_keys(const int* rows, int cnt, const float* qn, const float* __restrict__ store_f32,
                                                int ld, unsigned long long* out, int warp, int lane) {
    const float4* y = reinterpret_cast<const float4*>(qn);
    for (int c0 = warp; c0 < cnt; c0 += 4 * kWarps) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int c = c0 + u * kWarps;
            rr[u] = c < cnt ? rows[c] : -1;
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
            const int c = c0 + u * kWarps;
            // a NaN score never beats anything: its key sorts last and is never written out
            if (lane == 0 && c < cnt) out[c] = (t == t) ? (((unsigned long long)vq_score_key(t) << 32) | (unsigned)rr[u]) : ~0ull;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
exact_finish_kernel(const float* __restrict__ cand_s, const int* __restrict__ cand_r, const int* __restrict__ cand_cnt,
                    int cap, int k_sel, const float* __restrict__ qeps, const float* __restrict__ store_f32, int ld, int dim,
                    const float* __restrict__ queries, int query_norm, int k_out,
                    float* __restrict__ out_s, int* __restrict__ out_r, int* __restrict__ out_over, int* __restrict__ out_stats, int sort_cap, int kXMax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);          // [sort_cap] bf16-score keys
    unsigned long long* xkey = keys + sort_cap;                                           // [kXMax] exact-score keys
    int* xrow = reinterpret_cast<int*>(xkey + kXMax);                                     // [kXMax] rows to re-score
    float* qn = reinterpret_cast<float*>(xrow + kXMax);                                   // [ld]
    __shared__ int part[2][kWarps];
    __shared__ unsigned key_min, key_max, lb_key;
    __shared__ int x_cnt;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    vq_pdl_wait();
    vq_pdl_trigger();

    const int total = cand_cnt[q];
    const int n = total < cap ? total : cap;
    const size_t cbase = (size_t)q * cap;
    auto key_at = [&](int i) -> unsigned long long {
        return i < sort_cap ? keys[i] : (((unsigned long long)vq_score_key(cand_s[cbase + i]) << 32) | (unsigned)cand_r[cbase + i]);
    };
    if (tid == 0) { key_min = 0xffffffffu; key_max = 0u; x_cnt = 0; lb_key = 0xffffffffu; }
    __syncthreads();
    // ---- load (the last warp meanwhile normalises the query in fp32, same arithmetic as ingest_rows_kernel)
    if (warp == kWarps - 1) {
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
    {
        const size_t base = (size_t)q * cap;
        unsigned mn = 0xffffffffu, mx = 0u;
        for (int i = tid; i < n; i += kThreads) {
            const unsigned kh = vq_score_key(cand_s[base + i]);
            if (i < sort_cap) keys[i] = ((unsigned long long)kh << 32) | (unsigned)cand_r[base + i];
            mn = kh < mn ? kh : mn;
            mx = kh > mx ? kh : mx;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
        }
        if (lane == 0 && n > 0) { atomicMin(&key_min, mn); atomicMax(&key_max, mx); }
    }
    __syncthreads();

    // ---- A = the candidates with score key <= hi: at least min(k_sel, n) of them, about kSelStop at most
    unsigned hi = key_max;
    if (n > (k_sel > kSelStop ? k_sel : kSelStop)) {
        const int stop = k_sel > kSelStop ? k_sel : kSelStop;
        unsigned lo = key_min;
        int cnt_hi = n, it = 0;
        while (lo < hi && cnt_hi > stop) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int c = 0;
            for (int i = tid; i < n; i += kThreads) c += ((unsigned)(key_at(i) >> 32) <= mid) ? 1 : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) part[it & 1][warp] = c;
            __syncthreads();
            int tot = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tot += part[it & 1][w];
            if (tot >= k_sel) { hi = mid; cnt_hi = tot; } else lo = mid + 1;
            ++it;
        }
    }
    for (int i = tid; i < n; i += kThreads) {
        const unsigned long long key = key_at(i);
        if ((unsigned)(key >> 32) <= hi) {
            const int at = atomicAdd(&x_cnt, 1);
            if (at < kXMax) xrow[at] = (int)(unsigned)key;
        }
    }
    __syncthreads();
    const int x_cnt_a = x_cnt;                  // (the next write to x_cnt is two barriers away)
    const int n_a = x_cnt_a < kXMax ? x_cnt_a : kXMax;
    bool over = total > cap || x_cnt_a > kXMax;
    rescore_to_keys(xrow, n_a, qn, store_f32, ld, xkey, warp, lane);
    __syncthreads();

    // ---- s_lb: the k_out-th best exact score inside A (rank counting; keys are unique, NaN keys sort last)
    for (int i = tid; i < n_a; i += kThreads) {
        const unsigned long long mine = xkey[i];
        int rank = 0;
        for (int j = 0; j < n_a; ++j) rank += xkey[j] < mine ? 1 : 0;
        if (rank == k_out - 1 && mine != ~0ull) lb_key = (unsigned)(mine >> 32);
    }
    __syncthreads();
    // ---- every remaining candidate whose bf16 score reaches s_lb - eps (all of them if A holds fewer than k rows)
    int n_x = n_a;
    if (n > n_a && x_cnt_a <= kXMax) {
        const unsigned t_key = lb_key == 0xffffffffu ? 0xffffffffu : vq_score_key(vq_key_score(lb_key) - qeps[q]);
        for (int i = tid; i < n; i += kThreads) {
            const unsigned long long key = key_at(i);
            const unsigned kh = (unsigned)(key >> 32);
            if (kh > hi && kh <= t_key) {
                const int at = atomicAdd(&x_cnt, 1);
                if (at < kXMax) xrow[at] = (int)(unsigned)key;
            }
        }
        __syncthreads();
        n_x = x_cnt < kXMax ? x_cnt : kXMax;
        over = over || x_cnt > kXMax;
        rescore_to_keys(xrow + n_a, n_x - n_a, qn, store_f32, ld, xkey + n_a, warp, lane);
        __syncthreads();
    }

    // ---- best k_out of the n_x exact keys
    for (int o = tid; o < k_out; o += kThreads) {
        if (o >= n_x) { out_s[(size_t)q * k_out + o] = VQ_NEG_INF; out_r[(size_t)q * k_out + o] = -1; }
    }
    for (int i = tid; i < n_x; i += kThreads) {
        const unsigned long long mine = xkey[i];
        int rank = 0;
        for (int j = 0; j < n_x; ++j) rank += xkey[j] < mine ? 1 : 0;
        if (rank < k_out) {
            const bool ok = mine != ~0ull;
            out_s[(size_t)q * k_out + rank] = ok ? vq_key_score((unsigned)(mine >> 32)) : VQ_NEG_INF;
            out_r[(size_t)q * k_out + rank] = ok ? (int)(unsigned)mine : -1;
        }
    }
    if (tid == 0) {
        out_over[q] = over ? 1 : 0;
        if (out_stats) { out_stats[2 * q] = total; out_stats[2 * q + 1] = n_x; }     // rows gathered by the scan / re-scored in fp32
    }
}

}  // namespace

int vq_exact_finish_launch(const float* cand_s, const int* cand_r, const int* cand_cnt, int cap, int b, int k_sel,
                           const float* qeps, const float* store_f32, int ld, int dim, const float* queries, int query_norm,
                           int k_out, float* out_scores, int* out_rows, int* out_overflow, int* out_stats, cudaStream_t stream) {
    if (b <= 0) return VQ_OK;
    const int kXMax = k_sel <= 64 ? kXMaxSmall : kXMaxLarge;
    if (k_out > k_sel || k_sel > kXMax / 2) {
        vq_set_error("exact_finish: need k_out <= k_sel <= %d (k_sel=%d k_out=%d)", kXMax / 2, k_sel, k_out);
        return VQ_EUNSUPPORTED;
    }
    const int sort_cap = cap < kSortCapSmem ? cap : kSortCapSmem;
    const size_t smem = (size_t)sort_cap * 8 + (size_t)kXMax * 12 + (size_t)ld * 4 + 64;
    if (smem > 200 * 1024) {
        vq_set_error("exact_finish: %d candidate slots per query do not fit shared memory", cap);
        return VQ_EUNSUPPORTED;
    }
    static std::atomic<unsigned long long> attr_done{0};
    if (vq_first_use_on_device(&attr_done)) {
        VQ_CUDA(cudaFuncSetAttribute(exact_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        vq_mark_used(&attr_done);
    }
    const cudaError_t e = vq_launch(4, exact_finish_kernel, dim3(b), dim3(kThreads), smem, stream, cand_s, cand_r, cand_cnt, cap, k_sel,
                                    qeps, store_f32, ld, dim, queries, query_norm, k_out, out_scores, out_rows, out_overflow, out_stats, sort_cap, kXMax);
    if (e != cudaSuccess) {
        vq_set_error("launch of exact_finish_kernel failed: %s", cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    return VQ_OK;
}
