// Per-query selection of the best k_out out of g candidate lists of k_in entries.
// Used (1) to reduce the per-CTA candidate lists of the scan kernels, and (2) as the
// shard/merge layer after the all-gather of per-GPU top-k.  The reference
// has no counterpart (single process); the ordering rule is the engine's (score desc, row asc).
//
// One CTA per query.  The g*k_in candidates are split over the 8 warps; every warp filters its
// slice against its running k-th best (loads for 4 chunks of 32 candidates are issued before any is
// consumed, so the global-memory latency is paid once per 128 candidates) into a sorted list
// of (score, global row) in shared memory; warp 0 then merges the 8 sorted lists by repeated
// head selection, entirely out of shared memory.
#include "vq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kUnroll = 4;

__device__ __forceinline__ bool better64(float s, long long r, float s2, long long r2) {
    return (s > s2) || (s == s2 && r < r2);
}

// scores/rows: candidate c of query q lives at  sh * gs + q * k_in + j   (sh = c / k_in, j = c % k_in)
template <typename OutRow>
__global__ void __launch_bounds__(kThreads)
topk_merge_kernel(const float* __restrict__ scores, const int* __restrict__ rows, int g, long long gs, int k_in,
                  const long long* __restrict__ offsets, int k_out,
                  float* __restrict__ out_scores, OutRow* __restrict__ out_rows, int negate_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    long long* lr = reinterpret_cast<long long*>(smem_raw);                       // [kWarps][k_out] global rows
    float* ls = reinterpret_cast<float*>(lr + (size_t)kWarps * k_out);            // [kWarps][k_out] scores
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = g * k_in;

    float* mls = ls + (size_t)warp * k_out;
    long long* mlr = lr + (size_t)warp * k_out;
    for (int i = lane; i < k_out; i += 32) { mls[i] = VQ_NEG_INF; mlr[i] = LLONG_MAX; }
    __syncwarp();

    float ts = VQ_NEG_INF;          // running k-th best of this warp's list
    long long tr = LLONG_MAX;
    for (int base = warp * 32; base < total; base += kThreads * kUnroll) {
        float s[kUnroll];
        long long r[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int c = base + u * kThreads + lane;
            s[u] = VQ_NEG_INF;
            r[u] = -1;
            if (c < total) {
                const int sh = c / k_in, j = c - sh * k_in;
                const size_t at = (size_t)sh * gs + (size_t)q * k_in + j;
                const int lrow = rows[at];
                if (lrow >= 0) { s[u] = scores[at]; r[u] = (long long)lrow + (offsets ? offsets[sh] : 0); }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            unsigned m = __ballot_sync(0xffffffffu, r[u] >= 0 && better64(s[u], r[u], ts, tr));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cs = __shfl_sync(0xffffffffu, s[u], src);
                const long long cr = __shfl_sync(0xffffffffu, r[u], src);
                if (!better64(cs, cr, ts, tr)) continue;
                int pos = 0;
                for (int b0 = 0; b0 < k_out; b0 += 32) {
                    const int i = b0 + lane;
                    const bool b = (i < k_out) && better64(mls[i], mlr[i], cs, cr);
                    pos += __popc(__ballot_sync(0xffffffffu, b));
                }
                if (pos >= k_out) continue;
                for (int b0 = ((k_out - 1) / 32) * 32; b0 >= 0; b0 -= 32) {
                    const int i = b0 + lane;
                    float es = 0.f; long long er = 0;
                    const bool mv = (i < k_out) && (i > pos);
                    if (mv) { es = mls[i - 1]; er = mlr[i - 1]; }
                    __syncwarp();
                    if (mv) { mls[i] = es; mlr[i] = er; }
                    __syncwarp();
                    if (b0 <= pos) break;
                }
                if (lane == 0) { mls[pos] = cs; mlr[pos] = cr; }
                __syncwarp();
                ts = mls[k_out - 1];
                tr = mlr[k_out - 1];
            }
        }
    }
    __syncthreads();

    // warp 0 merges the kWarps sorted lists by repeated head selection (shared memory only)
    if (warp == 0) {
        int head = 0;                    // lanes 0..kWarps-1 each track one list
        for (int o = 0; o < k_out; ++o) {
            float s = VQ_NEG_INF;
            long long r = LLONG_MAX;
            if (lane < kWarps && head < k_out) {
                s = ls[(size_t)lane * k_out + head];
                r = lr[(size_t)lane * k_out + head];
            }
            float bs = s; long long br = r; int bl = lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, bs, off);
                const long long orr = __shfl_xor_sync(0xffffffffu, br, off);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
                const bool take = (os > bs) || (os == bs && (orr < br || (orr == br && ol < bl)));
                if (take) { bs = os; br = orr; bl = ol; }
            }
            if (lane == bl) {
                const bool valid = (r != LLONG_MAX);
                out_scores[(size_t)q * k_out + o] = valid ? (negate_out ? 1.0f - s : s) : (negate_out ? INFINITY : VQ_NEG_INF);
                out_rows[(size_t)q * k_out + o] = valid ? (OutRow)r : (OutRow)-1;
                ++head;
            }
        }
    }
}

}  // namespace

// g_stride: elements between consecutive candidate blocks (>= b_out*k_in); b_out queries are merged.
int vq_topk_merge_launch(const float* scores, const int* rows, int g, long long g_stride, int b_out, int k_in,
                         const long long* offsets, int k_out, float* out_scores, void* out_rows,
                         int rows64, int negate_out, cudaStream_t stream) {
    if (b_out <= 0) return VQ_OK;
    const size_t smem = (size_t)kWarps * k_out * 12;
    if (smem > 100 * 1024) {
        vq_set_error("topk_merge: k_out=%d too large", k_out);
        return VQ_EUNSUPPORTED;
    }
    static std::atomic<unsigned long long> attr_done{0};
    if (vq_first_use_on_device(&attr_done)) {
        VQ_CUDA(cudaFuncSetAttribute(topk_merge_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        VQ_CUDA(cudaFuncSetAttribute(topk_merge_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        vq_mark_used(&attr_done);
    }
    if (rows64)
        topk_merge_kernel<long long><<<b_out, kThreads, smem, stream>>>(scores, rows, g, g_stride, k_in, offsets, k_out,
                                                                        out_scores, (long long*)out_rows, negate_out);
    else
        topk_merge_kernel<int><<<b_out, kThreads, smem, stream>>>(scores, rows, g, g_stride, k_in, offsets, k_out,
                                                                  out_scores, (int*)out_rows, negate_out);
    VQ_LAUNCH_CHECK("topk_merge_kernel");
    return VQ_OK;
}
