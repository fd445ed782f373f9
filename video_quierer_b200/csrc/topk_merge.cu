// Per-query selection of the best k_out out of g candidate lists of k_in entries.
// Used (1) to reduce the per-CTA candidate lists of the scan kernels, and (2) as the
// shard/merge layer after the all-gather of per-GPU top-k.  The reference
// has no counterpart (single process); the ordering rule is the engine's (score desc, row asc).
#include "vq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// scores/rows: [g, b, k_in]; one CTA per query.
template <typename OutRow>
__global__ void __launch_bounds__(kThreads)
topk_merge_kernel(const float* __restrict__ scores, const int* __restrict__ rows, int g, long long gs /*shard stride*/, int k_in,
                  const long long* __restrict__ offsets, int k_out,
                  float* __restrict__ out_scores, OutRow* __restrict__ out_rows, int negate_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* ls = reinterpret_cast<float*>(smem_raw);                 // [kWarps][k_out]
    int* lr = reinterpret_cast<int*>(ls + (size_t)kWarps * k_out);   // [kWarps][k_out]  (candidate index, not row)
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = g * k_in;

    float* mls = ls + (size_t)warp * k_out;
    int* mlr = lr + (size_t)warp * k_out;
    for (int i = lane; i < k_out; i += 32) { mls[i] = VQ_NEG_INF; mlr[i] = VQ_EMPTY_ROW; }
    __syncwarp();

    // Order candidates by (score desc, global row asc).  With contiguous shards the global row
    // order equals (shard, local row) order, so the key below is (score, shard-major index)
    // only when offsets are absent; with offsets we compare true global rows.
    auto cand_row = [&](int c) -> long long {
        const int sh = c / k_in, j = c - sh * k_in;
        const int r = rows[(size_t)sh * gs + (size_t)q * k_in + j];
        if (r < 0) return -1;
        return (long long)r + (offsets ? offsets[sh] : 0);
    };

    // phase 1: each warp filters its slice.  Lists hold candidate indices; ties are broken on
    // the global row, looked up on demand (rare).
    float ts = VQ_NEG_INF;
    long long tr = LLONG_MAX;
    for (int base = warp * 32; base < total; base += kThreads) {
        const int c = base + lane;
        float s = VQ_NEG_INF;
        long long r = -1;
        if (c < total) {
            const int sh = c / k_in, j = c - sh * k_in;
            const size_t at = (size_t)sh * gs + (size_t)q * k_in + j;
            const int lr_ = rows[at];
            if (lr_ >= 0) { s = scores[at]; r = (long long)lr_ + (offsets ? offsets[sh] : 0); }
        }
        bool pass = (r >= 0) && ((s > ts) || (s == ts && r < tr));
        unsigned m = __ballot_sync(0xffffffffu, pass);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float cs = __shfl_sync(0xffffffffu, s, src);
            const long long cr = __shfl_sync(0xffffffffu, r, src);
            if (!((cs > ts) || (cs == ts && cr < tr))) continue;
            // insertion sort keyed on (score desc, global row asc); list stores candidate index
            int pos = 0;
            for (int b0 = 0; b0 < k_out; b0 += 32) {
                const int i = b0 + lane;
                bool better = false;
                if (i < k_out && mlr[i] != VQ_EMPTY_ROW) {
                    const float es = mls[i];
                    better = es > cs || (es == cs && cand_row(mlr[i]) < cr);
                }
                pos += __popc(__ballot_sync(0xffffffffu, better));
            }
            if (pos < k_out) {
                for (int b0 = ((k_out - 1) / 32) * 32; b0 >= 0; b0 -= 32) {
                    const int i = b0 + lane;
                    float es = 0.f; int er = 0;
                    const bool mv = (i < k_out) && (i > pos);
                    if (mv) { es = mls[i - 1]; er = mlr[i - 1]; }
                    __syncwarp();
                    if (mv) { mls[i] = es; mlr[i] = er; }
                    __syncwarp();
                    if (b0 <= pos) break;
                }
                if (lane == 0) { mls[pos] = cs; mlr[pos] = base + src; }
                __syncwarp();
                if (mlr[k_out - 1] != VQ_EMPTY_ROW) { ts = mls[k_out - 1]; tr = cand_row(mlr[k_out - 1]); }
            }
        }
    }
    __syncthreads();

    // phase 2: warp 0 merges the kWarps sorted lists by repeated head selection.
    if (warp == 0) {
        int head = 0;                    // lanes 0..kWarps-1 each track one list
        for (int o = 0; o < k_out; ++o) {
            float s = VQ_NEG_INF;
            long long r = LLONG_MAX;
            int c = VQ_EMPTY_ROW;
            if (lane < kWarps && head < k_out) {
                c = lr[(size_t)lane * k_out + head];
                if (c != VQ_EMPTY_ROW) { s = ls[(size_t)lane * k_out + head]; r = cand_row(c); }
            }
            // arg-best across lanes
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
                const bool valid = (c != VQ_EMPTY_ROW);
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
    const long long b = g_stride;
    if (b_out <= 0) return VQ_OK;
    const size_t smem = (size_t)kWarps * k_out * 8;
    if (smem > 96 * 1024) {
        vq_set_error("topk_merge: k_out=%d too large", k_out);
        return VQ_EUNSUPPORTED;
    }
    static bool attr_done = false;
    if (!attr_done) {
        VQ_CUDA(cudaFuncSetAttribute(topk_merge_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        VQ_CUDA(cudaFuncSetAttribute(topk_merge_kernel<int>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr_done = true;
    }
    if (rows64)
        topk_merge_kernel<long long><<<b_out, kThreads, smem, stream>>>(scores, rows, g, b, k_in, offsets, k_out,
                                                                    out_scores, (long long*)out_rows, negate_out);
    else
        topk_merge_kernel<int><<<b_out, kThreads, smem, stream>>>(scores, rows, g, b, k_in, offsets, k_out,
                                                              out_scores, (int*)out_rows, negate_out);
    VQ_LAUNCH_CHECK("topk_merge_kernel");
    return VQ_OK;
}
