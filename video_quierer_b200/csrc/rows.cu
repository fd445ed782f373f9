// (a) L2-normalisation, store ingest (fp32 -> fp32/bf16 device matrix) and exact re-scoring
// of candidate rows.  One warp per row, 128-bit loads; all HBM-bound element-wise work.
//
// Reference sites: `q / (np.linalg.norm(q) + 1e-10)` video_search_overhaul.py:49-50;
// `v / np.linalg.norm(v)` src/indexes/hnsw.py:157,250,499; `embedding.astype(np.float32)`
// video_search_overhaul.py:33.
#include "vq_common.cuh"

namespace {

__device__ __forceinline__ float inv_scale(float sumsq, int mode) {
    // IEEE sqrt and divide (not rsqrt) so the result matches numpy to <= 1 ulp of the norm.
    if (mode == VQ_NORM_NONE) return 1.0f;
    const float nrm = sqrtf(sumsq);
    return mode == VQ_NORM_EPS ? nrm + 1e-10f : nrm;
}

// dst may alias src (in-place normalise).  dst_ld >= dim; pad columns are zeroed.
template <bool OUT_BF16>
__global__ void __launch_bounds__(256)
ingest_rows_kernel(const float* src, long long rows, int dim, int src_ld,
                   void* dst_v, int dst_ld, int mode) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* s = src + row * src_ld;
    float sum = 0.f;
    if (mode != VQ_NORM_NONE) {
        for (int c = lane; c < dim; c += 32) { const float v = s[c]; sum = fmaf(v, v, sum); }
        sum = vq_warp_sum(sum);
    }
    const float d = inv_scale(sum, mode);
    if (OUT_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst_v) + row * dst_ld;
        for (int c = lane; c < dst_ld; c += 32)
            o[c] = __float2bfloat16_rn(c < dim ? (mode == VQ_NORM_NONE ? s[c] : s[c] / d) : 0.f);
    } else {
        float* o = reinterpret_cast<float*>(dst_v) + row * dst_ld;
        for (int c = lane; c < dst_ld; c += 32)
            o[c] = c < dim ? (mode == VQ_NORM_NONE ? s[c] : s[c] / d) : 0.f;
    }
}

// One warp per (query, candidate): exact fp32 dot from the fp32 store.
__global__ void __launch_bounds__(256)
rescore_rows_kernel(const float* __restrict__ store, int ld, const float* __restrict__ queries, int qld,
                    const int* __restrict__ cand, int b, int k_cand, float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (long long)b * k_cand) return;
    const int q = (int)(w / k_cand);
    const int r = cand[w];
    float acc = VQ_NEG_INF;
    if (r >= 0) {
        const float4* x = reinterpret_cast<const float4*>(store + (size_t)r * ld);
        const float4* y = reinterpret_cast<const float4*>(queries + (size_t)q * qld);
        acc = 0.f;
        for (int c = lane; c < ld / 4; c += 32) {
            const float4 a = x[c], bq = y[c];
            acc = fmaf(a.x, bq.x, acc); acc = fmaf(a.y, bq.y, acc);
            acc = fmaf(a.z, bq.z, acc); acc = fmaf(a.w, bq.w, acc);
        }
        acc = vq_warp_sum(acc);
    }
    if (lane == 0) out_scores[w] = acc;
}

// Per-store operand-rounding bounds of the exact search (vq_search_exact): bounds[0] = max over rows of the
// bf16 row's norm |x^|, bounds[1] = max |x^ - x| (x = the fp32 row).  One warp per row reads both copies;
// the maxima are merged with integer atomics on the (non-negative) float bit patterns.  Rows with a
// non-finite norm are skipped: their scores are NaN and never beat anything.
__global__ void __launch_bounds__(256)
store_bounds_kernel(const float* __restrict__ f32, const __nv_bfloat16* __restrict__ bf16, long long rows, int ld,
                    float* __restrict__ bounds) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* x = reinterpret_cast<const float4*>(f32 + row * ld);
    const uint2* h = reinterpret_cast<const uint2*>(bf16 + row * ld);
    float n2 = 0.f, e2 = 0.f;
    for (int c = lane; c < ld / 4; c += 32) {
        const float4 a = x[c];
        const uint2 w = h[c];
        const float h0 = vq_bf16lo(w.x), h1 = vq_bf16hi(w.x), h2 = vq_bf16lo(w.y), h3 = vq_bf16hi(w.y);
        n2 = fmaf(h0, h0, n2); n2 = fmaf(h1, h1, n2); n2 = fmaf(h2, h2, n2); n2 = fmaf(h3, h3, n2);
        const float d0 = h0 - a.x, d1 = h1 - a.y, d2 = h2 - a.z, d3 = h3 - a.w;
        e2 = fmaf(d0, d0, e2); e2 = fmaf(d1, d1, e2); e2 = fmaf(d2, d2, e2); e2 = fmaf(d3, d3, e2);
    }
    n2 = vq_warp_sum(n2);
    e2 = vq_warp_sum(e2);
    if (lane == 0) {
        const float nrm = sqrtf(n2), err = sqrtf(e2);
        if (nrm < INFINITY && err < INFINITY) {          // false for NaN as well
            atomicMax(reinterpret_cast<int*>(bounds), __float_as_int(nrm));
            atomicMax(reinterpret_cast<int*>(bounds) + 1, __float_as_int(err));
        }
    }
}

__global__ void fill_empty_kernel(float* scores, int* rows, long long count) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { scores[i] = VQ_NEG_INF; rows[i] = -1; }
}

}  // namespace

int vq_store_bounds_launch(const float* f32, const void* bf16, long long rows, int ld, float* bounds, cudaStream_t stream) {
    if (rows == 0) return VQ_OK;
    store_bounds_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(f32, reinterpret_cast<const __nv_bfloat16*>(bf16), rows, ld, bounds);
    VQ_LAUNCH_CHECK("store_bounds_kernel");
    return VQ_OK;
}

int vq_fill_empty_launch(float* scores, int* rows, long long count, cudaStream_t stream) {
    if (count == 0) return VQ_OK;
    fill_empty_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(scores, rows, count);
    VQ_LAUNCH_CHECK("fill_empty_kernel");
    return VQ_OK;
}

int vq_ingest_launch(const float* src, long long rows, int dim, int src_ld, void* dst, int dst_dtype,
                     int dst_ld, int mode, cudaStream_t stream) {
    if (rows == 0) return VQ_OK;
    const int wpb = 8;
    const long long blocks = (rows + wpb - 1) / wpb;
    if (dst_dtype == VQ_BF16)
        ingest_rows_kernel<true><<<(unsigned)blocks, wpb * 32, 0, stream>>>(src, rows, dim, src_ld, dst, dst_ld, mode);
    else
        ingest_rows_kernel<false><<<(unsigned)blocks, wpb * 32, 0, stream>>>(src, rows, dim, src_ld, dst, dst_ld, mode);
    VQ_LAUNCH_CHECK("ingest_rows_kernel");
    return VQ_OK;
}

int vq_rescore_launch(const float* store, int ld, const float* queries, int qld, const int* cand, int b,
                      int k_cand, float* out_scores, cudaStream_t stream) {
    const long long warps = (long long)b * k_cand;
    if (warps == 0) return VQ_OK;
    rescore_rows_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(store, ld, queries, qld, cand, b, k_cand,
                                                                        out_scores);
    VQ_LAUNCH_CHECK("rescore_rows_kernel");
    return VQ_OK;
}
