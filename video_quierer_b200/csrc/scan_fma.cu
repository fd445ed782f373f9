// (b)+(c) small-batch exact scan: HBM-streaming fp32-FMA "GEMV" path with fused top-k.
//
// Replaces the reference's  np.dot(embeddings_array, query_norm) + np.argsort(...)[::-1][:k]
// (video_search_overhaul.py:53,56) for query batches too small to be a dense contraction.
//
// Design (B200-first):
//   * one streaming pass over the store per batch of <= BT queries; every 16-byte load is a
//     coalesced `ld.global.nc.L1::no_allocate.v4` (8 lanes cover one 128-byte line of a row,
//     the 4 lane-groups of a warp take 4 different rows, R row-sets in flight per thread,
//     software-pipelined one chunk ahead => 2*R*16 B outstanding per thread);
//   * the query tile lives in shared memory as fp32 and is read with 128-bit broadcast loads
//     (8 distinct addresses per warp request => one wavefront), each feeding 4*R FMAs;
//   * arithmetic is fp32 multiply + fp32 accumulate like the reference's sgemv (bf16 stores are
//     widened exactly), so scores agree with the reference to summation-order noise (~1e-7);
//   * scores never go to HBM: a CTA stages its 32*R x BT score tile in shared memory, and each
//     warp keeps the running top-k of "its" queries as a sorted list in shared memory, filtered
//     by the current k-th best (insertions are rare after the first tiles);
//   * the grid is persistent: 2 CTAs per SM, each owning one contiguous row range; each CTA
//     writes one k-entry candidate list per query; `topk_merge` reduces them.
#include "vq_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <int BT, int R, bool BF16>
__global__ void __launch_bounds__(kThreads, 2)
scan_fma_kernel(const void* __restrict__ store_v, int n, int ld,
                const float* __restrict__ queries,   // [BT, ld] normalised, zero padded
                int k,
                float* __restrict__ part_scores,     // [gridDim.x, BT, k]
                int* __restrict__ part_rows) {
    constexpr int TILE = 32 * R;                       // rows per CTA step
    constexpr int CHUNK = BF16 ? 64 : 32;              // columns per warp-wide 16B load step
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sq = reinterpret_cast<float*>(smem_raw);             // [BT][ld]
    float* sscore = sq + (size_t)BT * ld;                      // [2][BT][TILE]
    float* ls = sscore + 2 * BT * TILE;                        // [BT][k]
    int* lr = reinterpret_cast<int*>(ls + (size_t)BT * k);      // [BT][k]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l8 = lane & 7, rg = lane >> 3;

    for (int i = tid * 4; i < BT * ld; i += kThreads * 4)
        *reinterpret_cast<float4*>(sq + i) = *reinterpret_cast<const float4*>(queries + i);
    for (int i = tid; i < BT * k; i += kThreads) { ls[i] = VQ_NEG_INF; lr[i] = VQ_EMPTY_ROW; }
    __syncthreads();

    // contiguous, balanced row range of this CTA (multiples of 4 rows)
    const long long groups = ((long long)n + 3) / 4;
    const int row_begin = (int)((groups * blockIdx.x) / gridDim.x) * 4;
    const int row_end_raw = (int)((groups * (blockIdx.x + 1)) / gridDim.x) * 4;
    const int row_end = row_end_raw < n ? row_end_raw : n;

    const unsigned char* store = reinterpret_cast<const unsigned char*>(store_v);
    const size_t row_bytes = (size_t)ld * (BF16 ? 2 : 4);

    int buf = 0;
    for (int tile = row_begin; tile < row_end; tile += TILE, buf ^= 1) {
        float acc[R][BT];
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int b = 0; b < BT; ++b) acc[i][b] = 0.f;

        const unsigned char* p[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            int row = tile + warp * (4 * R) + i * 4 + rg;
            row = row < n ? row : n - 1;                       // tail rows: clamp, masked at selection
            p[i] = store + (size_t)row * row_bytes + l8 * 16;
        }
        uint4 cur[R], nxt[R];
#pragma unroll
        for (int i = 0; i < R; ++i) cur[i] = vq_ldg_stream(p[i]);

        for (int j = 0; j < ld; j += CHUNK) {
            if (j + CHUNK < ld) {
#pragma unroll
                for (int i = 0; i < R; ++i) nxt[i] = vq_ldg_stream(p[i] + (size_t)(j + CHUNK) * (BF16 ? 2 : 4));
            }
            const float* qj = sq + j + l8 * (BF16 ? 8 : 4);
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                const float4 q0 = *reinterpret_cast<const float4*>(qj + (size_t)b * ld);
                if (BF16) {
                    const float4 q1 = *reinterpret_cast<const float4*>(qj + (size_t)b * ld + 4);
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        float a = acc[i][b];
                        a = fmaf(vq_bf16lo(cur[i].x), q0.x, a);
                        a = fmaf(vq_bf16hi(cur[i].x), q0.y, a);
                        a = fmaf(vq_bf16lo(cur[i].y), q0.z, a);
                        a = fmaf(vq_bf16hi(cur[i].y), q0.w, a);
                        a = fmaf(vq_bf16lo(cur[i].z), q1.x, a);
                        a = fmaf(vq_bf16hi(cur[i].z), q1.y, a);
                        a = fmaf(vq_bf16lo(cur[i].w), q1.z, a);
                        a = fmaf(vq_bf16hi(cur[i].w), q1.w, a);
                        acc[i][b] = a;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        float a = acc[i][b];
                        a = fmaf(__uint_as_float(cur[i].x), q0.x, a);
                        a = fmaf(__uint_as_float(cur[i].y), q0.y, a);
                        a = fmaf(__uint_as_float(cur[i].z), q0.z, a);
                        a = fmaf(__uint_as_float(cur[i].w), q0.w, a);
                        acc[i][b] = a;
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i) cur[i] = nxt[i];
        }

        // reduce over the 8 lanes that share a row
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int b = 0; b < BT; ++b) {
                float v = acc[i][b];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                acc[i][b] = v;
            }
        float* st = sscore + (size_t)buf * BT * TILE;
#pragma unroll
        for (int b = 0; b < BT; ++b)
            if ((b & 7) == l8) {
#pragma unroll
                for (int i = 0; i < R; ++i) st[b * TILE + warp * (4 * R) + i * 4 + rg] = acc[i][b];
            }
        __syncthreads();   // score tile complete; the other buffer is free again (see below)

        // fused top-k: warp w owns queries w, w+8, ...  Reads buffer `buf` while other warps may
        // already be producing the next tile into buffer `buf^1`; the barrier of that next tile
        // orders this read before buffer `buf` is written again two tiles later.
        for (int b = warp; b < BT; b += kWarps) {
            float* mls = ls + (size_t)b * k;
            int* mlr = lr + (size_t)b * k;
            float ts = mls[k - 1];
            int tr = mlr[k - 1];
#pragma unroll 1
            for (int c = 0; c < TILE; c += 32) {
                const int row = tile + c + lane;
                const float s = st[b * TILE + c + lane];
                unsigned m = __ballot_sync(0xffffffffu, row < row_end && vq_better(s, row, ts, tr));
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float cs = __shfl_sync(0xffffffffu, s, src);
                    const int cr = tile + c + src;
                    if (vq_better(cs, cr, ts, tr)) {
                        vq_list_insert(mls, mlr, k, cs, cr, lane);
                        ts = mls[k - 1];
                        tr = mlr[k - 1];
                    }
                }
            }
        }
    }
    __syncthreads();
    // publish this CTA's candidate lists
    float* ps = part_scores + (size_t)blockIdx.x * BT * k;
    int* pr = part_rows + (size_t)blockIdx.x * BT * k;
    for (int i = tid; i < BT * k; i += kThreads) {
        const int r = lr[i];
        ps[i] = ls[i];
        pr[i] = (r == VQ_EMPTY_ROW) ? -1 : r;
    }
}

template <int BT, int R, bool BF16>
cudaError_t launch(const void* store, int n, int ld, const float* q, int k, float* ps, int* pr,
                   int grid, size_t smem, cudaStream_t stream) {
    auto kern = scan_fma_kernel<BT, R, BF16>;
    static std::atomic<unsigned long long> attr_done{0};   // per instantiation, one bit per device
    if (vq_first_use_on_device(&attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        vq_mark_used(&attr_done);
    }
    kern<<<grid, kThreads, smem, stream>>>(store, n, ld, q, k, ps, pr);
    return cudaGetLastError();
}

}  // namespace

// Rows per CTA step for a query tile of `bt`.
int vq_scan_fma_tile_rows(int bt) { return bt <= 2 ? 256 : bt <= 16 ? 128 : 64; }

size_t vq_scan_fma_smem(int bt, int ld, int k) {
    const int tile = vq_scan_fma_tile_rows(bt);
    return (size_t)bt * ld * 4 + (size_t)2 * bt * tile * 4 + (size_t)bt * k * 8;
}

int vq_scan_fma_grid(int n, int bt) {
    const int tile = vq_scan_fma_tile_rows(bt);
    long long tiles = ((long long)n + tile - 1) / tile;
    long long g = 2LL * vq_num_sms();
    return (int)(tiles < g ? (tiles < 1 ? 1 : tiles) : g);
}

// Launch one pass for `bt` (1,2,4,8,16,32) queries.  q must be [bt, ld] zero padded.
int vq_scan_fma_launch(const void* store, int n, int ld, int store_dtype, const float* q, int bt, int k,
                       float* part_scores, int* part_rows, int grid, cudaStream_t stream) {
    const size_t smem = vq_scan_fma_smem(bt, ld, k);
    if (smem > 200 * 1024) {
        vq_set_error("scan_fma: shared memory %zu B too large (bt=%d ld=%d k=%d)", smem, bt, ld, k);
        return VQ_EUNSUPPORTED;
    }
    cudaError_t e = cudaErrorInvalidValue;
    const bool bf = store_dtype == VQ_BF16;
#define VQ_CASE(BT_, R_)                                                                            \
    case BT_:                                                                                       \
        e = bf ? launch<BT_, R_, true>(store, n, ld, q, k, part_scores, part_rows, grid, smem, stream)  \
               : launch<BT_, R_, false>(store, n, ld, q, k, part_scores, part_rows, grid, smem, stream); \
        break;
    switch (bt) {
        VQ_CASE(1, 8)
        VQ_CASE(2, 8)
        VQ_CASE(4, 4)
        VQ_CASE(8, 4)
        VQ_CASE(16, 4)
        VQ_CASE(32, 2)
        default:
            vq_set_error("scan_fma: unsupported query tile %d", bt);
            return VQ_EINVAL;
    }
#undef VQ_CASE
    if (e != cudaSuccess) {
        vq_set_error("launch of scan_fma_kernel<bt=%d> failed: %s", bt, cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    return VQ_OK;
}
