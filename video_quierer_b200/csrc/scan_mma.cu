// (b)+(c) dense-batch exact scan on the 5th-gen tensor cores: TMA -> shared memory ->
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld -> fused per-query top-k.
//
// Replaces the reference's per-query np.dot + argsort (video_search_overhaul.py:53,56 looped by
// src/api/routes.py:627-634) when the query batch makes the scan a dense contraction.
//
// Orientation: D[query, row] = Q[query, :] . S[row, :]  with  A = 128 queries (UMMA M = 128),
// B = 128 store rows (UMMA N = 128), both K-major, 128-byte swizzle.  After tcgen05.ld every
// epilogue thread owns ONE query (its TMEM lane) and sees the scores of 32 store rows at a time in
// registers: the running k-th best is a per-thread register, the top-k list is a private column
// of shared memory, no cross-thread traffic, and the scores never leave the SM.
//
// Warp roles (256 threads, 1 CTA / SM):
//   warp 0   TMA producer  : query tile once (resident for the whole kernel), then the store
//                            tiles k-block by k-block through a `stages`-deep mbarrier ring
//   warp 1   MMA issuer    : one thread, 4 x tcgen05.mma (K = 16) per k-block, tcgen05.commit
//                            frees the smem slot / publishes the accumulator
//   warp 2   TMEM allocator: 256 columns = two 128-column accumulators (MMA of tile i+1 overlaps
//                            the top-k epilogue of tile i)
//   warps 4-7 epilogue     : tcgen05.ld 32x32b.x32, threshold filter, insertion into the list
// Grid: persistent, gridDim = groups * n_qt; CTA c serves query tile c % n_qt and the store tiles
// c / n_qt, + groups, ... ; it writes one k-entry list per query, `topk_merge` reduces them.
#include <cuda.h>

#include "vq_common.cuh"

int vq_topk_merge_launch(const float* scores, const int* rows, int g, long long g_stride, int b_out, int k_in,
                         const long long* offsets, int k_out, float* out_scores, void* out_rows,
                         int rows64, int negate_out, cudaStream_t stream);
int vq_ingest_launch(const float* src, long long rows, int dim, int src_ld, void* dst, int dst_dtype,
                     int dst_ld, int mode, cudaStream_t stream);

namespace {

constexpr int QT = 128;                 // queries per tile   (UMMA M)
constexpr int NT = 128;                 // store rows per tile (UMMA N)
constexpr int KB_ELEMS = 64;            // bf16 per k-block = one 128-byte swizzle span
constexpr int A_KB_BYTES = QT * 128;    // 16 KB
constexpr int B_KB_BYTES = NT * 128;    // 16 KB
constexpr int TMEM_COLS = 2 * NT;       // double-buffered fp32 accumulator
constexpr int kThreads = 256;
constexpr int kMaxK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte off.  bits [16,30) (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset bits [32,46)
    d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = NT
__device__ __forceinline__ uint32_t umma_idesc_bf16() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
scan_mma_bf16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmS,
                     int n, int nkb, int n_qt, int k, int stages, int b_pad,
                     float* __restrict__ part_s, int* __restrict__ part_r) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                                   // nkb k-blocks of the query tile
    unsigned char* sB = sA + (size_t)nkb * A_KB_BYTES;          // ring of store-tile k-blocks
    float* ls = reinterpret_cast<float*>(sB + (size_t)stages * B_KB_BYTES);   // [k][128]
    int* lr = reinterpret_cast<int*>(ls + k * QT);                           // [k][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(lr + k * QT);
    uint64_t* empty = full + stages;
    uint64_t* a_full = empty + stages;
    uint64_t* tmem_full = a_full + 1;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x % n_qt, group = blockIdx.x / n_qt, n_groups = gridDim.x / n_qt;
    const int n_tiles = (n + NT - 1) / NT;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmS);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(a_full, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 4) {
        const int t = threadIdx.x - 128;
        for (int i = 0; i < k; ++i) { ls[i * QT + t] = VQ_NEG_INF; lr[i * QT + t] = VQ_EMPTY_ROW; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, (uint32_t)nkb * A_KB_BYTES);
            for (int kb = 0; kb < nkb; ++kb) tma_load_2d(sA + (size_t)kb * A_KB_BYTES, &tmQ, a_full, kb * KB_ELEMS, q_tile * QT);
            int stage = 0; uint32_t phase = 0;
            for (int tile = group; tile < n_tiles; tile += n_groups) {
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], B_KB_BYTES);
                    tma_load_2d(sB + (size_t)stage * B_KB_BYTES, &tmS, &full[stage], kb * KB_ELEMS, tile * NT);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16();
            mbar_wait(a_full, 0);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * NT;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + (size_t)kb * A_KB_BYTES);
                    const uint32_t b_addr = smem_u32(sB + (size_t)stage * B_KB_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < KB_ELEMS / 16; ++k4) {
                        const uint64_t ad = umma_desc_sw128(a_addr + k4 * 32);
                        const uint64_t bd = umma_desc_sw128(b_addr + k4 * 32);
                        umma_bf16(d_tmem, ad, bd, idesc, (kb | k4) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);           // smem slot reusable once these MMAs retire
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tmem_full[acc]);             // accumulator complete
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;                          // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        const int t = ew * 32 + lane;                     // query within the tile
        float tau = VQ_NEG_INF;
        int it = 0;
        for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row0 = tile * NT;
            const int valid = (n - row0) < NT ? (n - row0) : NT;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * NT + c0), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = __uint_as_float(v[j]);
                    if (s > tau && (c0 + j) < valid) {
                        // private insertion (rows arrive in ascending order, so equal scores keep row order)
                        int i = k - 1;
                        while (i > 0) {
                            const float p = ls[(i - 1) * QT + t];
                            if (!(p < s)) break;
                            ls[i * QT + t] = p;
                            lr[i * QT + t] = lr[(i - 1) * QT + t];
                            --i;
                        }
                        ls[i * QT + t] = s;
                        lr[i * QT + t] = row0 + c0 + j;
                        tau = ls[(k - 1) * QT + t];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const size_t dst = ((size_t)group * b_pad + (size_t)q_tile * QT + t) * k;
        for (int i = 0; i < k; ++i) {
            const int r = lr[i * QT + t];
            part_s[dst + i] = ls[i * QT + t];
            part_r[dst + i] = (r == VQ_EMPTY_ROW) ? -1 : r;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_map_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {KB_ELEMS, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct MmaPlan {
    int nkb, n_qt, b_pad, groups, grid, stages;
    size_t smem, qbf_bytes, part_bytes;
};
MmaPlan plan(int64_t n, int ld, int b, int k) {
    MmaPlan p;
    p.nkb = ld / KB_ELEMS;
    p.n_qt = (b + QT - 1) / QT;
    p.b_pad = p.n_qt * QT;
    const int sms = vq_num_sms();
    const long long n_tiles = (n + NT - 1) / NT;
    long long groups = sms / p.n_qt;
    if (groups < 1) groups = 1;
    if (groups > n_tiles) groups = n_tiles;
    p.groups = (int)groups;
    p.grid = p.groups * p.n_qt;
    const size_t fixed = 1024 + (size_t)p.nkb * A_KB_BYTES + (size_t)k * QT * 8 + 256;
    int st = (int)((227 * 1024 - fixed) / B_KB_BYTES);
    p.stages = st > 8 ? 8 : st;
    p.smem = fixed + (size_t)(p.stages > 0 ? p.stages : 0) * B_KB_BYTES;
    p.qbf_bytes = align256((size_t)p.b_pad * ld * 2);
    p.part_bytes = align256((size_t)p.groups * p.b_pad * k * 4);
    return p;
}

}  // namespace

bool vq_scan_mma_supported(int64_t n, int dim, int ld, int store_dtype, int b, int k) {
    (void)dim;
    if (store_dtype != VQ_BF16) return false;                 // kind::tf32 path for fp32 stores: not built yet
    if (ld % KB_ELEMS != 0 || n < 1 || b < 1 || k < 1 || k > kMaxK) return false;
    const MmaPlan p = plan(n, ld, b, k);
    if (p.n_qt > vq_num_sms()) return false;
    return p.stages >= 3;                                     // resident query tile + a useful ring must fit
}

size_t vq_scan_mma_workspace(int64_t n, int ld, int store_dtype, int b, int k) {
    if (store_dtype != VQ_BF16 || b < 1 || k < 1 || k > kMaxK || n < 1) return 0;
    const MmaPlan p = plan(n, ld, b, k);
    return p.qbf_bytes + 2 * p.part_bytes + 256;
}

int vq_scan_mma_run(const void* store, int64_t n, int dim, int ld, int store_dtype, const float* qnorm, int b, int k,
                    float* out_scores, int32_t* out_rows, void* ws_v, size_t ws_bytes, cudaStream_t stream, int* launches) {
    (void)dim;
    if (!vq_scan_mma_supported(n, dim, ld, store_dtype, b, k)) {
        vq_set_error("scan_mma: unsupported shape");
        return VQ_EUNSUPPORTED;
    }
    const MmaPlan p = plan(n, ld, b, k);
    if (ws_bytes < p.qbf_bytes + 2 * p.part_bytes) {
        vq_set_error("scan_mma: workspace too small");
        return VQ_EWORKSPACE;
    }
    unsigned char* ws = (unsigned char*)ws_v;
    __nv_bfloat16* qbf = (__nv_bfloat16*)ws;
    float* part_s = (float*)(ws + p.qbf_bytes);
    int* part_r = (int*)(ws + p.qbf_bytes + p.part_bytes);
    // queries: fp32 normalised [b, ld] -> bf16 [b_pad, ld], pad rows zero
    VQ_CUDA(cudaMemsetAsync(qbf, 0, (size_t)p.b_pad * ld * 2, stream));
    int rc = vq_ingest_launch(qnorm, b, ld, ld, qbf, VQ_BF16, ld, VQ_NORM_NONE, stream);
    if (rc) return rc;
    CUtensorMap tmQ, tmS;
    if (!make_map_bf16(&tmQ, qbf, (uint64_t)p.b_pad, (uint64_t)ld, QT) || !make_map_bf16(&tmS, store, (uint64_t)n, (uint64_t)ld, NT)) {
        vq_set_error("scan_mma: cuTensorMapEncodeTiled failed");
        return VQ_ECUDA;
    }
    static bool attr_done = false;
    if (!attr_done) {
        VQ_CUDA(cudaFuncSetAttribute(scan_mma_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done = true;
    }
    vq_prof_begin(stream);
    scan_mma_bf16_kernel<<<p.grid, kThreads, p.smem, stream>>>(tmQ, tmS, (int)n, p.nkb, p.n_qt, k, p.stages, p.b_pad, part_s, part_r);
    vq_prof_end(stream);
    VQ_LAUNCH_CHECK("scan_mma_bf16_kernel");
    rc = vq_topk_merge_launch(part_s, part_r, p.groups, (long long)p.b_pad * k, b, k, nullptr, k, out_scores, out_rows, 0, 0, stream);
    if (rc) return rc;
    *launches = 3;
    return VQ_OK;
}
