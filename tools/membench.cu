// Developer micro-benchmark: how fast can one B200 stream a row-major [n, ld] matrix from HBM
// with the access shapes the scan kernels use?  (Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o gpurun_out/membench tools/membench.cu ; run on the GPU box.)
//   ldg_rowwarp   : one warp reads 512 contiguous bytes per LDG.128 (whole rows, U loads in flight)
//   ldg_8lane     : 8 lanes x 16 B = one 128-byte line of a row, 4 rows per LDG.128 (scan_fma mapping)
//   tma2d_strided : TMA box [128 rows x 128 B] out of 1 KB/2 KB rows (scan_mma mapping), S stages
//   bulk1d        : cp.async.bulk of contiguous CHUNK bytes into an S-stage smem ring
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!done);
}

template <int U>
__global__ void __launch_bounds__(256) ldg_rowwarp(const uint4* __restrict__ src, size_t n_vec, unsigned* out) {
    // each warp takes contiguous 512 B pieces; U independent pieces in flight per lane
    const size_t warp = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (size_t)gridDim.x * 8;
    const int lane = threadIdx.x & 31;
    unsigned acc = 0;
    const size_t pieces = n_vec / 32;
    const size_t per = (pieces + nwarps - 1) / nwarps;
    const size_t lo = warp * per, hi = (lo + per < pieces) ? lo + per : pieces;
    for (size_t p = lo; p < hi; p += U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (p + u < hi) ? ldg_stream(src + (p + u) * 32 + lane) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int R>
__global__ void __launch_bounds__(256, 2) ldg_8lane(const unsigned char* __restrict__ src, int n, int row_bytes, unsigned* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane & 7, rg = lane >> 3;
    const long long groups = ((long long)n + 3) / 4;
    const int row_begin = (int)((groups * blockIdx.x) / gridDim.x) * 4;
    int row_end = (int)((groups * (blockIdx.x + 1)) / gridDim.x) * 4;
    if (row_end > n) row_end = n;
    unsigned acc = 0;
    for (int tile = row_begin; tile < row_end; tile += 32 * R) {
        const unsigned char* p[R];
#pragma unroll
        for (int i = 0; i < R; ++i) { int row = tile + warp * 4 * R + i * 4 + rg; row = row < n ? row : n - 1; p[i] = src + (size_t)row * row_bytes + l8 * 16; }
        uint4 cur[R], nxt[R];
#pragma unroll
        for (int i = 0; i < R; ++i) cur[i] = ldg_stream(p[i]);
        for (int j = 0; j < row_bytes; j += 128) {
            if (j + 128 < row_bytes) {
#pragma unroll
                for (int i = 0; i < R; ++i) nxt[i] = ldg_stream(p[i] + j + 128);
            }
#pragma unroll
            for (int i = 0; i < R; ++i) { acc += cur[i].x ^ cur[i].y ^ cur[i].z ^ cur[i].w; cur[i] = nxt[i]; }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

__global__ void __launch_bounds__(128, 1) tma2d_strided(const __grid_constant__ CUtensorMap map, int n_tiles, int nkb, int stages, int box_bytes, unsigned* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)stages * box_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], box_bytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(smem + (size_t)stage * box_bytes)), "l"(&map), "r"(smem_u32(&full[stage])), "r"(kb * 64), "r"(tile * 128) : "memory");
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0; unsigned acc = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[stage], phase);
                acc += *(volatile unsigned*)(smem + (size_t)stage * box_bytes);
                mbar_arrive(&empty[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

__global__ void __launch_bounds__(128, 1) bulk1d(const unsigned char* __restrict__ src, size_t total, int chunk, int stages, unsigned* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)stages * chunk);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const size_t n_chunks = total / chunk;
    // contiguous range per CTA
    const size_t lo = n_chunks * blockIdx.x / gridDim.x, hi = n_chunks * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (size_t c = lo; c < hi; ++c) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], chunk);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t)stage * chunk)), "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(&full[stage])) : "memory");
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0; unsigned acc = 0;
        for (size_t c = lo; c < hi; ++c) {
            mbar_wait(&full[stage], phase);
            acc += *(volatile unsigned*)(smem + (size_t)stage * chunk);
            mbar_arrive(&empty[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
float time_ms(F f, int iters = 10) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int n = 1000000;
    unsigned* out; CK(cudaMalloc(&out, 4));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void* p; EncodeTiledFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)); enc = (EncodeTiledFn)p;
    for (int row_bytes : {1024, 2048}) {
        const size_t bytes = (size_t)n * row_bytes;
        unsigned char* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
        printf("== row_bytes=%d total=%.2f GB\n", row_bytes, bytes / 1e9);
        float ms;
        ms = time_ms([&] { ldg_rowwarp<4><<<sms * 8, 256>>>((const uint4*)buf, bytes / 16, out); }); printf("ldg_rowwarp U=4  grid=%d : %.3f ms %.0f GB/s\n", sms * 8, ms, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_rowwarp<8><<<sms * 8, 256>>>((const uint4*)buf, bytes / 16, out); }); printf("ldg_rowwarp U=8  grid=%d : %.3f ms %.0f GB/s\n", sms * 8, ms, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_rowwarp<8><<<sms * 4, 256>>>((const uint4*)buf, bytes / 16, out); }); printf("ldg_rowwarp U=8  grid=%d : %.3f ms %.0f GB/s\n", sms * 4, ms, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_rowwarp<16><<<sms * 4, 256>>>((const uint4*)buf, bytes / 16, out); }); printf("ldg_rowwarp U=16 grid=%d : %.3f ms %.0f GB/s\n", sms * 4, ms, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_8lane<4><<<sms * 2, 256>>>(buf, n, row_bytes, out); }); printf("ldg_8lane R=4 (scan_fma)   : %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
        ms = time_ms([&] { ldg_8lane<8><<<sms * 2, 256>>>(buf, n, row_bytes, out); }); printf("ldg_8lane R=8              : %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
        // TMA 2D strided boxes of 128 rows x 128 B
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, (cuuint64_t)n}; cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
        cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        CK(cudaFuncSetAttribute(tma2d_strided, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        CK(cudaFuncSetAttribute(bulk1d, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        const int n_tiles = (n + 127) / 128, nkb = row_bytes / 128;
        for (int stages : {4, 8, 12}) {
            size_t smem = 1024 + (size_t)stages * 16384 + 256;
            ms = time_ms([&] { tma2d_strided<<<sms, 128, smem>>>(map, n_tiles, nkb, stages, 16384, out); });
            printf("tma2d_strided box=128x128B stages=%d : %.3f ms %.0f GB/s\n", stages, ms, bytes / ms / 1e6);
        }
        for (int chunk : {16384, 32768}) for (int stages : {2, 4, 6}) {
            if ((size_t)stages * chunk > 200 * 1024) continue;
            size_t smem = 1024 + (size_t)stages * chunk + 256;
            ms = time_ms([&] { bulk1d<<<sms, 128, smem>>>(buf, bytes, chunk, stages, out); });
            printf("bulk1d chunk=%d stages=%d (1 CTA/SM)   : %.3f ms %.0f GB/s\n", chunk, stages, ms, bytes / ms / 1e6);
        }
        CK(cudaFree(buf));
    }
    return 0;
}
