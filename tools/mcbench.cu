// Developer micro-benchmark: does TMA multicast inside a thread-block cluster lift the L2 -> SM
// delivery cap that bounds the batch-1024 scan?  Mimics scan_mma's access pattern: `n_qt` CTAs (one
// per query tile) walk the SAME sequence of store tiles; unicast makes each of them pull every
// tile through the crossbar, multicast lets a cluster of CS CTAs pull each tile once (every CTA
// loads 1/CS of the box and broadcasts it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mcbench tools/mcbench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(b)), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = groups * n_qt CTAs, cluster = CS consecutive CTAs (same group).  CTA c: q = c % n_qt, group = c / n_qt.
template <int CS>
__global__ void __launch_bounds__(128, 1)
tma_mc(const __grid_constant__ CUtensorMap map, int n_tiles, int nkb, int stages, int n_qt, unsigned* out) {
    constexpr int BOX = 16384, SLICE = BOX / CS, SLICE_ROWS = 128 / CS;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + (size_t)stages * BOX);
    uint64_t* empty = full + stages;
    const uint32_t rank = CS > 1 ? cluster_ctarank() : 0;
    const int group = blockIdx.x / n_qt, n_groups = gridDim.x / n_qt;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (CS > 1) cluster_sync(); else __syncthreads();
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint16_t mask = (uint16_t)((1u << CS) - 1);
        for (int tile = group; tile < n_tiles; tile += n_groups)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);          // slot free in EVERY CTA of the cluster
                mbar_expect_tx(&full[stage], BOX);
                unsigned char* dst = smem + (size_t)stage * BOX + rank * SLICE;
                if (CS == 1)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                 ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(&full[stage])), "r"(kb * 64), "r"(tile * 128) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                 ::"r"(smem_u32(dst)), "l"(&map), "r"(smem_u32(&full[stage])), "r"(kb * 64),
                                   "r"(tile * 128 + (int)rank * SLICE_ROWS), "h"(mask) : "memory");
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0; unsigned acc = 0;
        for (int tile = group; tile < n_tiles; tile += n_groups)
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full[stage], phase);
                acc += *(volatile unsigned*)(smem + (size_t)stage * BOX) + *(volatile unsigned*)(smem + (size_t)stage * BOX + BOX - 4);
#pragma unroll
                for (int c = 0; c < CS; ++c) mbar_arrive_remote(&empty[stage], c);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        if (acc == 0x12345678u) out[0] = acc;
    }
    if (CS > 1) cluster_sync();
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

template <int CS>
void run(const void* buf, int n, int row_bytes, int n_qt, int sms, EncodeTiledFn enc, unsigned* out) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, (cuuint64_t)n}; cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {64, (cuuint32_t)(128 / CS)}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(buf), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    CK(cudaFuncSetAttribute(tma_mc<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const int n_tiles = n / 128, nkb = row_bytes / 128, stages = 12;
    const int groups = sms / n_qt, grid = groups * n_qt;
    const size_t smem = 1024 + (size_t)stages * 16384 + 256;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = -1;
    if (CS > 1) cudaOccupancyMaxActiveClusters(&nclusters, tma_mc<CS>, &cfg);
    float ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, tma_mc<CS>, map, n_tiles, nkb, stages, n_qt, out)); });
    const double unique = (double)n * row_bytes, delivered = unique * n_qt;
    printf("n_qt=%d cluster=%d grid=%d (max active clusters %d): %.3f ms  unique %.0f GB/s  delivered-to-SMs %.0f GB/s\n",
           n_qt, CS, grid, nclusters, ms, unique / ms / 1e6, delivered / ms / 1e6);
}

int main() {
    const int n = 1000000 / 128 * 128, row_bytes = 1024;
    unsigned* out; CK(cudaMalloc(&out, 4));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    void* p; EncodeTiledFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)); enc = (EncodeTiledFn)p;
    unsigned char* buf; CK(cudaMalloc(&buf, (size_t)n * row_bytes)); CK(cudaMemset(buf, 1, (size_t)n * row_bytes));
    for (int n_qt : {2, 4, 8}) {
        run<1>(buf, n, row_bytes, n_qt, sms, enc, out);
        run<2>(buf, n, row_bytes, n_qt, sms, enc, out);
        if (n_qt >= 4) run<4>(buf, n, row_bytes, n_qt, sms, enc, out);
        if (n_qt >= 8) run<8>(buf, n, row_bytes, n_qt, sms, enc, out);
    }
    return 0;
}
