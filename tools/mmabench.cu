// Developer micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16 -> fp32) on B200 for the
// operand placements the scan kernel can choose between.  No TMA, operands are whatever is in
// shared / tensor memory (zeroed) — only the tensor pipe is exercised.
//   variants: A from shared memory (SS) or tensor memory (TS); N = 64 / 128 / 256;
//             cta_group::1 (M = 128) or cta_group::2 (M = 256 over a CTA pair)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mmabench.bin tools/mmabench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
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
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int CG, bool TS>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (CG == 1) {
        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
        if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
}

// N = UMMA N (for CG=2 each CTA holds N/2 rows of B); accumulators: 2 x N columns if they fit next to A, else 1.
template <int CG, bool TS, int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, long long* cycles_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                    // 128 x 64 bf16, SW128: 16 KB
    unsigned char* sB = smem + 16384;            // up to 256 x 64 bf16: 32 KB
    uint64_t* bar = (uint64_t*)(smem + 16384 + 32768);
    uint32_t* tmem_ptr = (uint32_t*)(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    constexpr int A_COLS = TS ? 32 : 0;                       // one 64-element k-block of A in TMEM
    constexpr int N_ACC = (2 * N + A_COLS <= 512) ? 2 : 1;
    long long t0 = 0, t1 = 0;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = idesc_bf16(128 * CG, N);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem_base + (uint32_t)((it % N_ACC) * N);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t ad = desc_sw128(smem_u32(sA) + k4 * 32);
                const uint64_t bd = desc_sw128(smem_u32(sB) + k4 * 32);
                mma<CG, TS>(d, tmem_base + N_ACC * N + k4 * 8, ad, bd, idesc, (it | k4) ? 1u : 0u);
            }
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    }
    // every thread waits for the MMAs (in both CTAs of a pair) before TMEM is released
    mbar_wait(bar, 0);
    if (warp == 1 && lane == 0 && rank == 0) {
        t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) cluster_sync(); else __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// Same MMA stream, but paced like the scan kernel: a helper thread plays the TMA producer
// (wait empty[s] -> arrive full[s]), the issuer waits full[s], issues MPS MMAs, commits to empty[s].
// mode 1: commit only (no full/empty handshake); mode 2: full handshake.
template <int MPS>
__global__ void __launch_bounds__(128, 1) mma_ring(int iters, int mode, int stages, long long* cycles_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = smem;                    // 128 x 64 bf16 k-block, reused
    uint64_t* full = (uint64_t*)(smem + 32768);
    uint64_t* empty = full + 16;
    uint64_t* done = empty + 16;
    uint32_t* tmem_ptr = (uint32_t*)(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < 16; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    if (warp == 2 && lane == 0 && mode == 2) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[stage])) : "memory");
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    }
    if (warp == 1 && lane == 0) {
        const uint32_t idesc = idesc_bf16(128, 128);
        int stage = 0; uint32_t phase = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t d = tmem_base + (uint32_t)(((it * MPS / 32) & 1) * 128);
            if (mode == 2) { mbar_wait(&full[stage], phase); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#pragma unroll
            for (int k4 = 0; k4 < MPS; ++k4) {
                const uint64_t bd = desc_sw128(smem_u32(sB) + (k4 & 3) * 32 + (k4 >> 2) * 16384);
                mma<1, true>(d, tmem_base + 256 + k4 * 8, 0, bd, idesc, (it | k4) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty[stage])) : "memory");
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
        mbar_wait(done, 0);
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int MPS>
void run_ring(int mode, int stages, int sms, long long* d_cycles) {
    const int iters = 16384 / MPS;
    auto kern = mma_ring<MPS>;
    const size_t smem = 1024 + 32768 + 512;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<sms, 128, smem>>>(iters, mode, stages, d_cycles);
    CK(cudaDeviceSynchronize());
    long long h[256]; CK(cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost));
    printf("ring MMAs/stage=%d mode=%d stages=%2d : %.1f cycles/MMA\n", MPS, mode, stages, (double)h[0] / (iters * MPS));
}

// Bisection of what slows the issue stream inside the real scan kernel.  flags: 1 = four extra warps
// spin on an mbarrier (like the epilogue warps on tmem_full), 2 = the A tile is written with tcgen05.st
// first, 4 = a commit (to a barrier somebody waits on) every 16 MMAs, 8 = A columns / B slot vary per k-block
__global__ void __launch_bounds__(256, 1) mma_var(int iters, const int flags, long long* cycles_out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = smem;                    // 8 slots x 16 KB
    uint64_t* bar = (uint64_t*)(smem + 8 * 16384);
    uint64_t* gbar = bar + 1;                    // [4] group barriers
    uint32_t* tmem_ptr = (uint32_t*)(gbar + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // two bf16 of magnitude ~0.03..0.06 with random sign and mantissa
        const uint32_t v = ((h & 0x807f807fu) | 0x3d003d00u);
        ((uint32_t*)smem)[i] = (flags & 16) ? v : 0;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&gbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr;
    if (warp >= 4 && (flags & 2)) {
        const uint32_t lane_base = tmem_base + ((uint32_t)((warp - 4) * 32) << 16);
        for (int c = 256; c < 512; c += 8) {
            uint32_t x = 0x3f803f80u;   // bf16 1.0, 1.0
            if (flags & 16) { uint32_t h = (uint32_t)(c * 131 + threadIdx.x) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; x = (h & 0x807f807fu) | 0x3d003d00u; }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_base + c), "r"(x) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1 || (warp == 3 && (flags & 512))) {
        const uint32_t idesc = idesc_bf16(128, 128);
        const uint32_t sB_addr = smem_u32(sB);
        uint64_t* mybar = warp == 1 ? bar : &gbar[3];
        if (flags & 512) iters /= 2;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {           // one "k-block" of 4 MMAs per iteration
            const int kb = (flags & 8) ? (it & 7) : 0;
            const uint32_t d = tmem_base + (uint32_t)((flags & 512) ? (warp == 1 ? 0 : 128) : ((it >> 3) & 1) * 128);
            if (flags & 32) mbar_wait(&gbar[2], 1);          // completes immediately (previous phase)
            if (flags & 64) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t el;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
            if (el) {
                const uint64_t bd0 = desc_sw128(sB_addr + kb * 16384);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                    mma<1, true>(d, tmem_base + 256 + kb * 32 + k4 * 8, 0, bd0 + (uint64_t)(k4 * 2), idesc, ((it & 7) | k4) ? 1u : 0u);
                if ((flags & 4) && (it & 3) == 3)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&gbar[(it >> 2) & 1])) : "memory");
                if (flags & 128)    // a commit after every 4 MMAs, to a barrier nobody waits on
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&gbar[1])) : "memory");
                if ((flags & 256) && (it & 7) == 7)    // a second commit at every tile end (tmem_full)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&gbar[1])) : "memory");
            }
            __syncwarp();
        }
        uint32_t el;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
        if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mybar)) : "memory");
        __syncwarp();
        mbar_wait(mybar, 0);
        if (lane == 0 && warp == 1) cycles_out[blockIdx.x] = clock64() - t0;
    } else if (warp >= 4 && (flags & 1)) {
        mbar_wait(bar, 0);
    } else if (warp == 0 && lane == 0 && (flags & 4)) {
        // a consumer of the group barriers (like the TMA producer waiting for free slots)
        uint32_t ph = 0;
        for (int g = 0; g < iters / 4; ++g) { mbar_wait(&gbar[g & 1], ph); if ((g & 1) == 1) ph ^= 1; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

void run_var(int flags, int sms, long long* d_cycles) {
    const int iters = 4096;
    const size_t smem = 1024 + 8 * 16384 + 512;
    CK(cudaFuncSetAttribute(mma_var, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_var<<<sms, 256, smem>>>(iters, flags, d_cycles);
    CK(cudaDeviceSynchronize());
    long long h[256]; CK(cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost));
    printf("var flags=%2d (4 commit/16, 8 moving operands, 32 try_wait/4, 64 fence/4, 128 commit/4, 256 commit/32): %.1f cycles/MMA\n", flags, (double)h[0] / (iters * 4));
}

template <int CG, bool TS, int N>
void run(const char* name, int sms, long long* d_cycles) {
    const int iters = 4096;
    auto kern = mma_rate<CG, TS, N>;
    const size_t smem = 1024 + 16384 + 32768 + 64;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / CG * CG); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaMemset(d_cycles, 0, sizeof(long long) * 256));
    CK(cudaLaunchKernelEx(&cfg, kern, iters, d_cycles));
    CK(cudaDeviceSynchronize());
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        CK(cudaLaunchKernelEx(&cfg, kern, iters, d_cycles));
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    long long h[256]; CK(cudaMemcpy(h, d_cycles, sizeof(h), cudaMemcpyDeviceToHost));
    const double n_mma = (double)iters * 4;
    const double flop_per_mma_per_sm = 2.0 * 128 * N * 16;        // per SM (a pair does twice this per instruction)
    const double tf = flop_per_mma_per_sm * n_mma * cfg.gridDim.x / (best * 1e-3) / 1e12;
    printf("%-28s grid=%3d  %.1f cycles/MMA  %.3f ms  %.0f TFLOP/s\n", name, cfg.gridDim.x, (double)h[0] / n_mma, best, tf);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_cycles; CK(cudaMalloc(&d_cycles, sizeof(long long) * 256));
    run<1, true, 64>("cg1 TS M128 N64", sms, d_cycles);
    run<1, true, 128>("cg1 TS M128 N128", sms, d_cycles);
    run<1, true, 256>("cg1 TS M128 N256 (1 acc)", sms, d_cycles);
    run<1, false, 128>("cg1 SS M128 N128", sms, d_cycles);
    run<1, false, 256>("cg1 SS M128 N256", sms, d_cycles);
    run<2, true, 128>("cg2 TS M256 N128", sms, d_cycles);
    run<2, true, 256>("cg2 TS M256 N256 (1 acc)", sms, d_cycles);
    run<2, false, 128>("cg2 SS M256 N128", sms, d_cycles);
    run<2, false, 256>("cg2 SS M256 N256", sms, d_cycles);
    for (int f : {0, 32, 512, 544, 672}) run_var(f, sms, d_cycles);
    for (int mode : {1, 2}) for (int stages : {4, 12}) { run_ring<4>(mode, stages, sms, d_cycles); run_ring<8>(mode, stages, sms, d_cycles); }
    return 0;
}
