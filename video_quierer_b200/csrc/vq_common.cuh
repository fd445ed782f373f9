// Shared device/host helpers for the frame-search kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <limits.h>
#include <math.h>

#include "../../include/vq_search.h"

#define VQ_WARP 32
#define VQ_NEG_INF (-INFINITY)
#define VQ_EMPTY_ROW INT_MAX   // sentinel row of an unused top-k slot (sorts last among equal scores)

// ----------------------------------------------------------------------------- host errors
void vq_set_error(const char* fmt, ...);
void vq_note_launch(const char* path_or_null, int launches);
#define VQ_CHECK_ARG(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            vq_set_error(__VA_ARGS__);          \
            return VQ_EINVAL;                   \
        }                                       \
    } while (0)
#define VQ_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            vq_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return VQ_ECUDA;                                                              \
        }                                                                                 \
    } while (0)
#define VQ_LAUNCH_CHECK(name)                                                             \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            vq_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));       \
            return VQ_ECUDA;                                                              \
        }                                                                                 \
    } while (0)

int vq_num_sms();
// Per-device once guard for a call site: cudaFuncSetAttribute (and __device__ symbols) apply to ONE device, a
// process may drive several.  `vq_first_use_on_device(&flags)` returns true until `vq_mark_used(&flags)` has
// run on the current device; the flags word is a static std::atomic at the call site (one bit per device).
#include <atomic>
inline unsigned long long vq_device_bit() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return 1ull << (dev & 63);
}
inline bool vq_first_use_on_device(const std::atomic<unsigned long long>* flags) {
    return (flags->load(std::memory_order_acquire) & vq_device_bit()) == 0;
}
inline void vq_mark_used(std::atomic<unsigned long long>* flags) { flags->fetch_or(vq_device_bit(), std::memory_order_release); }
int vq_pdl_mask();           // bit i set: kernels of PDL class i are launched with the attribute
int vq_launch_priority(int launch_class);   // CUDA priority of a launch class (0 = default), see api.cu

// Launch with programmatic dependent launch (PDL): the kernel may start while the previous kernel of
// the stream is still running; it must execute vq_pdl_wait() before touching anything an earlier
// kernel wrote and vq_pdl_trigger() AFTER that wait (so that completion stays transitive along the
// chain).  What overlaps is the launch latency and the kernel's prologue (barrier init, TMEM
// allocation, descriptor prefetch).  Without the attribute both device calls are no-ops.
template <typename... KArgs, typename... Args>
cudaError_t vq_launch_cluster(int pdl_class, int cluster, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[3];
    int na = 0;
    if (const int prio = vq_launch_priority(pdl_class)) {
        at[na].id = cudaLaunchAttributePriority;
        at[na].val.priority = prio;
        ++na;
    }
    if ((vq_pdl_mask() >> pdl_class) & 1) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (cluster > 1) {                         // thread-block cluster of `cluster` consecutive CTAs along x
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = (unsigned)cluster;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t vq_launch(int pdl_class, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    return vq_launch_cluster(pdl_class, 1, kernel, grid, block, smem, stream, args...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void vq_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void vq_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
void vq_prof_begin(cudaStream_t s);
void vq_prof_end(cudaStream_t s);

// ----------------------------------------------------------------------------- device utils
// Total order of the engine: higher score first, then lower row.  NaN never beats anything.
__device__ __forceinline__ bool vq_better(float s, int r, float s2, int r2) {
    return (s > s2) || (s == s2 && r < r2);
}

// 128-bit streaming load that does not pollute L1 (the store is read exactly once per pass).
__device__ __forceinline__ uint4 vq_ldg_stream(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// monotone map: larger score -> smaller 32-bit key (ascending key order = best first); -0 is folded into +0
__device__ __forceinline__ uint32_t vq_score_key(float s) {
    if (s == 0.f) s = 0.f;
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ~u;
}
__device__ __forceinline__ float vq_key_score(uint32_t k) {
    uint32_t u = ~k;
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

__device__ __forceinline__ float vq_bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float vq_bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ float vq_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Insert (cs, cr) into a warp-owned descending top-k list held in shared memory
// (ls/lr: k entries, sentinel-filled).  Must be called by all 32 lanes with uniform args.
// Cost: O(k/32) warp steps.
__device__ __forceinline__ void vq_list_insert(float* ls, int* lr, int k, float cs, int cr, int lane) {
    // position = number of entries strictly better than the candidate
    int pos = 0;
    for (int base = 0; base < k; base += 32) {
        int i = base + lane;
        bool b = (i < k) && vq_better(ls[i], lr[i], cs, cr);
        pos += __popc(__ballot_sync(0xffffffffu, b));
    }
    if (pos >= k) return;
    // shift [pos, k-1) up by one, highest chunk first so nothing is overwritten before it is read
    for (int base = ((k - 1) / 32) * 32; base >= 0; base -= 32) {
        int i = base + lane;                // destination index
        float s = 0.f; int r = 0;
        bool mv = (i < k) && (i > pos);
        if (mv) { s = ls[i - 1]; r = lr[i - 1]; }
        __syncwarp();
        if (mv) { ls[i] = s; lr[i] = r; }
        __syncwarp();
        if (base <= pos) break;
    }
    if (lane == 0) { ls[pos] = cs; lr[pos] = cr; }
    __syncwarp();
}
