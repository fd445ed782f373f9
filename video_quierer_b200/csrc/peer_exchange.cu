// Shard/merge layer over NVLink peer memory: ONE kernel per rank replaces
// `ncclAllGather` of the per-GPU top-k + `topk_merge_kernel`.  The reference has no counterpart
// (single process).
//
// Every rank owns an exchange *window* in its HBM that all peers have mapped (CUDA IPC, NVSwitch):
//
//   [ header 256 B | lines uint4 [2 parities][world][b_max][k_max] ]     line = {score, epoch, row, epoch}
//
// One launch = one epoch e (kept in the window header, so a CUDA-graph replay needs no new arguments).
// CTA c serves queries [8c, 8c+8):
//   push   its k local (score,row) pairs per query are stored straight into EVERY peer's window (slot
//          [e&1][my rank]) as 16-byte lines whose two 8-byte halves each carry the epoch next to the
//          payload.  An aligned 8-byte store lands atomically, so a half whose flag word reads e holds this
//          epoch's payload: the data validates itself — no separate flag, no system-scope fence on the
//          sender, no acquire on the receiver (the low-latency line protocol of NCCL's LL transport);
//   merge  one warp per query: every lane polls its lines of the world lists (volatile 16-byte loads from
//          the rank's OWN memory) until both halves carry e, stages the payload in shared memory, and the
//          world sorted lists are merged by repeated head selection (score desc, global row asc — the
//          engine's order), shard offsets added, global rows written as int64.  A warp depends only on the
//          same query of its peers: no grid-wide or cross-rank barrier.
// Two parities suffice: a peer can start epoch e+2 only after all my CTAs pushed e+1, i.e. after all my
// CTAs finished reading epoch e.  A poll that exceeds the timeout (a peer that never launched) sets the
// error word of the header and *out_status instead of hanging the GPU.
// (First version: plain stores + one release flag per (rank, CTA) + acquire polling: the two system-scope
// membars of the sender were ~half of the kernel's 23 us, profiles/r01c_peer_exchange_b1024.*.)
#include <stdlib.h>
#include <string.h>

#include "vq_common.cuh"

namespace {

constexpr int kQpc = 8;              // queries per CTA = warps per CTA
constexpr int kThreads = kQpc * 32;
constexpr size_t kHeaderBytes = 256;
constexpr int kMaxWorld = 32;

struct PeerHeader {
    unsigned epoch;      // last completed epoch of THIS rank
    unsigned done;       // CTAs of the running epoch that have finished
    unsigned error;      // sticky: a wait timed out
    unsigned pad;
};

__host__ __device__ inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }
__host__ __device__ inline int ctas_for(int b) { return (b + kQpc - 1) / kQpc; }

__device__ __forceinline__ void st_line(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_line(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kThreads)
peer_exchange_merge_kernel(void* const* __restrict__ windows, int world, int rank, int b_max, int k_max,
                           const float* __restrict__ scores, const int* __restrict__ rows, int b, int k,
                           const long long* __restrict__ offsets, int k_out,
                           float* __restrict__ out_scores, long long* __restrict__ out_rows,
                           int* __restrict__ out_status, unsigned long long timeout_ns) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2* staged = reinterpret_cast<int2*>(smem_raw);                     // [kQpc][world][k]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
    unsigned char* mine = static_cast<unsigned char*>(windows[rank]);
    PeerHeader* hdr = reinterpret_cast<PeerHeader*>(mine);
    const unsigned e = *reinterpret_cast<volatile unsigned*>(&hdr->epoch) + 1u;
    const int par = (int)(e & 1u);
    const int q0 = cta * kQpc;
    const int nq = min(kQpc, b - q0);

    // ---- push: this CTA's lines into every window (own rank included), peers visited in a
    // rank-rotated order so that the NVLink traffic of a step is spread over all ports at once
    const int per_peer = nq * k;
    for (int i = tid; i < world * per_peer; i += kThreads) {
        const int p = i / per_peer, x = i - p * per_peer;
        const int ql = x / k, j = x - ql * k;
        int pp = p + rank;
        if (pp >= world) pp -= world;
        const size_t src = (size_t)(q0 + ql) * k + j;
        const uint4 v = make_uint4(__float_as_uint(scores[src]), e, (unsigned)rows[src], e);
        uint4* dst = reinterpret_cast<uint4*>(static_cast<unsigned char*>(windows[pp]) + kHeaderBytes) +
                     (((size_t)par * world + rank) * b_max + (q0 + ql)) * k_max + j;
        st_line(dst, v);
    }

    // ---- merge: warp `warp` owns query q0 + warp
    if (warp < nq) {
        const int q = q0 + warp;
        int2* mys = staged + (size_t)warp * world * k;
        const uint4* data = reinterpret_cast<const uint4*>(mine + kHeaderBytes);
        unsigned long long t0 = 0;
        for (int i = lane; i < world * k; i += 32) {
            const int g = i / k, j = i - g * k;
            const uint4* src = data + (((size_t)par * world + g) * b_max + q) * k_max + j;
            uint4 v = ld_line(src);
            unsigned spins = 0;
            while (v.y != e || v.w != e) {                  // the line of rank g has not landed yet
                if (spins == 0) t0 = globaltimer_ns();
                __nanosleep(spins < 64u ? 20u : 200u);      // a late peer: stop competing with the co-resident scan
                if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
                    atomicExch(&hdr->error, 1u);
                    if (out_status) *out_status = 1;
                    v = make_uint4(__float_as_uint(VQ_NEG_INF), e, 0xffffffffu, e);     // empty slot
                    break;
                }
                v = ld_line(src);
            }
            mys[i] = make_int2((int)v.x, (int)v.z);
        }
        __syncwarp();
        int head = 0;
        float s = VQ_NEG_INF;
        long long r = LLONG_MAX;
        const long long off = (lane < world && offsets) ? offsets[lane] : 0;
        auto load_head = [&]() {
            s = VQ_NEG_INF;
            r = LLONG_MAX;
            if (lane < world && head < k) {
                const int2 v = mys[lane * k + head];
                if (v.y >= 0) { s = __int_as_float(v.x); r = (long long)v.y + off; }
            }
        };
        load_head();
        for (int o = 0; o < k_out; ++o) {
            float bs = s;
            long long br = r;
            int bl = lane;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, bs, d);
                const long long orr = __shfl_xor_sync(0xffffffffu, br, d);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, d);
                const bool take = (os > bs) || (os == bs && (orr < br || (orr == br && ol < bl)));
                if (take) { bs = os; br = orr; bl = ol; }
            }
            if (lane == bl) {
                const bool valid = (r != LLONG_MAX);
                out_scores[(size_t)q * k_out + o] = valid ? s : VQ_NEG_INF;
                out_rows[(size_t)q * k_out + o] = valid ? r : -1ll;
                if (valid) { ++head; load_head(); }
            }
        }
    }

    // ---- the last CTA closes the epoch (read by the next launch on this stream)
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&hdr->done, 1u);
        if (prev == gridDim.x - 1) {
            hdr->done = 0;
            __threadfence();
            hdr->epoch = e;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// All-gather of the query batch over peer memory (the ingest side of a sharded search step): every
// rank copies only ITS slice of the host batch over PCIe (rows [rank*per, ...), per = ceil(b/world))
// and the slices are exchanged over NVLink with the same self-validating lines, 2 floats per line.
// (Measured on 8 GPUs: with every rank pulling the whole 2 MB batch from host memory the end-to-end
// step is capped near 0.25 ms whatever the pipeline depth, while the host thread only waits.)
//   window: [ header 256 B | lines uint4 [2 parities][world][per_max][ld_max / 2] ]
// CTA c serves rows [kRowsPerCta*c, ...) of EVERY slice: it pushes those rows of its own slice into all
// windows, then unpacks the same rows of each rank's slice from its own window into the dense matrix.
constexpr int kRowsPerCta = 1;

__global__ void __launch_bounds__(256)
peer_allgather_rows_kernel(void* const* __restrict__ windows, int world, int rank, int per_max, int ld_max,
                           const float* __restrict__ slice,      // [rows of this rank, dim] dense
                           int b, int dim, float* __restrict__ out,   // [b, dim] dense, identical on every rank
                           int* __restrict__ out_status, unsigned long long timeout_ns) {
    const int tid = threadIdx.x;
    unsigned char* mine = static_cast<unsigned char*>(windows[rank]);
    PeerHeader* hdr = reinterpret_cast<PeerHeader*>(mine);
    const unsigned e = *reinterpret_cast<volatile unsigned*>(&hdr->epoch) + 1u;
    const int par = (int)(e & 1u);
    const int per = (b + world - 1) / world;                    // rows per slice (the last may be short)
    const int lpr = dim >> 1, lpr_max = ld_max >> 1;            // lines per row
    const int r0 = blockIdx.x * kRowsPerCta;
    const int my_rows = max(0, min(per, b - rank * per));

    // ---- push rows [r0, r0 + kRowsPerCta) of my slice into every window
    const int nr = max(0, min(kRowsPerCta, my_rows - r0));
    const int per_peer = nr * lpr;
    for (int i = tid; i < world * per_peer; i += blockDim.x) {
        const int p = i / per_peer, x = i - p * per_peer;
        const int rl = x / lpr, j = x - rl * lpr;
        int pp = p + rank;
        if (pp >= world) pp -= world;
        const float2 f = *reinterpret_cast<const float2*>(slice + (size_t)(r0 + rl) * dim + 2 * j);
        uint4* dst = reinterpret_cast<uint4*>(static_cast<unsigned char*>(windows[pp]) + kHeaderBytes) +
                     (((size_t)par * world + rank) * per_max + (r0 + rl)) * lpr_max + j;
        st_line(dst, make_uint4(__float_as_uint(f.x), e, __float_as_uint(f.y), e));
    }

    // ---- unpack the same rows of every rank's slice from my own window
    const uint4* data = reinterpret_cast<const uint4*>(mine + kHeaderBytes);
    unsigned long long t0 = 0;
    for (int g = 0; g < world; ++g) {
        const int g_rows = max(0, min(per, b - g * per));
        const int gnr = max(0, min(kRowsPerCta, g_rows - r0));
        for (int i = tid; i < gnr * lpr; i += blockDim.x) {
            const int rl = i / lpr, j = i - rl * lpr;
            const uint4* src = data + (((size_t)par * world + g) * per_max + (r0 + rl)) * lpr_max + j;
            uint4 v = ld_line(src);
            unsigned spins = 0;
            while (v.y != e || v.w != e) {
                if (spins == 0) t0 = globaltimer_ns();
                __nanosleep(spins < 64u ? 20u : 200u);
                if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > timeout_ns) {
                    atomicExch(&hdr->error, 1u);
                    if (out_status) *out_status = 1;
                    v = make_uint4(0u, e, 0u, e);
                    break;
                }
                v = ld_line(src);
            }
            *reinterpret_cast<float2*>(out + ((size_t)g * per + r0 + rl) * dim + 2 * j) =
                make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
        }
    }

    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&hdr->done, 1u);
        if (prev == gridDim.x - 1) {
            hdr->done = 0;
            __threadfence();
            hdr->epoch = e;
        }
    }
}

unsigned long long exchange_timeout_ns() {
    static const unsigned long long t =
        getenv("VQ_PEER_TIMEOUT_MS") ? strtoull(getenv("VQ_PEER_TIMEOUT_MS"), nullptr, 10) * 1000000ull : 10000000000ull;
    return t;
}

}  // namespace

extern "C" {

size_t vq_peer_window_bytes(int world, int b_max, int k_max) {
    if (world <= 0 || b_max <= 0 || k_max <= 0) return 0;
    return kHeaderBytes + align256((size_t)2 * world * b_max * k_max * sizeof(uint4));
}

int vq_peer_window_create(size_t bytes, void** local_ptr, unsigned char* handle_out) {
    VQ_CHECK_ARG(bytes >= kHeaderBytes && local_ptr && handle_out, "vq_peer_window_create: bad arguments");
    void* p = nullptr;
    VQ_CUDA(cudaMalloc(&p, bytes));
    VQ_CUDA(cudaMemset(p, 0, bytes));
    VQ_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        vq_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == VQ_PEER_HANDLE_BYTES, "handle size");
    memcpy(handle_out, &h, sizeof(h));
    *local_ptr = p;
    return VQ_OK;
}

int vq_peer_window_open(const unsigned char* handle, void** peer_ptr) {
    VQ_CHECK_ARG(handle && peer_ptr, "vq_peer_window_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    VQ_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VQ_OK;
}

int vq_peer_window_close(void* peer_ptr) {
    if (peer_ptr) VQ_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return VQ_OK;
}

int vq_peer_window_destroy(void* local_ptr) {
    if (local_ptr) VQ_CUDA(cudaFree(local_ptr));
    return VQ_OK;
}

int vq_peer_window_status(const void* local_ptr, uint32_t* epoch_host, uint32_t* error_host) {
    VQ_CHECK_ARG(local_ptr, "vq_peer_window_status: null window");
    PeerHeader h;
    VQ_CUDA(cudaMemcpy(&h, local_ptr, sizeof(h), cudaMemcpyDeviceToHost));
    if (epoch_host) *epoch_host = h.epoch;
    if (error_host) *error_host = h.error;
    return VQ_OK;
}

int vq_peer_exchange_merge(const void* windows_dev, int world, int rank, int b_max, int k_max,
                           const float* scores, const int32_t* rows, int b, int k,
                           const int64_t* shard_offsets, int k_out,
                           float* out_scores, int64_t* out_rows, int32_t* out_status, void* stream) {
    VQ_CHECK_ARG(windows_dev && scores && rows && out_scores && out_rows, "vq_peer_exchange_merge: null pointer");
    VQ_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "world=%d rank=%d out of range (world <= %d)",
                 world, rank, kMaxWorld);
    VQ_CHECK_ARG(b >= 0 && b <= b_max && k >= 1 && k <= k_max, "b=%d k=%d exceed the window (b_max=%d k_max=%d)", b, k, b_max, k_max);
    VQ_CHECK_ARG(k_out >= 1, "k_out=%d must be >= 1", k_out);
    if (b == 0) return VQ_OK;
    const size_t smem = (size_t)kQpc * world * k * sizeof(int2);
    VQ_CHECK_ARG(smem <= 200 * 1024, "world*k=%d too large for the merge stage", world * k);
    static std::atomic<unsigned long long> attr_done{0};
    if (vq_first_use_on_device(&attr_done)) {
        VQ_CUDA(cudaFuncSetAttribute(peer_exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        vq_mark_used(&attr_done);
    }
    peer_exchange_merge_kernel<<<ctas_for(b), kThreads, smem, (cudaStream_t)stream>>>(
        (void* const*)windows_dev, world, rank, b_max, k_max, scores, rows, b, k,
        (const long long*)shard_offsets, k_out, out_scores, (long long*)out_rows, out_status, exchange_timeout_ns());
    VQ_LAUNCH_CHECK("peer_exchange_merge_kernel");
    return VQ_OK;
}

size_t vq_peer_rows_window_bytes(int world, int b_max, int ld_max) {
    if (world <= 0 || b_max <= 0 || ld_max <= 0 || (ld_max & 1)) return 0;
    const size_t per_max = ((size_t)b_max + world - 1) / world;
    return kHeaderBytes + align256((size_t)2 * world * per_max * (ld_max / 2) * sizeof(uint4));
}

int vq_peer_allgather_rows(const void* windows_dev, int world, int rank, int b_max, int ld_max,
                           const float* slice, int b, int dim, float* out, int32_t* out_status, void* stream) {
    VQ_CHECK_ARG(windows_dev && out, "vq_peer_allgather_rows: null pointer");
    VQ_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "world=%d rank=%d out of range (world <= %d)",
                 world, rank, kMaxWorld);
    VQ_CHECK_ARG(b >= 0 && b <= b_max && dim >= 2 && dim <= ld_max && (dim & 1) == 0 && (ld_max & 1) == 0,
                 "b=%d dim=%d exceed the window (b_max=%d ld_max=%d) or dim is odd", b, dim, b_max, ld_max);
    if (b == 0) return VQ_OK;
    const int per = (b + world - 1) / world, per_max = (b_max + world - 1) / world;
    VQ_CHECK_ARG(per <= per_max, "slice of %d rows exceeds the window (%d)", per, per_max);
    VQ_CHECK_ARG(slice || rank * per >= b, "vq_peer_allgather_rows: null slice");
    const int grid = (per + kRowsPerCta - 1) / kRowsPerCta;
    peer_allgather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((void* const*)windows_dev, world, rank, per_max, ld_max, slice,
                                                                       b, dim, out, out_status, exchange_timeout_ns());
    VQ_LAUNCH_CHECK("peer_allgather_rows_kernel");
    return VQ_OK;
}

}  // extern "C"
