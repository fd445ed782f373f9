// (d) HNSW on the GPU: warp-per-query greedy/beam search, and a brute-force layer builder.
//
// Search replaces HNSWIndex.search / _search_layer (reference src/indexes/hnsw.py:76-121,
// 238-280, 488-528).  One warp owns one query:
//   * the query vector, the result list, the visited set and the neighbour staging buffer all
//     live in that warp's slice of shared memory;
//   * the reference's two heaps collapse into ONE list sorted by (distance, id) with an
//     "expanded" flag per entry: a candidate that has been evicted from the bounded result
//     heap can only ever terminate the search (its distance is >= the worst kept), so the
//     next node to expand is simply the first unexpanded entry of the list (same stop rule as
//     hnsw.py:103 and the same strict admit rule as :113, exact ties aside);
//   * neighbour ids of the expanded node are read by the first `deg` lanes, filtered through
//     an exact open-addressing visited set (atomicCAS in shared memory, per-layer reset through
//     a slot log like the reference's per-layer `visited = set()`), compacted with a ballot, and
//     their rows gathered 4 at a time: 8 lanes per row, 128-bit `ld.global.nc` loads, 16 loads
//     in flight per lane, fp32 FMA against the query in shared memory;
//   * distance evaluations and expansions are counted per query for the gather-bandwidth
//     roofline (bytes = evals * ld * elem + hops * deg * 4).
//
// Build: the reference inserts one node at a time in Python (~10 ms/insert, hnsw.py:150-229).
// Here a layer is built in three data-parallel steps: exact k-nearest members of every member
// (the fused scan + top-k kernels), reverse edges appended with atomics, and every list pruned
// back to the closest m (the batch analogue of :197-223, "closest M" selection of :123-148).
#include <stdlib.h>

#include <cub/device/device_radix_sort.cuh>

#include "vq_common.cuh"

int vq_scan_fma_grid(int n, int bt);
int vq_scan_fma_launch(const void* store, int n, int ld, int store_dtype, const float* q, int bt, int k,
                       float* part_scores, int* part_rows, int grid, cudaStream_t stream);
int vq_topk_merge_launch(const float* scores, const int* rows, int g, long long g_stride, int b_out, int k_in,
                         const long long* offsets, int k_out, float* out_scores, void* out_rows,
                         int rows64, int negate_out, cudaStream_t stream);
int vq_ingest_launch(const float* src, long long rows, int dim, int src_ld, void* dst, int dst_dtype,
                     int dst_ld, int mode, cudaStream_t stream);
bool vq_scan_mma_supported(int64_t n, int dim, int ld, int store_dtype, int b, int k);
size_t vq_scan_mma_prepared_workspace(int64_t n, int ld, int b, int k);
int vq_scan_mma_prepared(const void* store, int64_t n, int ld, const void* qbf, int b, int k, float* out_scores,
                         int32_t* out_rows, void* ws, size_t ws_bytes, cudaStream_t stream);

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kExpanded = 0x40000000;      // flag bit on list ids (node ids are < 2^30)
constexpr int kLogCap = 512;               // visited-slot log (per-layer reset without a full clear)

struct WarpState {
    float* q;        // [ld]
    float* ld_;      // list distances [ef]
    int* li;         // list ids (| kExpanded) [ef]
    int* hash;       // [cap]
    unsigned short* log;   // [kLogCap]
    int* nbr;        // [32]
    float* nd;       // [32]
};

// ascending (distance, id) insert into a list of `cnt` entries bounded by `cap`; returns new cnt
__device__ __forceinline__ int list_insert_asc(float* d, int* id, int cnt, int cap, float cd, int cid, int lane) {
    int pos = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        bool b = false;
        if (i < cnt) {
            const float e = d[i];
            const int eid = id[i] & ~kExpanded;
            b = (e < cd) || (e == cd && eid < cid);
        }
        pos += __popc(__ballot_sync(kFull, b));
    }
    if (pos >= cap) return cnt;
    const int ncnt = cnt < cap ? cnt + 1 : cap;
    for (int base = ((ncnt - 1) / 32) * 32; base >= 0; base -= 32) {
        const int i = base + lane;
        float s = 0.f; int r = 0;
        const bool mv = (i < ncnt) && (i > pos);
        if (mv) { s = d[i - 1]; r = id[i - 1]; }
        __syncwarp();
        if (mv) { d[i] = s; id[i] = r; }
        __syncwarp();
        if (base <= pos) break;
    }
    if (lane == 0) { d[pos] = cd; id[pos] = cid; }
    __syncwarp();
    return ncnt;
}


// The same list held in REGISTERS, striped over the warp: entry i lives in lane i % 32, slot i / 32
// (S slots per lane, capacity 32*S >= ef).  An insertion is S ballots to find the position and one
// shuffle-shift per slot instead of a shared-memory shift loop with two warp barriers per 32 entries
// (~5x fewer cycles at ef = 256, where the list work had grown as large as the row gathers).
template <int S>
struct RegList {
    float d[S];
    int id[S];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int s = 0; s < S; ++s) { d[s] = INFINITY; id[s] = -1; }
    }
    // value of entry `idx` (warp-uniform idx), broadcast to all lanes
    __device__ __forceinline__ float dist_at(int idx) const {
        float v = d[0];
#pragma unroll
        for (int s = 1; s < S; ++s) v = (idx >> 5) == s ? d[s] : v;
        return __shfl_sync(kFull, v, idx & 31);
    }
    __device__ __forceinline__ int id_at(int idx) const {
        int v = id[0];
#pragma unroll
        for (int s = 1; s < S; ++s) v = (idx >> 5) == s ? id[s] : v;
        return __shfl_sync(kFull, v, idx & 31);
    }
    __device__ __forceinline__ void mark_expanded(int idx, int lane) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if ((idx >> 5) == s && (idx & 31) == lane) id[s] |= kExpanded;
    }
    // first entry (< cnt) without the expanded flag, -1 if none
    __device__ __forceinline__ int first_open(int cnt, int lane) const {
        int pos = -1;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int i = s * 32 + lane;
            const unsigned mk = __ballot_sync(kFull, i < cnt && !(id[s] & kExpanded));
            if (pos < 0 && mk) pos = s * 32 + __ffs(mk) - 1;
        }
        return pos;
    }
    // ascending (distance, id) insert bounded by cap; returns the new count
    __device__ __forceinline__ int insert(int cnt, int cap, float cd, int cid, int lane) {
        int pos = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int i = s * 32 + lane;
            const int eid = id[s] & ~kExpanded;
            const bool b = i < cnt && ((d[s] < cd) || (d[s] == cd && eid < cid));
            pos += __popc(__ballot_sync(kFull, b));
        }
        if (pos >= cap) return cnt;
        const int ncnt = cnt < cap ? cnt + 1 : cap;
        float carry_d = 0.f; int carry_i = 0;              // lane 31 of the previous slot
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float up_d = __shfl_up_sync(kFull, d[s], 1);
            const int up_i = __shfl_up_sync(kFull, id[s], 1);
            const float last_d = __shfl_sync(kFull, d[s], 31);
            const int last_i = __shfl_sync(kFull, id[s], 31);
            const float sh_d = lane == 0 ? carry_d : up_d;    // the entry that sat one index below
            const int sh_i = lane == 0 ? carry_i : up_i;
            const int i = s * 32 + lane;
            if (i == pos) { d[s] = cd; id[s] = cid; }
            else if (i > pos && i < ncnt) { d[s] = sh_d; id[s] = sh_i; }
            carry_d = last_d; carry_i = last_i;
        }
        return ncnt;
    }
};

template <bool BF16>
__device__ __forceinline__ float row_dot_8lanes(const unsigned char* row, const float* q, int ld, int l8) {
    // 8 lanes cooperate on one row; lane l8 takes the 16-byte pieces l8, l8+8, ...
    constexpr int EPL = BF16 ? 8 : 4;            // elements per 16-byte load
    const int steps = ld / (8 * EPL);
    // All 16-byte pieces of a lane (up to kMLP at a time) are requested before the first one is used:
    // the search is bound by the latency of these dependent gathers (ncu: 64 % of the stall samples sat
    // on the first FMA after a 4-deep load batch), so memory-level parallelism is what buys bandwidth.
    constexpr int kMLP = 16;
    float acc = 0.f;
    for (int j0 = 0; j0 < steps; j0 += kMLP) {
        uint4 x[kMLP];
#pragma unroll
        for (int u = 0; u < kMLP; ++u)
            if (j0 + u < steps) x[u] = vq_ldg_stream(row + ((size_t)(j0 + u) * 8 + l8) * 16);
#pragma unroll
        for (int u = 0; u < kMLP; ++u) {
            if (j0 + u < steps) {
                const float* qq = q + ((j0 + u) * 8 + l8) * EPL;
                const float4 q0 = *reinterpret_cast<const float4*>(qq);
                if (BF16) {
                    const float4 q1 = *reinterpret_cast<const float4*>(qq + 4);
                    acc = fmaf(vq_bf16lo(x[u].x), q0.x, acc); acc = fmaf(vq_bf16hi(x[u].x), q0.y, acc);
                    acc = fmaf(vq_bf16lo(x[u].y), q0.z, acc); acc = fmaf(vq_bf16hi(x[u].y), q0.w, acc);
                    acc = fmaf(vq_bf16lo(x[u].z), q1.x, acc); acc = fmaf(vq_bf16hi(x[u].z), q1.y, acc);
                    acc = fmaf(vq_bf16lo(x[u].w), q1.z, acc); acc = fmaf(vq_bf16hi(x[u].w), q1.w, acc);
                } else {
                    acc = fmaf(__uint_as_float(x[u].x), q0.x, acc); acc = fmaf(__uint_as_float(x[u].y), q0.y, acc);
                    acc = fmaf(__uint_as_float(x[u].z), q0.z, acc); acc = fmaf(__uint_as_float(x[u].w), q0.w, acc);
                }
            }
        }
    }
    acc += __shfl_xor_sync(kFull, acc, 1);
    acc += __shfl_xor_sync(kFull, acc, 2);
    acc += __shfl_xor_sync(kFull, acc, 4);
    return acc;
}

// bf16 rows are half as long: FOUR lanes per row take the same 16 loads per lane, so a warp gathers 8 rows per round trip
// instead of 4 with the same registers in flight (the search is bound by the number of dependent round trips per hop).
__device__ __forceinline__ float row_dot_4lanes_bf16(const unsigned char* row, const float* q, int ld, int l4) {
    const int steps = ld / 32;                   // 16-byte pieces per lane (8 bf16 each, 4 lanes)
    constexpr int kMLP = 16;
    float acc = 0.f;
    for (int j0 = 0; j0 < steps; j0 += kMLP) {
        uint4 x[kMLP];
#pragma unroll
        for (int u = 0; u < kMLP; ++u)
            if (j0 + u < steps) x[u] = vq_ldg_stream(row + ((size_t)(j0 + u) * 4 + l4) * 16);
#pragma unroll
        for (int u = 0; u < kMLP; ++u) {
            if (j0 + u < steps) {
                const float* qq = q + ((j0 + u) * 4 + l4) * 8;
                const float4 q0 = *reinterpret_cast<const float4*>(qq);
                const float4 q1 = *reinterpret_cast<const float4*>(qq + 4);
                acc = fmaf(vq_bf16lo(x[u].x), q0.x, acc); acc = fmaf(vq_bf16hi(x[u].x), q0.y, acc);
                acc = fmaf(vq_bf16lo(x[u].y), q0.z, acc); acc = fmaf(vq_bf16hi(x[u].y), q0.w, acc);
                acc = fmaf(vq_bf16lo(x[u].z), q1.x, acc); acc = fmaf(vq_bf16hi(x[u].z), q1.y, acc);
                acc = fmaf(vq_bf16lo(x[u].w), q1.z, acc); acc = fmaf(vq_bf16hi(x[u].w), q1.w, acc);
            }
        }
    }
    acc += __shfl_xor_sync(kFull, acc, 1);
    acc += __shfl_xor_sync(kFull, acc, 2);
    return acc;
}

template <bool BF16, int S>        // S > 0: result list in registers (ef <= 32*S); S == 0: in shared memory
__global__ void __launch_bounds__(128)
hnsw_search_kernel(const void* __restrict__ store_v, int ld,
                   const int* __restrict__ adj0, int m0,
                   const int* __restrict__ upper_off, const int* __restrict__ upper_adj, int m,
                   int entry, int max_level, int ef, int cap,
                   const float* __restrict__ qn,    // [b, ld]
                   int b, int k,
                   float* __restrict__ out_dist, int* __restrict__ out_rows, unsigned* __restrict__ out_stats,
                   int warp_smem_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= b) return;                                   // whole warp exits together
    unsigned char* base = smem_raw + (size_t)warp * warp_smem_bytes;
    WarpState w;
    w.q = reinterpret_cast<float*>(base);
    w.ld_ = w.q + ld;
    w.li = reinterpret_cast<int*>(w.ld_ + ef);
    w.hash = w.li + ef;
    w.nbr = w.hash + cap;
    w.nd = reinterpret_cast<float*>(w.nbr + 32);
    w.log = reinterpret_cast<unsigned short*>(w.nd + 32);

    const unsigned char* store = reinterpret_cast<const unsigned char*>(store_v);
    const size_t row_bytes = (size_t)ld * (BF16 ? 2 : 4);
    const int l8 = lane & 7, grp = lane >> 3;
    const unsigned ucap = (unsigned)cap;                   // any size: slot = hi32(hash * cap)

    for (int i = lane * 4; i < ld; i += 128)
        *reinterpret_cast<float4*>(w.q + i) = *reinterpret_cast<const float4*>(qn + (size_t)qi * ld + i);
    for (int i = lane; i < cap; i += 32) w.hash[i] = -1;
    __syncwarp();

    unsigned evals = 0, hops = 0, overflow = 0;
    RegList<(S > 0 ? S : 1)> rl;
    rl.clear();
    int visited_cnt = 0;            // entries currently in the hash (all lanes hold the same value)
    int logged = 0;                 // slots recorded in the log (== visited_cnt while <= kLogCap)

    // exact visited-set insert; returns true if `v` was not present.  Divergent-safe (smem atomics).
    auto visit = [&](int v, int& slot_out) -> bool {
        unsigned h = __umulhi((unsigned)v * 2654435761u, ucap);
        for (;;) {
            const int prev = atomicCAS(&w.hash[h], -1, v);
            if (prev == -1) { slot_out = (int)h; return true; }
            if (prev == v) { slot_out = -1; return false; }
            h = h + 1 == ucap ? 0u : h + 1;
        }
    };

    // distance of the entry point
    float cur_d;
    {
        const float dot = row_dot_8lanes<BF16>(store + (size_t)entry * row_bytes, w.q, ld, l8);
        cur_d = 1.0f - __shfl_sync(kFull, dot, 0);
        evals += 1;
    }
    int cur = entry;

    for (int lv = max_level; lv >= 0; --lv) {
        const int ef_l = lv == 0 ? ef : 1;
        const int deg = lv == 0 ? m0 : m;
        // ---- per-layer reset of the visited set (hnsw.py:87 creates a fresh set per layer)
        if (visited_cnt > 0) {
            if (logged == visited_cnt && logged <= kLogCap) {
                for (int i = lane; i < logged; i += 32) w.hash[w.log[i]] = -1;
            } else {
                for (int i = lane; i < cap; i += 32) w.hash[i] = -1;
            }
            __syncwarp();
        }
        visited_cnt = 0; logged = 0;
        int cnt = 1;
        if (S > 0) {
            rl.clear();
            if (lane == 0) { rl.d[0] = cur_d; rl.id[0] = cur; }
        }
        if (lane == 0) {
            if (S == 0) { w.ld_[0] = cur_d; w.li[0] = cur; }
            int s; visit(cur, s);
            w.log[0] = (unsigned short)s;
        }
        visited_cnt = 1; logged = 1;
        __syncwarp();

        for (;;) {
            // first unexpanded entry of the list = closest open candidate
            int pos = -1;
            int u;
            if (S > 0) {
                pos = rl.first_open(cnt, lane);
                if (pos < 0) break;
                u = rl.id_at(pos);
                rl.mark_expanded(pos, lane);
            } else {
                for (int base0 = 0; base0 < cnt && pos < 0; base0 += 32) {
                    const int i = base0 + lane;
                    const unsigned mk = __ballot_sync(kFull, i < cnt && !(w.li[i] & kExpanded));
                    if (mk) pos = base0 + __ffs(mk) - 1;
                }
                if (pos < 0) break;
                u = w.li[pos];
                __syncwarp();
                if (lane == 0) w.li[pos] = u | kExpanded;
            }
            hops += 1;
            const int* arow = lv == 0 ? adj0 + (size_t)u * m0
                                      : upper_adj + ((size_t)upper_off[u] + lv - 1) * m;
            for (int c0 = 0; c0 < deg; c0 += 32) {
                int v = (c0 + lane < deg) ? arow[c0 + lane] : -1;
                int slot = -1;
                bool fresh = false;
                if (v >= 0 && !overflow) fresh = visit(v, slot);
                const unsigned mk = __ballot_sync(kFull, fresh);
                const int nnew = __popc(mk);
                const int my = __popc(mk & ((1u << lane) - 1u));
                if (fresh) {
                    w.nbr[my] = v;
                    if (logged + my < kLogCap) w.log[logged + my] = (unsigned short)slot;
                }
                visited_cnt += nnew;
                logged = (logged + nnew <= kLogCap) ? logged + nnew : kLogCap + 1;   // > cap => full clear next time
                if (visited_cnt > cap - cap / 8) overflow = 1;
                __syncwarp();
                // gather + distance: 4 rows per step (8 lanes each); bf16 rows (ld a multiple of 32): 8 rows per step, 4 lanes each
                if (BF16 && (ld & 31) == 0) {
                    for (int i0 = 0; i0 < nnew; i0 += 8) {
                        const int mine = i0 + (lane >> 2);
                        const int node = w.nbr[mine < nnew ? mine : nnew - 1];
                        const float dot = row_dot_4lanes_bf16(store + (size_t)node * row_bytes, w.q, ld, lane & 3);
                        if ((lane & 3) == 0 && mine < nnew) w.nd[mine] = 1.0f - dot;
                    }
                } else {
                    for (int i0 = 0; i0 < nnew; i0 += 4) {
                        const int mine = i0 + grp;
                        const int node = w.nbr[mine < nnew ? mine : nnew - 1];
                        const float dot = row_dot_8lanes<BF16>(store + (size_t)node * row_bytes, w.q, ld, l8);
                        if (l8 == 0 && mine < nnew) w.nd[mine] = 1.0f - dot;
                    }
                }
                evals += nnew;
                __syncwarp();
                // admit in neighbour order (hnsw.py:113): not full, or strictly better than the worst kept
                if (S > 0) {
                    float worst = rl.dist_at(ef_l - 1);
                    for (int i = 0; i < nnew; ++i) {
                        const float d = w.nd[i];
                        const int node = w.nbr[i];
                        if ((cnt < ef_l) || (d < worst)) {
                            cnt = rl.insert(cnt, ef_l, d, node, lane);
                            worst = rl.dist_at(ef_l - 1);
                        }
                    }
                } else {
                    for (int i = 0; i < nnew; ++i) {
                        const float d = w.nd[i];
                        const int node = w.nbr[i];
                        const bool admit = (cnt < ef_l) || (d < w.ld_[ef_l - 1]);
                        if (admit) cnt = list_insert_asc(w.ld_, w.li, cnt, ef_l, d, node, lane);
                    }
                }
                __syncwarp();
            }
        }
        if (S > 0) {
            cur = rl.id_at(0) & ~kExpanded;
            cur_d = rl.dist_at(0);
            if (lv == 0) {
#pragma unroll
                for (int s2 = 0; s2 < S; ++s2) {
                    const int i = s2 * 32 + lane;
                    if (i < k) {
                        const bool ok = i < cnt;
                        out_dist[(size_t)qi * k + i] = ok ? rl.d[s2] : INFINITY;
                        out_rows[(size_t)qi * k + i] = ok ? (rl.id[s2] & ~kExpanded) : -1;
                    }
                }
                for (int i = S * 32 + lane; i < k; i += 32) { out_dist[(size_t)qi * k + i] = INFINITY; out_rows[(size_t)qi * k + i] = -1; }
            }
        } else {
        cur = w.li[0] & ~kExpanded;
        cur_d = w.ld_[0];
        if (lv == 0) {
            for (int i = lane; i < k; i += 32) {
                const bool ok = i < cnt;
                out_dist[(size_t)qi * k + i] = ok ? w.ld_[i] : INFINITY;
                out_rows[(size_t)qi * k + i] = ok ? (w.li[i] & ~kExpanded) : -1;
            }
        }
        }
        __syncwarp();
    }
    if (out_stats && lane == 0) {
        out_stats[(size_t)qi * 4 + 0] = evals;
        out_stats[(size_t)qi * 4 + 1] = hops;
        out_stats[(size_t)qi * 4 + 2] = overflow;
        out_stats[(size_t)qi * 4 + 3] = 0;
    }
}

// ------------------------------------------------------------------------------------ build
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ members, long long n_members,
                                   int row_vec, uint4* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_members * row_vec) return;
    const long long r = i / row_vec;
    const int c = (int)(i - r * row_vec);
    dst[i] = src[(size_t)members[r] * row_vec + c];
}

// forward pick: first m non-self, valid entries of u's exact k-nearest list; every pick also
// lands in the target's reverse buffer (bounded, atomics).
__global__ void link_reverse_kernel(const float* __restrict__ knn_s, const int* __restrict__ knn_r, long long n_members,
                                    int kk, int m, int rcap, int* __restrict__ fwd, float* __restrict__ fwd_s,
                                    int* __restrict__ rev_cnt, int* __restrict__ rev, float* __restrict__ rev_s) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_members) return;
    int taken = 0;
    for (int j = 0; j < kk && taken < m; ++j) {
        const int v = knn_r[u * kk + j];
        if (v < 0 || v == (int)u) continue;
        const float s = knn_s[u * kk + j];
        fwd[u * m + taken] = v;
        fwd_s[u * m + taken] = s;
        ++taken;
        const int slot = atomicAdd(&rev_cnt[v], 1);
        if (slot < rcap) { rev[(size_t)v * rcap + slot] = (int)u; rev_s[(size_t)v * rcap + slot] = s; }
    }
    for (; taken < m; ++taken) { fwd[u * m + taken] = -1; fwd_s[u * m + taken] = VQ_NEG_INF; }
}

// one warp per node: union(forward, reverse) -> dedupe -> closest m by (score desc, id asc)
__global__ void __launch_bounds__(256)
prune_kernel(const int* __restrict__ fwd, const float* __restrict__ fwd_s, const int* __restrict__ rev_cnt,
             const int* __restrict__ rev, const float* __restrict__ rev_s, long long n_members, int m, int rcap,
             const int* __restrict__ members, int* __restrict__ adj_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (u >= n_members) return;
    const int tmax = m + rcap;
    float* cs = reinterpret_cast<float*>(smem_raw) + (size_t)warp * tmax * 2;
    int* ci = reinterpret_cast<int*>(cs + tmax);
    int nrev = rev_cnt[u];
    nrev = nrev < rcap ? nrev : rcap;
    const int total = m + nrev;
    for (int i = lane; i < total; i += 32) {
        if (i < m) { ci[i] = fwd[u * m + i]; cs[i] = fwd_s[u * m + i]; }
        else { ci[i] = rev[(size_t)u * rcap + (i - m)]; cs[i] = rev_s[(size_t)u * rcap + (i - m)]; }
    }
    for (int i = lane; i < m; i += 32) adj_out[u * m + i] = -1;
    __syncwarp();
    // pass 1: drop later copies of an id (an edge can be both a forward pick and a reverse add)
    bool dupf[8] = {false, false, false, false, false, false, false, false};
    for (int e = lane, t = 0; e < total && t < 8; e += 32, ++t) {
        const int id = ci[e];
        if (id < 0) continue;
        for (int f = 0; f < e; ++f)
            if (ci[f] == id) { dupf[t] = true; break; }
    }
    __syncwarp();
    for (int e = lane, t = 0; e < total && t < 8; e += 32, ++t)
        if (dupf[t]) ci[e] = -1;
    __syncwarp();
    // pass 2: rank by (score desc, id asc) and keep the closest m
    for (int e = lane; e < total; e += 32) {
        const int id = ci[e];
        if (id < 0) continue;
        const float s = cs[e];
        int rank = 0;
        for (int f = 0; f < total; ++f) {
            const int fid = ci[f];
            if (fid >= 0 && f != e && vq_better(cs[f], fid, s, id)) ++rank;
        }
        if (rank < m) adj_out[u * m + rank] = members ? members[id] : id;
    }
}


// ---- diversity selection (the HNSW "select neighbours" heuristic, evaluated from k-nearest lists)
// Candidates (member indices `cid`, scores `cs`, sorted best first, T of them) of node u are taken
// greedily; candidate c is rejected when some already selected a is closer to c than u is:
// s(c,a) > s(c,u).  s(c,a) is looked up in c's own exact k-nearest list — if a is not among c's
// k nearest then s(c,a) <= s_k(c), which cannot exceed s(c,u) whenever u itself is in that list —
// so no embedding row is gathered at all.  Rejected candidates refill the list while there is
// room ("keep pruned connections"), so every node keeps its full degree.
__device__ int diverse_select(const int* cid, const float* cs, int T, const int* __restrict__ knn_r,
                              const float* __restrict__ knn_s, int kk, int m, int* sel, float* sel_s,
                              unsigned char* rej, int lane) {
    int ns = 0;
    for (int j = 0; j < T && ns < m; ++j) {
        const int c = cid[j];
        const float sc = cs[j];
        int li[3]; float lsc[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int idx = lane + 32 * t;
            li[t] = idx < kk ? knn_r[(size_t)c * kk + idx] : -2;
            lsc[t] = idx < kk ? knn_s[(size_t)c * kk + idx] : 0.f;
        }
        bool reject = false;
        for (int a = 0; a < ns && !reject; ++a) {
            const int aid = sel[a];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const unsigned bm = __ballot_sync(kFull, li[t] == aid);
                if (bm) {
                    const float sca = __shfl_sync(kFull, lsc[t], __ffs(bm) - 1);
                    if (sca > sc) reject = true;
                }
            }
        }
        if (lane == 0) rej[j] = reject ? 1 : 0;
        if (!reject) {
            if (lane == 0) { sel[ns] = c; sel_s[ns] = sc; }
            ++ns;
        }
        __syncwarp();
    }
    for (int j = 0; j < T && ns < m; ++j) {      // refill with the closest rejected candidates
        if (rej[j]) {
            if (lane == 0) { sel[ns] = cid[j]; sel_s[ns] = cs[j]; }
            ++ns;
        }
    }
    __syncwarp();
    return ns;
}

// forward picks with the diversity heuristic (one warp per node) + reverse-buffer append
__global__ void __launch_bounds__(256)
select_forward_diverse_kernel(const float* __restrict__ knn_s, const int* __restrict__ knn_r, long long n_members, int kk,
                              int m, int rcap, int* __restrict__ fwd, float* __restrict__ fwd_s, int* __restrict__ rev_cnt,
                              int* __restrict__ rev, float* __restrict__ rev_s) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (u >= n_members) return;
    const int per = kk * 9 + m * 8 + 16;                       // cid, cs, rej | sel, sel_s
    unsigned char* base = smem_raw + (size_t)warp * ((per + 15) / 16 * 16);
    int* cid = reinterpret_cast<int*>(base);
    float* cs = reinterpret_cast<float*>(cid + kk);
    int* sel = reinterpret_cast<int*>(cs + kk);
    float* sel_s = reinterpret_cast<float*>(sel + m);
    unsigned char* rej = reinterpret_cast<unsigned char*>(sel_s + m);
    // compact the valid, non-self candidates (order preserved)
    int T = 0;
    for (int b0 = 0; b0 < kk; b0 += 32) {
        const int i = b0 + lane;
        const int v = i < kk ? knn_r[u * kk + i] : -1;
        const bool ok = v >= 0 && v != (int)u;
        const unsigned bm = __ballot_sync(kFull, ok);
        if (ok) { const int at = T + __popc(bm & ((1u << lane) - 1u)); cid[at] = v; cs[at] = knn_s[u * kk + i]; }
        T += __popc(bm);
    }
    __syncwarp();
    const int ns = diverse_select(cid, cs, T, knn_r, knn_s, kk, m, sel, sel_s, rej, lane);
    for (int i = lane; i < m; i += 32) {
        const bool ok = i < ns;
        fwd[u * m + i] = ok ? sel[i] : -1;
        fwd_s[u * m + i] = ok ? sel_s[i] : VQ_NEG_INF;
        if (ok) {
            const int v = sel[i];
            const int slot = atomicAdd(&rev_cnt[v], 1);
            if (slot < rcap) { rev[(size_t)v * rcap + slot] = (int)u; rev_s[(size_t)v * rcap + slot] = sel_s[i]; }
        }
    }
}

// Final adjacency of a node (one thread per node): the first half of its diversity-ordered forward
// picks, then reverse edges (nodes that picked it) for the other half, then whatever is left, never
// exceeding m.  Reverse candidates that hardly anybody links to (forward in-degree < kLowIn: the
// "anti-hubs" of high-dimensional data, which a plain closest-M rule leaves unreachable) are
// served first, closest first; the rest follow in closeness order.
constexpr int kLowIn = 4;
__global__ void __launch_bounds__(128)
merge_fwd_rev_kernel(const int* __restrict__ fwd, const int* __restrict__ rev_cnt, const int* __restrict__ rev,
                     const float* __restrict__ rev_s, long long n_members, int m, int rcap,
                     const int* __restrict__ members, int* __restrict__ adj_out) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_members) return;
    int out[32];
    int no = 0;
    const int half = (m + 1) / 2;
    auto has = [&](int v) { for (int i = 0; i < no; ++i) if (out[i] == v) return true; return false; };
    for (int i = 0; i < half; ++i) { const int v = fwd[u * m + i]; if (v >= 0 && !has(v)) out[no++] = v; }
    int nrev = rev_cnt[u];
    nrev = nrev < rcap ? nrev : rcap;
    const int* rv = rev + (size_t)u * rcap;
    const float* rs = rev_s + (size_t)u * rcap;
    // next reverse candidate after (last_s, last_id) in (score desc, id asc) order, restricted to a class
    auto next_rev = [&](float last_s, int last_id, int want_low, float& bs, int& bid) {
        bs = VQ_NEG_INF; bid = -1;
        for (int j = 0; j < nrev; ++j) {
            const float s = rs[j]; const int id = rv[j];
            if (want_low >= 0 && (int)(rev_cnt[id] < kLowIn) != want_low) continue;
            const bool after_last = (s < last_s) || (s == last_s && id > last_id);
            if (after_last && (bid < 0 || s > bs || (s == bs && id < bid))) { bs = s; bid = id; }
        }
    };
    int taken = 0;
    for (int cls = 1; cls >= 0; --cls) {                 // rarely-linked nodes first, then the others
        float last_s = INFINITY; int last_id = -1;
        while (no < m && taken < m / 2) {
            float bs; int bid;
            next_rev(last_s, last_id, cls, bs, bid);
            if (bid < 0) break;
            last_s = bs; last_id = bid;
            if (!has(bid)) { out[no++] = bid; ++taken; }
        }
    }
    for (int i = half; i < m && no < m; ++i) { const int v = fwd[u * m + i]; if (v >= 0 && !has(v)) out[no++] = v; }
    {   // still room: remaining reverse edges in closeness order
        float last_s = INFINITY; int last_id = -1;
        while (no < m) {
            float bs; int bid;
            next_rev(last_s, last_id, -1, bs, bid);
            if (bid < 0) break;
            last_s = bs; last_id = bid;
            if (!has(bid)) out[no++] = bid;
        }
    }
    for (int i = 0; i < m; ++i) adj_out[u * m + i] = i < no ? (members ? members[out[i]] : out[i]) : -1;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct SearchPlan { int cap, warp_bytes, warps; size_t smem; };
SearchPlan plan_search(int ld, int ef, int visited_capacity) {
    SearchPlan p;
    // visited set: 16 slots per beam entry covers the evaluations of a typical query with room to
    // spare (measured 4-9 evaluations per beam entry); a query that does fill it reports overflow
    // and is re-run by the caller with a larger table, so the common case keeps its occupancy.
    // measured on 1M clustered rows: 17 / 13 / 9 evaluations per beam entry at ef 64 / 128 / 256
    int want = visited_capacity > 0 ? visited_capacity : (ef * 20 < 2048 ? 2048 : ef * 20);
    p.cap = (want + 63) / 64 * 64;
    if (p.cap > 32768) p.cap = 32768;
    p.warp_bytes = (int)align256((size_t)ld * 4 + (size_t)ef * 8 + (size_t)p.cap * 4 + 32 * 8 + kLogCap * 2);
    p.warps = 4;
    while (p.warps > 1 && (size_t)p.warps * p.warp_bytes > 200 * 1024) p.warps >>= 1;
    p.smem = (size_t)p.warps * p.warp_bytes;
    return p;
}

}  // namespace

extern "C" {

size_t vq_hnsw_workspace_bytes(int b, int ld, int ef) {
    (void)ef;
    return align256((size_t)(b > 0 ? b : 1) * ld * 4) + 256;
}

int vq_hnsw_search(const void* store, int64_t n, int dim, int ld, int store_dtype, const int32_t* levels,
                   const int32_t* adj0, int m0, const int32_t* upper_off, const int32_t* upper_adj, int m,
                   int32_t entry, int max_level, int ef, const float* queries, int b, int k, int query_norm,
                   float* out_dist, int32_t* out_rows, uint32_t* out_stats, int visited_capacity, void* workspace,
                   size_t workspace_bytes, void* stream_v) {
    (void)levels;
    cudaStream_t stream = (cudaStream_t)stream_v;
    VQ_CHECK_ARG(store_dtype == VQ_F32 || store_dtype == VQ_BF16, "bad store_dtype %d", store_dtype);
    VQ_CHECK_ARG(n > 0 && n < (1 << 30), "n=%lld out of range (node ids must be < 2^30)", (long long)n);
    VQ_CHECK_ARG(dim > 0 && ld >= dim && ld % (store_dtype == VQ_BF16 ? 64 : 32) == 0, "bad dim/ld %d/%d", dim, ld);
    VQ_CHECK_ARG(m0 > 0 && m > 0 && m0 <= 1024 && m <= 1024, "bad degrees m0=%d m=%d", m0, m);
    VQ_CHECK_ARG(entry >= 0 && entry < n && max_level >= 0, "bad entry/max_level %d/%d", entry, max_level);
    VQ_CHECK_ARG(b >= 0 && k > 0, "bad b/k %d/%d", b, k);
    VQ_CHECK_ARG(query_norm >= VQ_NORM_NONE && query_norm <= VQ_NORM_PLAIN, "bad query_norm %d", query_norm);
    if (b == 0) return VQ_OK;
    VQ_CHECK_ARG(store && adj0 && queries && out_dist && out_rows && workspace, "NULL pointer argument");
    VQ_CHECK_ARG(max_level == 0 || (upper_off && upper_adj), "upper layers missing");
    if (ef < k) ef = k;                                       // hnsw.py:264  ef = max(ef_search, k)
    VQ_CHECK_ARG(ef <= 4096, "ef=%d too large (max 4096)", ef);
    if (workspace_bytes < vq_hnsw_workspace_bytes(b, ld, ef)) {
        vq_set_error("hnsw workspace too small");
        return VQ_EWORKSPACE;
    }
    float* qn = (float*)workspace;
    int rc = vq_ingest_launch(queries, b, dim, dim, qn, VQ_F32, ld, query_norm, stream);
    if (rc) return rc;
    const SearchPlan p = plan_search(ld, ef, visited_capacity);
    if ((size_t)p.warp_bytes > 200 * 1024) {
        vq_set_error("hnsw_search: per-query shared memory %d B too large (ld=%d ef=%d)", p.warp_bytes, ld, ef);
        return VQ_EUNSUPPORTED;
    }
    const int grid = (b + p.warps - 1) / p.warps;
    vq_prof_begin(stream);
    // result list in registers (32*S entries striped over the warp) up to ef = 256, in shared memory beyond
    static const bool reg_list = getenv("VQ_HNSW_REGLIST") ? atoi(getenv("VQ_HNSW_REGLIST")) != 0 : true;
    const int slots = !reg_list ? 0 : ef <= 64 ? 2 : ef <= 128 ? 4 : ef <= 256 ? 8 : 0;
#define VQ_HNSW_LAUNCH(BF, SL)                                                                                          \
    do {                                                                                                                \
        VQ_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<BF, SL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        hnsw_search_kernel<BF, SL><<<grid, p.warps * 32, p.smem, stream>>>(store, ld, adj0, m0, upper_off, upper_adj, m, entry, \
                                                                          max_level, ef, p.cap, qn, b, k, out_dist, out_rows, \
                                                                          out_stats, p.warp_bytes);                     \
    } while (0)
    if (store_dtype == VQ_BF16) {
        if (slots == 2) VQ_HNSW_LAUNCH(true, 2); else if (slots == 4) VQ_HNSW_LAUNCH(true, 4);
        else if (slots == 8) VQ_HNSW_LAUNCH(true, 8); else VQ_HNSW_LAUNCH(true, 0);
    } else {
        if (slots == 2) VQ_HNSW_LAUNCH(false, 2); else if (slots == 4) VQ_HNSW_LAUNCH(false, 4);
        else if (slots == 8) VQ_HNSW_LAUNCH(false, 8); else VQ_HNSW_LAUNCH(false, 0);
    }
#undef VQ_HNSW_LAUNCH
    vq_prof_end(stream);
    VQ_LAUNCH_CHECK("hnsw_search_kernel");
    vq_note_launch(store_dtype == VQ_BF16 ? "hnsw_search_bf16" : "hnsw_search_f32", 2);
    return VQ_OK;
}

// ---- "incremental" construction (diversify == 2): the reference's insertion ORDER without its per-insert search.
// hnsw.py:150-229 inserts node i into the graph of nodes 0..i-1: it links i to the M closest nodes it finds among
// them (:183-199, ef_construction-wide beam) and every touched neighbour keeps the closest max_conn of everything
// that was ever linked to it (:202-223).  Early nodes therefore start with long links and only lose them when closer
// nodes pick them later — the small-world structure an exact k-nearest graph of ALL nodes does not have (measured at
// 1M clustered rows: recall@10 0.64 / 0.73 / 0.79 for the exact 16-nearest graph vs 0.68 / 0.81 / 0.89 for the reference
// at ef 64 / 128 / 256).  Here: (1) the M nearest EARLIER nodes of every node, exactly, by the tensor-core scan over rows
// [0, batch start) — half the work of the all-pairs scan; (2) every pick becomes an edge in both directions;
// (3) every node keeps the closest M of its edges: one radix sort of the (node, score) keys + one selection kernel.
__global__ void causal_edges_kernel(const float* __restrict__ knn_s, const int* __restrict__ knn_r, long long n_members, int kk, int m,
                                    unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_members * m) return;
    const long long u = t / m;
    const int j = (int)(t - u * m);
    // the j-th pick of u that is neither empty nor u itself (the first batch scans itself: u is its own nearest)
    int taken = 0, v = -1;
    float sc = 0.f;
    for (int c = 0; c < kk; ++c) {
        const int w = knn_r[u * kk + c];
        if (w < 0 || w == (int)u) continue;
        if (taken == j) { v = w; sc = knn_s[u * kk + c]; break; }
        ++taken;
    }
    const unsigned long long none = ~0ull;
    if (v < 0) { keys[2 * t] = none; keys[2 * t + 1] = none; vals[2 * t] = -1; vals[2 * t + 1] = -1; return; }
    const unsigned long long sk = vq_score_key(sc);                     // larger score -> smaller key
    keys[2 * t] = ((unsigned long long)(unsigned)u << 32) | sk;          // u -> v
    vals[2 * t] = v;
    keys[2 * t + 1] = ((unsigned long long)(unsigned)v << 32) | sk;      // v -> u
    vals[2 * t + 1] = (int)u;
}

// one warp per node: the first m distinct neighbours of its (score-sorted) segment
__global__ void __launch_bounds__(256)
causal_select_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long n_items, long long n_members,
                     int m, const int* __restrict__ members, int* __restrict__ adj_out) {
    const int lane = threadIdx.x & 31;
    const long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n_members) return;
    // first item whose node is >= u
    long long lo = 0, hi = n_items;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((keys[mid] >> 32) < (unsigned long long)u) lo = mid + 1; else hi = mid;
    }
    int cnt = 0;
    for (long long base = lo; base < n_items && cnt < m; base += 32) {
        const long long i = base + lane;
        const bool mine = i < n_items && (keys[i] >> 32) == (unsigned long long)u;
        const int w = mine ? vals[i] : -1;
        const unsigned in_seg = __ballot_sync(kFull, mine);
        // duplicates (an edge picked from both ends inside the first batch): keep the first occurrence
        bool dup = false;
        for (int l = 0; l < 32; ++l) {
            const int o = __shfl_sync(kFull, w, l);
            if (l < lane && o == w && w >= 0) dup = true;
        }
        for (int c = 0; c < cnt; ++c) if (adj_out[u * m + c] == (members ? members[w < 0 ? 0 : w] : w) && w >= 0) dup = true;
        const unsigned keep = __ballot_sync(kFull, mine && w >= 0 && !dup);
        const int pos = cnt + __popc(keep & ((1u << lane) - 1u));
        if (((keep >> lane) & 1u) && pos < m) adj_out[u * m + pos] = members ? members[w] : w;
        cnt += __popc(keep);
        __syncwarp();
        if (in_seg != kFull) break;                                       // the segment ended inside this chunk
    }
    for (int c = (cnt < m ? cnt : m) + lane; c < m; c += 32) adj_out[u * m + c] = -1;
}

// ---- "sequential" construction (diversify == 3): the reference's add() with EXACT candidates, batch by batch.
// hnsw.py:183-223: node i links to its M closest candidates among the nodes already in the graph; every link is added
// in both directions; a neighbour whose list exceeds max_conn keeps its closest max_conn and the dropped link is removed
// at BOTH ends (:221-223) — which leaves most lists short of max_conn (mean degree 8.5 of 16 at 1M rows), so the long
// links a node made when the graph was still sparse survive: the small-world structure that makes the reference's layer 0
// more navigable than an exact 16-nearest graph (measured at 1M clustered rows with the same upper layers: recall@10
// 0.692 / 0.824 / 0.901 on the reference's layer 0 vs 0.612 / 0.757 / 0.870 on the diversity-pruned exact one).
// The candidates of node i do not depend on the graph (exact M nearest among the nodes before i's batch, by the tensor-core
// scan), so only the link / prune bookkeeping is sequential: per batch of 2048 nodes (1) every new node writes its own
// list and emits one request per pick, (2) the requests are sorted by target and ONE thread per target merges them into
// the target's list (closest max_conn of list + requests), emitting a removal for every dropped link, (3) the removals
// are applied at the other end.  Within a batch this is the reference's rule in a different order; across batches it is
// the reference's order.
__global__ void seq_emit_kernel(const float* __restrict__ knn_s, const int* __restrict__ knn_r, int q0, int bq, int kk, int m, int first,
                                int* __restrict__ adj, float* __restrict__ adj_s, unsigned long long* __restrict__ req_key,
                                int* __restrict__ req_src) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= bq) return;
    const int i = q0 + t;
    int taken = 0;
    for (int c = 0; c < kk && taken < m; ++c) {
        const int w = knn_r[(size_t)i * kk + c];
        if (w < 0 || w == i) continue;
        if (first && w >= i) continue;                       // first batch scans itself: keep the insertion order inside it
        const float sc = knn_s[(size_t)i * kk + c];
        adj[(size_t)i * m + taken] = w;
        adj_s[(size_t)i * m + taken] = sc;
        req_key[(size_t)t * m + taken] = ((unsigned long long)(unsigned)w << 32) | vq_score_key(sc);
        req_src[(size_t)t * m + taken] = i;
        ++taken;
    }
    for (int c = taken; c < m; ++c) {
        adj[(size_t)i * m + c] = -1;
        req_key[(size_t)t * m + c] = ~0ull;
        req_src[(size_t)t * m + c] = -1;
    }
}

__global__ void seq_merge_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ src, int n_items, int m,
                                 int* __restrict__ adj, float* __restrict__ adj_s, int2* __restrict__ rm_pairs, int* __restrict__ rm_cnt) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items) return;
    const unsigned long long k0 = key[idx];
    if (k0 == ~0ull) return;
    const int j = (int)(k0 >> 32);
    if (idx > 0 && (int)(key[idx - 1] >> 32) == j) return;       // not the head of its target's segment
    int ids[25];
    float sc[25];
    int cnt = 0;
    for (int t = 0; t < m; ++t) {                                // current list of j, holes squeezed out
        const int v = adj[(size_t)j * m + t];
        if (v >= 0) { ids[cnt] = v; sc[cnt] = adj_s[(size_t)j * m + t]; ++cnt; }
    }
    for (int p = idx; p < n_items && key[p] != ~0ull && (int)(key[p] >> 32) == j; ++p) {
        const int i = src[p];
        const float s = vq_key_score((unsigned)key[p]);
        if (cnt < m) { ids[cnt] = i; sc[cnt] = s; ++cnt; continue; }
        int w = 0;                                               // worst kept: lowest score, then highest id (hnsw.py:138 sorts (distance, id))
        for (int t = 1; t < m; ++t)
            if (sc[t] < sc[w] || (sc[t] == sc[w] && ids[t] > ids[w])) w = t;
        int drop = i;
        if (s > sc[w] || (s == sc[w] && i < ids[w])) { drop = ids[w]; ids[w] = i; sc[w] = s; }
        rm_pairs[atomicAdd(rm_cnt, 1)] = make_int2(drop, j);     // the dropped node loses its link to j as well (:221-223)
    }
    for (int t = 0; t < m; ++t) {
        adj[(size_t)j * m + t] = t < cnt ? ids[t] : -1;
        adj_s[(size_t)j * m + t] = t < cnt ? sc[t] : VQ_NEG_INF;
    }
}

__global__ void seq_remove_kernel(const int2* __restrict__ rm_pairs, const int* __restrict__ rm_cnt, int m, int* __restrict__ adj) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= *rm_cnt) return;
    const int2 pr = rm_pairs[idx];
    for (int t = 0; t < m; ++t)
        if (adj[(size_t)pr.x * m + t] == pr.y) adj[(size_t)pr.x * m + t] = -1;
}

// lists squeezed to the front and translated from member indices to node ids
__global__ void seq_finish_kernel(int* __restrict__ adj, long long n_members, int m, const int* __restrict__ members) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_members) return;
    int out[25];
    int cnt = 0;
    for (int t = 0; t < m; ++t) {
        const int v = adj[u * m + t];
        if (v >= 0) out[cnt++] = members ? members[v] : v;
    }
    for (int t = 0; t < m; ++t) adj[u * m + t] = t < cnt ? out[t] : -1;
}

// workspace layout of one layer build
constexpr int kBuildQB = 2048;      // queries per tensor-core k-nearest pass (16 query tiles)
struct BuildPlan {
    int kk, rcap, grid;
    size_t compact, knn_s, knn_r, part_s, part_r, fwd, fwd_s, rev_cnt, rev, rev_s, qpad, mma_ws, mma_ws_bytes, total;
    size_t sort_k0, sort_k1, sort_v0, sort_v1, sort_tmp, sort_tmp_bytes;      // incremental mode (k_cand == m_out)
};
static BuildPlan plan_build(int64_t n_members, int ld, int k_cand, int m_out, bool need_compact, int elem = 4) {
    BuildPlan p;
    p.kk = (k_cand > m_out ? k_cand : m_out) + 1;           // +1: the node itself is its own nearest
    p.rcap = 8 * m_out;
    p.grid = vq_scan_fma_grid((int)n_members, 16);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    p.compact = take(need_compact ? (size_t)n_members * ld * elem : 0);
    p.knn_s = take((size_t)n_members * p.kk * 4);
    p.knn_r = take((size_t)n_members * p.kk * 4);
    p.part_s = take((size_t)p.grid * 16 * p.kk * 4);
    p.part_r = take((size_t)p.grid * 16 * p.kk * 4);
    p.fwd = take((size_t)n_members * m_out * 4);
    p.fwd_s = take((size_t)n_members * m_out * 4);
    p.rev_cnt = take((size_t)n_members * 4);
    p.rev = take((size_t)n_members * p.rcap * 4);
    p.rev_s = take((size_t)n_members * p.rcap * 4);
    p.qpad = take(elem == 2 ? (size_t)kBuildQB * ld * 2 : 0);
    p.mma_ws_bytes = elem == 2 ? vq_scan_mma_prepared_workspace(n_members, ld, kBuildQB, p.kk) : 0;
    p.mma_ws = take(p.mma_ws_bytes);
    p.sort_tmp_bytes = 0;
    p.sort_k0 = p.sort_k1 = p.sort_v0 = p.sort_v1 = p.sort_tmp = 0;
    if (k_cand == m_out) {                       // closest / incremental selection: room for the edge sort
        const size_t items = (size_t)n_members * m_out * 2;
        cub::DeviceRadixSort::SortPairs(nullptr, p.sort_tmp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                        (const int*)nullptr, (int*)nullptr, (long long)items, 0, 64);
        p.sort_k0 = take(items * 8);
        p.sort_k1 = take(items * 8);
        p.sort_v0 = take(items * 4);
        p.sort_v1 = take(items * 4);
        p.sort_tmp = take(p.sort_tmp_bytes);
    }
    p.total = off + 256;
    return p;
}

// one batch of the sequential construction: own lists + requests, sort by target, merge, removals
static int seq_link_batch(unsigned char* ws, const BuildPlan& p, const float* knn_s, const int* knn_r, int q0, int bq, int m,
                          int* adj, cudaStream_t stream) {
    float* adj_s = (float*)(ws + p.fwd_s);
    unsigned long long* k0 = (unsigned long long*)(ws + p.sort_k0);
    unsigned long long* k1 = (unsigned long long*)(ws + p.sort_k1);
    int* v0 = (int*)(ws + p.sort_v0);
    int* v1 = (int*)(ws + p.sort_v1);
    int2* rm = (int2*)(ws + p.rev);
    int* rm_cnt = (int*)(ws + p.rev_cnt);
    const int items = bq * m;
    seq_emit_kernel<<<(bq + 127) / 128, 128, 0, stream>>>(knn_s, knn_r, q0, bq, m + 1, m, q0 == 0 ? 1 : 0, adj, adj_s, k0, v0);
    VQ_LAUNCH_CHECK("seq_emit_kernel");
    size_t tmp_bytes = p.sort_tmp_bytes;
    VQ_CUDA(cub::DeviceRadixSort::SortPairs(ws + p.sort_tmp, tmp_bytes, k0, k1, v0, v1, items, 0, 64, stream));
    VQ_CUDA(cudaMemsetAsync(rm_cnt, 0, 4, stream));
    seq_merge_kernel<<<(items + 127) / 128, 128, 0, stream>>>(k1, v1, items, m, adj, adj_s, rm, rm_cnt);
    VQ_LAUNCH_CHECK("seq_merge_kernel");
    seq_remove_kernel<<<(items + 127) / 128, 128, 0, stream>>>(rm, rm_cnt, m, adj);
    VQ_LAUNCH_CHECK("seq_remove_kernel");
    return VQ_OK;
}

size_t vq_hnsw_layer_workspace_bytes(int64_t n_members, int dim, int ld, int store_dtype, int k_cand, int m_out) {
    (void)dim;
    if (n_members <= 0) return 256;
    return plan_build(n_members, ld, k_cand, m_out, true, store_dtype == VQ_BF16 ? 2 : 4).total;
}

int vq_hnsw_build_layer(const void* store, int64_t n, int dim, int ld, int store_dtype, const int32_t* members,
                        int64_t n_members, int k_cand, int m_out, int diversify, int32_t* adj_out, void* workspace,
                        size_t workspace_bytes, void* stream_v) {
    cudaStream_t stream = (cudaStream_t)stream_v;
    VQ_CHECK_ARG(store_dtype == VQ_F32 || store_dtype == VQ_BF16, "bad store dtype %d", store_dtype);
    const bool bf = store_dtype == VQ_BF16;
    VQ_CHECK_ARG(n > 0 && n < (1 << 30) && dim > 0 && ld >= dim && ld % (bf ? 64 : 32) == 0, "bad shape n=%lld dim=%d ld=%d", (long long)n, dim, ld);
    VQ_CHECK_ARG(n_members >= 0 && n_members <= n, "bad n_members %lld", (long long)n_members);
    VQ_CHECK_ARG(m_out > 0 && m_out <= 25 && k_cand > 0 && k_cand <= 512, "bad m_out/k_cand %d/%d", m_out, k_cand);
    VQ_CHECK_ARG(diversify >= 0 && diversify <= 3, "diversify must be 0 (closest), 1 (diversity heuristic), 2 (incremental) or 3 (sequential), got %d", diversify);
    VQ_CHECK_ARG(diversify != 1 || k_cand + 1 <= 96, "diversify needs k_cand <= 95 (got %d)", k_cand);
    VQ_CHECK_ARG(diversify < 2 || k_cand == m_out, "incremental / sequential construction takes k_cand == m_out (got %d / %d)", k_cand, m_out);
    const bool causal = diversify >= 2;
    const bool sequential = diversify == 3;
    if (n_members == 0) return VQ_OK;
    VQ_CHECK_ARG(store && adj_out && workspace, "NULL pointer argument");
    VQ_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
    const bool compact = members != nullptr;
    const BuildPlan p = plan_build(n_members, ld, k_cand, m_out, true, bf ? 2 : 4);
    // tensor-core k-nearest pass needs the bf16 store and lists that fit the register top-k
    const bool use_mma = bf && vq_scan_mma_supported(n_members, dim, ld, VQ_BF16, kBuildQB, p.kk);
    if (bf && !use_mma) {
        vq_set_error("hnsw_build_layer: bf16 store needs ld <= 768 and k_cand <= 63 (ld=%d k_cand=%d)", ld, k_cand);
        return VQ_EUNSUPPORTED;
    }
    if (workspace_bytes < p.total) {
        vq_set_error("hnsw build workspace too small: %zu < %zu", workspace_bytes, p.total);
        return VQ_EWORKSPACE;
    }
    unsigned char* ws = (unsigned char*)workspace;
    const float* mat = (const float*)store;
    int launches = 0;
    if (compact) {
        const int row_vec = ld * (bf ? 2 : 4) / 16;
        const long long items = (long long)n_members * row_vec;
        gather_rows_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>((const uint4*)store, members, n_members,
                                                                              row_vec, (uint4*)(ws + p.compact));
        VQ_LAUNCH_CHECK("gather_rows_kernel");
        mat = (const float*)(ws + p.compact);
        ++launches;
    }
    float* knn_s = (float*)(ws + p.knn_s);
    int* knn_r = (int*)(ws + p.knn_r);
    // exact k-nearest members of every member: the rows themselves are the (already unit-norm,
    // zero-padded) query tiles, 16 per pass of the fused scan + top-k kernel.
    const int nm = (int)n_members;
    if (use_mma) {
        // bf16 rows are their own unit-norm, zero-padded query tiles: 2048 queries per tensor-core pass
        const unsigned char* matb = (const unsigned char*)mat;
        for (int q0 = 0; q0 < nm; q0 += kBuildQB) {
            const int bq = nm - q0 < kBuildQB ? nm - q0 : kBuildQB;
            const void* qptr = matb + (size_t)q0 * ld * 2;
            if (bq % 128 != 0) {                               // ragged tail: copy into a zero-padded tile
                VQ_CUDA(cudaMemsetAsync(ws + p.qpad, 0, (size_t)kBuildQB * ld * 2, stream));
                VQ_CUDA(cudaMemcpyAsync(ws + p.qpad, qptr, (size_t)bq * ld * 2, cudaMemcpyDeviceToDevice, stream));
                qptr = ws + p.qpad;
            }
            // incremental construction: node i only sees the nodes inserted before its batch (the first batch sees itself)
            const int n_scan = causal ? (q0 == 0 ? bq : q0) : nm;
            int rc = vq_scan_mma_prepared(mat, n_scan, ld, qptr, bq, p.kk, knn_s + (size_t)q0 * p.kk, knn_r + (size_t)q0 * p.kk,
                                          ws + p.mma_ws, p.mma_ws_bytes, stream);
            if (rc) return rc;
            launches += 3;
            if (sequential) {
                rc = seq_link_batch(ws, p, knn_s, knn_r, q0, bq, m_out, adj_out, stream);
                if (rc) return rc;
                launches += 5;
            }
        }
    } else
    for (int q0 = 0; q0 < nm;) {
        int bt, start;
        if (nm - q0 >= 16) { bt = 16; start = q0; }
        else if (nm >= 16) { bt = 16; start = nm - 16; }      // last tile shifted back (recomputes a few rows)
        else { bt = 1; start = q0; }                           // tiny top layers: one query per pass
        const int n_scan = causal ? (start < 16 ? (nm < 16 ? nm : 16) : start) : nm;
        const int grid = vq_scan_fma_grid(n_scan, bt);
        int rc = vq_scan_fma_launch(mat, n_scan, ld, VQ_F32, mat + (size_t)start * ld, bt, p.kk, (float*)(ws + p.part_s),
                                    (int*)(ws + p.part_r), grid, stream);
        if (rc) return rc;
        rc = vq_topk_merge_launch((float*)(ws + p.part_s), (int*)(ws + p.part_r), grid, (long long)bt * p.kk, bt, p.kk, nullptr,
                                  p.kk, knn_s + (size_t)start * p.kk, knn_r + (size_t)start * p.kk, 0, 0, stream);
        if (rc) return rc;
        launches += 2;
        if (sequential && start + bt > q0) {                  // fp32 path: tiles of 16 nodes (a shifted last tile re-links nothing new)
            rc = seq_link_batch(ws, p, knn_s, knn_r, q0, start + bt - q0, m_out, adj_out, stream);
            if (rc) return rc;
            launches += 5;
        }
        q0 = start + bt;
    }
    if (sequential) {
        seq_finish_kernel<<<(unsigned)((n_members + 255) / 256), 256, 0, stream>>>(adj_out, n_members, m_out, members);
        VQ_LAUNCH_CHECK("seq_finish_kernel");
        vq_note_launch("hnsw_build_layer<sequential>", launches + 1);
        return VQ_OK;
    }
    if (causal) {
        const long long items = (long long)n_members * m_out * 2;
        unsigned long long* k0 = (unsigned long long*)(ws + p.sort_k0);
        unsigned long long* k1 = (unsigned long long*)(ws + p.sort_k1);
        int* v0 = (int*)(ws + p.sort_v0);
        int* v1 = (int*)(ws + p.sort_v1);
        causal_edges_kernel<<<(unsigned)((n_members * m_out + 255) / 256), 256, 0, stream>>>(knn_s, knn_r, n_members, p.kk, m_out, k0, v0);
        VQ_LAUNCH_CHECK("causal_edges_kernel");
        size_t tmp_bytes = p.sort_tmp_bytes;
        VQ_CUDA(cub::DeviceRadixSort::SortPairs(ws + p.sort_tmp, tmp_bytes, k0, k1, v0, v1, items, 0, 64, stream));
        causal_select_kernel<<<(unsigned)((n_members + 7) / 8), 256, 0, stream>>>(k1, v1, items, n_members, m_out, members, adj_out);
        VQ_LAUNCH_CHECK("causal_select_kernel");
        vq_note_launch("hnsw_build_layer<incremental>", launches + 3);
        return VQ_OK;
    }
    VQ_CUDA(cudaMemsetAsync(ws + p.rev_cnt, 0, (size_t)n_members * 4, stream));
    if (diversify) {
        const size_t per_f = ((size_t)p.kk * 9 + m_out * 8 + 16 + 15) / 16 * 16;
        select_forward_diverse_kernel<<<(unsigned)((n_members + 7) / 8), 256, 8 * per_f, stream>>>(
            knn_s, knn_r, n_members, p.kk, m_out, p.rcap, (int*)(ws + p.fwd), (float*)(ws + p.fwd_s),
            (int*)(ws + p.rev_cnt), (int*)(ws + p.rev), (float*)(ws + p.rev_s));
        VQ_LAUNCH_CHECK("select_forward_diverse_kernel");
        merge_fwd_rev_kernel<<<(unsigned)((n_members + 127) / 128), 128, 0, stream>>>(
            (int*)(ws + p.fwd), (int*)(ws + p.rev_cnt), (int*)(ws + p.rev), (float*)(ws + p.rev_s), n_members, m_out,
            p.rcap, members, adj_out);
        VQ_LAUNCH_CHECK("merge_fwd_rev_kernel");
    } else {
        link_reverse_kernel<<<(unsigned)((n_members + 255) / 256), 256, 0, stream>>>(
            knn_s, knn_r, n_members, p.kk, m_out, p.rcap, (int*)(ws + p.fwd), (float*)(ws + p.fwd_s), (int*)(ws + p.rev_cnt),
            (int*)(ws + p.rev), (float*)(ws + p.rev_s));
        VQ_LAUNCH_CHECK("link_reverse_kernel");
        const size_t psmem = (size_t)8 * (m_out + p.rcap) * 8;
        prune_kernel<<<(unsigned)((n_members + 7) / 8), 256, psmem, stream>>>(
            (int*)(ws + p.fwd), (float*)(ws + p.fwd_s), (int*)(ws + p.rev_cnt), (int*)(ws + p.rev), (float*)(ws + p.rev_s),
            n_members, m_out, p.rcap, members, adj_out);
        VQ_LAUNCH_CHECK("prune_kernel");
    }
    vq_note_launch("hnsw_build_layer", launches + 2);
    return VQ_OK;
}

}  // extern "C"
