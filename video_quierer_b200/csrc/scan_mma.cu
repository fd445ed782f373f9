// (b)+(c) dense-batch exact scan on the 5th-gen tensor cores: TMA -> shared memory ->
// tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld -> fused per-query top-k.
//
// Replaces the reference's per-query np.dot + argsort (video_search_overhaul.py:53,56 looped by
// src/api/routes.py:627-634) when the query batch makes the scan a dense contraction.
//
// Orientation: D[query, row] = Q[query, :] . S[row, :]  with  A = 128 queries (UMMA M = 128) and
// B = NT store rows (UMMA N = NT, K-major, 128-byte swizzle).  The query tile never changes while
// a CTA lives, so it is written ONCE into tensor memory (tcgen05.st, packed bf16 pairs, lane =
// query) and every MMA takes its A operand from TMEM: shared memory carries only the streamed
// store tiles (deep TMA ring, half the smem operand traffic per MMA).  After tcgen05.ld every
// epilogue thread owns ONE query (its TMEM lane) and sees the scores of 32 store rows at a time in
// registers: the running k-th best and the top-k list are registers of that thread, so there is
// no cross-thread traffic and the scores never leave the SM.  The k-th best of every query is also
// shared between CTAs through a global array (atomic max), so a row is dropped as soon as ANY CTA
// has k better ones — insertions fall from ~k*ln(rows per CTA) per CTA to ~k*ln(N) in total.
//
// TMEM (512 columns): [0, 2*NT) two fp32 accumulators (MMA of tile i+1 overlaps the epilogue of
// tile i), [2*NT, 2*NT + ld/2) the resident query tile.
// Warp roles (256 or 384 threads, 1 CTA / SM):
//   warp 0    TMA producer : store tiles through a ring that is filled, consumed and released in GROUPS of
//                            k-blocks (64 KB at dim 512): one expect_tx + the group's loads per group
//   warps 1,2 MMA issuers  : own alternate tiles (issuer r: accumulator r); per group ONE barrier wait,
//                            4 * gs back-to-back tcgen05.mma (K = 16) and one tcgen05.commit; a `turn`
//                            barrier keeps the tiles in issue order (tile t completes while t+1 runs)
//   warp 2    also the TMEM allocator; warp 3: bound refresher (exact mode)
//   warps 4-7 (4-11 in the exact mode above 128 queries) epilogue: load the query tile into TMEM, then per
//                            tile tcgen05.ld 32x32b.x32, accumulator release, threshold filter, insertion
// Grid: persistent, gridDim = groups * n_qt; CTA c serves query tile c % n_qt and the store tiles
// c / n_qt, + groups, ... .
// CTA pairs (exact mode, even number of query tiles; template flag CG2): the two CTAs of a cluster hold
// neighbouring query tiles and issue ONE tcgen05.mma.cta_group::2 (M = 256) per step — each keeps its own
// queries in its TMEM and HALF of every store tile in its shared memory (see the helpers below).
//
// Measured design rules (tools/mmabench.cu, tools/issue_rate.sh, tools/scan_trace.sh; B200): a lone
// tcgen05.mma M128 N128 K16 retires every 64 cycles.  A barrier wait + fence + election + commit per
// k-block made the ISSUE LOOP the bound (~100 cycles per MMA even with loads and MMAs switched off);
// with one wait and one commit per group and the hand-offs described in the kernel the traced steady
// state is one 128 x 128 x 512 tile per 2045 cycles = 64.0 cycles per MMA.
//
// Threshold bootstrap: the per-query insertion cost is ~k*(1+ln(rows per CTA / k)) serial list
// insertions per CTA.  For large stores a BOOT pass of the same kernel first scores a sample of
// `boot_tiles` store tiles and writes only each tile's per-query MAXIMUM; the k-th largest of those
// maxima (boot_select_kernel) is a valid lower bound of the global k-th best score (k distinct rows
// reach it), so the main pass starts with that bound in `gtau` and inserts ~10x fewer rows.
// Queries are dealt round-robin over the four epilogue warps (query i of a tile -> TMEM lane
// (i%4)*32 + i/4) so that a batch of 32 keeps all four warps' schedulers busy instead of one.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "vq_common.cuh"

int vq_topk_merge_launch(const float* scores, const int* rows, int g, long long g_stride, int b_out, int k_in,
                         const long long* offsets, int k_out, float* out_scores, void* out_rows,
                         int rows64, int negate_out, cudaStream_t stream);
int vq_ingest_launch(const float* src, long long rows, int dim, int src_ld, void* dst, int dst_dtype,
                     int dst_ld, int mode, cudaStream_t stream);

namespace {

constexpr int QT = 128;                 // queries per tile   (UMMA M)
constexpr int KB_ELEMS = 64;            // bf16 per k-block = one 128-byte swizzle span
constexpr int TMEM_COLS = 512;          // whole tensor memory: 2 accumulators + resident query tile

// Epilogue warps: 4 (one per quarter of the 128 TMEM lanes) or, in the exact mode with lists of <= 32 entries, 8 — warps w and
// w + 4 serve the same 32 queries and split the columns (store rows) of every tile, halving the per-tile critical path of the
// epilogue, which is as long as the tile's MMAs as soon as the gather slow path runs (12 warps x 168 registers still fit).
[[maybe_unused]] constexpr int epi_warps(int kl, int mode) { return (mode == 3 && kl <= 32) ? 8 : 4; }
constexpr int kThreadsFor(int epi) { return (4 + epi) * 32; }
constexpr int kMaxK = 64;
constexpr int kModeList = 0, kModeBoot = 1, kModeCollect = 2, kModeExact = 3;   // epilogue of scan_mma_bf16_kernel
constexpr int kStage = 8;               // collect mode: candidates a thread stages in shared memory per global atomic
constexpr int kMaxBootTiles = 256;     // sample tiles of the threshold bootstrap (8 per lane in boot_select)

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
__device__ __forceinline__ void tma_load_2d_addr(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// multicast variants for a 2-CTA cluster that shares the store tiles (each CTA loads half a box and
// writes it into both CTAs' shared memory; a released ring group is signalled to both CTAs)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// ---- cta_group::2: the CTA pair of a cluster issues ONE M = 256 MMA per step.  Each CTA holds its own 128 queries in tensor
// memory and HALF of the store tile (64 rows) in shared memory, so a store tile crosses the L2 -> SM fabric once per 256
// queries instead of once per 128 (the fabric, ~6300 B/clk chip-wide, is what bounds the 1-CTA kernel: see DESIGN.md).
__device__ __forceinline__ uint32_t mapa_rank(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope), as for a local arrive: what the waiting issuer depends on — the tcgen05.ld /
    // TMA completions before this arrive — is ordered by the tcgen05 fences and the barrier itself.  A .release.cluster
    // here made every arrive wait for the thread's outstanding global stores and atomics (~1000s of cycles per tile).
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {     // arrives on the same barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_cg2(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one lane of the (converged) warp; always the same lane, so tcgen05.commit sees the MMAs it tracks
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
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
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = nt
__device__ __forceinline__ uint32_t umma_idesc_bf16(int nt, int m = QT) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]   (A operand resident in tensor memory)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
// fire-and-forget maximum (no value returned: the warp never waits for the L2 round trip)
__device__ __forceinline__ void red_max_float(float* addr, float v) {
    if (v >= 0.f) asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" ::"l"(addr), "r"(__float_as_int(v)) : "memory");
    else asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(addr), "r"(__float_as_uint(v)) : "memory");
}
// Bound of the NEXT tile, loaded a whole tile ahead.  `after` is a value the caller has just computed from the previous bound:
// naming it as an operand keeps the load behind the last use of the register the result is carried in — with a plain volatile
// load ptxas hoisted the load, had to copy the result into the loop-carried register right away, and every tile stalled
// on the L2 round trip it was meant to hide (~430 cycles).
__device__ __forceinline__ float ld_bound_after(const float* p, float after) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p), "f"(after) : "memory");
    return v;
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
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

// tcgen05.ld split into issue + wait so that the next chunk's load overlaps the filter of this one.
// The wait names the destination registers as in/out operands: their uses cannot be scheduled above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
          "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
          "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
          "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
        :: "memory");
}

// Filter 32 fresh scores (store rows row_base .. row_base+31, the first `left` of them real) against
// the running threshold and insert the survivors into the thread's register top-k list.
template <int KL>
__device__ __forceinline__ void filter_insert(const uint32_t (&v)[32], int row_base, int left, float g_keep, float& thr,
                                              float (&ls)[KL], int (&lr)[KL]) {
    // Once the lists are warm almost no chunk holds a candidate: one max-reduction (3-input max, ~16
    // instructions) decides that, the per-score mask (64 instructions) is built only for the rare chunk
    // that passes.  A stale threshold only lets more through; every candidate is re-checked below.
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int j = 1; j + 1 < 32; j += 2) m = fmaxf(fmaxf(m, __uint_as_float(v[j])), __uint_as_float(v[j + 1]));
    m = fmaxf(m, __uint_as_float(v[31]));
    if (!(m > thr)) return;
    unsigned mask = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
    if (left < 32) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        float s = __uint_as_float(v[0]);
#pragma unroll
        for (int jj = 1; jj < 32; ++jj) s = (j == jj) ? __uint_as_float(v[jj]) : s;
        if (s > thr) {
            // replace the current k-th best, then one bubble pass restores the order.  Rows arrive in
            // ascending order inside a CTA, so with the strict comparison equal scores end up in
            // ascending-row order (the engine's tie rule).
            ls[0] = s;
            lr[0] = row_base + j;
#pragma unroll
            for (int i = 0; i + 1 < KL; ++i) {
                const bool sw = ls[i] > ls[i + 1];
                const float a0 = ls[i], a1 = ls[i + 1];
                const int r0 = lr[i], r1 = lr[i + 1];
                ls[i] = sw ? a1 : a0;
                ls[i + 1] = sw ? a0 : a1;
                lr[i] = sw ? r1 : r0;
                lr[i + 1] = sw ? r0 : r1;
            }
            thr = fmaxf(ls[0], g_keep);
        }
    }
}

// Exact mode: the register list only tracks the running k-th best bf16-operand score (thr); EVERY row whose
// score reaches thr_c = thr - 2*eps is appended to the query's candidate buffer (staged in shared memory,
// `flush` publishes kStage of them with one atomicAdd).  Why 2*eps: let S_k be the final k-th best bf16
// score.  The k rows that reach it have exact score >= S_k - eps, so the exact k-th best s_k >= S_k - eps;
// a row of the exact top-k has exact score >= s_k, i.e. bf16 score >= s_k - eps >= S_k - 2*eps >= thr_c at
// any time (thr only grows towards S_k).  The collected set therefore contains the exact top-k.
template <int KL, bool LIST, typename Flush>
__device__ __forceinline__ float filter_collect(const uint32_t (&v)[32], int row_base, int left, float g_keep, const float& eps2,
                                               float& thr, float& thr_c, float (&ls)[KL], int (&lr)[KL],
                                               float* stage_s, int* stage_r, int sstride, int& staged, Flush&& flush) {
    // maxima of the four groups of 8 scores, then of the chunk: almost no chunk holds a candidate once the bound is warm
    float g[4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const int o = 8 * q4;
        g[q4] = fmaxf(fmaxf(fmaxf(__uint_as_float(v[o]), __uint_as_float(v[o + 1])), fmaxf(__uint_as_float(v[o + 2]), __uint_as_float(v[o + 3]))),
                      fmaxf(fmaxf(__uint_as_float(v[o + 4]), __uint_as_float(v[o + 5])), fmaxf(__uint_as_float(v[o + 6]), __uint_as_float(v[o + 7]))));
    }
    const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
    if (!(m >= thr_c)) return m;
    // slow path (a warp takes it when ANY of its 32 queries has a candidate, i.e. in most chunks while the bound is still
    // loose): the per-score mask is built only for the groups of 8 that hold a candidate, and a lone candidate is the
    // chunk maximum itself — no 32-way extraction
    unsigned mask = 0;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        if (g[q4] >= thr_c) {
#pragma unroll
            for (int j = 8 * q4; j < 8 * q4 + 8; ++j) mask |= (__uint_as_float(v[j]) >= thr_c) ? (1u << j) : 0u;
        }
    }
    if (left < 32) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
    const bool lone = (mask & (mask - 1u)) == 0u;
    while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        float s = m;
        if (!lone) {
            s = __uint_as_float(v[0]);
#pragma unroll
            for (int jj = 1; jj < 32; ++jj) s = (j == jj) ? __uint_as_float(v[jj]) : s;
        }
        stage_s[staged * sstride] = s;
        stage_r[staged * sstride] = row_base + j;
        if (++staged == kStage) flush();
        if (LIST && s > thr) {
            ls[0] = s;
            lr[0] = row_base + j;
#pragma unroll
            for (int i = 0; i + 1 < KL; ++i) {
                const bool sw = ls[i] > ls[i + 1];
                const float a0 = ls[i], a1 = ls[i + 1];
                const int r0 = lr[i], r1 = lr[i + 1];
                ls[i] = sw ? a1 : a0;
                ls[i + 1] = sw ? a0 : a1;
                lr[i] = sw ? r1 : r0;
                lr[i + 1] = sw ? r0 : r1;
            }
            thr = fmaxf(ls[0], g_keep);
            thr_c = thr - eps2;
        }
    }
    return m;
}

// Cooperative bound of the exact mode (one per launch, device pointers into the workspace):
//   cmax    [b_pad][groups * ms]  running maximum of sub-stream j (tiles it % ms == j) of CTA `group` for the query
//                                 — every entry is the score of a row no other entry covers, so the k-th largest
//                                 entry of a query's row is a lower bound of its k-th best score over the store
//   arrived [n_qt][4]             CTAs of a query tile whose epilogue warp w has published the maxima of its first boot_T tiles
struct XShared { float* cmax; int* arrived; int ms; int boot_T; int refresh_ns;      // refresh_ns: first sleep of the refresher (0 = no periodic refresh)
                 const float* boot_max; int boot_ext_T; };                        // sample-pass maxima [b_pad][boot_ext_T] the scan selects its first bound from (0: gtau holds it)
constexpr int kMaxSub = 8;              // sub-streams per CTA at most

// Lower bounds of the k-th largest of the V published maxima of 8 queries at once (warp-cooperative; all lanes
// get all 8 results; -inf while fewer than k maxima are known).  The loads of the 8 queries (up to 5 per lane and
// query: V <= 160) are issued together so that one L2 round trip serves the batch.  Every lane keeps the two
// largest of its strided share per query, then k rounds of (warp max, retire it on the lowest lane holding it,
// promote that lane's second value).  Where a lane would have needed a third value the result is the k-th largest of
// a SUBSET of the maxima — smaller or equal, i.e. still a valid bound.
constexpr int kMaxPublished = 160;
__device__ __forceinline__ void kth_largest_batch8(const float* __restrict__ cm_tile, int V, int k, int lane, int ql0, int ql_step,
                                                   float (&out)[8], int nq = 8) {      // only the first nq queries are computed
    float t1[8], t2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { t1[u] = VQ_NEG_INF; t2[u] = VQ_NEG_INF; }
#pragma unroll
    for (int ii = 0; ii < kMaxPublished / 32; ++ii) {
        const int i = lane + 32 * ii;
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i < V && u < nq) ? __ldcg(cm_tile + (size_t)(ql0 + u * ql_step) * V + i) : VQ_NEG_INF;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (v[u] > t1[u]) { t2[u] = t1[u]; t1[u] = v[u]; } else if (v[u] > t2[u]) t2[u] = v[u];
        }
    }
    // the selections advance in lockstep: independent dependency chains per round hide the redux / ballot latency
#pragma unroll
    for (int u = 0; u < 8; ++u) out[u] = VQ_NEG_INF;
    for (int j = 0; j < k; ++j) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (u < nq) {
                const unsigned key = ~vq_score_key(t1[u]);            // monotone increasing in the score
                const unsigned mx = __reduce_max_sync(0xffffffffu, key);
                const unsigned who = __ballot_sync(0xffffffffu, key == mx);
                out[u] = vq_key_score(~mx);
                if (lane == __ffs(who) - 1) { t1[u] = t2[u]; t2[u] = VQ_NEG_INF; }
            }
        }
    }
}

#ifdef VQ_SCAN_TRACE
// developer build: clock64 stamps of CTA 0's pipeline events over four consecutive tiles (printed when the kernel ends)
__device__ long long g_tr[4 * 16];
#define TR(itv, slot) do { if (lane == 0 && blockIdx.x == 0 && (itv) >= 200 && (itv) < 204) g_tr[((itv) - 200) * 16 + (slot)] = clock64(); } while (0)
#else
#define TR(itv, slot) do { } while (0)
#endif
template <int KL, int NT, int MODE, int EPI, bool CG2>
__global__ void __launch_bounds__(kThreadsFor(EPI), 1)
scan_mma_bf16_kernel(const __grid_constant__ CUtensorMap tmS,
                     const __nv_bfloat16* __restrict__ qbf,   // [b_pad, ld] normalised, zero padded
                     float* __restrict__ gtau,                // [b_pad] shared k-th best per query (-inf initialised)
                     int n, int ld, int nkb, int n_qt, int k, int stages, int grp_log2, int cl, int tile_mul, int b_pad,
                     float* __restrict__ cand_s,              // [b_pad, cap] surviving candidates (BOOT: [tiles, b_pad] maxima)
                     int* __restrict__ cand_r, int* __restrict__ cand_cnt, int cap, int dbg,
                     const float* __restrict__ qeps,              // [b_pad] per-query score-error bound (exact mode)
                     const XShared xs) {                          // exact mode: cooperative bound (see XShared)
    constexpr int B_KB_BYTES = NT * 128;                        // one k-block of a store tile
    constexpr int B_ST_BYTES = CG2 ? B_KB_BYTES / 2 : B_KB_BYTES;   // ... and what one ring slot of this CTA holds of it
    constexpr uint32_t A_COL0 = 2 * NT;                         // first TMEM column of the query tile
    constexpr int kEpi = EPI;                                   // epilogue warps
    constexpr int QTS = QT * (kEpi / 4);                        // epilogue threads = staging slices
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = smem;                                   // ring of store-tile k-blocks
    const size_t ring_bytes = (size_t)stages * B_ST_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + ring_bytes);
    uint64_t* empty = full + stages;
    uint64_t* a_full = empty + stages;
    uint64_t* tmem_full = a_full + 1;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint64_t* turn = tmem_empty + 2;       // [2] turn[r]: the other issuer has issued the tile before issuer r's next one
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(turn + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q_tile = blockIdx.x % n_qt, group = blockIdx.x / n_qt, n_groups = gridDim.x / n_qt;
    const int n_tiles = (n + NT - 1) / NT;
    // This CTA's tiles: group, group + n_groups, ...  Exact mode scans its first boot_T tiles twice: first for
    // their maxima only (threshold bootstrap inside the kernel), again at the very end with the bound in place.
    const int n_local = group < n_tiles ? (n_tiles - group + n_groups - 1) / n_groups : 0;
    const int boot_T = MODE == kModeExact ? (xs.boot_T < n_local ? xs.boot_T : n_local) : 0;
    const int n_iter = n_local + boot_T;
    auto tile_at = [&](int it) { return group + (it < n_local ? it : it - n_local) * n_groups; };
    // exact mode: bound per query of this tile (refreshed by warp 3), flags
    float* sbound = reinterpret_cast<float*>(smem + ring_bytes + 512 + (size_t)2 * QT * kStage * 8);
    int* sflags = reinterpret_cast<int*>(sbound + QT);          // [0] epilogue warps past the boot tiles, [1] bound ready, [2] epilogue warps done

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmS);
    if (warp == 1 && lane == 0) {
        // cta_group::2: only the leader's issuers commit (to both CTAs); the leader's a_full / tmem_empty count both CTAs' warps
        // (one full / empty barrier per ring GROUP; a group is committed by the one issuer that owns its tile)
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CG2 ? 1 : cl); }
        mbar_init(a_full, CG2 ? 8 : 4);
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], CG2 ? 2 * kEpi : kEpi); mbar_init(&turn[a], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"((uint32_t)TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (cl > 1) cluster_sync_all(); else __syncthreads();     // the peer's barriers are live before anything lands on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t cl_rank = cl > 1 ? cluster_ctarank() : 0u;
    const uint16_t cl_mask = (uint16_t)((1u << cl) - 1u);
    // everything above touched no global memory: it overlaps the previous kernel of the stream (PDL)
    vq_pdl_wait();
    vq_pdl_trigger();
    if (MODE == kModeExact) {
        if (threadIdx.x < 4) sflags[threadIdx.x] = 0;
        __syncthreads();
    }

    // Producer and MMA warps run their loops with ALL lanes (addresses, counters and descriptors stay
    // in uniform registers) and elect one lane only for the asynchronous instruction itself: a loop
    // executed by a single lane makes ptxas wrap every UTCHMMA in an ELECT / R2UR waterfall
    // (~25 dependent instructions, measured 131 cycles per MMA instead of 64).
    //
    // The ring is handed over in GROUPS of gs = 2^grp_log2 k-blocks (gs divides nkb; 64 KB per group at dim 512: half a tile
    // with 1-CTA MMAs, a whole tile for a CTA pair): one expect_tx and gs loads per group on the producer side, ONE barrier
    // wait, 4 * gs back-to-back MMAs and one commit on the issuer side.  With a wait, a fence, an election and a commit per
    // k-block the issue loop itself was the bound (~100 cycles per MMA, with or without loads and MMAs switched off).
    const int gs = 1 << grp_log2;
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        const uint32_t sB_addr = smem_u32(sB), full_addr = smem_u32(full);
        for (int it_p = 0; it_p < n_iter; ++it_p) {
            const int tile = tile_at(it_p);
            for (int kb0 = 0; kb0 < nkb; kb0 += gs) {
                const int grp = stage >> grp_log2;
                mbar_wait(&empty[grp], phase ^ 1);                         // the whole group is free
                if (elect_one()) {
                    if (dbg & 2) { mbar_arrive(&full[grp]); }
                    else if (CG2) {
                        // each CTA keeps its own half of every k-block (plain loads on its own barrier; the peer's relay warp
                        // forwards the completion to the leader: .cta_group::2 loads counted on the leader's barrier streamed
                        // at half the rate, 0.72 ms per 1M-row pass against 0.46 ms)
                        mbar_expect_tx(&full[grp], (uint32_t)gs * B_ST_BYTES);
                        for (int j = 0; j < gs; ++j)
                            tma_load_2d_addr(sB_addr + (uint32_t)(stage + j) * B_ST_BYTES, &tmS, full_addr + (uint32_t)grp * 8,
                                             (kb0 + j) * KB_ELEMS, tile * NT + (int)cl_rank * (NT / 2));
                    } else {
                        mbar_expect_tx(&full[grp], (uint32_t)gs * B_KB_BYTES);  // whole boxes: own part + the peer's multicast
                        for (int j = 0; j < gs; ++j) {
                            if (cl > 1)
                                tma_load_2d_mc(sB_addr + (uint32_t)(stage + j) * B_KB_BYTES + cl_rank * (B_KB_BYTES / 2), &tmS,
                                               full_addr + (uint32_t)grp * 8, (kb0 + j) * KB_ELEMS, tile * NT + (int)cl_rank * (NT / 2), cl_mask);
                            else
                                tma_load_2d_addr(sB_addr + (uint32_t)(stage + j) * B_KB_BYTES, &tmS, full_addr + (uint32_t)grp * 8,
                                                 (kb0 + j) * KB_ELEMS, tile * tile_mul * NT);   // boot pass: sample tile j is store tile j * tile_mul
                        }
                    }
                }
                __syncwarp();
                stage += gs;
                if (stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (CG2 && cl_rank != 0 && warp == 1) {
        // relay (peer CTA of a pair): "my half of group g has landed" -> the leader's pfull[g]
        const uint32_t pfull_leader = mapa_rank(smem_u32(full + stages / 2), 0);
        int stage = 0; uint32_t phase = 0;
        for (int it_r = 0; it_r < n_iter; ++it_r) {
            for (int kb0 = 0; kb0 < nkb; kb0 += gs) {
                const int grp = stage >> grp_log2;
                mbar_wait(&full[grp], phase);
                if (lane == 0) mbar_arrive_cluster(pfull_leader + (uint32_t)grp * 8);
                __syncwarp();
                stage += gs;
                if (stage == stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if ((warp == 1 || warp == 2) && !(CG2 && cl_rank != 0)) {
        // Two issuing warps on two SM sub-partitions own alternate tiles (issuer r: accumulator r), so the waits, the fence
        // and the descriptor arithmetic of one tile hide behind the other issuer's MMAs.  The tensor pipe runs MMAs in issue
        // order: issuer r starts a tile only after the other one has issued ALL MMAs of the tile before it (turn[]), so that
        // tile t completes — and its epilogue starts — while tile t + 1 is still being multiplied.
        const int role = warp - 1;
        const uint32_t idesc = umma_idesc_bf16(NT, CG2 ? 2 * QT : QT);
        const uint32_t sB_addr = smem_u32(sB);
        mbar_wait(a_full, 0);                          // query tile is in tensor memory
        tc_fence_after();
        int stage = 0; uint32_t phase = 0; int it = 0;
        long long dbg_c0 = 0, dbg_t0 = 0;
        if (dbg & 128) { dbg_c0 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0)); }
        for (; it < n_iter; ++it) {
            const int acc = it & 1;
            const bool mine = acc == role;
            const int nth = it >> 1;                   // this issuer's nth tile
            if (mine) {
                TR(it, 0);
                // The tile before this one is in the pipe.  This wait comes BEFORE the ring waits for a reason: an issuer
                // skips the other one's tiles, and a parity wait two phases ahead of its barrier passes at once.
                if (it > 0) mbar_wait(&turn[role], (uint32_t)((role == 1 ? nth : nth - 1) & 1));
                TR(it, 2);
            }
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * NT;
            for (int kb0 = 0; kb0 < nkb; kb0 += gs) {
                const int grp = stage >> grp_log2;
                if (mine) {
                    mbar_wait(&full[grp], phase);
                    if (CG2) mbar_wait(&full[stages / 2 + grp], phase);      // pfull: the peer's half (see the relay warp)
                    // the accumulator last: the epilogue's release -> first MMA of this tile is the chain that has to fit
                    // into the other tile's MMAs
                    if (kb0 == 0) mbar_wait(&tmem_empty[acc], (uint32_t)(nth & 1) ^ 1);
                    tc_fence_after();
                    if (kb0 == 0) TR(it, 1);
                    TR(it, kb0 == 0 ? 3 : 5);
                    if (elect_one()) {
                        if (!(dbg & 1)) {
                            for (int j = 0; j < gs; ++j) {
                                const uint64_t bd0 = umma_desc_sw128(sB_addr + (uint32_t)(stage + j) * B_ST_BYTES);
                                const uint32_t a_tmem = tmem_base + A_COL0 + (uint32_t)(kb0 + j) * (KB_ELEMS / 2);
#pragma unroll
                                for (int k4 = 0; k4 < KB_ELEMS / 16; ++k4) {   // +32 bytes per K=16 step = +2 in the (addr >> 4) field
                                    const uint32_t accum = ((kb0 + j) | k4) != 0 ? 1u : 0u;
                                    if (CG2) umma_bf16_ts_cg2(d_tmem, a_tmem + k4 * 8, bd0 + (uint64_t)(k4 * 2), idesc, accum);
                                    else umma_bf16_ts(d_tmem, a_tmem + k4 * 8, bd0 + (uint64_t)(k4 * 2), idesc, accum);
                                }
                            }
                        }
                        // the group goes back to the producer(s), the accumulator to the epilogue after the last k-block
                        const bool last = kb0 + gs == nkb;
                        if (CG2) {
                            umma_commit_cg2(&empty[grp]);
                            if (last) umma_commit_cg2(&tmem_full[acc]);
                        } else {
                            if (cl > 1) umma_commit_mc(&empty[grp], cl_mask); else umma_commit(&empty[grp]);
                            if (last) umma_commit(&tmem_full[acc]);
                        }
                        if (last) mbar_arrive(&turn[role ^ 1]);
                    }
                    __syncwarp();
                    TR(it, kb0 == 0 ? 4 : 6);
                }
                stage += gs;
                if (stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        if ((dbg & 128) && blockIdx.x == 0 && lane == 0 && role == 0) {
            long long c1 = clock64(), t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            printf("[scan_mma dbg] issue loop: %d tiles, %lld cycles, %lld ns -> %.1f cycles/MMA, %.0f MHz\n", it, c1 - dbg_c0,
                   t1 - dbg_t0, (double)(c1 - dbg_c0) / ((double)it * nkb * 4), (double)(c1 - dbg_c0) / (double)(t1 - dbg_t0) * 1e3);
        }
    } else if (warp == 3 && MODE == kModeExact) {
        // Bound refresher.  The k-th largest of the maxima all CTAs of a query tile have published is the same number
        // whichever CTA computes it, so the work is SPLIT: CTA `group` refreshes only its share of the tile's 128 queries
        // (~128 / n_groups of them, one batch of 8 per round) and publishes the result with an atomic max in the global
        // bound array every epilogue thread reads (one tile ahead) anyway.  A round is a handful of L2 loads plus ~100
        // instructions — cheap enough to run every few microseconds next to an epilogue warp, for scans of any length
        // (a private refresher over all 128 queries per CTA cost 16 such batches per round and slowed short scans down).
        const int V = n_groups * xs.ms;
        const float* cm_tile = xs.cmax + (size_t)(q_tile * QT) * V;
        const int q_lo = (int)(((long long)group * QT) / n_groups), q_hi = (int)(((long long)(group + 1) * QT) / n_groups);
        if (boot_T > 0) {
            const long long t0 = clock64();
            while (*reinterpret_cast<volatile int*>(&sflags[0]) < kEpi && clock64() - t0 < 8000000) __nanosleep(500);
        }
        unsigned sleep_ns = (unsigned)xs.refresh_ns;
        while (sleep_ns > 0 && q_hi > q_lo && *reinterpret_cast<volatile int*>(&sflags[2]) < kEpi) {
            // sleep in 1 us slices: the CTA cannot retire before this warp has seen the epilogue's done flag
            for (unsigned slept = 0; slept < sleep_ns && *reinterpret_cast<volatile int*>(&sflags[2]) < kEpi; slept += 1000) __nanosleep(1000);
            if (*reinterpret_cast<volatile int*>(&sflags[2]) >= kEpi) break;
            bool changed = false;
            for (int q0 = q_lo; q0 < q_hi; q0 += 8) {
                const int ql = q0 + lane < q_hi ? q0 + lane : q_hi - 1;              // lanes 0..7: the batch's queries (clamped)
                float cur = VQ_NEG_INF;
                if (lane < 8) cur = *reinterpret_cast<volatile float*>(gtau + q_tile * QT + ql);
                if (__all_sync(0xffffffffu, lane >= 8 || cur == INFINITY)) continue;   // padding queries
                float kth[8];
                const int shift = q0 + 7 < QT ? 0 : q0 - (QT - 8);                    // batch start pulled back at the tile's end
                const int nq = (q_hi - q0 < 8 ? q_hi - q0 : 8) + shift;               // CTAs of a 148-group launch own ONE query each
                kth_largest_batch8(cm_tile, V, k, lane, q0 - shift, 1, kth, nq);
                float mine = VQ_NEG_INF;
#pragma unroll
                for (int u = 0; u < 8; ++u) mine = (lane + shift == u) ? kth[u] : mine;
                const bool up = lane < 8 && q0 + lane < q_hi && mine > cur;
                if (up) atomic_max_float(gtau + q_tile * QT + ql, mine);
                changed = changed || __any_sync(0xffffffffu, up);
            }
            sleep_ns = changed ? (unsigned)xs.refresh_ns : (sleep_ns < 16u * (unsigned)xs.refresh_ns ? sleep_ns * 2 : sleep_ns);
        }
    } else if (warp >= 4 && warp < 4 + kEpi) {
        const int ew = (warp - 4) & 3;                    // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        const int half = (warp - 4) >> 2;                 // 8 epilogue warps: which half of a tile's columns this warp filters
        // TMEM lane ew*32+lane serves query (lane*4 + ew) of the tile: a small batch is spread over all
        // four lane quarters
        const int q = q_tile * QT + lane * 4 + ew;
        const uint32_t lane_base = tmem_base + ((uint32_t)(ew * 32) << 16);
        // an accumulator is handed back to the issuers — of the pair's leader when the pair shares its MMAs
        const uint32_t tmem_empty_leader = CG2 ? mapa_rank(smem_u32(tmem_empty), 0) : 0u;
        auto release_acc = [&](int a) { if (CG2) mbar_arrive_cluster(tmem_empty_leader + (uint32_t)a * 8); else mbar_arrive(&tmem_empty[a]); };
        // ---- query tile -> tensor memory: lane t holds query q, column c holds elements 2c, 2c+1
        if (half == 0) {
            const uint4* qrow = reinterpret_cast<const uint4*>(qbf + (size_t)q * ld);
            // 4 k-blocks (32 x 16 B loads) in flight per thread before the first tcgen05.st
            for (int kb0 = 0; kb0 < nkb; kb0 += 4) {
                uint32_t w[4][32];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (kb0 + u < nkb) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint4 x = qrow[(kb0 + u) * 8 + i];
                            w[u][4 * i] = x.x; w[u][4 * i + 1] = x.y; w[u][4 * i + 2] = x.z; w[u][4 * i + 3] = x.w;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (kb0 + u < nkb) tmem_st32(lane_base + A_COL0 + (uint32_t)(kb0 + u) * (KB_ELEMS / 2), w[u]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (CG2) mbar_arrive_cluster(mapa_rank(smem_u32(a_full), 0)); else mbar_arrive(a_full); }
        }
        if (MODE == kModeCollect) {
            // collect pass: EVERY row whose score reaches the query's fixed threshold is appended to the
            // query's candidate buffer (no lists).  Used to resolve queries the certificate rejected.
            const float thr = gtau[q];
            const size_t dst = (size_t)q * cap;
            // candidates are staged kStage at a time in this thread's slice of shared memory and flushed
            // with ONE atomicAdd: waiting for a global atomic per candidate (an L2 round trip inside the
            // tile loop, hit by some lane in almost every 32-score chunk) cost as much as the MMAs
            float* stage_s = reinterpret_cast<float*>(smem + ring_bytes + 512) + (size_t)(ew * 32 + lane) * kStage;
            int* stage_r = reinterpret_cast<int*>(smem + ring_bytes + 512 + (size_t)QT * kStage * 4) + (size_t)(ew * 32 + lane) * kStage;
            int staged = 0;
            auto flush = [&]() {
                const int at = atomicAdd(cand_cnt + q, staged);
                for (int i = 0; i < staged; ++i)
                    if (at + i < cap) { cand_s[dst + at + i] = stage_s[i]; cand_r[dst + at + i] = stage_r[i]; }
                staged = 0;
            };
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
                    tmem_ld32(lane_base + (uint32_t)(acc * NT + c0), v);
                    unsigned mask = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) >= thr) ? (1u << j) : 0u;
                    const int left = valid - c0;
                    if (left < 32) mask &= left > 0 ? ((1u << left) - 1u) : 0u;
                    while (mask) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        float sc = __uint_as_float(v[0]);
#pragma unroll
                        for (int jj = 1; jj < 32; ++jj) sc = (j == jj) ? __uint_as_float(v[jj]) : sc;
                        stage_s[staged] = sc;
                        stage_r[staged] = row0 + c0 + j;
                        if (++staged == kStage) flush();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
            }
            if (staged) flush();
        } else if (MODE == kModeExact) {
            // single-pass exact search: register list for the running k-th best + collection of every row
            // within 2*eps of the bound (see filter_collect).  The bound is max(own k-th best, sbound[query]):
            // sbound is the k-th largest of the maxima all CTAs of this query tile have published (warp 3).
            // KL == 1: LISTLESS (k beyond the register lists, up to 128): the bound is the cooperative one alone —
            // with about as many published maxima as k, their k-th largest sits within a few hundred ranks of the true
            // k-th best, which is as tight as a gather for k ~ 100 needs to be.
            constexpr bool LIST = KL > 1;
            float ls[KL];
            int lr[KL];
#pragma unroll
            for (int i = 0; i < KL; ++i) { ls[i] = (LIST && i < k) ? VQ_NEG_INF : INFINITY; lr[i] = VQ_EMPTY_ROW; }
            float eps2 = 2.f * qeps[q];              // -inf once the buffer has overflowed: thr_c = +inf, nothing passes
            const size_t dst = (size_t)q * cap;
            float* stage_s = reinterpret_cast<float*>(smem + ring_bytes + 512) + (half * QT + ew * 32 + lane);
            int* stage_r = reinterpret_cast<int*>(smem + ring_bytes + 512 + (size_t)QTS * kStage * 4) + (half * QT + ew * 32 + lane);
            int staged = 0;
            auto flush = [&]() {
                const int at = atomicAdd(cand_cnt + q, staged);
                for (int i = 0; i < staged; ++i)
                    if (at + i < cap) { cand_s[dst + at + i] = stage_s[i * QTS]; cand_r[dst + at + i] = stage_r[i * QTS]; }
                staged = 0;
                if (at + kStage > cap) eps2 = VQ_NEG_INF;    // overflow (mass ties): the finish kernel flags the query
            };
            // published maxima of this thread's query: sub-stream j = first-pass tiles with it % ms == j
            float smax[kMaxSub];
#pragma unroll
            for (int j = 0; j < kMaxSub; ++j) smax[j] = VQ_NEG_INF;
            const int ms = xs.ms;
            float* my_cmax = xs.cmax + ((size_t)q * n_groups + group) * ms;
            auto publish = [&](int sub, float m) {
#pragma unroll
                for (int j = 0; j < kMaxSub; ++j)
                    if (j == sub && m > smax[j]) { smax[j] = m; red_max_float(my_cmax + j, m); }   // (two column halves share a slot)
            };
            int sub_it = 0;                                                   // it % ms, kept incrementally
            float g_next = *reinterpret_cast<volatile float*>(gtau + q);      // static bound (+inf: padding query), then the refreshed one
            // First bound = the k-th largest of the sample pass's tile maxima of the query (k distinct rows reach it).  Selected
            // HERE, while the first tiles are being loaded and multiplied, instead of by a kernel of its own between the sample
            // pass and the scan: sample pass -> select -> scan was the serial chain a step could not overlap with its
            // neighbours (12 us of ~155 on a 125k-row shard).  The halves of a quarter split its four batches of 8 queries.
            float g0 = VQ_NEG_INF;
            if (xs.boot_ext_T > 0) {
                const unsigned real = __ballot_sync(0xffffffffu, g_next != INFINITY);      // bit j: lane j's query is not padding
                const float* bm = xs.boot_max + (size_t)(q_tile * QT) * xs.boot_ext_T;
                const int bq_lo = kEpi == 8 ? 2 * half : 0, bq_hi = kEpi == 8 ? 2 * half + 2 : 4;
                for (int bq = bq_lo; bq < bq_hi; ++bq) {
                    if (((real >> (8 * bq)) & 0xffu) == 0) continue;
                    float kth[8];
                    kth_largest_batch8(bm, xs.boot_ext_T, k, lane, (8 * bq) * 4 + ew, 4, kth);
                    float mine = kth[0];
#pragma unroll
                    for (int u = 1; u < 8; ++u) mine = (lane == u) ? kth[u] : mine;
                    if (lane < 8) sbound[(8 * bq + lane) * 4 + ew] = mine;
                }
                asm volatile("bar.sync 1, %0;" ::"r"(kEpi * 32) : "memory");           // the epilogue warps only
                if (g_next != INFINITY) g0 = sbound[lane * 4 + ew];
            }
            for (int it = 0; it < n_iter; ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                const int tile = tile_at(it);
                const int row0 = tile * NT;
                const int valid = (n - row0) < NT ? (n - row0) : NT;
                constexpr int n_chunks = NT / 32;
                if (it < boot_T) {
                    // ---- bootstrap tiles: only the maximum is taken (they are scanned again at the end)
                    mbar_wait(&tmem_full[acc], acc_phase);
                    tc_fence_after();
                    float m = VQ_NEG_INF;
#pragma unroll 1
                    for (int c0 = half * (NT / (kEpi / 4)); c0 < (half + 1) * (NT / (kEpi / 4)); c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(lane_base + (uint32_t)(acc * NT + c0), v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) m = (c0 + j < valid) ? fmaxf(m, __uint_as_float(v[j])) : m;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                    publish(sub_it, m);
                    if (++sub_it == ms) sub_it = 0;
                    if (it == boot_T - 1) {
                        // The maxima of this warp's 32 queries are out.  Arrive on the (query tile, epilogue warp)
                        // counter and wait — bounded — for the same warp of every other CTA: the k-th largest of what
                        // they published is the first bound of these 32 queries (the threshold bootstrap).
                        __threadfence();
                        __syncwarp();
                        int* arr = xs.arrived + q_tile * 8 + (warp - 4);
                        if (lane == 0) atomicAdd(arr, 1);
                        const long long t0 = clock64();
                        while (*reinterpret_cast<volatile int*>(arr) < n_groups && clock64() - t0 < 4000000) __nanosleep(64);
                        __threadfence();
                        const float cur = *reinterpret_cast<volatile float*>(gtau + q);
                        const unsigned real = __ballot_sync(0xffffffffu, cur != INFINITY);      // bit j: lane j's query is not padding
                        const float* cm_tile = xs.cmax + (size_t)(q_tile * QT) * (n_groups * ms);
                        float mine = cur;
                        for (int bq = 0; bq < 4; ++bq) {
                            if (((real >> (8 * bq)) & 0xffu) == 0) continue;
                            float kth[8];
                            kth_largest_batch8(cm_tile, n_groups * ms, k, lane, (8 * bq) * 4 + ew, 4, kth);
#pragma unroll
                            for (int u = 0; u < 8; ++u) mine = (lane == 8 * bq + u) ? fmaxf(mine, kth[u]) : mine;
                        }
                        if (cur != INFINITY && mine > cur) atomic_max_float(gtau + q, mine);
                        g_next = fmaxf(g_next, mine);
                        __syncwarp();
                        if (lane == 0) atomicAdd(&sflags[0], 1);
                    }
                    continue;
                }
                const float g = fmaxf(g_next, g0);                             // loaded one tile ago: no L2 round trip here
                const float g_keep = (g == VQ_NEG_INF) ? g : nextafterf(g, VQ_NEG_INF);
                g_next = ld_bound_after(gtau + q, g_keep);
                float thr = LIST ? fmaxf(ls[0], g_keep) : g_keep;
                float thr_c = thr - eps2;
                if (ew == 0) TR(it, 8 + 4 * half);
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                if (ew == 0) TR(it, 9 + 4 * half);
                uint32_t va[32], vb[32];
                float tmax = VQ_NEG_INF;
                constexpr int cpw = n_chunks / (kEpi / 4);           // chunks of 32 columns this warp filters per tile
                const int c_lo = half * cpw;
                if (cpw <= 2) {
                    // the warp's whole share of the tile fits the two register buffers: copy it out and hand the accumulator
                    // back to the issuers BEFORE filtering (the release -> next MMA chain was on the critical path)
                    tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + c_lo * 32), va);
                    if (cpw == 2) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + (c_lo + 1) * 32), vb);
                    tmem_ld_wait(va);
                    if (cpw == 2) tmem_ld_wait(vb);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                    if (!(dbg & 4)) {
                        tmax = filter_collect<KL, LIST>(va, row0 + c_lo * 32, valid - c_lo * 32, g_keep, eps2, thr, thr_c, ls, lr, stage_s, stage_r, QTS, staged, flush);
                        if (cpw == 2)
                            tmax = fmaxf(tmax, filter_collect<KL, LIST>(vb, row0 + (c_lo + 1) * 32, valid - (c_lo + 1) * 32, g_keep, eps2, thr, thr_c, ls, lr, stage_s, stage_r, QTS, staged, flush));
                    }
                } else {
                if (!(dbg & 4)) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + c_lo * 32), va);
#pragma unroll 1
                for (int c = (dbg & 4) ? cpw : 0; c < cpw; c += 2) {
                    tmem_ld_wait(va);
                    if (c + 1 < cpw) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + (c_lo + c + 1) * 32), vb);
                    tmax = fmaxf(tmax, filter_collect<KL, LIST>(va, row0 + (c_lo + c) * 32, valid - (c_lo + c) * 32, g_keep, eps2, thr, thr_c, ls, lr, stage_s, stage_r, QTS, staged, flush));
                    if (c + 1 < cpw) {
                        tmem_ld_wait(vb);
                        if (c + 2 < cpw) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + (c_lo + c + 2) * 32), va);
                        tmax = fmaxf(tmax, filter_collect<KL, LIST>(vb, row0 + (c_lo + c + 1) * 32, valid - (c_lo + c + 1) * 32, g_keep, eps2, thr, thr_c, ls, lr, stage_s, stage_r, QTS, staged, flush));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                }
                if (ew == 0) TR(it, 10 + 4 * half);
                // first pass over a full tile: its maximum feeds the cooperative bound (a partial tile's maximum would
                // include the zero scores of the padding rows; re-scanned tiles were counted in their first pass)
                if (it < n_local && valid == NT && !(dbg & 8)) publish(sub_it, tmax);
                if (it < n_local && ++sub_it == ms) sub_it = 0;
            }
            if (staged) flush();
            __syncwarp();
            if (lane == 0) atomicAdd(&sflags[2], 1);
        } else if (MODE == kModeBoot) {
            // threshold bootstrap: only the per-query maximum of every sample tile is kept
            int it = 0;
            for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tmem_full[acc], acc_phase);
                tc_fence_after();
                const int valid = (n - tile * NT) < NT ? (n - tile * NT) : NT;
                float m = VQ_NEG_INF;
#pragma unroll 1
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + (uint32_t)(acc * NT + c0), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) m = (c0 + j < valid) ? fmaxf(m, __uint_as_float(v[j])) : m;
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(acc);
                // (cap > 0: query-major [b_pad][cap] — the layout the exact scan selects its first bound from)
                if (cap > 0) cand_s[(size_t)q * cap + tile] = m; else cand_s[(size_t)tile * b_pad + q] = m;
            }
        } else {
        // top-k list of this thread's query, in registers, WORST first: slots [0,k) are live and
        // ascending, slots [k,KL) hold +inf so a bubble pass stops in front of them.  The running
        // k-th best is therefore always ls[0] (a static register, no dynamic indexing).
        float ls[KL];
        int lr[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) { ls[i] = i < k ? VQ_NEG_INF : INFINITY; lr[i] = VQ_EMPTY_ROW; }
        float published = VQ_NEG_INF;
        int it = 0;
        // k-th best any CTA has published for this query (or the bootstrap bound).  It is (re)loaded one
        // tile ahead so that the L2 round trip never sits between an accumulator becoming ready and
        // its first filter step; a bound that is one tile stale only lets a few more rows through.
        float g_next = *reinterpret_cast<volatile float*>(gtau + q);
        for (int tile = group; tile < n_tiles; tile += n_groups, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            // ties with the bound are kept (>=) so the (score desc, row asc) rule still sees every
            // candidate it needs
            const float g = g_next;
            g_next = *reinterpret_cast<volatile float*>(gtau + q);
            const float g_keep = (g == VQ_NEG_INF) ? g : nextafterf(g, VQ_NEG_INF);
            float thr = fmaxf(ls[0], g_keep);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row0 = tile * NT;
            const int valid = (n - row0) < NT ? (n - row0) : NT;
            const int n_chunks = (dbg & 4) ? 0 : NT / 32;
            // two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is filtered
            uint32_t va[32], vb[32];
            if (n_chunks > 0) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT), va);
#pragma unroll 1
            for (int c = 0; c < n_chunks; c += 2) {
                tmem_ld_wait(va);
                if (c + 1 < n_chunks) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + (c + 1) * 32), vb);
                filter_insert<KL>(va, row0 + c * 32, valid - c * 32, g_keep, thr, ls, lr);
                if (c + 1 < n_chunks) {
                    tmem_ld_wait(vb);
                    if (c + 2 < n_chunks) tmem_ld32_issue(lane_base + (uint32_t)(acc * NT + (c + 2) * 32), va);
                    filter_insert<KL>(vb, row0 + (c + 1) * 32, valid - (c + 1) * 32, g_keep, thr, ls, lr);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(acc);
            if (ls[0] > published) {                      // list full and improved: share the new k-th best
                published = ls[0];
                atomic_max_float(gtau + q, published);
            }
        }
        // Write-out: only entries that can still belong to the global top-k, i.e. that reach the latest
        // shared bound, are appended to the query's candidate buffer (one atomicAdd per thread).  The
        // ~k..2k survivors per query replace groups*k list entries as the input of the final selection.
        {
            const float g = *reinterpret_cast<volatile float*>(gtau + q);
            const float g_keep = (g == VQ_NEG_INF) ? g : nextafterf(g, VQ_NEG_INF);
            int n_surv = 0;
#pragma unroll
            for (int i = 0; i < KL; ++i) n_surv += (i < k && lr[i] != VQ_EMPTY_ROW && ls[i] > g_keep) ? 1 : 0;
            int at = n_surv ? atomicAdd(cand_cnt + q, n_surv) : 0;
            const size_t dst = (size_t)q * cap;
#pragma unroll
            for (int i = 0; i < KL; ++i) {
                if (i < k && lr[i] != VQ_EMPTY_ROW && ls[i] > g_keep) {
                    if (at < cap) { cand_s[dst + at] = ls[i]; cand_r[dst + at] = lr[i]; }
                    ++at;
                }
            }
        }
        }
    }
    tc_fence_before();
    if (cl > 1) cluster_sync_all(); else __syncthreads();     // no CTA leaves while its peer may still write into it
#ifdef VQ_SCAN_TRACE
    if (threadIdx.x == 0 && blockIdx.x == 0 && n_iter > 204) {
        for (int t = 0; t < 4; ++t) {
            printf("[scan trace] tile %d:", 200 + t);
            for (int i = 0; i < 15; ++i) printf(" %d:%lld", i, g_tr[t * 16 + i] ? g_tr[t * 16 + i] - g_tr[1] : 0);
            printf("\n");
        }
    }
#endif
    if (warp == 2) {
        tc_fence_after();
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// fp32 queries -> L2-normalised bf16 [b_pad, ld] (rows >= b and columns >= dim are zero) and the
// shared threshold array reset; one warp per row.
__global__ void __launch_bounds__(256)
prep_queries_bf16_kernel(const float* __restrict__ src, int b, int dim, int src_ld, __nv_bfloat16* __restrict__ dst, int ld,
                         int b_pad, int mode, float* __restrict__ gtau, int* __restrict__ cand_cnt,
                         const float* __restrict__ thr_in,     // collect pass: per-query thresholds (NULL otherwise)
                         const float* __restrict__ bounds,     // exact mode: {max |x^|, max |x^ - x|} over the store's rows
                         float* __restrict__ qeps,             // exact mode: per-query bound on |bf16 score - fp32 score|
                         float* __restrict__ cmax, int cmax_v, // exact mode: published maxima [b_pad][cmax_v], reset to -inf
                         int* __restrict__ arrived, int n_qt,
                         float eps_const) {                    // bounds == NULL: qeps[q] = eps_const (caller-supplied bound)
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    vq_pdl_wait();                     // the previous search's kernels still read gtau / cand_cnt / dst
    vq_pdl_trigger();
    if (row >= b_pad) return;
    if (cmax) {
        for (int c = lane; c < cmax_v; c += 32) cmax[(size_t)row * cmax_v + c] = VQ_NEG_INF;
        if (lane == 0 && row < 8 * n_qt) arrived[row] = 0;
    }
    __nv_bfloat16* o = dst + (size_t)row * ld;
    // padding queries (zero rows of the last query tile) must neither keep nor gather anything: bound +inf
    if (lane == 0) { gtau[row] = row < b ? (thr_in ? thr_in[row] : VQ_NEG_INF) : INFINITY; cand_cnt[row] = 0; }
    if (row >= b) {
        for (int c = lane; c < ld; c += 32) o[c] = __float2bfloat16_rn(0.f);
        if (qeps && lane == 0) qeps[row] = 0.f;
        return;
    }
    const float* s = src + (size_t)row * src_ld;
    float d = 1.f;
    if (mode != VQ_NORM_NONE) {
        float sum = 0.f;
        for (int c = lane; c < dim; c += 32) { const float v = s[c]; sum = fmaf(v, v, sum); }
        sum = vq_warp_sum(sum);
        d = sqrtf(sum);
        if (mode == VQ_NORM_EPS) d += 1e-10f;
    }
    float e2 = 0.f, n2 = 0.f;          // |q^ - q|^2 and |q|^2 of the normalised fp32 query q (what the re-score uses)
    for (int c = lane; c < ld; c += 32) {
        const float x = c < dim ? (mode == VQ_NORM_NONE ? s[c] : s[c] / d) : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        o[c] = h;
        const float df = __bfloat162float(h) - x;
        e2 = fmaf(df, df, e2);
        n2 = fmaf(x, x, n2);
    }
    if (qeps) {
        // score error of the tensor-core scan against the fp32 re-score, per query (Cauchy-Schwarz on
        //   q^.x^ - q.x = (q^ - q).x^ + q.(x^ - x)):  |q^ - q| * max|x^| + |q| * max|x^ - x|,
        // plus the fp32 accumulation error of both sums (<= 3 * ld * 2^-24 * |q| |x^|), inflated by 1e-3 for
        // the rounding of this very computation.  Non-finite input: eps = +inf would gather everything, the
        // scores are NaN anyway and nothing is gathered.
        e2 = vq_warp_sum(e2);
        n2 = vq_warp_sum(n2);
        if (lane == 0) {
            if (bounds) {
                const float qn = sqrtf(n2), b0 = bounds[0], b1 = bounds[1];
                qeps[row] = 1.001f * (sqrtf(e2) * b0 + qn * b1 + 3.f * (float)ld * 5.9604645e-8f * qn * b0);
            } else {
                qeps[row] = eps_const;
            }
        }
    }
}

// Threshold bootstrap, step 2: the k-th largest of the n_t per-tile maxima of a query is reached by k
// distinct store rows, i.e. it is a valid lower bound of the global k-th best score.  One warp per
// query: the maxima go to shared memory and every lane ranks its own values against all of them
// (n_t <= kMaxBootTiles, no serial dependency chain); the value of rank k-1 is the bound.
__global__ void __launch_bounds__(256)
boot_select_kernel(const float* __restrict__ boot_max, int n_t, int b, int b_pad, int k, float* __restrict__ gtau) {
    __shared__ float vals[8][kMaxBootTiles];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * 8 + w;
    vq_pdl_wait();
    vq_pdl_trigger();
    if (q >= b_pad) return;
    for (int t = lane; t < n_t; t += 32) vals[w][t] = boot_max[(size_t)t * b_pad + q];
    __syncwarp();
    float kth = VQ_NEG_INF;
    for (int t = lane; t < n_t; t += 32) {
        const float v = vals[w][t];
        int rank = 0;                                      // values strictly better (ties broken by index)
        for (int j = 0; j < n_t; ++j) {
            const float o = vals[w][j];
            rank += (o > v || (o == v && j < t)) ? 1 : 0;
        }
        if (rank == k - 1) kth = v;                        // exactly one (t, lane) has this rank
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
    // padding queries (zero vectors) must not keep or gather anything: their bound is +inf
    if (lane == 0) gtau[q] = q >= b ? INFINITY : (n_t >= k) ? kth : VQ_NEG_INF;
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

// Tensor maps are pure functions of (base, rows, ld, box): keep the last few (encoding costs ~10 us of
// host time per call, which is visible at batch 1).
struct MapKey { const void* base; uint64_t rows, ld; uint32_t box_rows; };
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
std::mutex g_map_mu;
MapSlot g_maps[8];
int g_map_next = 0;

bool get_map_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    std::lock_guard<std::mutex> lock(g_map_mu);
    for (auto& sl : g_maps)
        if (sl.used && sl.key.base == base && sl.key.rows == rows && sl.key.ld == ld && sl.key.box_rows == box_rows) {
            *out = sl.map;
            return true;
        }
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {KB_ELEMS, box_rows};
    cuuint32_t estr[2] = {1, 1};
    MapSlot& sl = g_maps[g_map_next];
    g_map_next = (g_map_next + 1) % 8;
    if (enc(&sl.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        sl.used = false;
        return false;
    }
    sl.key = {base, rows, ld, box_rows};
    sl.used = true;
    *out = sl.map;
    return true;
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct MmaPlan {
    int nkb, nt, n_qt, b_pad, groups, grid, stages, grp_log2, cl;
    int boot_tiles, boot_groups, boot_mul;       // threshold bootstrap (0 = off); sample tile j = store tile j * boot_mul
    int boot_qmajor = 0;                         // sample maxima stored [b_pad][boot_tiles] and selected by the scan itself (exact mode)
    int cap;                                     // candidate slots per query (= groups * k, cannot overflow)
    size_t smem;
    // workspace layout
    size_t off_tau, off_cnt, off_cand_s, off_cand_r, off_boot, off_qbf, total;
};
MmaPlan plan(int64_t n, int ld, int b, int k) {
    MmaPlan p;
    p.nkb = ld / KB_ELEMS;
    // TMEM budget: 2 accumulators of nt columns + ld/2 columns of resident queries <= 512
    p.nt = (ld <= 512) ? 128 : (ld <= 768 ? 64 : 0);
    p.n_qt = (b + QT - 1) / QT;
    p.b_pad = p.n_qt * QT;
    const int sms = vq_num_sms();
    const long long n_tiles = p.nt ? (n + p.nt - 1) / p.nt : 1;
    long long groups = sms / p.n_qt;
    if (groups < 1) groups = 1;
    if (groups > n_tiles) groups = n_tiles;
    p.groups = (int)groups;
    p.grid = p.groups * p.n_qt;
    const size_t stage_bytes = (size_t)(p.nt ? p.nt : 128) * 128;
    int st = (int)((200 * 1024) / stage_bytes);
    // The ring is filled, consumed and released in groups of up to 4 k-blocks (one barrier wait and one commit per group,
    // see the kernel); the group size must divide the k-blocks of a tile.
    static const int grp_env = getenv("VQ_MMA_GROUP") ? atoi(getenv("VQ_MMA_GROUP")) : -1;
    p.grp_log2 = grp_env >= 0 && grp_env <= 2 ? grp_env : 2;
    while (p.grp_log2 > 0 && p.nkb % (1 << p.grp_log2) != 0) --p.grp_log2;
    static const int st_env = getenv("VQ_MMA_STAGES") ? atoi(getenv("VQ_MMA_STAGES")) : 0;      // experiments: a shallower ring
    if (st_env > 0 && st_env < st) st = st_env;
    p.stages = (st > 12 ? 12 : st) >> p.grp_log2 << p.grp_log2;
    // an even number of query tiles: two CTAs with neighbouring query tiles form a cluster and share every
    // store tile (each loads half a box and multicasts it), halving the L2 -> SM traffic per FLOP
    // Short runs gain 1-5 %; in a long tensor-bound run the board sits at its 1000 W cap and the halved
    // L2 -> SM stream buys clock: 400 steps at batch 1024, 0.896 -> 0.930 M QPS (SM clock under load 1.60 ->
    // 1.67 GHz), measured twice back to back.  VQ_MMA_CLUSTER=0 switches it off.
    static const int cl_env = getenv("VQ_MMA_CLUSTER") ? atoi(getenv("VQ_MMA_CLUSTER")) : 2;
    p.cl = (cl_env == 2 && p.n_qt % 2 == 0) ? 2 : 1;
    // Bootstrap the per-query threshold from a sample of tiles when the scan is long enough to pay for two
    // extra (tiny) launches; the sample holds >= 4k tiles so that the k-th largest tile maximum is a strong
    // bound.  With several query tiles per store tile the list insertions dominate short scans as well
    // (125k rows x 1024 queries, one shard of 8: 0.26 ms cold vs the 0.07 ms of MMA work), so there the
    // pass is already worth a quarter of the tiles.
    int bt = sms > 4 * k ? sms : 4 * k;
    if (bt > kMaxBootTiles) bt = kMaxBootTiles;
    static const bool boot_on = getenv("VQ_MMA_BOOT") ? atoi(getenv("VQ_MMA_BOOT")) != 0 : true;
    const long long min_tiles = (p.n_qt >= 2 ? 4LL : 16LL) * bt;
    p.boot_tiles = (boot_on && p.nt && n_tiles >= min_tiles && bt >= k) ? bt : 0;
    p.boot_groups = p.boot_tiles ? (p.boot_tiles < (int)groups ? p.boot_tiles : (int)groups) : 0;
    // the sample tiles are spread evenly over the FULL tiles of the store: frames of a video sit next to
    // each other, a prefix of the store would only know the first few videos
    p.boot_mul = p.boot_tiles ? (int)((n / p.nt) / p.boot_tiles) : 1;
    if (p.boot_mul < 1) p.boot_mul = 1;
    p.smem = 1024 + (size_t)p.stages * stage_bytes + 512 + (size_t)2 * QT * kStage * 8 + (size_t)QT * 4 + 64;
    // candidate capacity is sized for the largest group count any batch <= b can get (the HNSW builder
    // reuses one workspace for a shrinking last batch)
    const int max_groups = sms > p.n_qt ? sms : p.n_qt;
    p.cap = p.groups * k;
    const size_t cand_bytes = align256((size_t)max_groups * QT * k * 4);
    size_t o = 0;
    p.off_tau = o;    o += align256((size_t)p.b_pad * 4);
    p.off_cnt = o;    o += align256((size_t)p.b_pad * 4);
    p.off_cand_s = o; o += cand_bytes;
    p.off_cand_r = o; o += cand_bytes;
    p.off_boot = o;   o += align256((size_t)kMaxBootTiles * p.b_pad * 4);
    p.off_qbf = o;    o += align256((size_t)p.b_pad * ld * 2);
    p.total = o;
    return p;
}

struct MmaWs {
    float* gtau; int* cnt; float* cand_s; int* cand_r; float* boot_max; __nv_bfloat16* qbf; float* qeps; XShared xs;
};
MmaWs carve(const MmaPlan& p, void* ws_v) {
    unsigned char* ws = (unsigned char*)ws_v;
    MmaWs w;
    w.qeps = nullptr;
    w.xs = XShared{nullptr, nullptr, 1, 0, 0, nullptr, 0};
    w.gtau = (float*)(ws + p.off_tau);
    w.cnt = (int*)(ws + p.off_cnt);
    w.cand_s = (float*)(ws + p.off_cand_s);
    w.cand_r = (int*)(ws + p.off_cand_r);
    w.boot_max = (float*)(ws + p.off_boot);
    w.qbf = (__nv_bfloat16*)(ws + p.off_qbf);
    w.qeps = nullptr;
    w.xs = XShared{nullptr, nullptr, 1, 0, 0, nullptr, 0};
    return w;
}

template <int KL, int NT, int MODE, int EPI = epi_warps(KL, MODE), bool CG2 = false>
cudaError_t launch_mma(const MmaPlan& p, const CUtensorMap& tmS, const __nv_bfloat16* qbf, const MmaWs& w, int n, int ld, int k,
                       int dbg, cudaStream_t stream) {
    auto kern = scan_mma_bf16_kernel<KL, NT, MODE, EPI, CG2>;
    constexpr bool BOOT = MODE == kModeBoot;
    static std::atomic<unsigned long long> attr_done{0};   // per instantiation, one bit per device
    if (vq_first_use_on_device(&attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        vq_mark_used(&attr_done);
    }
    const int grid = BOOT ? p.boot_groups * p.n_qt : p.grid;
    const int cl = (MODE == kModeList || MODE == kModeExact) ? p.cl : 1;
    // The sample pass runs a handful of tiles per CTA: two ring groups are enough, and with ~130 KB of shared memory instead of
    // 215 KB its CTAs share their SMs with the exact_finish CTAs of the step before (the two kernels become runnable at the
    // same moment — when the scan between them retires — and used to run one after the other).
    static const int boot_ring_env = getenv("VQ_BOOT_RING") ? atoi(getenv("VQ_BOOT_RING")) : 2;
    int stages = CG2 ? 2 * p.stages : p.stages;
    size_t smem = p.smem;
    if (BOOT && boot_ring_env > 0 && (boot_ring_env << p.grp_log2) < p.stages) {
        stages = boot_ring_env << p.grp_log2;
        smem -= (size_t)(p.stages - stages) * NT * 128;
    }
    return vq_launch_cluster(BOOT ? 1 : 3, cl, kern, dim3(grid), dim3(kThreadsFor(EPI)), smem, stream, tmS, qbf, w.gtau, n, ld, p.nkb, p.n_qt,
                             k, stages, (CG2 && p.nkb % (2 << p.grp_log2) == 0) ? p.grp_log2 + 1 : p.grp_log2, cl, BOOT ? p.boot_mul : 1, p.b_pad, BOOT ? w.boot_max : w.cand_s, w.cand_r, w.cnt, BOOT ? (p.boot_qmajor ? p.boot_tiles : 0) : p.cap, dbg,
                             (const float*)w.qeps, w.xs);
}

// [boot pass ->] main pass; gtau / cnt must have been reset by the caller's prologue kernel.
int run_scan(const MmaPlan& p, const void* store, int64_t n, int ld, const __nv_bfloat16* qbf, const MmaWs& w, int b, int k,
             cudaStream_t stream, int* launches, bool exact = false) {
    static const int dbg = getenv("VQ_MMA_DEBUG") ? atoi(getenv("VQ_MMA_DEBUG")) : 0;
    CUtensorMap tmS;
    if (!get_map_bf16(&tmS, store, (uint64_t)n, (uint64_t)ld, (uint32_t)p.nt)) {
        vq_set_error("scan_mma: cuTensorMapEncodeTiled failed");
        return VQ_ECUDA;
    }
    cudaError_t e;
    *launches = 1;
    if (p.boot_tiles) {
        // sample pass over boot_tiles full tiles spread over the store: per-tile maxima -> boot_max, then gtau
        const int n_boot = p.boot_tiles * p.nt;
        e = p.nt == 128 ? launch_mma<1, 128, kModeBoot>(p, tmS, qbf, w, n_boot, ld, k, dbg, stream)
                        : launch_mma<1, 64, kModeBoot>(p, tmS, qbf, w, n_boot, ld, k, dbg, stream);
        if (e != cudaSuccess) {
            vq_set_error("launch of scan_mma_bf16_kernel<boot> failed: %s", cudaGetErrorString(e));
            return VQ_ECUDA;
        }
        *launches = 2;
        if (!p.boot_qmajor) {
            e = vq_launch(2, boot_select_kernel, dim3((p.b_pad + 7) / 8), dim3(256), 0, stream, (const float*)w.boot_max, p.boot_tiles,
                          b, p.b_pad, k, w.gtau);
            if (e != cudaSuccess) {
                vq_set_error("launch of boot_select_kernel failed: %s", cudaGetErrorString(e));
                return VQ_ECUDA;
            }
            *launches = 3;
        }
    }
    CUtensorMap tmM = tmS;                       // main pass: half-height boxes when two CTAs share a tile
    if (p.cl > 1 && !get_map_bf16(&tmM, store, (uint64_t)n, (uint64_t)ld, (uint32_t)(p.nt / p.cl))) {
        vq_set_error("scan_mma: cuTensorMapEncodeTiled failed");
        return VQ_ECUDA;
    }
    vq_prof_begin(stream);
    if (exact) {
        if (p.nt == 128) {
            // 8 epilogue warps pay when the gather slow path runs for many of the tile's 128 queries; a batch of a few
            // queries is bound by the store stream and the four extra warps only add to launch and drain (measured: b <= 128
            // is 1-5 us slower with them, b >= 256 8-12 % faster)
            static const int epi_env = getenv("VQ_EXACT_EPI") ? atoi(getenv("VQ_EXACT_EPI")) : 0;
            const bool epi4 = epi_env ? epi_env == 4 : b <= 128;
            // CTA pairs issue cta_group::2 MMAs (one store tile per 256 queries through the L2 -> SM fabric) wherever two query
            // tiles share a cluster and the 8-warp epilogue keeps up; VQ_MMA_CG2=0 falls back to multicast + 1-CTA MMAs
            static const int cg2_env = getenv("VQ_MMA_CG2") ? atoi(getenv("VQ_MMA_CG2")) : 1;
            const bool cg2 = cg2_env && p.cl == 2 && !epi4 && k <= 32;
            if (cg2)
                e = k == 10 ? launch_mma<10, 128, kModeExact, 8, true>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                  : k <= 16 ? launch_mma<16, 128, kModeExact, 8, true>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                            : launch_mma<32, 128, kModeExact, 8, true>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream);
            else
            e = k == 10 ? (epi4 ? launch_mma<10, 128, kModeExact, 4>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                                : launch_mma<10, 128, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream))      // the contract's k: no sentinel slots to bubble through
              : k <= 16 ? launch_mma<16, 128, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
              : k <= 32 ? launch_mma<32, 128, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
              : k <= 64 ? launch_mma<64, 128, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                        : launch_mma<1, 128, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream);
        } else
            e = k <= 16 ? launch_mma<16, 64, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
              : k <= 32 ? launch_mma<32, 64, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
              : k <= 64 ? launch_mma<64, 64, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                        : launch_mma<1, 64, kModeExact>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream);
    } else if (p.nt == 128)
        e = k <= 16 ? launch_mma<16, 128, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
          : k <= 32 ? launch_mma<32, 128, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                    : launch_mma<64, 128, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream);
    else
        e = k <= 16 ? launch_mma<16, 64, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
          : k <= 32 ? launch_mma<32, 64, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream)
                    : launch_mma<64, 64, kModeList>(p, tmM, qbf, w, (int)n, ld, k, dbg, stream);
    vq_prof_end(stream);
    if (e != cudaSuccess) {
        vq_set_error("launch of scan_mma_bf16_kernel failed: %s", cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    return VQ_OK;
}

__global__ void reset_scan_state_kernel(float* gtau, int* cnt, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    vq_pdl_wait();
    vq_pdl_trigger();
    if (i < n) { gtau[i] = VQ_NEG_INF; cnt[i] = 0; }
}

}  // namespace

// scan_finish.cu: final selection out of the candidate buffers (+ optional exact fp32 re-score)
int vq_scan_finish_launch(int mode, const float* cand_s, const int* cand_r, const int* cand_cnt, int cap, int b, int k_sel,
                          const float* store_f32, int ld, int dim, const float* queries, int query_norm, float eps,
                          int k_out, float* out_scores, int* out_rows, int* out_bad, cudaStream_t stream);

bool vq_scan_mma_supported(int64_t n, int dim, int ld, int store_dtype, int b, int k) {
    (void)dim;
    if (store_dtype != VQ_BF16) return false;                 // kind::tf32 path for fp32 stores: not built yet
    if (ld % KB_ELEMS != 0 || ld > 768 || n < 1 || b < 1 || k < 1 || k > kMaxK) return false;
    return (b + QT - 1) / QT <= vq_num_sms();
}

// Scan with queries that are ALREADY unit-norm bf16 rows [b_pad, ld] (b_pad % 128 == 0, rows >= b zero):
// used by the HNSW builder, whose queries are the stored rows themselves.  The workspace of a batch b
// is valid for every smaller batch.
size_t vq_scan_mma_prepared_workspace(int64_t n, int ld, int b, int k) {
    return plan(n, ld, b, k).total + 256;
}
int vq_scan_mma_prepared(const void* store, int64_t n, int ld, const void* qbf, int b, int k, float* out_scores,
                         int32_t* out_rows, void* ws_v, size_t ws_bytes, cudaStream_t stream) {
    if (!vq_scan_mma_supported(n, ld, ld, VQ_BF16, b, k)) {
        vq_set_error("scan_mma_prepared: unsupported shape n=%lld ld=%d b=%d k=%d", (long long)n, ld, b, k);
        return VQ_EUNSUPPORTED;
    }
    const MmaPlan p = plan(n, ld, b, k);
    if (ws_bytes < p.off_qbf) {                                // the caller supplies the bf16 queries
        vq_set_error("scan_mma_prepared: workspace too small (%zu < %zu)", ws_bytes, p.off_qbf);
        return VQ_EWORKSPACE;
    }
    const MmaWs w = carve(p, ws_v);
    if (vq_launch(0, reset_scan_state_kernel, dim3((p.b_pad + 255) / 256), dim3(256), 0, stream, w.gtau, w.cnt, p.b_pad) != cudaSuccess) {
        vq_set_error("launch of reset_scan_state_kernel failed");
        return VQ_ECUDA;
    }
    int nl = 0;
    int rc = run_scan(p, store, n, ld, (const __nv_bfloat16*)qbf, w, b, k, stream, &nl);
    if (rc) return rc;
    return vq_scan_finish_launch(0, w.cand_s, w.cand_r, w.cnt, p.cap, b, k, nullptr, ld, ld, nullptr, VQ_NORM_NONE, 0.f, k,
                                 out_scores, out_rows, nullptr, stream);
}

size_t vq_scan_mma_workspace(int64_t n, int ld, int store_dtype, int b, int k) {
    if (store_dtype != VQ_BF16 || b < 1 || k < 1 || k > kMaxK || n < 1 || ld > 768) return 0;
    return plan(n, ld, b, k).total + 256;
}

// queries: raw fp32 [b, dim]; normalisation (query_norm) is fused into the bf16 conversion.
// store_f32 == NULL: plain top-k of the bf16 scores.  store_f32 != NULL: two-stage exact search — the
// scan selects k_sel candidates, they are re-scored from the fp32 copy, the best k_out are returned and
// out_bad[q] says whether the result could NOT be certified (see vq_search_two_stage).
int vq_scan_mma_run(const void* store, int64_t n, int dim, int ld, int store_dtype, const float* queries, int query_norm,
                    int b, int k_sel, const float* store_f32, float eps, int k_out, float* out_scores, int32_t* out_rows,
                    int32_t* out_bad, void* ws_v, size_t ws_bytes, cudaStream_t stream, int* launches) {
    if (!vq_scan_mma_supported(n, dim, ld, store_dtype, b, k_sel)) {
        vq_set_error("scan_mma: unsupported shape n=%lld dim=%d ld=%d b=%d k=%d", (long long)n, dim, ld, b, k_sel);
        return VQ_EUNSUPPORTED;
    }
    const MmaPlan p = plan(n, ld, b, k_sel);
    if (ws_bytes < p.total) {
        vq_set_error("scan_mma: workspace too small (%zu < %zu)", ws_bytes, p.total);
        return VQ_EWORKSPACE;
    }
    const MmaWs w = carve(p, ws_v);
    if (vq_launch(0, prep_queries_bf16_kernel, dim3((p.b_pad + 7) / 8), dim3(256), 0, stream, queries, b, dim, dim, w.qbf, ld,
                  p.b_pad, query_norm, w.gtau, w.cnt, (const float*)nullptr, (const float*)nullptr, (float*)nullptr, (float*)nullptr, 0, (int*)nullptr, 0, 0.f) != cudaSuccess) {
        vq_set_error("launch of prep_queries_bf16_kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        return VQ_ECUDA;
    }
    int nl = 0;
    int rc = run_scan(p, store, n, ld, w.qbf, w, b, k_sel, stream, &nl);
    if (rc) return rc;
    rc = vq_scan_finish_launch(store_f32 ? 1 : 0, w.cand_s, w.cand_r, w.cnt, p.cap, b, k_sel, store_f32, ld, dim, queries, query_norm,
                               eps, k_out, out_scores, out_rows, out_bad, stream);
    if (rc) return rc;
    *launches = 2 + nl;
    return VQ_OK;
}

// Collect pass (resolves queries the two-stage certificate rejected): every row whose bf16-operand score
// reaches thresholds[q] is gathered (at most `cap` per query), ALL of them are re-scored exactly from the
// fp32 copy and the best k by exact score are returned.  With thresholds[q] = (a lower bound of the exact
// k-th best score) - score_eps the gathered set contains the exact top-k by construction.
// out_overflow[q] = 1 if more than `cap` rows reached the threshold (result incomplete).
size_t vq_scan_mma_collect_workspace(int64_t n, int ld, int b, int cap) {
    const MmaPlan p = plan(n, ld, b, 1);
    return p.off_cand_s + 2 * align256((size_t)p.b_pad * cap * 4) + align256((size_t)p.b_pad * ld * 2) +
           align256((size_t)kMaxBootTiles * p.b_pad * 4) + align256((size_t)p.b_pad * 4) + 256;
}
// thresholds == NULL: the threshold of every query is derived from the store itself — a boot pass scores
// min(256, tiles) sample tiles and the k-th largest tile maximum (reached by k distinct rows) is used, so at
// least k rows are gathered.  The exact top-k of THOSE rows is then a set of k real rows: its k-th exact
// score is a valid lower bound of the k-th best of any store that contains them (large-k search, step 1).
int vq_exact_finish_launch(const float* cand_s, const int* cand_r, const int* cand_cnt, int cap, int b, int k_sel,
                           const float* qeps, const float* store_f32, int ld, int dim, const float* queries, int query_norm,
                           int k_out, float* out_scores, int* out_rows, int* out_overflow, int* out_stats, cudaStream_t stream);
// bounds != NULL (the store's {max |x^|, max |x^ - x|}, as for the exact mode): the final stage re-scores only the best
// candidates by bf16 score and those within the per-query rounding bound of the k-th (exact_finish) instead of
// everything gathered.
int vq_scan_mma_collect(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld, const float* queries,
                        int query_norm, int b, const float* thresholds, int cap, const float* bounds, int k, float* out_scores,
                        int32_t* out_rows, int32_t* out_overflow, void* ws_v, size_t ws_bytes, cudaStream_t stream, int* launches) {
    if (!vq_scan_mma_supported(n, dim, ld, VQ_BF16, b, 1) || cap < k) {
        vq_set_error("scan_mma_collect: unsupported shape n=%lld dim=%d ld=%d b=%d cap=%d k=%d", (long long)n, dim, ld, b, cap, k);
        return VQ_EUNSUPPORTED;
    }
    MmaPlan p = plan(n, ld, b, 1);
    if (ws_bytes < vq_scan_mma_collect_workspace(n, ld, b, cap) - 256) {
        vq_set_error("scan_mma_collect: workspace too small");
        return VQ_EWORKSPACE;
    }
    unsigned char* ws = (unsigned char*)ws_v;
    MmaWs w;
    w.qeps = nullptr;
    w.xs = XShared{nullptr, nullptr, 1, 0, 0, nullptr, 0};
    w.gtau = (float*)(ws + p.off_tau);
    w.cnt = (int*)(ws + p.off_cnt);
    const size_t cand_bytes = align256((size_t)p.b_pad * cap * 4);
    w.cand_s = (float*)(ws + p.off_cand_s);
    w.cand_r = (int*)(ws + p.off_cand_s + cand_bytes);
    w.qbf = (__nv_bfloat16*)(ws + p.off_cand_s + 2 * cand_bytes);
    w.boot_max = (float*)(ws + p.off_cand_s + 2 * cand_bytes + align256((size_t)p.b_pad * ld * 2));
    float* qeps = (float*)(ws + p.off_cand_s + 2 * cand_bytes + align256((size_t)p.b_pad * ld * 2) + align256((size_t)kMaxBootTiles * p.b_pad * 4));
    p.cap = cap;
    if (vq_launch(0, prep_queries_bf16_kernel, dim3((p.b_pad + 7) / 8), dim3(256), 0, stream, queries, b, dim, dim, w.qbf, ld,
                  p.b_pad, query_norm, w.gtau, w.cnt, thresholds, bounds, bounds ? qeps : (float*)nullptr, (float*)nullptr, 0, (int*)nullptr, 0, 0.f) != cudaSuccess) {
        vq_set_error("launch of prep_queries_bf16_kernel failed");
        return VQ_ECUDA;
    }
    static const int dbg = 0;
    CUtensorMap tmS;
    if (!get_map_bf16(&tmS, store_bf16, (uint64_t)n, (uint64_t)ld, (uint32_t)p.nt)) {
        vq_set_error("scan_mma: cuTensorMapEncodeTiled failed");
        return VQ_ECUDA;
    }
    cudaError_t e;
    *launches = 3;
    if (thresholds == nullptr) {
        const long long full_tiles = n / p.nt;
        p.boot_tiles = (int)(full_tiles < kMaxBootTiles ? full_tiles : kMaxBootTiles);
        if (p.boot_tiles < k) {
            vq_set_error("scan_mma_collect: automatic thresholds need at least k=%d full tiles of %d rows (n=%lld)", k, p.nt, (long long)n);
            return VQ_EUNSUPPORTED;
        }
        p.boot_groups = p.boot_tiles < p.groups ? p.boot_tiles : p.groups;
        p.boot_mul = (int)(full_tiles / p.boot_tiles);
        const int n_boot = p.boot_tiles * p.nt;
        e = p.nt == 128 ? launch_mma<1, 128, kModeBoot>(p, tmS, w.qbf, w, n_boot, ld, k, dbg, stream)
                        : launch_mma<1, 64, kModeBoot>(p, tmS, w.qbf, w, n_boot, ld, k, dbg, stream);
        if (e == cudaSuccess)
            e = vq_launch(2, boot_select_kernel, dim3((p.b_pad + 7) / 8), dim3(256), 0, stream, (const float*)w.boot_max, p.boot_tiles,
                          b, p.b_pad, k, w.gtau);
        if (e != cudaSuccess) {
            vq_set_error("launch of the boot pass failed: %s", cudaGetErrorString(e));
            return VQ_ECUDA;
        }
        *launches = 5;
    }
    vq_prof_begin(stream);
    e = p.nt == 128 ? launch_mma<1, 128, kModeCollect>(p, tmS, w.qbf, w, (int)n, ld, 1, dbg, stream)
                    : launch_mma<1, 64, kModeCollect>(p, tmS, w.qbf, w, (int)n, ld, 1, dbg, stream);
    vq_prof_end(stream);
    if (e != cudaSuccess) {
        vq_set_error("launch of scan_mma_bf16_kernel<collect> failed: %s", cudaGetErrorString(e));
        return VQ_ECUDA;
    }
    int rc;
    const int k_sel = k <= 16 ? 32 : (k + (k / 2 > 22 ? k / 2 : 22));
    if (bounds != nullptr && k_sel <= 512 && cap >= k_sel)
        rc = vq_exact_finish_launch(w.cand_s, w.cand_r, w.cnt, cap, b, k_sel, qeps, store_f32, ld, dim, queries, query_norm, k,
                                    out_scores, out_rows, out_overflow, nullptr, stream);
    else
        rc = vq_scan_finish_launch(2, w.cand_s, w.cand_r, w.cnt, cap, b, k, store_f32, ld, dim, queries, query_norm, 0.f, k,
                                   out_scores, out_rows, out_overflow, stream);
    if (rc) return rc;
    return VQ_OK;
}

// ---------------------------------------------------------------------------------------- exact search
// exact_finish.cu: selection + fp32 re-score of the candidates the exact-mode scan gathered
int vq_exact_finish_launch(const float* cand_s, const int* cand_r, const int* cand_cnt, int cap, int b, int k_sel,
                           const float* qeps, const float* store_f32, int ld, int dim, const float* queries, int query_norm,
                           int k_out, float* out_scores, int* out_rows, int* out_overflow, int* out_stats, cudaStream_t stream);

namespace {
// Exact-mode plan: the scan plan for lists of k entries with a candidate buffer of `cap` slots per query
// (appends of the whole scan, not just list survivors) and the per-query eps array.
struct ExactPlan { MmaPlan p; int k_sel, ms, boot_T, refresh_ns; size_t off_eps, off_cmax, off_arrived; };
ExactPlan plan_exact(int64_t n, int ld, int b, int k) {
    ExactPlan x;
    MmaPlan& p = x.p;
    p = plan(n, ld, b, k);
    const int sms = vq_num_sms();
    const long long n_tiles = (n + p.nt - 1) / p.nt;
    // rows a query may gather: the rows within 2 eps of its k-th best plus the transient while the bound warms up
    // (measured: a few hundred on iid data, 1-2 thousand on tightly clustered data)
    int cap = p.n_qt <= 16 ? 4096 : 2048;        // (config 5, 32 query tiles over 12.5M clustered rows: 700 rows per query on average, 235 of 4096 queries beyond 1024)
    static const int cap_env = getenv("VQ_EXACT_CAP") ? atoi(getenv("VQ_EXACT_CAP")) : 0;
    if (cap_env > 0) cap = cap_env;
    const long long n8 = (n + 7) / 8 * 8;
    if (cap > n8) cap = (int)n8;                  // a row is gathered at most once per query: cannot overflow
    static const int boot_env = getenv("VQ_EXACT_BOOT") ? atoi(getenv("VQ_EXACT_BOOT")) : -1;
    static const int ms_env = getenv("VQ_EXACT_SUB") ? atoi(getenv("VQ_EXACT_SUB")) : 0;
    int bt = sms > 4 * k ? sms : 4 * k;           // bootstrap sample: max(#SMs, 4k) tiles per query tile
    if (bt > kMaxPublished) bt = kMaxPublished;
    // Where the threshold bootstrap runs: as a separate sample pass + boot_select in front of the scan (like the list
    // mode).  The alternative — INSIDE the scan: the first boot_T tiles of every CTA only publish their maxima, the
    // CTAs of a query tile meet at a counter, and those tiles are scanned again at the end; no extra launches — was
    // built and measured (VQ_EXACT_INSIDE=1 selects it) and is NOT the default: with three steps in flight the two
    // small launches hide behind the neighbouring steps, while the in-kernel barrier stalls every CTA right after
    // its first tiles and the re-scan is paid in full (1M rows x 1024 queries: 1.039 ms separate vs 1.06-1.09 ms inside;
    // 125k-row shard: 0.225 vs 0.254 ms; batch 1: 0.167 vs 0.182 ms).
    static const int inside_env = getenv("VQ_EXACT_INSIDE") ? atoi(getenv("VQ_EXACT_INSIDE")) : -1;
    const bool inside = inside_env > 0 || (inside_env < 0 && boot_env > 0);
    // with the bootstrap inside, short stores run fewer groups so that every CTA has 16 tiles to take its share from
    if (inside && (long long)p.groups * 16 > n_tiles) {
        p.groups = (int)(n_tiles / 16 > 0 ? n_tiles / 16 : 1);
        p.grid = p.groups * p.n_qt;
    }
    const long long per_cta = n_tiles / p.groups;
    int T = 0;
    if (inside) {
        p.boot_tiles = p.boot_groups = 0;
        p.boot_mul = 1;
        T = boot_env > 0 ? boot_env : (bt + p.groups - 1) / p.groups;
        const long long t_max = per_cta / 16 > 1 ? per_cta / 16 : (per_cta >= 4 ? 1 : 0);
        if (T > t_max) T = (int)t_max;
    } else {
        const long long full_tiles = n / p.nt;
        int bs = bt < kMaxBootTiles ? bt : kMaxBootTiles;
        if (bs > full_tiles / 2) bs = (int)(full_tiles / 2);
        if (bs >= k && cap < n8 && boot_env != 0) {
            p.boot_tiles = bs;
            p.boot_qmajor = k <= 16 ? 1 : 0;   // (the in-scan selection keeps two values per lane: not enough for a larger k — boot_select_kernel ranks all)
            p.boot_groups = bs < p.groups ? bs : p.groups;
            p.boot_mul = (int)(full_tiles / bs);
            if (p.boot_mul < 1) p.boot_mul = 1;
        } else {
            p.boot_tiles = p.boot_groups = 0;
            p.boot_mul = 1;
        }
    }
    // sub-streams per CTA: as many published maxima per query as the bootstrap sample holds tiles, so that their
    // k-th largest is as strong a bound as the k-th largest tile maximum of that sample; with the bootstrap inside
    // the kernel every sub-stream must see one of the boot_T tiles
    {
        int ms = ms_env > 0 ? ms_env : (bt + p.groups - 1) / p.groups;
        if (ms > kMaxPublished / p.groups) ms = kMaxPublished / p.groups;
        if (inside && T > 0 && ms > T) ms = T;
        x.ms = ms < 1 ? 1 : (ms > kMaxSub ? kMaxSub : ms);
    }
    if (inside && ((long long)p.groups * x.ms < k || cap >= n8)) T = 0;       // no usable bound / nothing to protect
    x.boot_T = T;
    if (x.boot_T == 0 && p.boot_tiles == 0 && cap < n8 && (long long)p.groups * k > cap / 2) {
        p.groups = cap / (2 * k) > 0 ? cap / (2 * k) : 1;     // no bound at all: groups * k unconditional appends
        p.grid = p.groups * p.n_qt;
    }
    // Periodic refresh of the cooperative bound (warp 3 of every CTA, each for its share of the tile's queries): first
    // sleep 4 us, backing off to 64 us while nothing changes.  VQ_EXACT_REFRESH_NS overrides (0 = off).
    static const int refresh_env = getenv("VQ_EXACT_REFRESH_NS") ? atoi(getenv("VQ_EXACT_REFRESH_NS")) : -1;
    // A single query tile (HBM-bound scan, 148 CTAs that each see 1/148 of the store) refreshes at a slower cadence: at
    // 4 us the refresher cost batch 1 / 32 at 1M rows 0.166 -> 0.184 ms of kernel time, but without it the CTA-local
    // bounds let the gather grow with the store (2900 rows per query at 1M, buffer overflow at 12.5M rows per GPU).
    x.refresh_ns = refresh_env >= 0 ? refresh_env : (p.n_qt >= 2 ? 4000 : 16000);
    p.cap = cap;
    x.k_sel = k <= 16 ? 32 : (k + (k / 2 > 22 ? k / 2 : 22));
    const size_t cand_bytes = align256((size_t)p.b_pad * cap * 4);
    size_t o = p.off_cand_s;
    p.off_cand_r = o + cand_bytes;
    o += 2 * cand_bytes;
    p.off_boot = o;    o += align256((size_t)kMaxBootTiles * p.b_pad * 4);     // tile maxima of the separate sample pass (one query tile)
    p.off_qbf = o;     o += align256((size_t)p.b_pad * ld * 2);
    x.off_eps = o;     o += align256((size_t)p.b_pad * 4);
    x.off_cmax = o;    o += align256((size_t)p.b_pad * sms * kMaxSub * 4 / (p.n_qt > 0 ? p.n_qt : 1) + 1024);
    x.off_arrived = o; o += align256((size_t)p.n_qt * 8 * 4);
    p.total = o;
    return x;
}
}  // namespace

constexpr int kMaxExactK = 128;          // 64 < k <= 128: listless exact mode (needs a bootstrap bound)
// k beyond the register lists is served only where a bound exists before the first row is gathered
bool vq_scan_mma_exact_supported(int64_t n, int ld, int b, int k) {
    if (b < 1 || k < 1 || k > kMaxExactK || n < 1 || ld > 768 || ld % KB_ELEMS != 0) return false;
    if ((b + QT - 1) / QT > vq_num_sms()) return false;
    if (k <= kMaxK) return true;
    // Listless mode (64 < k <= 128): the bound is only the k-th largest of <= 160 published maxima, which at bootstrap
    // time cover a small sample — measured on 1.25M x 768, k = 100: tens of thousands of rows pass before the bound
    // tightens and every gather overflows.  Served only where that cannot happen (the buffer holds a tenth of the
    // store) unless VQ_EXACT_LISTLESS=1; larger stores go through the two-pass route (vq_search_collect).
    static const bool listless_env = getenv("VQ_EXACT_LISTLESS") ? atoi(getenv("VQ_EXACT_LISTLESS")) != 0 : false;
    const ExactPlan x = plan_exact(n, ld, b, k);
    if (!(x.boot_T > 0 || x.p.boot_tiles > 0 || x.p.cap >= n)) return false;
    return listless_env || (long long)x.p.cap * 10 >= n;
}
size_t vq_scan_mma_exact_workspace(int64_t n, int ld, int b, int k) {
    if (!vq_scan_mma_exact_supported(n, ld, b, k)) return 0;
    return plan_exact(n, ld, b, k).p.total + 256;
}

// Single-pass exact top-k: prep (normalise + bf16 + per-query eps) -> [boot -> boot_select] -> exact-mode
// scan (list + gather) -> exact_finish (select, fp32 re-score, final top-k).  Exact by construction; the only
// failure mode is a gather that overflows (mass ties), reported per query in out_overflow.
int vq_scan_mma_exact(const void* store_bf16, const float* store_f32, int64_t n, int dim, int ld, const float* queries,
                      int query_norm, int b, int k, const float* bounds, float* out_scores, int32_t* out_rows,
                      int32_t* out_overflow, int32_t* out_stats, void* ws_v, size_t ws_bytes, cudaStream_t stream, int* launches) {
    if (!vq_scan_mma_exact_supported(n, ld, b, k)) {
        vq_set_error("exact search: unsupported shape n=%lld dim=%d ld=%d b=%d k=%d", (long long)n, dim, ld, b, k);
        return VQ_EUNSUPPORTED;
    }
    const ExactPlan x = plan_exact(n, ld, b, k);
    const MmaPlan& p = x.p;
    if (ws_bytes < p.total) {
        vq_set_error("exact search: workspace too small (%zu < %zu)", ws_bytes, p.total);
        return VQ_EWORKSPACE;
    }
    MmaWs w = carve(p, ws_v);
    w.qeps = (float*)((unsigned char*)ws_v + x.off_eps);
    w.xs = XShared{(float*)((unsigned char*)ws_v + x.off_cmax), (int*)((unsigned char*)ws_v + x.off_arrived), x.ms, x.boot_T, x.refresh_ns,
                   p.boot_qmajor ? (const float*)w.boot_max : nullptr, p.boot_qmajor ? p.boot_tiles : 0};
    if (vq_launch(0, prep_queries_bf16_kernel, dim3((p.b_pad + 7) / 8), dim3(256), 0, stream, queries, b, dim, dim, w.qbf, ld,
                  p.b_pad, query_norm, w.gtau, w.cnt, (const float*)nullptr, bounds, w.qeps, w.xs.cmax, p.groups * w.xs.ms, w.xs.arrived, p.n_qt, 0.f) != cudaSuccess) {
        vq_set_error("launch of prep_queries_bf16_kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        return VQ_ECUDA;
    }
    int nl = 0;
    int rc = run_scan(p, store_bf16, n, ld, w.qbf, w, b, k, stream, &nl, true);
    if (rc) return rc;
    rc = vq_exact_finish_launch(w.cand_s, w.cand_r, w.cnt, p.cap, b, x.k_sel, w.qeps, store_f32, ld, dim, queries, query_norm,
                                k, out_scores, out_rows, out_overflow, out_stats, stream);
    if (rc) return rc;
    *launches = 2 + nl;
    return VQ_OK;
}
