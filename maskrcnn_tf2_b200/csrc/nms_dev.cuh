// nms_dev.cuh -- device helpers of the NMS kernel (nms.cu): warp ballots, mbarrier + distributed-shared-memory PTX
// wrappers, debug cycle counters, output epilogue.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace mrcnn {

#ifdef MRCNN_NMS_PROFILE   // debug build only (scripts/profile_nms.py): per-phase cycle counters of cluster 0
static __device__ long long g_nms_prof[8];  // one copy per translation unit
#define PROF_DECL long long p_t0 = clock64(), p_acc[4] = {0, 0, 0, 0}; int p_tiles = 0
#define PROF_MARK(i) do { const long long p_now = clock64(); p_acc[i] += p_now - p_t0; p_t0 = p_now; } while (0)
#define PROF_TILE ++p_tiles
#define PROF_DUMP do { if (blockIdx.x == 0 && threadIdx.x == 0) { for (int q = 0; q < 4; ++q) g_nms_prof[q] = p_acc[q]; \
                       g_nms_prof[4] = p_tiles; g_nms_prof[5] = nkept; } } while (0)
static __device__ long long g_nms_profw[8];  // far warp 1 of CTA 0: prologue, wait release, far, barrier, send far
#define PROFW_DECL long long w_t0 = clock64(), w_acc[7] = {0, 0, 0, 0, 0, 0, 0}
#define PROFW_COUNT ++w_acc[6]
#define PROFW_MARK(i) do { const long long w_now = clock64(); w_acc[i] += w_now - w_t0; w_t0 = w_now; } while (0)
static __device__ long long g_nms_tl[16];    // spare slots for ad-hoc clock64 stamps
static __device__ long long g_nms_timeline[128 * 8];    // clock64 stamps of CTA 0, [tile][event], scripts/profile_nms.py
#define PROF_TL(tile, ev) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile) < 128) g_nms_timeline[(tile) * 8 + (ev)] = clock64(); } while (0)
static __device__ unsigned long long g_nms_fallbacks;   // exact-division fallbacks taken (all warps, all CTAs)
#define PROF_FALLBACK do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_nms_fallbacks, 1ull); } while (0)
#define PROFW_DUMP do { if (blockIdx.x == 0 && threadIdx.x == 32) { for (int q = 0; q < 7; ++q) g_nms_profw[q] = w_acc[q]; } } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_TILE
#define PROF_DUMP
#define PROFW_DECL
#define PROFW_MARK(i)
#define PROFW_DUMP
#define PROFW_COUNT
#define PROF_FALLBACK
#define PROF_TL(tile, ev)
#endif

constexpr int kTile = 64;
constexpr int kSmallN = 256;                        // fused-ordering problems up to this many candidates skip the pipeline
// Pipeline depth: far(u) covers the boxes kept in tiles <= u - kDepth, so the far warps run kDepth tiles ahead of the
// resolver; the tiles in between are covered by kDepth - 1 precomputed cross blocks.
constexpr int kDepth = 2;
constexpr int kRows = kDepth * kTile;              // rows per tile: diag + (kDepth - 1) cross blocks
// Per-tile slots (rows, far partials, mbarriers, kept counts) are rings of kRing entries, tile t uses slot t % kRing.
// What the ring depth has to cover (cluster size > 1): a sender writes the set of tile v into every peer once its OWN
// resolver has finished tile v - 3 (rows, sent one tile ahead) or v - 2 (far partials).  For that the sender needed the
// peer's far partials of tile v - 3, which the peer produced after ITS resolver finished tile v - 5: the peer is then
// at tile v - 4 or later, and slot v % 8 was last read (and its mbarrier phase last completed) for tile v - 8.
constexpr int kRing = 8;
constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;

constexpr int kMaxFarSrc = 8 * 12;                 // far partials per tile: cluster size (<= 8) x far warps per CTA

constexpr int kCompactSlack = 32;                  // the far warps' private kept lists: sum of the roundings-up (<= 2 each)
static size_t nms_smem_bytes(int M, int max_out, bool compact) {
    const size_t cap = (size_t)(max_out < M ? max_out : M);
    return (size_t)M * (sizeof(float4) + sizeof(float)) + cap * sizeof(int32_t) +
           (compact ? (cap + kCompactSlack) * (sizeof(float4) + sizeof(float)) : 0);
}
// Fused candidate ordering (single-CTA problems): the kernel first builds (score key, ~index) composites of the
// candidates in shared memory, sorts them and stages the boxes in candidate order itself.
// Layout: [composites: sort_n x 8 B][candidate -> original row: M x 4 B][union(NMS arrays, sort exchange buffers)]
static int nms_sort_n(int M) { int p = 32; while (p < M) p <<= 1; return p; }
static size_t nms_fused_smem_bytes(int M, int max_out, bool compact) {
    const int sort_n = nms_sort_n(M);
    const size_t xch = sort_n >= 1024 ? block_sort_xch_bytes(sort_n / 1024) : 0;
    const size_t body = nms_smem_bytes(M, max_out, compact);
    return (size_t)sort_n * 8 + align_up((size_t)M * 4, 16) + (body > xch ? body : xch);
}

__device__ __forceinline__ uint64_t ballot64(bool lo, bool hi) {
    return (uint64_t)__ballot_sync(0xffffffffu, lo) | ((uint64_t)__ballot_sync(0xffffffffu, hi) << 32);
}

// ---- mbarrier / distributed-shared-memory primitives (sm_90+ PTX) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arm(uint32_t bar, uint32_t tx_bytes) {  // one arrival + expected bytes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
}
// CTA-scope acquire (the default) on purpose: the awaited data is written into THIS CTA's shared memory by the peers'
// st.async / bulk copies, whose complete_tx makes it visible; a cluster-scope acquire makes ptxas add CCTL.IVALL (an
// L1 invalidate, ~700 cycles of long-scoreboard stall per wait -- half of the resolver's time when it was measured)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// the same with a suspend-time hint: the warp is parked by the hardware until the phase completes (or the hint, in ns,
// expires) instead of re-issuing try_wait -- a dozen warps polling in a tight loop take the issue slots of the warps
// that work (measured in nms_sweep_kernel: every working warp ran at 3-5 cycles per instruction)
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!done);
}
// 8-byte store into a peer CTA's shared memory that completes 8 transaction bytes on the peer's mbarrier
__device__ __forceinline__ void st_async_u64(uint32_t peer_addr, uint64_t v, uint32_t peer_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(peer_addr), "l"(v), "r"(peer_bar) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from this CTA's shared memory into a peer's, completing `bytes` transaction
// bytes on the peer's mbarrier (SASS UBLKCP); both addresses 16-byte aligned
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t peer_dst, uint32_t local_src, uint32_t bytes, uint32_t peer_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(peer_dst), "r"(local_src), "r"(bytes), "r"(peer_bar) : "memory");
}
// bulk copy of `bytes` (multiple of 16) from global memory into this CTA's shared memory, completing on its mbarrier
__device__ __forceinline__ void bulk_copy_from_global(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// completes `bytes` of a (possibly remote) mbarrier's transaction count without moving data: the stand-in for a send
// that is no longer needed
__device__ __forceinline__ void mbar_complete_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
    asm volatile("mbarrier.complete_tx.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void or_into(unsigned long long* word, uint64_t bits) {  // two native 32-bit shared atomics
    unsigned* w = reinterpret_cast<unsigned*>(word);
    if ((unsigned)bits) atomicOr(w, (unsigned)bits);
    if ((unsigned)(bits >> 32)) atomicOr(w + 1, (unsigned)(bits >> 32));
}

// fixed-size padded outputs, no host round trip; `sel` = kept candidate positions (shared memory), `total` of them
// `orig` (shared memory, fused ordering) or epi.orig_idx (global) maps a candidate position to its input row
__device__ __forceinline__ void nms_write_outputs(const NmsEpilogue& epi, const float4* __restrict__ bx, int b, int M,
                                                  int max_out, int total, const int32_t* sel, const int32_t* orig,
                                                  int tid, int nthreads) {
    auto row_of = [&](int pos) { return orig ? orig[pos] : (epi.orig_idx ? epi.orig_idx[(size_t)b * M + pos] : pos); };
    if (epi.mode == 0) {
        for (int r = tid; r < max_out; r += nthreads) {
            int32_t v = -1;
            if (r < total) v = row_of(sel[r]);
            epi.keep[(size_t)b * max_out + r] = v;
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    } else if (epi.mode == 1) {  // ProposalLayer.nms L:227-230: gather + zero pad
        for (int r = tid; r < max_out; r += nthreads) {
            epi.proposals[(size_t)b * max_out + r] = (r < total) ? __ldg(bx + sel[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (epi.keep) epi.keep[(size_t)b * max_out + r] = (r < total) ? sel[r] : -1;
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    } else {  // refine_detections L:494-500: [y1,x1,y2,x2,class,score] rows + zero pad
        for (int r = tid; r < max_out; r += nthreads) {
            float* o = epi.detections + ((size_t)b * max_out + r) * 6;
            if (r < total) {
                const int i = row_of(sel[r]);
                const float4 v = epi.refined[(size_t)b * epi.N + i];
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
                o[4] = (float)epi.class_ids[(size_t)b * epi.N + i];
                o[5] = epi.scores[(size_t)b * epi.N + i];
                if (epi.det_boxes) epi.det_boxes[(size_t)b * max_out + r] = v;
            } else {
                o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f;
                if (epi.det_boxes) epi.det_boxes[(size_t)b * max_out + r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    }
}

// nms_sweep.cu: the cluster kernel for ordered candidate lists; -1 = the problem does not fit it (keep nms_lazy_kernel)
__attribute__((visibility("hidden"))) int launch_nms_sweep(const float4* boxes_sorted, const int32_t* valid, int B, int M,
                                                           int max_out, float thr, const NmsEpilogue& epi,
                                                           cudaStream_t stream, bool unit_boxes);

}  // namespace mrcnn
