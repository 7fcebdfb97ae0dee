// topk_cluster.cu -- exact per-image top-K (K <= 8192, TF TopKV2 order: value desc, index asc) as ONE kernel.
// Replaces tf.nn.top_k at mrcnn_layers.py:246, the per-image gathers at L:247-250 and (fused epilogue) the box
// decode / clip of utils.py:830-869.
//
// A thread-block cluster of 1..16 CTAs (1024 threads each; 10 at batch 8 on a B200: sizes need not be powers of two) owns
// an image.  The scores are dealt to the CTAs in units of one warp-wide 128-bit load (64 anchors of rpn_probs), round
// robin -- NOT in contiguous slices: the anchors are level-major and a trained RPN puts most of its top scores on the
// coarse levels at the end of the list (5000 of the 6000 winners sat in the last eighth on the synthetic COCO-shape
// input), which would leave one CTA with all the candidates.  Every CTA streams its share once from HBM (eight 128-bit
// loads per thread in flight: the pass is bound by bytes in flight) and everything else stays on chip:
//   1. radix select on the 64-bit composite (order-preserving score key, ~index), 12 + 12 + 8 bits of the key and, only
//      if more than 8192 candidates share all 32 key bits (tie floods, e.g. saturated probabilities), 12 + 12 + 8 bits
//      of the index -- lower index first, TopKV2's tie rule, by the same mechanism: per-CTA shared-memory histogram,
//      whose 64 coarse group sums every CTA pushes into every peer BEFORE the level's one cluster barrier; after it a
//      CTA finds the group of the K-th element from local memory and reads only that group's bins from its peers
//      (2 KB; summing whole histograms through distributed shared memory cost 11 k cycles per level).  Identical integer
//      sums everywhere, so no broadcast of the result; histograms and coarse sums are double buffered.  The descent stops
//      as soon as "above + boundary bin" fits 8192 entries;
//   2. compaction of the CTA's candidates (composite >= the resolved prefix) into its own shared-memory list;
//   3. one cluster barrier, then every CTA copies ALL candidate lists (<= 8192 entries, unsorted) into its shared memory
//      and ranks by counting: order-preserving bins over [smallest candidate key, largest key], one histogram + scan for
//      "candidates in higher bins", plus the greater composites inside the own bin (a handful on real score
//      distributions).  No sorting network, no sorted-list exchange, no binary searches -- they were 45 k of the kernel's
//      79 k cycles.  A bin with more than 256 entries (tie floods) or a select that went into the index bits sends the
//      whole list through the register-blocked bitonic network instead, in every CTA, same decision everywhere;
//   4. epilogue, every CTA for its own candidates at their global rank: indices / values, and for ProposalLayer the
//      gather + std-dev scale + decode + clip of the winners.
// Cluster barriers per launch: one per radix level (one on COCO-shape RPN scores) + one before the gather + the exit
// barrier (split arrive / wait, so it costs no waiting).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mrcnn {

#ifdef MRCNN_NMS_PROFILE   // debug build only (scripts/profile_topk.py): clock64 stamps of thread 0 of CTA 0
static __device__ long long g_tk_timeline[32];
#define TK_TL(ev) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tk_timeline[ev] = clock64(); } while (0)
#else
#define TK_TL(ev)
#endif

constexpr int kTkThreads = 1024;
constexpr size_t kTkListBytes = (size_t)kMaxSort * 8;                 // the CTA's own candidates
constexpr size_t kTkScratchBytes = block_sort_xch_bytes(8);           // sort exchange buffers / gathered peer lists
constexpr int kTkBinLimit = 256;                                      // largest bin the counting rank accepts
constexpr int kTkCacheKeys = 30720;                                   // key cache: the first 120 KB of the scratch
constexpr size_t kTkHistOffset = kTkScratchBytes - 2 * 4096 * sizeof(uint32_t);  // two histograms at the scratch's end,
// out of reach of every sort's exchange buffers (<= 96 KB from the scratch's start for <= 4096 keys; the 8192-key sort
// holds the list in registers and uses the list's own 64 KB + the first 96 KB of the scratch): a slow peer may still
// be reading this CTA's last histogram while it sorts

struct TkControl {                 // static shared memory, one per CTA; cnt / cnt_eq are written by the peers
    uint32_t sel[4];               // digit, need left, boundary-bin count
    uint32_t cnt[16];              // [source CTA] length of that CTA's candidate list
    uint32_t n_list;
    int warp_sums[32];
    int scan_total;
    uint32_t group_tot[64];
    uint32_t fine[64];
};

__device__ __forceinline__ uint32_t tk_load_key(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                                int a) {
    return score_key(__ldg(scores + ((size_t)b * A + a) * stride + offset));
}

// Visits every score index of this CTA's share once (order unspecified).  Unit = 32 consecutive vectors (one warp-wide
// load: 512 B); unit q belongs to CTA q % csize, and the CTA deals its units to its 32 warps round robin.
// The first pass (FIRST) computes the order-preserving keys from global memory and parks them in shared memory
// (`cache`: kTkCacheKeys keys, the part of the scratch area that is idle until the sort); later passes read the keys
// back (one LDS.64 per two anchors) instead of re-reading and re-encoding the scores.  Shares larger than the cache
// re-read their tail from global memory (L2).
template <int MODE, bool FIRST, typename F>
__device__ __forceinline__ void for_each_key(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                             int crank, int csize, uint32_t* __restrict__ cache, F f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int PER = (MODE == 2) ? 2 : (MODE == 1) ? 4 : 1;   // scores per vector
    constexpr int kSlots = kTkCacheKeys / PER;                    // vectors the cache holds
    const int V = A / PER;                                        // MODE 1 / 2 are only selected when PER divides A
    const int units = (V + 31) >> 5;
    int slot = threadIdx.x;                                       // this thread's vectors sit at slot, slot + 1024, ...
    if (MODE == 2 && FIRST) {  // the first pass over rpn_probs [B,A,2] comes from HBM: eight 128-bit loads per thread in
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A * 2);   // flight before the first use
        uint2* c2 = reinterpret_cast<uint2*>(cache);   // (measured: the pass is bound by bytes in flight, not by its atomics)
        constexpr int U = 8;
        for (int q = warp * csize + crank; q < units; q += 32 * csize * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int g = (q + u * 32 * csize) * 32 + lane;
                v[u] = (g < V) ? __ldg(p4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u, slot += kTkThreads) {
                const int g = (q + u * 32 * csize) * 32 + lane;
                if (g < V) {
                    const uint2 k = make_uint2(score_key(offset ? v[u].y : v[u].x), score_key(offset ? v[u].w : v[u].z));
                    if (slot < kSlots) c2[slot] = k;
                    f(k.x, 2 * g);
                    f(k.y, 2 * g + 1);
                }
            }
        }
    } else if (MODE == 2) {  // two interleaved columns (rpn_probs [B,A,2]): one float4 = two anchors
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A * 2);
        uint2* c2 = reinterpret_cast<uint2*>(cache);
#pragma unroll 4
        for (int q = warp * csize + crank; q < units; q += 32 * csize, slot += kTkThreads) {
            const int g = q * 32 + lane;
            if (g < V) {
                uint2 k;
                if (!FIRST && slot < kSlots) k = c2[slot];
                else {
                    const float4 v = __ldg(p4 + g);
                    k = make_uint2(score_key(offset ? v.y : v.x), score_key(offset ? v.w : v.z));
                    if (FIRST && slot < kSlots) c2[slot] = k;
                }
                f(k.x, 2 * g);
                f(k.y, 2 * g + 1);
            }
        }
    } else if (MODE == 1) {  // dense [B,A]
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A);
        uint4* c4 = reinterpret_cast<uint4*>(cache);
#pragma unroll 4
        for (int q = warp * csize + crank; q < units; q += 32 * csize, slot += kTkThreads) {
            const int g = q * 32 + lane;
            if (g < V) {
                uint4 k;
                if (!FIRST && slot < kSlots) k = c4[slot];
                else {
                    const float4 v = __ldg(p4 + g);
                    k = make_uint4(score_key(v.x), score_key(v.y), score_key(v.z), score_key(v.w));
                    if (FIRST && slot < kSlots) c4[slot] = k;
                }
                f(k.x, 4 * g);
                f(k.y, 4 * g + 1);
                f(k.z, 4 * g + 2);
                f(k.w, 4 * g + 3);
            }
        }
    } else {
#pragma unroll 4
        for (int q = warp * csize + crank; q < units; q += 32 * csize, slot += kTkThreads) {
            const int g = q * 32 + lane;
            if (g < V) {
                uint32_t k;
                if (!FIRST && slot < kSlots) k = cache[slot];
                else {
                    k = tk_load_key(scores, stride, offset, A, b, g);
                    if (FIRST && slot < kSlots) cache[slot] = k;
                }
                f(k, g);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kTkThreads, 1)
topk_cluster_kernel(const float* __restrict__ scores, int stride, int offset, int A, int K,
                    int32_t* __restrict__ idx_out, float* __restrict__ vals_out, TopkDecode dec, bool has_dec) {
    extern __shared__ __align__(16) unsigned char tk_smem[];
    __shared__ TkControl ctl;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize, tid = threadIdx.x;
    uint64_t* list = reinterpret_cast<uint64_t*>(tk_smem);                                  // [8192] own candidates
    unsigned char* scratch = tk_smem + kTkListBytes;
    uint32_t* hist2 = reinterpret_cast<uint32_t*>(scratch + kTkHistOffset);                 // [2][4096]
    uint32_t* cache = reinterpret_cast<uint32_t*>(scratch);                                  // [kTkCacheKeys] (until the sort)
    // [level parity][source CTA][group of nb / 64 bins] coarse sums, written by the peers; between the key cache and the
    // histograms (static shared memory has no room left next to 224 KB of dynamic)
    uint32_t* coarse = reinterpret_cast<uint32_t*>(scratch + (size_t)kTkCacheKeys * 4);
    static_assert((size_t)kTkCacheKeys * 4 + 2 * 16 * 64 * 4 <= kTkHistOffset, "key cache + coarse sums must end before the histograms");
    TK_TL(0);
    pdl_launch_dependents();
    if (tid == 0) ctl.n_list = 0;
    pdl_wait();
    TK_TL(1);

    // ---- 1. radix select on the composite (key << 32 | ~index) -----------------------------------------
    uint64_t prefix = 0;               // the resolved leading bits of the K-th composite (== composite >> shift)
    uint32_t need = (uint32_t)K, above_total = 0, in_bin = 0;
    int shift = 64;
    for (int level = 0; level < 6; ++level) {
        const int bits = (level % 3 < 2) ? 12 : 8;   // key: 12 + 12 + 8, then (tie floods only) index: 12 + 12 + 8
        const int nb = 1 << bits;
        const int pshift = shift;  // bits above `pshift` are resolved (== prefix)
        shift -= bits;
        uint32_t* hist = hist2 + (level & 1) * 4096;
        for (int i = tid; i < nb; i += kTkThreads) hist[i] = 0;
        __syncthreads();
        if (level == 0) {
            for_each_key<MODE, true>(scores, stride, offset, A, b, crank, csize, cache,
                                     [&](uint32_t key, int) { atomicAdd(&hist[key >> 20], 1u); });
        } else {
            for_each_key<MODE, false>(scores, stride, offset, A, b, crank, csize, cache, [&](uint32_t key, int a) {
                const uint64_t c = make_composite(key, (uint32_t)a);
                if ((c >> pshift) == prefix) atomicAdd(&hist[(uint32_t)(c >> shift) & (uint32_t)(nb - 1)], 1u);
            });
        }
        // Coarse sums first: 64 groups of nb / 64 bins, pushed into every peer BEFORE the cluster barrier, so that after it
        // every CTA finds the group of the K-th element from local memory and reads only that group's bins from its peers
        // (2 KB instead of the whole histogram of every peer: 128 KB of distributed-shared-memory reads per CTA and level
        // took 11 k cycles, measured).
        __syncthreads();   // this CTA's histogram is complete
        {
            const int gsz = nb >> 6;   // bins per group: 64 (12-bit levels) or 4 (8-bit levels)
            uint32_t part = 0u;
            if (nb == 4096) {
                const uint4 c4 = *reinterpret_cast<const uint4*>(hist + 4 * tid);   // 16 threads per group
                part = c4.x + c4.y + c4.z + c4.w;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            } else if (tid < 64) {
                for (int i = 0; i < gsz; ++i) part += hist[tid * gsz + i];
            }
            const bool owner = (nb == 4096) ? ((tid & 15) == 0) : (tid < 64);
            const int g = (nb == 4096) ? (tid >> 4) : tid;
            if (owner)
                for (int r = 0; r < csize; ++r) cluster.map_shared_rank(coarse + ((level & 1) * 16 + crank) * 64, r)[g] = part;
        }
        TK_TL(2 + 4 * level);
        cluster.sync();  // every CTA's histogram of this level is complete and its coarse sums are everywhere
        TK_TL(3 + 4 * level);
        {
            const int gsz = nb >> 6;
            if (tid < 64) {
                uint32_t t = 0u;
                for (int r = 0; r < csize; ++r) t += coarse[((level & 1) * 16 + r) * 64 + tid];
                ctl.group_tot[tid] = t;
            }
            __syncthreads();
            if (tid < 64) {   // the group that holds the K-th element (counting from the top)
                uint32_t above = 0u;
                for (int g = 63; g > tid; --g) above += ctl.group_tot[g];
                if (above < need && need <= above + ctl.group_tot[tid]) { ctl.sel[0] = (uint32_t)tid; ctl.sel[1] = need - above; }
            }
            __syncthreads();
            const int grp = (int)ctl.sel[0];
            const uint32_t need_g = ctl.sel[1];
            if (tid < gsz) {   // that group's bins, summed over the cluster
                uint32_t t = 0u;
                for (int r = 0; r < csize; ++r) t += cluster.map_shared_rank(hist, r)[grp * gsz + tid];
                ctl.fine[tid] = t;
            }
            __syncthreads();
            TK_TL(4 + 4 * level);
            if (tid < gsz) {
                uint32_t above = 0u;
                for (int i = gsz - 1; i > tid; --i) above += ctl.fine[i];
                if (above < need_g && need_g <= above + ctl.fine[tid]) {
                    ctl.sel[0] = (uint32_t)(grp * gsz + tid);   // nobody reads sel[0] as the group any more
                    ctl.sel[1] = need_g - above;
                    ctl.sel[2] = ctl.fine[tid];
                    ctl.sel[3] = (need - need_g) + above;        // candidates above the digit's bin at this level
                }
            }
        }
        __syncthreads();
        const uint32_t digit = ctl.sel[0], need_left = ctl.sel[1];
        in_bin = ctl.sel[2];
        prefix = (prefix << bits) | (uint64_t)digit;
        above_total += need - need_left;
        need = need_left;
        TK_TL(5 + 4 * level);
        if (above_total + in_bin <= (uint32_t)kMaxSort) break;   // certain at the last level: composites are distinct
    }

    // ---- 2. compaction of this CTA's candidates into its own list ---------------------------------------
    const uint64_t bound = prefix << shift;   // (c >> shift) >= prefix  <=>  c >= prefix << shift; shift < 64 here
    for_each_key<MODE, false>(scores, stride, offset, A, b, crank, csize, cache, [&](uint32_t key, int a) {
        const uint64_t c = make_composite(key, (uint32_t)a);
        if (c >= bound) list[atomicAdd(&ctl.n_list, 1u)] = c;
    });
    __syncthreads();
    const int n_mine = (int)ctl.n_list;
    TK_TL(26);
    if (has_dec) {
        // the epilogue gathers each winner's anchor and delta row from HBM by index: ask the L2 for them now, they arrive
        // while the lists are exchanged and ranked (a few of the candidates end up below rank K: harmless)
        for (int i = tid; i < n_mine; i += kTkThreads) {
            const uint32_t a = composite_idx(list[i]);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(dec.anchors + (size_t)b * A + a));
            if (dec.deltas) asm volatile("prefetch.global.L2 [%0];" ::"l"(dec.deltas + (size_t)b * A + a));
        }
    }

    TK_TL(27);
    // ---- 3. every CTA gathers ALL candidate lists (unsorted, <= 8192 entries) and ranks its own by counting ----
    if (tid < csize) cluster.map_shared_rank(&ctl.cnt[0], tid)[crank] = (uint32_t)n_mine;
    cluster.sync();  // every list is complete, every length is known everywhere
    TK_TL(28);
    uint64_t* all = reinterpret_cast<uint64_t*>(scratch);                 // [total] the peers' lists and this CTA's, back to back
    uint64_t* byb = reinterpret_cast<uint64_t*>(scratch + kTkListBytes);  // [total] the same, grouped by bin
    int total = 0;
    for (int r = 0; r < csize; ++r) {
        const uint64_t* src = cluster.map_shared_rank(list, r);
        const int n = (int)ctl.cnt[r];
        for (int i = tid; i < n; i += kTkThreads) all[total + i] = src[i];
        total += n;
    }
    // this CTA no longer reads its peers' shared memory: arrive now, wait at exit
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    __syncthreads();
    TK_TL(29);
    // Order-preserving bins over the candidates' key range [key_lo, 2^32): bin = (key - key_lo) >> s, 4096 bins.  The rank
    // of a candidate is the number of candidates in higher bins (one scan) plus the number of greater composites in its
    // own bin (a handful on real score distributions).  A bin with more than kTkBinLimit entries (tie floods: thousands of
    // equal scores) or a select that had to descend into the index bits sends the whole list through the sorting network
    // instead -- every CTA takes the same decision from the same histogram.
    uint32_t* hist = hist2;             // [4096] bin counts, then the bins' first ranks
    uint32_t* cursor = hist2 + 4096;    // [4096]
    bool fast = shift >= 32;
    const uint32_t key_lo = fast ? (uint32_t)(prefix << (shift - 32)) : 0u;
    int bshift = 0;
    if (fast) {
        for (int i = tid; i < 4096; i += kTkThreads) { hist[i] = 0u; cursor[i] = 0u; }
        if (tid == 0) { ctl.sel[3] = 0u; ctl.sel[2] = 0u; }
        __syncthreads();
        uint32_t kmax = 0u;   // the largest key: the bins cover [key_lo, kmax]
        for (int i = tid; i < total; i += kTkThreads) kmax = max(kmax, composite_key(all[i]));
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        if ((tid & 31) == 0) atomicMax(&ctl.sel[2], kmax);
        __syncthreads();
        kmax = max(ctl.sel[2], key_lo);
        while (bshift < 32 && ((kmax - key_lo) >> bshift) >= 4096u) ++bshift;
        for (int i = tid; i < total; i += kTkThreads) atomicAdd(&hist[(composite_key(all[i]) - key_lo) >> bshift], 1u);
        __syncthreads();
        // thread t owns bins 4t .. 4t+3; first rank of a bin = candidates in the bins above it
        const uint4 c4 = *reinterpret_cast<const uint4*>(hist + 4 * tid);
        const uint32_t tsum = c4.x + c4.y + c4.z + c4.w;
        if (max(max(c4.x, c4.y), max(c4.z, c4.w)) > (uint32_t)kTkBinLimit) ctl.sel[3] = 1u;
        const int before = block_exclusive_scan((int)tsum, ctl.warp_sums, &ctl.scan_total);
        uint32_t acc = (uint32_t)ctl.scan_total - (uint32_t)before - tsum;   // candidates in the bins of higher threads
        uint4 st;
        st.w = acc; acc += c4.w;
        st.z = acc; acc += c4.z;
        st.y = acc; acc += c4.y;
        st.x = acc;
        __syncthreads();   // every thread has read its counts
        *reinterpret_cast<uint4*>(hist + 4 * tid) = st;
        __syncthreads();
        fast = ctl.sel[3] == 0u;
    }
    if (fast) {
        for (int i = tid; i < total; i += kTkThreads) {
            const uint64_t v = all[i];
            const uint32_t bin = (composite_key(v) - key_lo) >> bshift;
            byb[hist[bin] + atomicAdd(&cursor[bin], 1u)] = v;
        }
        __syncthreads();
    } else {
        // rare path: all candidates through the sorting network, in every CTA; nobody may still be reading this CTA's list
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        uint64_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (tid * 8 + e < total) ? all[tid * 8 + e] : 0ull;   // padding sorts last
        __syncthreads();
        block_sort_desc_blocked<8>(v, reinterpret_cast<uint64_t*>(tk_smem));
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; ++e) list[tid * 8 + e] = v[e];
        __syncthreads();
    }
    TK_TL(30);

    // ---- 4. emit at the global rank ------------------------------------------------------------------------
    auto emit = [&](uint64_t v, int r) {
        const uint32_t a = composite_idx(v);
        if (idx_out) idx_out[(size_t)b * K + r] = (int32_t)a;
        if (vals_out) vals_out[(size_t)b * K + r] = key_score(composite_key(v));
        if (has_dec) {
            // L:247-261: gather anchors / deltas by index, scale by std_dev (L:238), decode, clip to [0,1]
            const float4 an = __ldg(dec.anchors + (size_t)b * A + a);
            const float4 dl = scale_deltas(topk_load_delta(dec, b, A, a), dec.std_dev);
            const float4 bx = clip_box(apply_box_deltas(an, dl), make_float4(0.f, 0.f, 1.f, 1.f));
            dec.boxes_sorted[(size_t)b * K + r] = bx;
            if (dec.pre_nms_boxes) dec.pre_nms_boxes[(size_t)b * K + r] = bx;
        }
    };
    if (fast) {
        // this CTA's own candidates; the order inside a bin of `byb` is whatever the atomics produced: count the bin's
        // greater composites
        for (int i = tid; i < n_mine; i += kTkThreads) {
            const uint64_t v = list[i];
            const uint32_t bin = (composite_key(v) - key_lo) >> bshift;
            const int first = (int)hist[bin], n_bin = (int)cursor[bin];
            int r = first;
            for (int j = first; j < first + n_bin; ++j) r += (byb[j] > v);
            if (r < K) emit(v, r);
        }
        TK_TL(31);
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // no CTA leaves while a peer may still read its list
    } else {
        for (int r = crank + csize * tid; r < min(total, K); r += csize * kTkThreads) emit(list[r], r);
    }
}

static size_t tk_smem_bytes() { return kTkListBytes + kTkScratchBytes; }

template <int M> static int tk_cluster_size_for(int B) {
    return pick_cluster_size_any((const void*)topk_cluster_kernel<M>, kTkThreads, B, tuning_knob("MRCNN_TOPK_MAX_CLUSTER", 16),
                                 [](int) { return tk_smem_bytes(); });
}

size_t topk_ws_bytes(int /*B*/) { return 256; }  // the cluster kernel keeps everything on chip

int launch_topk(const float* scores, int stride, int offset, int B, int A, int K, int32_t* idx, float* vals,
                const TopkDecode* dec, void* /*ws*/, cudaStream_t stream) {
    int mode = 0;
    if (aligned16(scores)) {
        if (stride == 2 && (A % 2) == 0) mode = 2;
        else if (stride == 1 && (A % 4) == 0) mode = 1;
    }
    const int cs = (mode == 2) ? tk_cluster_size_for<2>(B) : (mode == 1) ? tk_cluster_size_for<1>(B) : tk_cluster_size_for<0>(B);
    const size_t smem = tk_smem_bytes();
    TopkDecode d{};
    if (dec) d = *dec;
    const bool has_dec = dec != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kTkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1, stream);
    cudaError_t e;
#define MRCNN_TK(M)                                                                                                   \
    do {                                                                                                              \
        e = cudaFuncSetAttribute(topk_cluster_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return (int)e;                                                                          \
        e = cudaLaunchKernelEx(&cfg, topk_cluster_kernel<M>, scores, stride, offset, A, K, idx, vals, d, has_dec); \
    } while (0)
    if (mode == 2) MRCNN_TK(2);
    else if (mode == 1) MRCNN_TK(1);
    else MRCNN_TK(0);
#undef MRCNN_TK
    if (e != cudaSuccess) return (int)e;
    return last_error();
}

}  // namespace mrcnn

using namespace mrcnn;

#ifdef MRCNN_NMS_PROFILE
MRCNN_EXPORT int mrcnn_debug_topk_timeline(long long* host_out32) {
    return (int)cudaMemcpyFromSymbol(host_out32, mrcnn::g_tk_timeline, sizeof(long long) * 32);
}
#endif

MRCNN_EXPORT int mrcnn_topk_workspace_bytes(int B, int A, int K, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || A < 1 || K < 1 || K > A || K > kMaxSort) return MRCNN_ERR_RANGE;
    *bytes = topk_ws_bytes(B);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_topk_forward(const float* scores, int stride, int offset, int B, int A, int K, int32_t* idx,
                                    float* vals, void* ws, size_t ws_bytes, void* stream) {
    if (!scores || !idx || !ws) return MRCNN_ERR_NULL;
    if (B < 1 || A < 1 || K < 1 || K > A || K > kMaxSort || stride < 1 || offset < 0 || offset >= stride)
        return MRCNN_ERR_RANGE;
    if ((size_t)B * A * stride > 0x7fffffffull * 4) return MRCNN_ERR_RANGE;
    if (ws_bytes < topk_ws_bytes(B)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(ws)) return MRCNN_ERR_ALIGN;
    return launch_topk(scores, stride, offset, B, A, K, idx, vals, nullptr, ws, (cudaStream_t)stream);
}
