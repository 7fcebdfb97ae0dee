// topk_cluster.cu -- exact per-image top-K (K <= 8192, TF TopKV2 order: value desc, index asc) as ONE kernel.
// Replaces tf.nn.top_k at mrcnn_layers.py:246, the per-image gathers at L:247-250 and (fused epilogue) the box
// decode / clip of utils.py:830-869.
//
// A thread-block cluster of 1..8 CTAs (1024 threads each) owns an image; every CTA streams one contiguous slice
// of the scores (128-bit loads, L2-resident after the first pass) and everything else stays on chip:
//   1. radix select, up to three levels (12 + 12 + 8 bits of an order-preserving key): per-CTA shared-memory
//      histogram, cluster-wide reduction where CTA s sums bin slice s from all peers through distributed shared
//      memory, the CTA whose slice holds the K-th element resolves the digit and stores it into every CTA; the
//      descent stops as soon as "above + boundary bin" fits the 8192-entry sort;
//   2. compaction of the candidates as 64-bit (key, ~index) composites into shared memory, exclusive offsets from
//      the per-CTA counts, then a scatter of every candidate to its sort slot in the owning CTA (DSMEM stores);
//      if more than 8192 candidates share all 32 key bits the equal keys are taken in index order instead (ordered
//      block scans per CTA, quotas from the per-CTA counts);
//   3. cluster-wide bitonic sort: 8192 / (1024 * CTAs) keys per thread in registers, shuffles inside a warp, padded
//      shared memory across warps, DSMEM reads across CTAs (log2(CTAs)*(log2(CTAs)+1)/2 cluster barriers in total);
//   4. epilogue: indices / values, and for ProposalLayer the gather + std-dev scale + decode + clip of the winners.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace mrcnn {

constexpr int kTkThreads = 1024;

struct TkControl {                 // static shared memory, one per CTA; the *_from arrays are written by the peers
    uint32_t slice_sum[8];         // [source CTA] total of that CTA's bin slice
    uint32_t sel[4];               // digit, need left, boundary-bin count, -- (stored by the owner into every CTA)
    uint32_t cnt_a[8], cnt_b[8];   // [source CTA] candidates above / on the boundary
    uint32_t cnt_eq[8];            // [source CTA] tie path: elements equal to the K-th key
    uint32_t n_a, n_b;
    int warp_sums[32];
    int scan_total;
};

__device__ __forceinline__ uint32_t tk_load_key(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                                int a) {
    return score_key(__ldg(scores + ((size_t)b * A + a) * stride + offset));
}

// visits every score index in [lo, hi) once across the block (order unspecified); lo, hi multiples of 4 or == A
template <int MODE, typename F>
__device__ __forceinline__ void for_each_key(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                             int lo, int hi, F f) {
    const int tid = threadIdx.x;
    if (MODE == 2) {  // two interleaved columns (rpn_probs [B,A,2]): one float4 = two anchors
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A * 2);
#pragma unroll 4
        for (int a = lo + tid * 2; a < hi; a += kTkThreads * 2) {
            const float4 q = __ldg(p4 + (a >> 1));
            f(score_key(offset ? q.y : q.x), a);
            f(score_key(offset ? q.w : q.z), a + 1);
        }
    } else if (MODE == 1) {  // dense [B,A]
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A);
#pragma unroll 4
        for (int a = lo + tid * 4; a < hi; a += kTkThreads * 4) {
            const float4 q = __ldg(p4 + (a >> 2));
            f(score_key(q.x), a);
            f(score_key(q.y), a + 1);
            f(score_key(q.z), a + 2);
            f(score_key(q.w), a + 3);
        }
    } else {
#pragma unroll 4
        for (int a = lo + tid; a < hi; a += kTkThreads) f(tk_load_key(scores, stride, offset, A, b, a), a);
    }
}

// cluster-wide bitonic sort, descending; thread t of CTA r holds ranks (r*1024 + t)*E .. +E-1
template <int E>
__device__ __forceinline__ void cluster_sort_desc(uint64_t (&v)[E], uint64_t* xch_local, uint64_t* xch_remote,
                                                  cg::cluster_group& cluster, int crank, int csize) {
    constexpr int LOG_E = (E == 1) ? 0 : (E == 2) ? 1 : (E == 4) ? 2 : 3;
    constexpr int ROW = E + 2;
    constexpr int NCTA = 1024 * E;
    const int N = NCTA * csize;
    const int t = threadIdx.x;
    const int gbase = (crank * 1024 + t) << LOG_E;
    int buf = 0, rbuf = 0;
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= NCTA) {  // partner element lives in another CTA of the cluster
                const int dc = j / NCTA;
                uint64_t* mine = xch_remote + (size_t)rbuf * NCTA + (size_t)t * E;
#pragma unroll
                for (int e = 0; e < E; ++e) mine[e] = v[e];
                cluster.sync();
                const uint64_t* theirs = cluster.map_shared_rank(mine, crank ^ dc);
                const bool lower = (crank & dc) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t p = theirs[e];
                    const bool desc = ((gbase | e) & k) == 0;
                    const bool take_max = (lower == desc);
                    v[e] = take_max ? (v[e] > p ? v[e] : p) : (v[e] < p ? v[e] : p);
                }
                rbuf ^= 1;
            } else if (j >= 32 * E) {  // another warp of this CTA
                const int dm = j >> LOG_E;
                uint64_t* mine = xch_local + (size_t)buf * (1024 * ROW) + (size_t)t * ROW;
#pragma unroll
                for (int e = 0; e < E; ++e) mine[e] = v[e];
                __syncthreads();
                const uint64_t* theirs = xch_local + (size_t)buf * (1024 * ROW) + (size_t)(t ^ dm) * ROW;
                const bool lower = (t & dm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t p = theirs[e];
                    const bool desc = ((gbase | e) & k) == 0;
                    const bool take_max = (lower == desc);
                    v[e] = take_max ? (v[e] > p ? v[e] : p) : (v[e] < p ? v[e] : p);
                }
                buf ^= 1;
            } else if (j >= E) {  // another lane of this warp
                const int lm = j >> LOG_E;
                const bool lower = (t & lm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t p = __shfl_xor_sync(0xffffffffu, v[e], lm);
                    const bool desc = ((gbase | e) & k) == 0;
                    const bool take_max = (lower == desc);
                    v[e] = take_max ? (v[e] > p ? v[e] : p) : (v[e] < p ? v[e] : p);
                }
            } else {  // both elements in this thread's registers
                // sort_local_stage derives the direction from (t << LOG_E | e) & k; add the CTA offset through t
                const int tt = crank * 1024 + t;
                if (j == 4) sort_local_stage<E, 4>(v, k, tt, LOG_E);
                else if (j == 2) sort_local_stage<E, 2>(v, k, tt, LOG_E);
                else sort_local_stage<E, 1>(v, k, tt, LOG_E);
            }
        }
    }
}

template <int E>
__device__ __forceinline__ void tk_sort_emit(const uint64_t* slots, uint64_t* xch_local, uint64_t* xch_remote,
                                             cg::cluster_group& cluster, int crank, int csize, int A, int K, int b,
                                             int32_t* idx_out, float* vals_out, const TopkDecode& dec, bool has_dec) {
    const int tid = threadIdx.x;
    uint64_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) v[e] = slots[tid * E + e];
    __syncthreads();  // the exchange buffers alias the slot array
    cluster_sort_desc<E>(v, xch_local, xch_remote, cluster, crank, csize);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int r = (crank * 1024 + tid) * E + e;
        if (r >= K) continue;
        const uint32_t a = composite_idx(v[e]);
        if (idx_out) idx_out[(size_t)b * K + r] = (int32_t)a;
        if (vals_out) vals_out[(size_t)b * K + r] = key_score(composite_key(v[e]));
        if (has_dec) {
            // L:247-261: gather anchors / deltas by index, scale by std_dev (L:238), decode, clip to [0,1]
            const float4 an = __ldg(dec.anchors + (size_t)b * A + a);
            const float4 dl = scale_deltas(topk_load_delta(dec, b, A, a), dec.std_dev);
            const float4 bx = clip_box(apply_box_deltas(an, dl), make_float4(0.f, 0.f, 1.f, 1.f));
            dec.boxes_sorted[(size_t)b * K + r] = bx;
            if (dec.pre_nms_boxes) dec.pre_nms_boxes[(size_t)b * K + r] = bx;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(kTkThreads, 1)
topk_cluster_kernel(const float* __restrict__ scores, int stride, int offset, int A, int K, int chunk,
                    int32_t* __restrict__ idx_out, float* __restrict__ vals_out, TopkDecode dec, bool has_dec) {
    extern __shared__ __align__(16) unsigned char tk_smem[];
    __shared__ TkControl ctl;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize, tid = threadIdx.x;
    const int E = kMaxSort / (csize * kTkThreads);
    uint32_t* hist = reinterpret_cast<uint32_t*>(tk_smem);                  // [4096]
    uint32_t* slice_tot = hist + 4096;                                      // [4096 / csize]
    uint64_t* region = reinterpret_cast<uint64_t*>(slice_tot + 4096);
    uint64_t* staging = region;                                             // [8192] above from the front, boundary from the back
    uint64_t* slots = region + kMaxSort;                                    // [1024 * E] this CTA's sort input
    uint64_t* xch_local = region;                                           // sort: aliases staging / slots (both dead by then)
    uint64_t* xch_remote = region + 2 * 1024 * (E + 2);
    const int lo = min(A, crank * chunk), hi = min(A, lo + chunk);

    // ---- 1. radix select -----------------------------------------------------------------------------
    uint32_t prefix = 0, need = (uint32_t)K, above_total = 0, in_bin = 0;
    int shift = 32;
    bool fits = false;
    for (int level = 0; level < 3; ++level) {
        const int bits = (level < 2) ? 12 : 8;
        const int nb = 1 << bits;
        const int W = nb / csize;  // bins per slice
        const int pshift = shift;  // bits above `pshift` are resolved (== prefix)
        shift -= bits;
        for (int i = tid; i < nb; i += kTkThreads) hist[i] = 0;
        __syncthreads();
        for_each_key<MODE>(scores, stride, offset, A, b, lo, hi, [&](uint32_t key, int) {
            if (level == 0 || (key >> pshift) == prefix) atomicAdd(&hist[(key >> shift) & (uint32_t)(nb - 1)], 1u);
        });
        __syncthreads();
        cluster.sync();  // every CTA's histogram is complete
        // CTA `crank` sums bin slice `crank` over the cluster
        uint32_t part = 0;
        for (int i = tid; i < W; i += kTkThreads) {
            uint32_t tot = 0;
            for (int r = 0; r < csize; ++r) tot += cluster.map_shared_rank(hist, r)[crank * W + i];
            slice_tot[i] = tot;
            part += tot;
        }
        {
            const int excl = block_exclusive_scan((int)part, ctl.warp_sums, &ctl.scan_total);
            (void)excl;
        }
        if (tid < csize) cluster.map_shared_rank(&ctl.slice_sum[0], tid)[crank] = (uint32_t)ctl.scan_total;
        cluster.sync();  // slice totals of every CTA are visible everywhere
        uint32_t above_s = 0;
        for (int r = crank + 1; r < csize; ++r) above_s += ctl.slice_sum[r];
        const uint32_t mine = ctl.slice_sum[crank];
        if (above_s < need && need <= above_s + mine) {  // this CTA's slice holds the K-th element (CTA-uniform)
            const int per = (W + kTkThreads - 1) / kTkThreads;  // bins per thread (1..4)
            uint32_t tsum = 0;
            for (int i = tid * per; i < min(W, (tid + 1) * per); ++i) tsum += slice_tot[i];
            const int before = block_exclusive_scan((int)tsum, ctl.warp_sums, &ctl.scan_total);
            const uint32_t above_t = above_s + (mine - (uint32_t)before - tsum);  // bins of higher threads
            if (above_t < need && need <= above_t + tsum) {
                uint32_t acc = above_t;
                for (int i = min(W, (tid + 1) * per) - 1; i >= tid * per; --i) {
                    const uint32_t c = slice_tot[i];
                    if (need <= acc + c) {
                        for (int r = 0; r < csize; ++r) {
                            uint32_t* dst = cluster.map_shared_rank(&ctl.sel[0], r);
                            dst[0] = (uint32_t)(crank * W + i);
                            dst[1] = need - acc;
                            dst[2] = c;
                        }
                        break;
                    }
                    acc += c;
                }
            }
        }
        cluster.sync();  // the selection is visible in every CTA
        const uint32_t digit = ctl.sel[0], need_left = ctl.sel[1];
        in_bin = ctl.sel[2];
        prefix = (level == 0) ? digit : ((prefix << bits) | digit);
        above_total += need - need_left;
        need = need_left;
        if (above_total + in_bin <= (uint32_t)kMaxSort) { fits = true; break; }
    }

    // ---- 2. compaction into shared memory, offsets, scatter to the sort slots ----------------------------
    if (tid == 0) { ctl.n_a = 0; ctl.n_b = 0; }
    for (int i = tid; i < kTkThreads * E; i += kTkThreads) slots[i] = 0ull;
    __syncthreads();
    if (fits) {
        for_each_key<MODE>(scores, stride, offset, A, b, lo, hi, [&](uint32_t key, int a) {
            const uint32_t kp = key >> shift;
            if (kp > prefix) staging[atomicAdd(&ctl.n_a, 1u)] = make_composite(key, (uint32_t)a);
            else if (kp == prefix) staging[kMaxSort - 1 - atomicAdd(&ctl.n_b, 1u)] = make_composite(key, (uint32_t)a);
        });
        __syncthreads();
    } else {
        // tie flood: all 32 key bits resolved (prefix = K-th key, need = how many of its copies belong to the
        // top K); copies are taken in index order
        uint32_t eq_local = 0;
        for_each_key<MODE>(scores, stride, offset, A, b, lo, hi, [&](uint32_t key, int a) {
            if (key > prefix) staging[atomicAdd(&ctl.n_a, 1u)] = make_composite(key, (uint32_t)a);
            else if (key == prefix) ++eq_local;
        });
        (void)block_exclusive_scan((int)eq_local, ctl.warp_sums, &ctl.scan_total);
        if (tid < csize) cluster.map_shared_rank(&ctl.cnt_eq[0], tid)[crank] = (uint32_t)ctl.scan_total;
        cluster.sync();
        uint32_t eq_before = 0;
        for (int r = 0; r < crank; ++r) eq_before += ctl.cnt_eq[r];
        const uint32_t quota = (need > eq_before) ? min(need - eq_before, ctl.cnt_eq[crank]) : 0u;
        uint32_t taken = 0;  // uniform
        for (int a0 = lo; a0 < hi && taken < quota; a0 += kTkThreads) {
            const int a = a0 + tid;
            const bool eq = (a < hi) && tk_load_key(scores, stride, offset, A, b, a) == prefix;
            const int rank = block_exclusive_scan(eq ? 1 : 0, ctl.warp_sums, &ctl.scan_total);
            if (eq && taken + (uint32_t)rank < quota)
                staging[kMaxSort - 1 - (taken + (uint32_t)rank)] = make_composite(prefix, (uint32_t)a);
            taken += (uint32_t)ctl.scan_total;
            __syncthreads();
        }
        if (tid == 0) ctl.n_b = quota;
        __syncthreads();
    }
    const uint32_t n_a = ctl.n_a, n_b = ctl.n_b;
    if (tid < csize) {
        cluster.map_shared_rank(&ctl.cnt_a[0], tid)[crank] = n_a;
        cluster.map_shared_rank(&ctl.cnt_b[0], tid)[crank] = n_b;
    }
    cluster.sync();  // counts visible everywhere, every CTA's slot array is zeroed
    uint32_t off_a = 0, tot_a = 0, off_b = 0;
    for (int r = 0; r < csize; ++r) {
        if (r < crank) { off_a += ctl.cnt_a[r]; off_b += ctl.cnt_b[r]; }
        tot_a += ctl.cnt_a[r];
    }
    off_b += tot_a;
    const uint32_t per_cta = (uint32_t)(kTkThreads * E);
    for (uint32_t i = tid; i < n_a + n_b; i += kTkThreads) {
        const bool is_a = i < n_a;
        const uint32_t pos = is_a ? off_a + i : off_b + (i - n_a);
        const uint64_t comp = is_a ? staging[i] : staging[kMaxSort - 1 - (i - n_a)];
        if (pos < (uint32_t)kMaxSort) cluster.map_shared_rank(slots, pos / per_cta)[pos % per_cta] = comp;
    }
    cluster.sync();  // every candidate sits in its slot

    // ---- 3 + 4. sort and emit ---------------------------------------------------------------------------
    if (E == 1) tk_sort_emit<1>(slots, xch_local, xch_remote, cluster, crank, csize, A, K, b, idx_out, vals_out, dec, has_dec);
    else if (E == 2) tk_sort_emit<2>(slots, xch_local, xch_remote, cluster, crank, csize, A, K, b, idx_out, vals_out, dec, has_dec);
    else if (E == 4) tk_sort_emit<4>(slots, xch_local, xch_remote, cluster, crank, csize, A, K, b, idx_out, vals_out, dec, has_dec);
    else tk_sort_emit<8>(slots, xch_local, xch_remote, cluster, crank, csize, A, K, b, idx_out, vals_out, dec, has_dec);
    cluster.sync();  // no CTA leaves while a peer may still read its exchange buffers
}

static size_t tk_smem_bytes(int cs);
template <int M> static int tk_cluster_size_for(int B) {
    static int cache[4][2] = {};
    return pick_cluster_size(topk_cluster_kernel<M>, kTkThreads, B, 8, [](int cs) { return tk_smem_bytes(cs); }, cache);
}

static size_t tk_smem_bytes(int cs) {
    const int E = kMaxSort / (cs * kTkThreads);
    const size_t stage = (size_t)(kMaxSort + kTkThreads * E) * 8;
    const size_t sort = (size_t)2 * 1024 * (E + 2) * 8 + (cs > 1 ? (size_t)2 * 1024 * E * 8 : 0);
    return 2 * 4096 * sizeof(uint32_t) + (stage > sort ? stage : sort);
}

int launch_topk_cluster(const float* scores, int stride, int offset, int B, int A, int K, int32_t* idx, float* vals,
                        const TopkDecode* dec, cudaStream_t stream) {
    int mode = 0;
    if (aligned16(scores)) {
        if (stride == 2 && (A % 2) == 0) mode = 2;
        else if (stride == 1 && (A % 4) == 0) mode = 1;
    }
    const int cs = (mode == 2) ? tk_cluster_size_for<2>(B) : (mode == 1) ? tk_cluster_size_for<1>(B) : tk_cluster_size_for<0>(B);
    const size_t smem = tk_smem_bytes(cs);
    int chunk = (A + cs - 1) / cs;
    chunk = (chunk + 3) & ~3;
    TopkDecode d{};
    if (dec) d = *dec;
    const bool has_dec = dec != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kTkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e;
#define MRCNN_TK(M)                                                                                                   \
    do {                                                                                                              \
        e = cudaFuncSetAttribute(topk_cluster_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e != cudaSuccess) return (int)e;                                                                          \
        e = cudaLaunchKernelEx(&cfg, topk_cluster_kernel<M>, scores, stride, offset, A, K, chunk, idx, vals, d, has_dec); \
    } while (0)
    if (mode == 2) MRCNN_TK(2);
    else if (mode == 1) MRCNN_TK(1);
    else MRCNN_TK(0);
#undef MRCNN_TK
    if (e != cudaSuccess) return (int)e;
    return last_error();
}

}  // namespace mrcnn
