// topk.cu -- per-image exact top-K (K <= 8192) of fp32 scores with TF TopKV2 ordering (value desc, index asc).
// Replaces tf.nn.top_k at mrcnn_layers.py:246 and the three per-image tf.gather's that follow (L:247-250).
//
// Radix select on an order-preserving 32-bit key, two 12-bit histogram passes over the scores (multi-CTA,
// shared-memory histograms merged with global atomics; the last CTA of an image resolves the digit), one
// compaction pass (block-aggregated slots) and one single-CTA bitonic sort of the <= 8192 survivors as 64-bit
// (key, ~index) composites.  Adversarial tie floods (more than 8192 candidates sharing the 24-bit prefix) take
// an in-order rescan path in the final kernel that resolves the last 8 key bits and the index tie-break.
#include "common.cuh"

namespace mrcnn {

constexpr int kChunk = 4096;        // scores per CTA in the streaming passes
constexpr int kStreamThreads = 256;
constexpr int kBins = 4096;
constexpr int kCtl = 16;            // u32 control words per image
// ctl: 0 done(pass0) 1 done(pass1) 2 prefix12 3 need1 4 prefix (12 or 24 bit) 5 need2 6 n_above 7 n_bound
//      8 skip-pass-1 flag 9 compaction shift (20 or 8)

struct TopkWs {
    uint32_t* hist;   // [B,4096]
    uint32_t* ctl;    // [B,16]
    uint64_t* above;  // [B,8192]
    uint64_t* bound;  // [B,8192]
};

size_t topk_ws_bytes(int B) {
    return align_up((size_t)B * (kBins + kCtl) * sizeof(uint32_t), 256) + 2 * (size_t)B * kMaxSort * sizeof(uint64_t);
}

static TopkWs carve(void* ws, int B) {
    TopkWs w;
    char* p = (char*)ws;
    w.hist = (uint32_t*)p;
    w.ctl = w.hist + (size_t)B * kBins;
    p += align_up((size_t)B * (kBins + kCtl) * sizeof(uint32_t), 256);
    w.above = (uint64_t*)p;
    w.bound = w.above + (size_t)B * kMaxSort;
    return w;
}

__device__ __forceinline__ uint32_t load_key(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                             int a) {
    return score_key(__ldg(scores + ((size_t)b * A + a) * stride + offset));
}

// The streaming passes give every thread kPerThread = 16 scores of its CTA's 4096-score chunk.  MODE 2: scores
// interleaved two per anchor (rpn_probs [B,A,2]), 8 x 128-bit loads per thread; MODE 1: dense [B,A], 4 x 128-bit
// loads; MODE 0: any stride, scalar loads.  All loads are issued before the first key is consumed.
constexpr int kPerThread = kChunk / kStreamThreads;

template <int MODE>
__device__ __forceinline__ void load_chunk(const float* __restrict__ scores, int stride, int offset, int A, int b,
                                           int chunk, uint32_t (&key)[kPerThread], int (&idx)[kPerThread]) {
    const int tid = threadIdx.x;
    const int base = chunk * kChunk;
    if (MODE == 2) {
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A * 2) + (base >> 1);
        float4 q[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int a = base + (it * kStreamThreads + tid) * 2;
            q[it] = (a + 1 < A) ? __ldg(p4 + it * kStreamThreads + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (a + 1 >= A && a < A) {  // odd tail anchor
                const float* r = scores + ((size_t)b * A + a) * 2;
                q[it].x = __ldg(r); q[it].y = __ldg(r + 1);
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int a = base + (it * kStreamThreads + tid) * 2;
            idx[2 * it] = (a < A) ? a : -1;
            idx[2 * it + 1] = (a + 1 < A) ? a + 1 : -1;
            key[2 * it] = score_key(offset ? q[it].y : q[it].x);
            key[2 * it + 1] = score_key(offset ? q[it].w : q[it].z);
        }
    } else if (MODE == 1) {
        const float4* p4 = reinterpret_cast<const float4*>(scores + (size_t)b * A) + (base >> 2);
        float4 q[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int a = base + (it * kStreamThreads + tid) * 4;
            if (a + 3 < A) q[it] = __ldg(p4 + it * kStreamThreads + tid);
            else {
                const float* r = scores + (size_t)b * A;
                q[it].x = (a < A) ? __ldg(r + a) : 0.f;
                q[it].y = (a + 1 < A) ? __ldg(r + a + 1) : 0.f;
                q[it].z = (a + 2 < A) ? __ldg(r + a + 2) : 0.f;
                q[it].w = 0.f;
            }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int a = base + (it * kStreamThreads + tid) * 4;
            const float f[4] = {q[it].x, q[it].y, q[it].z, q[it].w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                idx[4 * it + h] = (a + h < A) ? a + h : -1;
                key[4 * it + h] = score_key(f[h]);
            }
        }
    } else {
#pragma unroll
        for (int it = 0; it < kPerThread; ++it) {
            const int a = base + it * kStreamThreads + tid;
            idx[it] = (a < A) ? a : -1;
            key[it] = (a < A) ? load_key(scores, stride, offset, A, b, a) : 0u;
        }
    }
}

static int stream_mode(const float* scores, int stride, int A) {
    if (!aligned16(scores)) return 0;
    if (stride == 2 && (A % 2) == 0) return 2;
    if (stride == 1 && (A % 4) == 0) return 1;
    return 0;
}

// ctl[8] = 1: the 12-bit boundary bin plus everything above it fits the sort buffer, pass 1 is skipped and the
// compaction compares 12-bit prefixes (ctl[9] = shift used by the compaction: 20 or 8)
template <int PASS, int MODE>
__global__ void __launch_bounds__(kStreamThreads)
topk_hist_kernel(const float* __restrict__ scores, int stride, int offset, int A, int K, uint32_t* __restrict__ hist,
                 uint32_t* __restrict__ ctl, int chunks) {
    __shared__ uint32_t sh[kBins];
    __shared__ uint32_t part[kStreamThreads];
    __shared__ int s_last;
    const int b = blockIdx.y, tid = threadIdx.x;
    uint32_t* gh = hist + (size_t)b * kBins;
    uint32_t* c = ctl + (size_t)b * kCtl;
    if (PASS == 1 && c[8]) return;
    uint32_t key[kPerThread];
    int idx[kPerThread];
    load_chunk<MODE>(scores, stride, offset, A, b, blockIdx.x, key, idx);
    for (int i = tid; i < kBins; i += kStreamThreads) sh[i] = 0;
    const uint32_t prefix12 = (PASS == 1) ? c[2] : 0u;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kPerThread; ++it) {
        if (idx[it] >= 0) {
            if (PASS == 0) atomicAdd(&sh[key[it] >> 20], 1u);
            else if ((key[it] >> 20) == prefix12) atomicAdd(&sh[(key[it] >> 8) & 0xfffu], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < kBins; i += kStreamThreads) {
        const uint32_t v = sh[i];
        if (v) atomicAdd(&gh[i], v);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&c[PASS], 1u) == (uint32_t)(chunks - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA of this image: find the digit d with count(bins > d) < need <= count(bins >= d)
    const uint32_t need = (PASS == 0) ? (uint32_t)K : c[3];
    constexpr int kPer = kBins / kStreamThreads;  // 16 bins per thread
    uint32_t loc[kPer];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < kPer; ++i) { loc[i] = __ldcg(&gh[tid * kPer + i]); tsum += loc[i]; }
    part[tid] = tsum;
    __syncthreads();
    uint32_t above = 0;
    for (int u = tid + 1; u < kStreamThreads; ++u) above += part[u];
    if (above < need && need <= above + tsum) {
        uint32_t acc = above, in_bin = 0;
        int digit = 0;
#pragma unroll
        for (int i = kPer - 1; i >= 0; --i) {
            if (acc < need && need <= acc + loc[i]) { digit = tid * kPer + i; above = acc; in_bin = loc[i]; }
            acc += loc[i];
        }
        if (PASS == 0) {
            c[2] = (uint32_t)digit;
            c[3] = need - above;
            if (above + in_bin <= (uint32_t)kMaxSort) {  // whole boundary bin fits: no second pass
                c[4] = (uint32_t)digit; c[5] = need - above; c[8] = 1u; c[9] = 20u;
            } else {
                c[9] = 8u;
            }
        } else {
            c[4] = (prefix12 << 12) | (uint32_t)digit;
            c[5] = need - above;
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) gh[tid * kPer + i] = 0;
    }
}

// candidates above the selected prefix go to `above` (all of them are in the top-K), candidates sharing it go to
// `bound`; slots are reserved per CTA (one global atomic per list per CTA)
template <int MODE>
__global__ void __launch_bounds__(kStreamThreads)
topk_compact_kernel(const float* __restrict__ scores, int stride, int offset, int A, uint32_t* __restrict__ ctl,
                    uint64_t* __restrict__ above, uint64_t* __restrict__ bound) {
    __shared__ uint64_t stage[kChunk];  // above from the front, bound from the back
    __shared__ uint32_t n_a, n_b, base_a, base_b;
    const int b = blockIdx.y, tid = threadIdx.x;
    uint32_t* c = ctl + (size_t)b * kCtl;
    uint32_t key[kPerThread];
    int idx[kPerThread];
    load_chunk<MODE>(scores, stride, offset, A, b, blockIdx.x, key, idx);
    if (tid == 0) { n_a = 0; n_b = 0; }
    const uint32_t pfx = c[4], shift = c[9];
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kPerThread; ++it) {
        if (idx[it] >= 0) {
            const uint32_t kp = key[it] >> shift;
            if (kp > pfx) stage[atomicAdd(&n_a, 1u)] = make_composite(key[it], (uint32_t)idx[it]);
            else if (kp == pfx) stage[kChunk - 1 - atomicAdd(&n_b, 1u)] = make_composite(key[it], (uint32_t)idx[it]);
        }
    }
    __syncthreads();
    if (tid == 0) {
        base_a = n_a ? atomicAdd(&c[6], n_a) : 0u;
        base_b = n_b ? atomicAdd(&c[7], n_b) : 0u;
    }
    __syncthreads();
    uint64_t* ga = above + (size_t)b * kMaxSort;
    uint64_t* gb = bound + (size_t)b * kMaxSort;
    for (uint32_t i = tid; i < n_a; i += kStreamThreads) ga[base_a + i] = stage[i];  // n_above < K <= 8192
    for (uint32_t i = tid; i < n_b; i += kStreamThreads) {
        const uint32_t slot = base_b + i;
        if (slot < (uint32_t)kMaxSort) gb[slot] = stage[kChunk - 1 - i];  // overflow -> rescan path
    }
}

// one CTA per image: gather the survivors, sort (register-blocked bitonic), emit indices (+ fused ProposalLayer
// gather / decode / clip)
template <int E>
__device__ __forceinline__ void topk_sort_emit(uint64_t* s, int n, bool in_smem, const uint64_t* ga, int n_above,
                                               const uint64_t* gb, int A, int K, int b, int32_t* idx_out,
                                               float* vals_out, const TopkDecode& dec, bool has_dec) {
    const int tid = threadIdx.x;
    uint64_t v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = tid * E + e;
        uint64_t c = 0ull;
        if (i < n) c = in_smem ? s[i] : (i < n_above ? ga[i] : gb[i - n_above]);
        v[e] = c;
    }
    __syncthreads();  // s[] (if used) is consumed; the exchange buffers alias it
    block_sort_desc_blocked<E>(v, s);
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int r = tid * E + e;
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

__global__ void __launch_bounds__(1024)
topk_final_kernel(const float* __restrict__ scores, int stride, int offset, int A, int K,
                  const uint32_t* __restrict__ ctl, const uint64_t* __restrict__ above,
                  const uint64_t* __restrict__ bound, int32_t* __restrict__ idx_out, float* __restrict__ vals_out,
                  TopkDecode dec, bool has_dec) {
    extern __shared__ __align__(16) uint64_t s[];  // candidate staging (tie-flood path) / sort exchange buffers
    __shared__ uint32_t h8[256];
    __shared__ int warp_sums[32];
    __shared__ int s_total;
    __shared__ uint32_t s_kth, s_need3, s_fill;
    const int b = blockIdx.x, tid = threadIdx.x;
    const uint32_t* c = ctl + (size_t)b * kCtl;
    const uint32_t n_above = c[6], n_bound = c[7], p24 = c[4], need2 = c[5];
    const uint64_t* ga = above + (size_t)b * kMaxSort;
    const uint64_t* gb = bound + (size_t)b * kMaxSort;
    int n;
    bool in_smem = false;
    if (n_above + n_bound <= (uint32_t)kMaxSort) {
        n = (int)(n_above + n_bound);
    } else {
        // tie flood: resolve the last 8 key bits, then take equal keys in index order
        in_smem = true;
        for (int i = tid; i < (int)n_above; i += blockDim.x) s[i] = ga[i];
        if (tid < 256) h8[tid] = 0;
        if (tid == 0) s_fill = n_above;
        __syncthreads();
        for (int a = tid; a < A; a += blockDim.x) {
            const uint32_t key = load_key(scores, stride, offset, A, b, a);
            if ((key >> 8) == p24) atomicAdd(&h8[key & 0xffu], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t acc = 0;
            int d = 255;
            for (; d > 0; --d) {
                if (acc + h8[d] >= need2) break;
                acc += h8[d];
            }
            s_kth = (p24 << 8) | (uint32_t)d;
            s_need3 = need2 - acc;
        }
        __syncthreads();
        const uint32_t kth = s_kth, need3 = s_need3;
        int taken = 0;  // equal-key elements accepted so far (uniform)
        for (int a0 = 0; a0 < A; a0 += blockDim.x) {
            const int a = a0 + tid;
            uint32_t key = 0;
            bool in = false, gt = false, eq = false;
            if (a < A) {
                key = load_key(scores, stride, offset, A, b, a);
                in = (key >> 8) == p24;
                gt = in && key > kth;
                eq = key == kth;
            }
            if (gt) s[atomicAdd(&s_fill, 1u)] = make_composite(key, (uint32_t)a);
            const int rank = block_exclusive_scan(eq ? 1 : 0, warp_sums, &s_total);
            if (eq && (uint32_t)(taken + rank) < need3) s[atomicAdd(&s_fill, 1u)] = make_composite(key, (uint32_t)a);
            taken += s_total;
            __syncthreads();
        }
        n = K;
    }
    if (n <= 1024) topk_sort_emit<1>(s, n, in_smem, ga, (int)n_above, gb, A, K, b, idx_out, vals_out, dec, has_dec);
    else if (n <= 2048) topk_sort_emit<2>(s, n, in_smem, ga, (int)n_above, gb, A, K, b, idx_out, vals_out, dec, has_dec);
    else if (n <= 4096) topk_sort_emit<4>(s, n, in_smem, ga, (int)n_above, gb, A, K, b, idx_out, vals_out, dec, has_dec);
    else topk_sort_emit<8>(s, n, in_smem, ga, (int)n_above, gb, A, K, b, idx_out, vals_out, dec, has_dec);
}

int launch_topk(const float* scores, int stride, int offset, int B, int A, int K, int32_t* idx, float* vals,
                const TopkDecode* dec, void* ws, cudaStream_t stream) {
#ifndef MRCNN_TOPK_MULTIKERNEL
    // default: the single cluster kernel of topk_cluster.cu; the multi-kernel pipeline below (global histograms,
    // candidate lists in the workspace, one sorting CTA per image) is kept as a build-time alternative
    (void)ws;
    return launch_topk_cluster(scores, stride, offset, B, A, K, idx, vals, dec, stream);
#endif
    TopkWs w = carve(ws, B);
    cudaError_t e = cudaMemsetAsync(w.hist, 0, (size_t)B * (kBins + kCtl) * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return (int)e;
    const int chunks = (A + kChunk - 1) / kChunk;
    const dim3 grid(chunks, B);
    const int mode = stream_mode(scores, stride, A);
#define MRCNN_STREAM(M)                                                                                             \
    do {                                                                                                            \
        topk_hist_kernel<0, M><<<grid, kStreamThreads, 0, stream>>>(scores, stride, offset, A, K, w.hist, w.ctl, chunks); \
        topk_hist_kernel<1, M><<<grid, kStreamThreads, 0, stream>>>(scores, stride, offset, A, K, w.hist, w.ctl, chunks); \
        topk_compact_kernel<M><<<grid, kStreamThreads, 0, stream>>>(scores, stride, offset, A, w.ctl, w.above, w.bound); \
    } while (0)
    if (mode == 2) MRCNN_STREAM(2);
    else if (mode == 1) MRCNN_STREAM(1);
    else MRCNN_STREAM(0);
#undef MRCNN_STREAM
    const size_t smem = block_sort_xch_bytes(8);  // >= kMaxSort * 8: also stages the tie-flood candidates
    e = cudaFuncSetAttribute(topk_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    TopkDecode d{};
    if (dec) d = *dec;
    topk_final_kernel<<<B, 1024, smem, stream>>>(scores, stride, offset, A, K, w.ctl, w.above, w.bound, idx, vals, d,
                                                 dec != nullptr);
    return last_error();
}

}  // namespace mrcnn

using namespace mrcnn;

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
