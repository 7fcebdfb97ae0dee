// nms.cu -- batched greedy hard NMS with TF NonMaxSuppressionV3 (CPU kernel) semantics.
// Replaces tf.image.non_max_suppression at mrcnn_layers.py:225 (RPN, thr 0.7) and :455 (detections, thr 0.3).
//
// Three stages, all asynchronous on one stream:
//   1. (generic entry only) per-image sort of the candidates by (score desc, index asc);
//   2. nms_mask_kernel: 64x64 tiles of the upper-triangular "IoU > thr" bit matrix, column boxes staged in
//      shared memory, one 64-bit word per (row, column tile); diagonal tiles are computed symmetric;
//   3. nms_sweep_kernel: one CTA per image walks the tiles in order.  Inside a tile the keep/suppress
//      decisions are resolved by a warp with ballots (fixed-point over the symmetric diagonal block: a
//      candidate is kept once every earlier overlapping candidate is decided-removed, removed once one is
//      decided-kept); the rows of the kept boxes are then OR-ed into the per-word "removed" registers.
//      Stops as soon as max_out boxes are kept.
#include "common.cuh"

namespace mrcnn {

constexpr int kTile = 64;
constexpr int kSweepThreads = 128;  // one removed-word per thread: up to 128 * 64 = 8192 candidates

size_t nms_mask_bytes(int B, int M) {
    const size_t words = (size_t)((M + kTile - 1) / kTile);
    return align_up((size_t)B * M * words * sizeof(uint64_t), 256);
}

__global__ void __launch_bounds__(kTile)
nms_mask_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ valid, int M, int words, float thr,
                uint64_t* __restrict__ mask) {
    const int c = blockIdx.x, r = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
    if (c < r) return;
    const int n = valid ? min(max(valid[b], 0), M) : M;
    if (r * kTile >= n || c * kTile >= n) return;
    __shared__ float4 cb[kTile];
    __shared__ float ca[kTile];
    const float4* bx = boxes + (size_t)b * M;
    {
        const int j = c * kTile + tid;
        float4 nb = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);  // never overlaps, never ambiguous
        float a = 1.0f;
        if (j < n) {
            float aj;
            const float4 t = normalise_box(__ldg(bx + j), aj);
            if (aj > 0.0f) { nb = t; a = aj; }  // TF: area <= 0 -> IoU 0
        }
        cb[tid] = nb;
        ca[tid] = a;
    }
    __syncthreads();
    const int i = r * kTile + tid;
    if (i >= n) return;
    float ai;
    const float4 bi = normalise_box(__ldg(bx + i), ai);
    uint64_t bits = 0;
    if (ai > 0.0f) {
        const int ncol = min(kTile, n - c * kTile);
#pragma unroll 8
        for (int j = 0; j < ncol; ++j) {
            if (iou_gt(bi, ai, cb[j], ca[j], thr)) bits |= (1ull << j);
        }
        if (c == r) bits &= ~(1ull << tid);
    }
    mask[((size_t)b * M + i) * words + c] = bits;
}

__device__ __forceinline__ uint64_t ballot64(bool lo, bool hi) {
    return (uint64_t)__ballot_sync(0xffffffffu, lo) | ((uint64_t)__ballot_sync(0xffffffffu, hi) << 32);
}

__global__ void __launch_bounds__(kSweepThreads)
nms_sweep_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ valid, int M, int words, int max_out,
                 const uint64_t* __restrict__ mask, NmsEpilogue epi) {
    extern __shared__ int32_t sel[];  // min(max_out, M) selected sorted positions
    __shared__ uint64_t s_cur, s_kept;
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = valid ? min(max(valid[b], 0), M) : M;
    const int tiles = (n + kTile - 1) / kTile;
    const uint64_t* mk = mask + (size_t)b * M * words;
    uint64_t my_removed = 0;  // removed bits of word `tid`
    int total = 0;
    // warp 0 keeps the diagonal words of the next tile in flight
    uint64_t nlo = 0, nhi = 0;
    if (tid < 32 && tiles > 0) {
        if (lane < n) nlo = __ldg(mk + (size_t)lane * words);
        if (lane + 32 < n) nhi = __ldg(mk + (size_t)(lane + 32) * words);
    }
    for (int t = 0; t < tiles && total < max_out; ++t) {
        if (tid == t) s_cur = my_removed;
        __syncthreads();
        if (tid < 32) {
            const uint64_t rlo = nlo, rhi = nhi;
            if (t + 1 < tiles) {  // prefetch next diagonal block
                const int r0 = (t + 1) * kTile + lane, r1 = r0 + 32;
                nlo = (r0 < n) ? __ldg(mk + (size_t)r0 * words + t + 1) : 0ull;
                nhi = (r1 < n) ? __ldg(mk + (size_t)r1 * words + t + 1) : 0ull;
            }
            const int rem = n - t * kTile;
            const uint64_t validbits = (rem >= kTile) ? ~0ull : ((1ull << rem) - 1ull);
            uint64_t und = ~s_cur & validbits, kept = 0;
            const uint64_t blk0 = rlo & ((1ull << lane) - 1ull);          // earlier overlapping candidates
            const uint64_t blk1 = rhi & ((1ull << (lane + 32)) - 1ull);
            while (und) {
                const bool u0 = (und >> lane) & 1ull, u1 = (und >> (lane + 32)) & 1ull;
                const bool d0 = u0 && (blk0 & kept), d1 = u1 && (blk1 & kept);          // removed
                const bool k0 = u0 && !d0 && !(blk0 & und), k1 = u1 && !d1 && !(blk1 & und);  // kept
                const uint64_t nk = ballot64(k0, k1), nd = ballot64(d0, d1);
                kept |= nk;
                und &= ~(nk | nd);
            }
            const int room = max_out - total;
            while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll(kept)));
            if ((kept >> lane) & 1ull) sel[total + __popcll(kept & ((1ull << lane) - 1ull))] = t * kTile + lane;
            if ((kept >> (lane + 32)) & 1ull)
                sel[total + __popcll(kept & ((1ull << (lane + 32)) - 1ull))] = t * kTile + lane + 32;
            if (lane == 0) { s_kept = kept; s_total = total + __popcll(kept); }
        }
        __syncthreads();
        uint64_t kept = s_kept;
        total = s_total;
        if (total >= max_out) break;
        if (tid > t && tid < tiles) {
            const uint64_t* col = mk + (size_t)t * kTile * words + tid;
            while (kept) {  // up to 4 independent row loads in flight
                const int j0 = __ffsll((long long)kept) - 1; kept &= kept - 1;
                uint64_t v = __ldg(col + (size_t)j0 * words);
                if (kept) { const int j1 = __ffsll((long long)kept) - 1; kept &= kept - 1; v |= __ldg(col + (size_t)j1 * words); }
                if (kept) { const int j2 = __ffsll((long long)kept) - 1; kept &= kept - 1; v |= __ldg(col + (size_t)j2 * words); }
                if (kept) { const int j3 = __ffsll((long long)kept) - 1; kept &= kept - 1; v |= __ldg(col + (size_t)j3 * words); }
                my_removed |= v;
            }
        }
    }
    __syncthreads();
    total = (tiles > 0) ? s_total : 0;
    // ---- epilogue: fixed-size padded outputs, no host round trip --------------------------------
    if (epi.mode == 0) {
        for (int r = tid; r < max_out; r += blockDim.x) {
            int32_t v = -1;
            if (r < total) v = epi.orig_idx ? epi.orig_idx[(size_t)b * M + sel[r]] : sel[r];
            epi.keep[(size_t)b * max_out + r] = v;
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    } else if (epi.mode == 1) {  // ProposalLayer.nms L:227-230: gather + zero pad
        for (int r = tid; r < max_out; r += blockDim.x) {
            epi.proposals[(size_t)b * max_out + r] =
                (r < total) ? boxes[(size_t)b * M + sel[r]] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (epi.keep) epi.keep[(size_t)b * max_out + r] = (r < total) ? sel[r] : -1;
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    } else {  // refine_detections L:494-500: [y1,x1,y2,x2,class,score] rows + zero pad
        for (int r = tid; r < max_out; r += blockDim.x) {
            float* o = epi.detections + ((size_t)b * max_out + r) * 6;
            if (r < total) {
                const int i = epi.orig_idx[(size_t)b * M + sel[r]];
                const float4 bx = epi.refined[(size_t)b * epi.N + i];
                o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
                o[4] = (float)epi.class_ids[(size_t)b * epi.N + i];
                o[5] = epi.scores[(size_t)b * epi.N + i];
            } else {
                o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.0f;
            }
        }
        if (tid == 0 && epi.count) epi.count[b] = total;
    }
}

int launch_nms_sorted(const float4* boxes_sorted, const int32_t* valid, int B, int M, int max_out, float thr,
                      uint64_t* mask, const NmsEpilogue& epi, cudaStream_t stream) {
    const int words = (M + kTile - 1) / kTile;
    const dim3 grid(words, words, B);
    nms_mask_kernel<<<grid, kTile, 0, stream>>>(boxes_sorted, valid, M, words, thr, mask);
    const size_t smem = (size_t)(max_out < M ? max_out : M) * sizeof(int32_t);  // never more than M kept
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    nms_sweep_kernel<<<B, kSweepThreads, smem, stream>>>(boxes_sorted, valid, M, words, max_out, mask, epi);
    return last_error();
}

// generic entry: sort candidates (score > -inf) by (score desc, index asc); one CTA per image
__global__ void __launch_bounds__(1024)
nms_sort_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ valid,
                int M, float4* __restrict__ boxes_sorted, int32_t* __restrict__ orig_idx, int32_t* __restrict__ ncand) {
    extern __shared__ __align__(16) uint64_t s[];
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = valid ? min(max(valid[b], 0), M) : M;
    const int sort_n = max(32, 1 << (32 - __clz(max(M, 1) - 1)));
    if (tid == 0) s_n = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < sort_n; i += blockDim.x) {
        uint64_t comp = 0ull;
        if (i < n) {
            const uint32_t key = score_key(__ldg(scores + (size_t)b * M + i));
            if (key > kKeyNegInf) { comp = make_composite(key, (uint32_t)i); ++local; }
        }
        s[i] = comp;
    }
    if (local) atomicAdd(&s_n, local);
    __syncthreads();
    block_bitonic_sort_desc(s, sort_n);
    const int nc = s_n;
    for (int r = tid; r < M; r += blockDim.x) {
        if (r < nc) {
            const int i = (int)composite_idx(s[r]);
            orig_idx[(size_t)b * M + r] = i;
            boxes_sorted[(size_t)b * M + r] = __ldg(boxes + (size_t)b * M + i);
        }
    }
    if (tid == 0) ncand[b] = nc;
}

struct NmsWs {
    float4* boxes_sorted;
    int32_t* orig_idx;
    int32_t* ncand;
    uint64_t* mask;
};
static size_t nms_ws_bytes(int B, int M) {
    return align_up((size_t)B * M * sizeof(float4), 256) + align_up((size_t)B * M * sizeof(int32_t), 256) +
           align_up((size_t)B * sizeof(int32_t), 256) + nms_mask_bytes(B, M);
}

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_nms_workspace_bytes(int B, int M, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || M < 1 || M > kMaxSort) return MRCNN_ERR_RANGE;
    *bytes = nms_ws_bytes(B, M);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_nms_forward(const float* boxes, const float* scores, const int32_t* valid, int B, int M,
                                   int max_out, float thr, int32_t* keep, int32_t* count, void* ws, size_t ws_bytes,
                                   void* stream) {
    if (!boxes || !scores || !keep || !ws) return MRCNN_ERR_NULL;
    if (B < 1 || M < 1 || M > kMaxSort || max_out < 1 || !(thr >= 0.0f && thr <= 1.0f)) return MRCNN_ERR_RANGE;
    if (ws_bytes < nms_ws_bytes(B, M)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(boxes) || !aligned16(ws)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    NmsWs w;
    char* p = (char*)ws;
    w.boxes_sorted = (float4*)p; p += align_up((size_t)B * M * sizeof(float4), 256);
    w.orig_idx = (int32_t*)p;    p += align_up((size_t)B * M * sizeof(int32_t), 256);
    w.ncand = (int32_t*)p;       p += align_up((size_t)B * sizeof(int32_t), 256);
    w.mask = (uint64_t*)p;
    const int sort_n = next_pow2(M < 32 ? 32 : M);
    const size_t smem = (size_t)sort_n * sizeof(uint64_t);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    nms_sort_kernel<<<B, 1024, smem, st>>>((const float4*)boxes, scores, valid, M, w.boxes_sorted, w.orig_idx, w.ncand);
    NmsEpilogue epi{};
    epi.mode = 0;
    epi.orig_idx = w.orig_idx;
    epi.keep = keep;
    epi.count = count;
    return launch_nms_sorted(w.boxes_sorted, w.ncand, B, M, max_out, thr, w.mask, epi, st);
}
