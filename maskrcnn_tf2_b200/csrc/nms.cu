// nms.cu -- batched greedy hard NMS with TF NonMaxSuppressionV3 (CPU kernel) semantics.
// Replaces tf.image.non_max_suppression at mrcnn_layers.py:225 (RPN, thr 0.7) and :455 (detections, thr 0.3).
//
// nms_lazy_kernel: a thread-block CLUSTER of 1..16 CTAs (1024 threads each) per image; every CTA stages the image's
// candidate boxes (already in candidate order) in its own shared memory and walks the candidates in 64-box tiles.
// Only the IoU tests that can matter are evaluated, and neither the cluster nor the CTA blocks on a full barrier
// inside the loop.  Each CTA is warp-specialised:
//   worker warps (16..31), up to two tiles ahead of the resolvers:
//     far(u)  = tile u's candidates suppressed by boxes kept in tiles <= u-2: the kept list is dealt round-robin to
//               the worker warps of every CTA of the cluster; each CTA sends its 64-bit partial to every peer with
//               st.async (a distributed-shared-memory store that completes a transaction on the receiver's mbarrier);
//     diag(u) = the tile's own symmetric 64x64 block: row i is computed by CTA i % CTAs and sent the same way;
//   resolver warps (0..15), every CTA for itself, identically:
//     near(u) = suppressed by the boxes kept in tile u-1 (<= 64 x 64 tests);
//     resolve = warp 0 waits for the tile's mbarrier phase and decides the tile with ballots (fixed point over diag:
//               a candidate is kept once every earlier overlapping candidate is decided-removed, removed once one is
//               decided-kept), appends the kept ones to the CTA's copy of the kept list and releases the workers
//               for tile u+2 through a named barrier.
// Work is sum_t kept(t) * 64 + M * 64 pair tests instead of the M^2/2 of a full bit matrix, nothing is written to
// global memory but the result, and the loop stops as soon as max_out boxes are kept.
#include "nms_dev.cuh"

namespace cg = cooperative_groups;

namespace mrcnn {

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_lazy_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ valid, int M, int max_out, float thr,
                NmsEpilogue epi) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    float4* sb = reinterpret_cast<float4*>(nms_smem);          // [M] min/max-normalised corners
    float* sa = reinterpret_cast<float*>(sb + M);               // [M] areas
    int32_t* sel = reinterpret_cast<int32_t*>(sa + M);          // [min(max_out, M)] kept candidate positions
    __shared__ unsigned long long s_diag[4][kTile];             // [tile & 3][row]: symmetric in-tile block, from the peers
    __shared__ unsigned long long s_far[4][16];                 // [tile & 3][source CTA], written by the peers
    __shared__ unsigned long long s_near[2];                    // [tile parity]
    __shared__ unsigned long long s_farpart[2];                 // [tile parity] this CTA's partial of far(u)
    __shared__ __align__(8) uint64_t s_bar[4];                  // mbarriers, [tile & 3]: a peer can run at most two
                                                                // tiles ahead, so four phases never alias
    __shared__ int s_nk[4];                                     // [tile & 3] kept count after that tile's resolve
    __shared__ int s_final[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = valid ? min(max(valid[b], 0), M) : M;
    const int tiles = (n + kTile - 1) / kTile;
    const float4* bx = boxes + (size_t)b * M;
    const float4 kNone = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);  // overlaps nothing, never ambiguous
    for (int i = tid; i < n; i += kNmsThreads) {
        float a;
        float4 t = normalise_box(__ldg(bx + i), a);
        if (!(a > 0.0f)) { t = kNone; a = 1.0f; }  // TF: area <= 0 -> IoU 0 with everything
        sb[i] = t;
        sa[i] = a;
    }
    const uint32_t bar_base = smem_u32(&s_bar[0]);
    // per tile every CTA receives one 64-bit far partial from each CTA and the 64 rows of the tile's diag block
    const uint32_t far_bytes = (uint32_t)csize * 8u + (uint32_t)kTile * 8u;
    if (tid == 0) {
        s_near[0] = 0ull; s_near[1] = 0ull; s_farpart[0] = 0ull; s_farpart[1] = 0ull;
        for (int j = 0; j < 4; ++j) { mbar_init(bar_base + 8u * j, 1); s_nk[j] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 1; j < 4; ++j) mbar_arm(bar_base + 8u * j, far_bytes);  // tiles 1..3; tile 4 is armed in tile 0
    }
    __syncthreads();
    // diag(0): rows 2*warp, 2*warp+1 of tile 0, every CTA for itself
    {
        const float4 b0 = (lane < n) ? sb[lane] : kNone, b1 = (lane + 32 < n) ? sb[lane + 32] : kNone;
        const float a0 = (lane < n) ? sa[lane] : 1.0f, a1 = (lane + 32 < n) ? sa[lane + 32] : 1.0f;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int i = 2 * warp + rr;
            const float4 bi = (i < n) ? sb[i] : kNone;
            const float ai = (i < n) ? sa[i] : 1.0f;
            const uint64_t row = ballot64(iou_gt(bi, ai, b0, a0, thr), iou_gt(bi, ai, b1, a1, thr)) & ~(1ull << i);
            if (lane == 0) s_diag[0][i] = row;
        }
    }
    cluster.sync();  // every CTA of the cluster is resident and its mbarriers are initialised before any st.async
    int nkept = 0, t = 0;
    PROF_DECL;
    if (warp < kResolvers) {
        // ================= resolver warps: near(t) + resolve(t), tile by tile =================
        for (; t < tiles && nkept < max_out; ++t) {
            PROF_TILE;
            const int p = t & 1;
            const int base = t * kTile;
            const int nkept_before = nkept;  // boxes kept in tiles <= t-1
            if (warp == 0) {
                uint64_t removed = (uint64_t)s_near[p];
                const uint32_t bar_t = bar_base + 8u * (uint32_t)(t & 3);
                if (t >= 1) {  // barrier (t & 3) serves tiles t&3, t&3 + 4, ...; tile 0 has no far set
                    mbar_wait(bar_t, (uint32_t)(((t >> 2) - ((t & 3) == 0 ? 1 : 0)) & 1));
                    for (int r = 0; r < csize; ++r) removed |= (uint64_t)s_far[t & 3][r];
                }
                PROF_MARK(0);
                const int rem = n - base;
                const uint64_t validbits = (rem >= kTile) ? ~0ull : ((1ull << rem) - 1ull);
                uint64_t und = ~removed & validbits, kept = 0;
                const uint64_t blk0 = (uint64_t)s_diag[t & 3][lane] & ((1ull << lane) - 1ull);  // earlier overlapping candidates
                const uint64_t blk1 = (uint64_t)s_diag[t & 3][lane + 32] & ((1ull << (lane + 32)) - 1ull);
                while (und) {
                    const bool u0 = (und >> lane) & 1ull, u1 = (und >> (lane + 32)) & 1ull;
                    const bool d0 = u0 && (blk0 & kept), d1 = u1 && (blk1 & kept);              // removed
                    const bool k0 = u0 && !d0 && !(blk0 & und), k1 = u1 && !d1 && !(blk1 & und);  // kept
                    const uint64_t nk = ballot64(k0, k1), nd = ballot64(d0, d1);
                    kept |= nk;
                    und &= ~(nk | nd);
                }
                const int room = max_out - nkept;
                while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll(kept)));
                if ((kept >> lane) & 1ull) sel[nkept + __popcll(kept & ((1ull << lane) - 1ull))] = base + lane;
                if ((kept >> (lane + 32)) & 1ull)
                    sel[nkept + __popcll(kept & ((1ull << (lane + 32)) - 1ull))] = base + lane + 32;
                __syncwarp();
                if (lane == 0) {
                    s_nk[t & 3] = nkept + __popcll(kept);
                    s_near[p] = 0ull;            // next accumulated for tile t+2, two resolver barriers later
                    mbar_arm(bar_t, far_bytes);  // phase of tile t+4
                }
                __threadfence_block();
                // release the workers for tile t+2 (they need the kept list through tile t)
                if (t + 2 < tiles) asm volatile("bar.arrive %0, %1;" ::"r"(3 + p), "r"(32 + kWorkers * 32) : "memory");
                PROF_MARK(1);
            }
            named_barrier(1, kResolvers * 32);
            nkept = s_nk[t & 3];
            PROF_MARK(2);
            // near(t+1): boxes kept in tile t x tile t+1
            if (t + 1 < tiles && nkept < max_out && nkept_before + warp < nkept) {
                const int c0 = base + kTile + lane, c1 = c0 + 32;
                const float4 b0 = (c0 < n) ? sb[c0] : kNone, b1 = (c1 < n) ? sb[c1] : kNone;
                const float a0 = (c0 < n) ? sa[c0] : 1.0f, a1 = (c1 < n) ? sa[c1] : 1.0f;
                bool r0 = false, r1 = false, u0 = false, u1 = false;
                for (int k = nkept_before + warp; k < nkept; k += kResolvers) {
                    const int ki = sel[k];
                    const float4 bk = sb[ki];
                    const float ak = sa[ki];
                    iou_screen(bk, ak, b0, a0, thr, r0, u0);
                    iou_screen(bk, ak, b1, a1, thr, r1, u1);
                }
                if (__any_sync(0xffffffffu, u0 || u1)) {  // a pair within 2^-21 of the threshold: exact division
                    r0 = false; r1 = false;
                    for (int k = nkept_before + warp; k < nkept; k += kResolvers) {
                        const int ki = sel[k];
                        r0 |= iou_gt(sb[ki], sa[ki], b0, a0, thr);
                        r1 |= iou_gt(sb[ki], sa[ki], b1, a1, thr);
                    }
                }
                const uint64_t hit = ballot64(r0, r1);
                if (lane == 0 && hit) or_into(&s_near[p ^ 1], hit);
            }
            named_barrier(1, kResolvers * 32);
            PROF_MARK(3);
        }
    } else {
        // ================= worker warps: far(u) share + diag(u) rows, u = 1 .. tiles-1 =================
        const int wk = warp - kResolvers;  // 0..15
        for (int u = 1; u < tiles; ++u) {
            int nk = 0;  // boxes kept in tiles <= u-2
            if (u >= 2) {
                asm volatile("bar.sync %0, %1;" ::"r"(3 + (u & 1)), "r"(32 + kWorkers * 32) : "memory");
                nk = s_nk[(u - 2) & 3];
                if (nk >= max_out) break;  // the resolvers stop after tile u-2
            }
            const int ubase = u * kTile;
            const int c0 = ubase + lane, c1 = c0 + 32;
            const float4 b0 = (c0 < n) ? sb[c0] : kNone, b1 = (c1 < n) ? sb[c1] : kNone;
            const float a0 = (c0 < n) ? sa[c0] : 1.0f, a1 = (c1 < n) ? sa[c1] : 1.0f;
            bool r0 = false, r1 = false, u0 = false, u1 = false;
#pragma unroll 4
            for (int k = crank * kWorkers + wk; k < nk; k += csize * kWorkers) {
                const int ki = sel[k];
                const float4 bk = sb[ki];
                const float ak = sa[ki];
                iou_screen(bk, ak, b0, a0, thr, r0, u0);
                iou_screen(bk, ak, b1, a1, thr, r1, u1);
            }
            if (__any_sync(0xffffffffu, u0 || u1)) {  // a pair within 2^-21 of the threshold: exact division
                r0 = false; r1 = false;
                for (int k = crank * kWorkers + wk; k < nk; k += csize * kWorkers) {
                    const int ki = sel[k];
                    r0 |= iou_gt(sb[ki], sa[ki], b0, a0, thr);
                    r1 |= iou_gt(sb[ki], sa[ki], b1, a1, thr);
                }
            }
            const uint64_t hit = ballot64(r0, r1);
            if (lane == 0 && hit) or_into(&s_farpart[u & 1], hit);
            const uint32_t bar_u = bar_base + 8u * (uint32_t)(u & 3);
            for (int i = crank + csize * wk; i < kTile; i += csize * kWorkers) {  // this CTA's rows of diag(u)
                const int ci = ubase + i;
                const float4 bi = (ci < n) ? sb[ci] : kNone;
                const float ai = (ci < n) ? sa[ci] : 1.0f;
                const uint64_t row = ballot64(iou_gt(bi, ai, b0, a0, thr), iou_gt(bi, ai, b1, a1, thr)) & ~(1ull << i);
                if (lane < csize)
                    st_async_u64(mapa_u32(smem_u32(&s_diag[u & 3][i]), (uint32_t)lane), row, mapa_u32(bar_u, (uint32_t)lane));
            }
            named_barrier(2, kWorkers * 32);  // all partials of this CTA are in s_farpart[u & 1]
            if (wk == 0) {
                const unsigned long long part = s_farpart[u & 1];
                __syncwarp();
                if (lane < csize)
                    st_async_u64(mapa_u32(smem_u32(&s_far[u & 3][crank]), (uint32_t)lane), part, mapa_u32(bar_u, (uint32_t)lane));
                if (lane == 0) s_farpart[u & 1] = 0ull;  // next used for tile u+2, one worker barrier later
            }
        }
    }
    if (tid == 0) { s_final[0] = t; s_final[1] = nkept; }  // the resolvers' loop state, for everybody
    __syncthreads();
    t = s_final[0];
    nkept = s_final[1];
    PROF_DUMP;
    // far(t) was sent during the last iteration but never consumed: drain it, so that no st.async is in flight
    // towards this CTA when it exits; the cluster barrier then keeps every CTA alive until its peers have drained
    if (warp == 0 && t >= 1 && t < tiles)
        mbar_wait(bar_base + 8u * (uint32_t)(t & 3), (uint32_t)(((t >> 2) - ((t & 3) == 0 ? 1 : 0)) & 1));
    cluster.sync();
    if (crank != 0) return;
    const int total = nkept;
    nms_write_outputs(epi, bx, b, M, max_out, total, sel);
}

// cluster size: spread one image over as many SMs as the batch leaves free (148 SMs, 1 CTA per SM), up to the
// portable maximum of 8 (16-CTA clusters do not co-schedule for 8 images on this part: measured); small candidate
// sets do not amortise the exchange
static int nms_cluster_size(int B, int M, size_t smem) {
    if (M <= 2048) return 1;
    static int cache[4][2] = {};
    return pick_cluster_size(nms_lazy_kernel, kNmsThreads, B, 8, [smem](int) { return smem; }, cache);
}

int launch_nms_sorted(const float4* boxes_sorted, const int32_t* valid, int B, int M, int max_out, float thr,
                      const NmsEpilogue& epi, cudaStream_t stream) {
    const size_t smem = nms_smem_bytes(M, max_out);
    cudaError_t e = cudaFuncSetAttribute(nms_lazy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int cs = nms_cluster_size(B, M, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, nms_lazy_kernel, boxes_sorted, valid, M, max_out, thr, epi);
    if (e != cudaSuccess) return (int)e;
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
    block_sort_desc_any(s, sort_n, s + sort_n);
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
};
static size_t nms_ws_bytes(int B, int M) {
    return align_up((size_t)B * M * sizeof(float4), 256) + align_up((size_t)B * M * sizeof(int32_t), 256) +
           align_up((size_t)B * sizeof(int32_t), 256);
}

}  // namespace mrcnn

using namespace mrcnn;

#ifdef MRCNN_NMS_PROFILE
MRCNN_EXPORT int mrcnn_debug_nms_profile(long long* host_out8) {
    return (int)cudaMemcpyFromSymbol(host_out8, g_nms_prof, sizeof(long long) * 8);
}
#endif

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
    w.ncand = (int32_t*)p;
    const int sort_n = next_pow2(M < 32 ? 32 : M);
    const size_t smem = (size_t)sort_n * sizeof(uint64_t) + (sort_n >= 1024 ? block_sort_xch_bytes(sort_n / 1024) : 0);
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
    return launch_nms_sorted(w.boxes_sorted, w.ncand, B, M, max_out, thr, epi, st);
}
