// nms.cu -- batched greedy hard NMS with TF NonMaxSuppressionV3 (CPU kernel) semantics.
// Replaces tf.image.non_max_suppression at mrcnn_layers.py:225 (RPN, thr 0.7) and :455 (detections, thr 0.3).
//
// Three code paths.  Ordered candidate lists of more than 2048 boxes (ProposalLayer) go to nms_sweep.cu, the round-2
// redesign of this file's cluster kernel.  Single-CTA problems that order their own candidates and end up with at most
// 256 of them (DetectionLayer) build the whole strict-lower-triangle overlap matrix and let one warp resolve it
// (nms_lazy_kernel, "few candidates").  Everything else -- and whatever does not fit the sweep kernel's shared memory --
// runs the pipeline below.
//
// nms_lazy_kernel: a thread-block CLUSTER of 1..8 CTAs per image; every CTA stages the image's candidate boxes
// (already in candidate order) in its own shared memory and walks the candidates in 64-box tiles.  Only the IoU tests
// that can matter are evaluated, and neither the cluster nor the CTA blocks on a full barrier inside the loop.  Each CTA
// is warp-specialised (the other warps only help with the staging and exit):
//   far warps, up to kDepth tiles ahead of the resolver: far(u) = tile u's candidates suppressed by boxes kept in tiles
//     <= u-kDepth.  The kept list is dealt round-robin to the far warps of the whole cluster; each streams its boxes
//     from dense shared-memory copies (box, thr * area) four at a time through a multiplicative IoU screen (one FFMA
//     per pair; the exact division only inside a 2^-20 band) and sends its 64-bit partial straight to every CTA;
//   row warps, one more tile ahead: diag(u) = the tile's own symmetric 64x64 block and cross_d(u) = tile u-d (rows) x
//     tile u (columns), d < kDepth -- independent of what is kept.  The rows are dealt over the cluster in contiguous
//     blocks, staged locally and shipped as ONE bulk copy per peer;
//     everything travels through distributed shared memory (st.async / cp.async.bulk) and completes transactions on
//     the receiver's mbarrier;
//   resolver warp (0), every CTA for itself, identically, alone on its scheduler: waits for the tile's mbarrier phase,
//     ORs the far partials and the cross rows of the boxes kept in the previous kDepth-1 tiles (one warp-wide redux, no
//     IoU test on the critical path), then decides the tile with ballots (fixed point over diag: a candidate is kept
//     once every earlier overlapping candidate is decided-removed, removed once one is decided-kept), appends the kept
//     ones to the CTA's copy of the kept list and releases the workers for tile u+kDepth through a named barrier.
// Work is sum_t kept(t) * 64 + kDepth * M * 64 pair tests instead of the M^2/2 of a full bit matrix, nothing is written
// to global memory but the result, and the loop stops as soon as max_out boxes are kept.  What bounds it (measured,
// DESIGN.md): the per-tile instruction count of the far and row warps on the SM's four schedulers and the resolver's
// serial chain of ~300 dependent instructions per tile.
#include <cstdlib>

#include "nms_dev.cuh"

namespace cg = cooperative_groups;

namespace mrcnn {

// Where the candidates come from.  Ordered input: `boxes` [B,M] is already in candidate order (ProposalLayer: the
// top-k output) and `valid` [B] (or NULL = M) says how many rows count.  Unordered input (FUSED kernels, single-CTA
// problems): `boxes` is in input order and the kernel orders the candidates itself, by `keys` [B,M] (order-preserving
// score keys, 0 = not a candidate: DetectionLayer's keep set, L:402-414) or else by `scores` [B,M] (candidates:
// score > -inf and not NaN, as NonMaxSuppressionV3).
struct NmsInput {
    const float4* boxes;
    const int32_t* valid;
    const float* scores;
    const uint32_t* keys;
    // ordered input only: the diag / cross rows of every tile, precomputed by nms_rows_kernel ([B][tiles][kRows] words);
    // NULL = the kernel's own row warps compute them
    const unsigned long long* rows;
    bool unit;   // every box lies inside the unit square (ProposalLayer's clipped boxes)
};

// The rows of a tile -- its own symmetric 64 x 64 block and tile v-1 (rows) x tile v (columns) -- do not depend on what is
// kept, only on the candidate order.  For ordered input they CAN be computed ahead of the sweep by this kernel, over the
// whole GPU (one CTA per tile, one warp per 32 rows), instead of by row warps inside the 8-CTA clusters of the sweep
// (25 % of the sweep's instructions are row work); the sweep then fetches a tile's 1 KB of rows with one bulk copy,
// eight tiles ahead.  Opt-in (MRCNN_NMS_GLOBAL_ROWS=1): measured on B200 at config 2 it is NOT faster (ProposalLayer
// 132-136 us against 130 us with the row warps: the sweep is bound by the far warps' loop, not by the rows).
__global__ void __launch_bounds__(128)
nms_rows_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ valid, int M, int tiles_max, float thr,
                unsigned long long* __restrict__ rows) {
    __shared__ float4 s_b[2 * kTile];
    __shared__ float s_a[2 * kTile];
    const int v = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int n = valid ? min(max(valid[b], 0), M) : M;
    const float4 kNone = make_float4(1.0e18f, 1.0e18f, -1.0e18f, -1.0e18f);
    {   // s_b[0..63] = tile v, s_b[64..127] = tile v-1
        const int c = (tid < kTile) ? v * kTile + tid : (v - 1) * kTile + (tid - kTile);
        float a = 1.0f;
        float4 t = kNone;
        if (c >= 0 && c < n) {
            t = normalise_box(__ldg(boxes + (size_t)b * M + c), a);
            if (!(a > 0.0f)) { t = kNone; a = 1.0f; }
        }
        s_b[tid] = t;
        s_a[tid] = a;
    }
    __syncthreads();
    const float4 b0 = s_b[lane], b1 = s_b[lane + 32];
    const float a0 = s_a[lane], a1 = s_a[lane + 32];
    const float tc0 = __fmul_rn(thr, a0), tc1 = __fmul_rn(thr, a1);
    const float cthr = __fadd_rn(1.0f, thr);
    CandPair cp;
    cp.y1a = b0.x; cp.x1a = b0.y; cp.y2a = b0.z; cp.x2a = b0.w;
    cp.y1b = b1.x; cp.x1b = b1.y; cp.y2b = b1.z; cp.x2b = b1.w;
    cp.ntac = pack_f2(-tc0, -tc1);
    const unsigned long long c1_2 = pack_f2(cthr, cthr);
    unsigned long long mine = 0ull;   // lane j keeps row 32 * warp + j
    for (int j = 0; j < 32; ++j) {
        const int r = warp * 32 + j;  // 0..63 diag rows, 64..127 cross rows
        const float4 bi = s_b[r];
        const float ai = s_a[r], ti = __fmul_rn(thr, ai);
        float e0, e1;
        iou_screen_d2(bi, -ti, cp, c1_2, e0, e1);
        const float m0 = __fmul_rn(__fadd_rn(ti, tc0), kScreenBand), m1 = __fmul_rn(__fadd_rn(ti, tc1), kScreenBand);
        bool h0 = e0 > m0, h1 = e1 > m1;
        if (__any_sync(0xffffffffu, fabsf(e0) <= m0 || fabsf(e1) <= m1)) {
            PROF_FALLBACK;
            h0 = iou_gt(bi, ai, b0, a0, thr);
            h1 = iou_gt(bi, ai, b1, a1, thr);
        }
        uint64_t row = ballot64(h0, h1);
        if (r < kTile) row &= ~(1ull << r);
        if (lane == j) mine = row;
    }
    rows[((size_t)b * tiles_max + v) * kRows + warp * 32 + lane] = mine;
}

// COMPACT: the kept boxes are also appended, by the resolver, to dense arrays (box, thr * area) so that the far loop
// streams them with plain strided shared-memory loads; false (shared memory too small for the copies): the far loop
// goes through the kept index list.
// FUSED: order the candidates in the kernel (cluster size 1 only).
template <bool COMPACT, bool FUSED>
__global__ void __launch_bounds__(kNmsThreads, 1)
nms_lazy_kernel(NmsInput in, int M, int max_out, float thr, int nfar, int nrow, int roles, NmsEpilogue epi) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    const int cap = min(max_out, M);
    // fused ordering: composites and the candidate -> row map sit in front of the NMS arrays
    uint64_t* s_comp = reinterpret_cast<uint64_t*>(nms_smem);
    const int sort_n_max = FUSED ? max(32, 1 << (32 - __clz(max(M, 1) - 1))) : 0;
    int32_t* s_orig = reinterpret_cast<int32_t*>(nms_smem + (size_t)sort_n_max * 8);
    unsigned char* body = nms_smem + (FUSED ? (size_t)sort_n_max * 8 + (((size_t)M * 4 + 15) & ~(size_t)15) : 0);
    float4* sb = reinterpret_cast<float4*>(body);              // [M] min/max-normalised corners
    // COMPACT: kept box k of the selection order belongs to far warp (k % nsrc) of the cluster; the far warps of THIS CTA
    // get dense private copies [far warp][k / nsrc] of (box, -thr * area), so that a far warp streams its share with
    // consecutive shared-memory loads (capacity kCompactSlack entries above cap: nms_smem_bytes)
    float4* kb = sb + M;                                        // [nfar][wcap]
    float* sa = reinterpret_cast<float*>(kb + (COMPACT ? cap + kCompactSlack : 0));  // [M] areas
    float* kt = sa + M;                                         // [nfar][wcap]
    int32_t* sel = reinterpret_cast<int32_t*>(kt + (COMPACT ? cap + kCompactSlack : 0));  // [cap] kept candidate positions
    __shared__ __align__(16) unsigned long long s_rows[kRing][kRows];   // [tile % kRing][0..63 diag rows, 64 d + i: cross_d rows]
    __shared__ __align__(16) unsigned long long s_stage[kRing][kRows];  // this CTA's rows on their way out
    __shared__ unsigned long long s_far[kRing][kMaxFarSrc];     // [tile % kRing][source CTA * nfar + far warp], written by the peers
    __shared__ __align__(8) uint64_t s_bar[kRing];              // mbarriers, [tile % kRing] (ring depth: nms_dev.cuh)
    __shared__ __align__(8) uint64_t s_rbar[kRing];             // precomputed rows (in.rows): their bulk copies land here
    __shared__ int s_nk[kRing];                                 // [tile % kRing] kept count after that tile's resolve
    __shared__ int s_final[2];
    __shared__ int s_ncand;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_base = smem_u32(&s_bar[0]);
    const int nsrc = csize * nfar;                   // far partials per tile: one per far warp of the cluster
    const int wcap = (cap + nsrc - 1) / nsrc + 1;    // entries of one far warp's private kept list
    const uint32_t nsrc_magic = ((1u << 20) + (uint32_t)nsrc - 1u) / (uint32_t)nsrc;  // k / nsrc == (k * magic) >> 20, k < 2^13
    // per tile every CTA receives one 64-bit far partial from each far warp of the cluster and, as one bulk copy per
    // CTA, the rows of the tile's diag + cross blocks
    const bool grows = in.rows != nullptr;          // rows precomputed in global memory: no row warps in this kernel
    if (grows) nrow = 0;
    const int rel_threads = 32 * (1 + nfar + nrow);  // the release barrier: resolver (arrives) + far and row warps (wait)
    const uint32_t far_bytes = (uint32_t)nsrc * 8u + (grows ? 0u : (uint32_t)kRows * 8u);
    const uint32_t rbar_base = smem_u32(&s_rbar[0]);
    PROF_TL(100, 0);
    pdl_launch_dependents();
    if (tid == 0) {
        for (int j = 0; j < kRing; ++j) { mbar_init(bar_base + 8u * j, 1); mbar_init(rbar_base + 8u * j, 1); s_nk[j] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 1; j < kRing; ++j) mbar_arm(bar_base + 8u * j, far_bytes);  // tiles 1..7; tile 8 is armed in tile 0
        s_ncand = 0;
    }
    pdl_wait();  // everything above overlaps the tail of the producing kernel; global memory is touched from here on
    PROF_TL(100, 1);
    int n = in.valid ? min(max(in.valid[b], 0), M) : M;
    const float4* bx = in.boxes + (size_t)b * M;
    const float4 kNone = make_float4(1.0e18f, 1.0e18f, -1.0e18f, -1.0e18f);  // overlaps nothing, never ambiguous, no inf
    if (FUSED) {
        // candidate order = (score desc, input row asc): TF's NonMaxSuppressionV3 priority queue (L:455; L:225 via
        // mrcnn_nms_forward).  Only the candidates are sorted (DetectionLayer: ~15 % of the ROIs).
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
            const int i = i0 + tid;
            uint32_t key = 0u;
            if (i < n) {
                if (in.keys) key = in.keys[(size_t)b * M + i];
                else { key = score_key(__ldg(in.scores + (size_t)b * M + i)); if (key <= kKeyNegInf) key = 0u; }
            }
            const unsigned vote = __ballot_sync(0xffffffffu, key != 0u);
            int base = 0;
            if (lane == 0 && vote) base = atomicAdd(&s_ncand, __popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (key) s_comp[base + __popc(vote & ((1u << lane) - 1u))] = make_composite(key, (uint32_t)i);
        }
        __syncthreads();
        PROF_TL(100, 2);
        const int nc = s_ncand;
        n = nc;
        if (nc <= 512) {
            // few candidates (DetectionLayer: ~150 of 1000 ROIs): rank by counting -- thread r counts the composites above
            // its own (broadcast shared-memory loads, no barrier) and stages its box at that rank directly
            for (int r = tid; r < nc; r += kNmsThreads) {
                const uint64_t x = s_comp[r];
                int rank = 0;
                for (int j = 0; j < nc; ++j) rank += (s_comp[j] > x);
                const int i = (int)composite_idx(x);
                s_orig[rank] = i;
                float a;
                float4 t = normalise_box(__ldg(bx + i), a);
                if (!(a > 0.0f)) { t = kNone; a = 1.0f; }
                sb[rank] = t;
                sa[rank] = a;
            }
        } else {
            const int sort_n = max(32, 1 << (32 - __clz(max(nc, 1) - 1)));
            for (int i = nc + tid; i < sort_n; i += kNmsThreads) s_comp[i] = 0ull;
            __syncthreads();
            block_sort_desc_any(s_comp, sort_n, reinterpret_cast<uint64_t*>(body));  // exchange buffers alias the NMS arrays
            __syncthreads();
            for (int r = tid; r < n; r += kNmsThreads) {
                const int i = (int)composite_idx(s_comp[r]);
                s_orig[r] = i;
                float a;
                float4 t = normalise_box(__ldg(bx + i), a);
                if (!(a > 0.0f)) { t = kNone; a = 1.0f; }
                sb[r] = t;
                sa[r] = a;
            }
        }
    } else {
        for (int i = tid; i < n; i += kNmsThreads) {
            float a;
            float4 t = normalise_box(__ldg(bx + i), a);
            if (!(a > 0.0f)) { t = kNone; a = 1.0f; }  // TF: area <= 0 -> IoU 0 with everything
            sb[i] = t;
            sa[i] = a;
        }
    }
    PROF_TL(100, 3);
    if (FUSED && n <= kSmallN) {
        // ---- few candidates (DetectionLayer: ~150 per image): the whole strict-lower-triangle overlap matrix first, all
        // threads, then ONE warp decides tile after tile from its own columns -- no far / row pipeline, whose fill and
        // drain cost 14 k cycles for three tiles (measured); the matrix sits in the row ring the pipeline would use
        uint32_t* colm = reinterpret_cast<uint32_t*>(&s_rows[0][0]);   // [candidate c][8] bit i of word w: candidate 32 w + i < c overlaps c
        static_assert(sizeof(s_rows) >= (size_t)kSmallN * 8 * 4, "the row ring holds the small-problem matrix");
        __syncthreads();   // sb / sa staged
        {
            // one warp per 32-row job (column tile t, row tile r <= t, half h): lanes own two candidates of tile t and keep
            // their own masks (bit i = row 64 r + 32 h + i overlaps my candidate), rows are broadcast loads
            const int T = (n + kTile - 1) / kTile;
            const float cthr = __fadd_rn(1.0f, thr);
            const unsigned long long c1_2 = pack_f2(cthr, cthr);
            for (int job = warp; job < T * (T + 1); job += kNmsWarps) {
                int t = 0, rem = job >> 1;
                while (rem > t) { rem -= t + 1; ++t; }   // jobs of tile t: (t + 1) row tiles x 2 halves
                const int r = rem, h = job & 1;
                const int c0 = t * kTile + lane, c1 = c0 + 32;
                const float4 b0 = (c0 < n) ? sb[c0] : kNone, b1 = (c1 < n) ? sb[c1] : kNone;
                const float a0 = (c0 < n) ? sa[c0] : 1.0f, a1 = (c1 < n) ? sa[c1] : 1.0f;
                const float tc0 = __fmul_rn(thr, a0), tc1 = __fmul_rn(thr, a1);
                CandPair cp;
                cp.y1a = b0.x; cp.x1a = b0.y; cp.y2a = b0.z; cp.x2a = b0.w;
                cp.y1b = b1.x; cp.x1b = b1.y; cp.y2b = b1.z; cp.x2b = b1.w;
                cp.ntac = pack_f2(-tc0, -tc1);
                const int row0 = r * kTile + 32 * h;
                uint32_t m0 = 0u, m1 = 0u;
                float amin = 1.0f;   // min over the pairs of |d| - band: <= 0 = some pair within the band
#pragma unroll 8
                for (int i = 31; i >= 0; --i) {   // downwards: every row shifts its bit in from the right
                    const int ri = row0 + i;
                    const float4 bi = (ri < n) ? sb[ri] : kNone;
                    const float nti = -__fmul_rn(thr, (ri < n) ? sa[ri] : 1.0f);
                    float e0, e1;
                    iou_screen_d2(bi, nti, cp, c1_2, e0, e1);
                    const float g0 = __fmul_rn(__fsub_rn(tc0, nti), kScreenBand), g1 = __fmul_rn(__fsub_rn(tc1, nti), kScreenBand);
                    m0 = __funnelshift_l(__float_as_uint(__fsub_rn(g0, e0)), m0, 1);   // sign(band - d) = (d > band)
                    m1 = __funnelshift_l(__float_as_uint(__fsub_rn(g1, e1)), m1, 1);
                    amin = fmin3(amin, __fsub_rn(fabsf(e0), g0), __fsub_rn(fabsf(e1), g1));
                }
                if (__any_sync(0xffffffffu, amin <= 0.0f)) {   // a pair within 2^-20 of the threshold: exact division
                    PROF_FALLBACK;
                    m0 = 0u; m1 = 0u;
                    for (int i = 0; i < 32; ++i) {
                        const int ri = row0 + i;
                        if (ri < n) {
                            m0 |= iou_gt(sb[ri], sa[ri], b0, a0, thr) ? (1u << i) : 0u;
                            m1 |= iou_gt(sb[ri], sa[ri], b1, a1, thr) ? (1u << i) : 0u;
                        }
                    }
                }
                if (r == t) {   // the tile's own block: strictly earlier candidates only
                    const uint32_t lt = (1u << lane) - 1u;
                    m0 = h ? 0u : (m0 & lt);
                    m1 = h ? (m1 & lt) : m1;
                }
                if (c0 < n) colm[c0 * 8 + 2 * r + h] = m0;
                if (c1 < n) colm[c1 * 8 + 2 * r + h] = m1;
            }
        }
        __syncthreads();
        PROF_TL(100, 4);
        if (warp == 0) {
            uint32_t keptw[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) keptw[w] = 0u;
            int nkept = 0;
#pragma unroll
            for (int t = 0; t < kSmallN / kTile; ++t) {
                if (t * kTile < n && nkept < max_out) {
                    const int c0 = t * kTile + lane, c1 = c0 + 32;
                    const uint32_t* m0 = colm + c0 * 8;
                    const uint32_t* m1 = colm + c1 * 8;
                    uint32_t rem0 = 0u, rem1 = 0u;
#pragma unroll
                    for (int w = 0; w < 2 * t; ++w) {   // removed by a box kept in an earlier tile
                        if (c0 < n) rem0 |= m0[w] & keptw[w];
                        if (c1 < n) rem1 |= m1[w] & keptw[w];
                    }
                    const uint32_t blk0 = (c0 < n) ? m0[2 * t] : 0u;
                    const uint32_t blk1l = (c1 < n) ? m1[2 * t] : 0u, blk1h = (c1 < n) ? m1[2 * t + 1] : 0u;
                    uint32_t und_lo = __ballot_sync(0xffffffffu, c0 < n && !rem0), und_hi = __ballot_sync(0xffffffffu, c1 < n && !rem1);
                    uint32_t kept_lo = 0u, kept_hi = 0u;
                    while (und_lo | und_hi) {   // fixed point, as the resolver of the pipeline
                        const bool u0 = (und_lo >> lane) & 1u, u1 = (und_hi >> lane) & 1u;
                        const bool d0 = u0 && (blk0 & kept_lo), d1 = u1 && ((blk1l & kept_lo) | (blk1h & kept_hi));
                        const bool k0 = u0 && !d0 && !(blk0 & und_lo), k1 = u1 && !d1 && !((blk1l & und_lo) | (blk1h & und_hi));
                        const uint32_t nk_lo = __ballot_sync(0xffffffffu, k0), nk_hi = __ballot_sync(0xffffffffu, k1);
                        const uint32_t nd_lo = __ballot_sync(0xffffffffu, d0), nd_hi = __ballot_sync(0xffffffffu, d1);
                        kept_lo |= nk_lo; kept_hi |= nk_hi;
                        und_lo &= ~(nk_lo | nd_lo); und_hi &= ~(nk_hi | nd_hi);
                    }
                    const uint32_t lt = (1u << lane) - 1u;
                    const int room = max_out - nkept;
                    if (__popc(kept_lo) + __popc(kept_hi) > room) {   // the first `room` of them
                        const bool f0 = ((kept_lo >> lane) & 1u) && __popc(kept_lo & lt) < room;
                        const bool f1 = ((kept_hi >> lane) & 1u) && __popc(kept_lo) + __popc(kept_hi & lt) < room;
                        kept_lo = __ballot_sync(0xffffffffu, f0);
                        kept_hi = __ballot_sync(0xffffffffu, f1);
                    }
                    if ((kept_lo >> lane) & 1u) sel[nkept + __popc(kept_lo & lt)] = c0;
                    if ((kept_hi >> lane) & 1u) sel[nkept + __popc(kept_lo) + __popc(kept_hi & lt)] = c1;
                    nkept += __popc(kept_lo) + __popc(kept_hi);
                    keptw[2 * t] = kept_lo;
                    keptw[2 * t + 1] = kept_hi;
                }
            }
            if (lane == 0) s_final[1] = nkept;
        }
        PROF_TL(100, 5);
        __syncthreads();
        PROF_TL(100, 6);
        nms_write_outputs(epi, bx, b, M, max_out, s_final[1], sel, s_orig, tid, kNmsThreads);
        PROF_TL(100, 7);
        return;
    }
    const int tiles = (n + kTile - 1) / kTile;
    const int tiles_max = (M + kTile - 1) / kTile;
    const unsigned long long* rows_b = grows ? in.rows + (size_t)b * tiles_max * kRows : nullptr;
    if (grows && tid == 0) {
        for (int v = 1; v < kRing && v < tiles; ++v) {
            mbar_arm(rbar_base + 8u * v, (uint32_t)kRows * 8u);
            bulk_copy_from_global(smem_u32(&s_rows[v][0]), rows_b + (size_t)v * kRows, (uint32_t)kRows * 8u, rbar_base + 8u * v);
        }
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
            if (lane == 0) s_rows[0][i] = row;
        }
    }
    cluster.sync();  // every CTA of the cluster is resident and its mbarriers are initialised before any st.async
    PROF_TL(100, 4);
    // Roles.  Warp 0 resolves; it keeps its scheduler (warps 4, 8, ... share it) to itself, because its per-tile chain
    // is the serial part of the kernel.  The other 24 warps are numbered 0..23 in order: the first nfar are far warps,
    // the next nrow row warps; the rest only helped with the staging and exit (exited threads count as arrived at the
    // barriers below).
    // roles == 1: the workers are warps 1, 2, 3, ... in order, i.e. two far and two row warps share the resolver's
    // scheduler (the resolver waits for its workers about half of the time: measured).
    const int ri = roles ? warp - 1 : ((warp & 3) ? (warp >> 2) * 3 + (warp & 3) - 1 : -1);
    if (warp != 0 && (ri < 0 || ri >= nfar + nrow)) return;
    const int eidx = (warp == 0) ? lane : (1 + ri) * 32 + lane;  // dense index of the surviving threads (epilogue)
    int nkept = 0, t = 0;
    PROF_DECL;
    if (warp == 0) {
        // ================= resolver warp: tile by tile =================
        uint64_t kept_hist[kDepth - 1];  // kept masks of tiles t-1 .. t-(kDepth-1)
#pragma unroll
        for (int d = 0; d < kDepth - 1; ++d) kept_hist[d] = 0ull;
        for (; t < tiles && nkept < max_out; ++t) {
            PROF_TILE;
            PROF_TL(t, 0);
            const int base = t * kTile;
            const int slot = t & (kRing - 1);
            const uint32_t bar_t = bar_base + 8u * (uint32_t)slot;
            // this lane's two candidates, for the compact kept arrays (loaded while the far set is still in flight)
            const int c0 = base + lane, c1 = c0 + 32;
            float4 mb0 = kNone, mb1 = kNone;
            float mt0 = thr, mt1 = thr;
            if (COMPACT) {
                if (c0 < n) { mb0 = sb[c0]; mt0 = -__fmul_rn(thr, sa[c0]); }
                if (c1 < n) { mb1 = sb[c1]; mt1 = -__fmul_rn(thr, sa[c1]); }
            }
            uint64_t removed = 0;
            if (t >= 1) {  // slot j serves tiles j, j + 8, ...; tile 0 has no far / cross set
                const uint32_t parity = (uint32_t)(((t >> 3) - (slot == 0 ? 1 : 0)) & 1);
                if (grows) mbar_wait(rbar_base + 8u * (uint32_t)slot, parity);   // fetched eight tiles ago: long complete
                mbar_wait(bar_t, parity);
                PROF_MARK(0);
                PROF_TL(t, 1);
                uint64_t v = 0ull;
#pragma unroll
                for (int q = 0; q < kMaxFarSrc / 32; ++q)  // far(t): fixed trip count, predicated loads (a generic
                    if (lane + 32 * q < nsrc) v |= (uint64_t)s_far[slot][lane + 32 * q];  // loop costs ~60 instructions)
#pragma unroll
                for (int d = 1; d < kDepth; ++d) {  // near(t) = cross_d rows of the boxes kept in tile t-d
                    if ((kept_hist[d - 1] >> lane) & 1ull) v |= (uint64_t)s_rows[slot][d * kTile + lane];
                    if ((kept_hist[d - 1] >> (lane + 32)) & 1ull) v |= (uint64_t)s_rows[slot][d * kTile + lane + 32];
                }
                removed = (uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)v) |
                          ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(v >> 32)) << 32);
            }
            const int rem = n - base;
            const uint64_t validbits = (rem >= kTile) ? ~0ull : ((1ull << rem) - 1ull);
            uint64_t und = ~removed & validbits, kept = 0;
            const uint64_t blk0r = (uint64_t)s_rows[slot][lane] & ((1ull << lane) - 1ull);  // earlier overlapping candidates
            const uint64_t blk1 = (uint64_t)s_rows[slot][lane + 32] & ((1ull << (lane + 32)) - 1ull);
            while (und) {
                const bool u0 = (und >> lane) & 1ull, u1 = (und >> (lane + 32)) & 1ull;
                const bool d0 = u0 && (blk0r & kept), d1 = u1 && (blk1 & kept);              // removed
                const bool k0 = u0 && !d0 && !(blk0r & und), k1 = u1 && !d1 && !(blk1 & und);  // kept
                const uint64_t nk = ballot64(k0, k1), nd = ballot64(d0, d1);
                kept |= nk;
                und &= ~(nk | nd);
            }
            const int room = max_out - nkept;
            while (__popcll(kept) > room) kept &= ~(1ull << (63 - __clzll(kept)));
            PROF_TL(t, 2);
            if ((kept >> lane) & 1ull) {
                const int pos = nkept + __popcll(kept & ((1ull << lane) - 1ull));
                sel[pos] = c0;
                if (COMPACT) {
                    const uint32_t q = ((uint32_t)pos * nsrc_magic) >> 20;
                    const uint32_t lw = (uint32_t)pos - q * (uint32_t)nsrc - (uint32_t)(crank * nfar);
                    if (lw < (uint32_t)nfar) { kb[lw * wcap + q] = mb0; kt[lw * wcap + q] = mt0; }
                }
            }
            if ((kept >> (lane + 32)) & 1ull) {
                const int pos = nkept + __popcll(kept & ((1ull << (lane + 32)) - 1ull));
                sel[pos] = c1;
                if (COMPACT) {
                    const uint32_t q = ((uint32_t)pos * nsrc_magic) >> 20;
                    const uint32_t lw = (uint32_t)pos - q * (uint32_t)nsrc - (uint32_t)(crank * nfar);
                    if (lw < (uint32_t)nfar) { kb[lw * wcap + q] = mb1; kt[lw * wcap + q] = mt1; }
                }
            }
            nkept += __popcll(kept);
#pragma unroll
            for (int d = kDepth - 2; d > 0; --d) kept_hist[d] = kept_hist[d - 1];
            kept_hist[0] = kept;
            __syncwarp();
            if (lane == 0) {
                s_nk[slot] = nkept;
                mbar_arm(bar_t, far_bytes);  // phase of tile t + kRing
                if (grows && t + kRing < tiles) {   // this tile's rows have been consumed: fetch those of tile t + kRing
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arm(rbar_base + 8u * (uint32_t)slot, (uint32_t)kRows * 8u);
                    bulk_copy_from_global(smem_u32(&s_rows[slot][0]), rows_b + (size_t)(t + kRing) * kRows,
                                          (uint32_t)kRows * 8u, rbar_base + 8u * (uint32_t)slot);
                }
            }
            __threadfence_block();
            // release the workers for tile t+kDepth (they need the kept list through tile t)
            if (t + kDepth < tiles) asm volatile("bar.arrive %0, %1;" ::"r"(3 + (t % kDepth)), "r"(rel_threads) : "memory");
            PROF_MARK(1);
            PROF_TL(t, 3);
        }
    } else if (ri < nfar) {
        // ================= far warps: far(u) share, u = 1 .. tiles-1 =================
        // Kept box k of the list belongs to far warp (k mod nsrc) of the cluster.  Every far warp sends its own 64-bit
        // partial straight to every CTA (no reduction inside the CTA on the critical path).  Per tile the share is
        // tested in two instalments: the boxes this warp already knew of after the previous release (kept in tiles
        // <= u-3) BEFORE waiting for the release of tile u, so that only the boxes kept in tile u-2 -- at most 64 for
        // the whole cluster, one or two per far warp -- are left between the release and the send.
        const int fw = ri;  // 0..nfar-1
        const int k0 = crank * nfar + fw;
        const float inv_src = __frcp_rn((float)nsrc), cthr = __fadd_rn(1.0f, thr);
        auto share = [&](int nk) {  // ceil((nk - k0) / nsrc), >= 0, without the integer-division sequence (operands < 2^14)
            int cnt = 0;
            if (nk > k0) {
                const int x = nk - k0 + nsrc - 1;
                cnt = (int)__fmul_rn((float)x, inv_src);
                cnt += ((cnt + 1) * nsrc <= x) - (cnt * nsrc > x);
            }
            return cnt;
        };
        PROFW_DECL;
        int nk_known = 0;  // kept count after tile u-3 (read at the previous release)
        float tkmax = 0.0f;  // largest thr * area over this warp's share so far (the list only grows)
        const float4* kbw = kb + fw * wcap;
        const float* ktw = kt + fw * wcap;
        const unsigned long long c1_2 = pack_f2(cthr, cthr);
        for (int u = 1; u < tiles; ++u) {
            const int ubase = u * kTile;
            const int c0 = ubase + lane, c1 = c0 + 32;
            const float4 b0 = (c0 < n) ? sb[c0] : kNone, b1 = (c1 < n) ? sb[c1] : kNone;
            const float a0 = (c0 < n) ? sa[c0] : 1.0f, a1 = (c1 < n) ? sa[c1] : 1.0f;
            const float tc0 = __fmul_rn(thr, a0), tc1 = __fmul_rn(thr, a1);
            CandPair cp;
            cp.y1a = b0.x; cp.x1a = b0.y; cp.y2a = b0.z; cp.x2a = b0.w;
            cp.y1b = b1.x; cp.x1b = b1.y; cp.y2b = b1.z; cp.x2b = b1.w;
            cp.ntac = pack_f2(-tc0, -tc1);
            const uint32_t bar_u = bar_base + 8u * (uint32_t)(u & (kRing - 1));
            const uint32_t my_far = smem_u32(&s_far[u & (kRing - 1)][k0]);
            float dmax0 = -3.0e38f, dmax1 = -3.0e38f;
            auto screen = [&](int j_begin, int j_end, int nk_lim, bool fresh) {  // this warp's boxes j_begin <= j < j_end
                for (int j = j_begin; j < j_end; j += 4) {  // four kept boxes per round: independent IoU chains; a round
                    float d0[4], d1[4];                      // that runs past the list repeats its last box (harmless)
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        if (COMPACT) {
                            const int jj = min(j + qq, j_end - 1);
                            const float4 bk = kbw[jj];
                            const float ntk = ktw[jj];
                            if (fresh) tkmax = fmaxf(tkmax, -ntk);
                            iou_screen_d2(bk, ntk, cp, c1_2, d0[qq], d1[qq]);
                        } else {
                            const int ki = sel[min(k0 + (j + qq) * nsrc, nk_lim - 1)];
                            const float4 bk = sb[ki];
                            const float tk = __fmul_rn(thr, sa[ki]);
                            tkmax = fmaxf(tkmax, tk);
                            d0[qq] = iou_screen_d(bk, tk, b0, tc0, cthr);
                            d1[qq] = iou_screen_d(bk, tk, b1, tc1, cthr);
                        }
                    }
                    dmax0 = fmax3(fmax3(dmax0, d0[0], d0[1]), d0[2], d0[3]);
                    dmax1 = fmax3(fmax3(dmax1, d1[0], d1[1]), d1[2], d1[3]);
                }
            };
            PROFW_MARK(0);
            if (fw == 0) PROF_TL(u, 7);
            const int cnt_known = share(nk_known);
            screen(0, cnt_known, nk_known, false);   // first instalment: no dependence on the release
            PROFW_MARK(4);
            int nk = 0;  // boxes kept in tiles <= u-kDepth
            if (u >= kDepth) {
                asm volatile("bar.sync %0, %1;" ::"r"(3 + (u % kDepth)), "r"(rel_threads) : "memory");
                nk = s_nk[(u - kDepth) & (kRing - 1)];
                if (nk >= max_out) {
                    // the resolver stops before tile u-kDepth+1.  The row warps have already sent rows(u): complete the tile's
                    // transaction set with empty partials so that the final drain can wait for it
                    if (lane < csize) st_async_u64(mapa_u32(my_far, (uint32_t)lane), 0ull, mapa_u32(bar_u, (uint32_t)lane));
                    break;
                }
            }
            PROFW_MARK(1);
            if (fw == 0) PROF_TL(u, 4);
            const int cnt = share(nk);
            if (COMPACT) {   // second instalment: the boxes kept in tile u-2 -- one or two per far warp, taken one by one
                for (int j = cnt_known; j < cnt; ++j) {
                    const float4 bk = kbw[j];
                    const float ntk = ktw[j];
                    tkmax = fmaxf(tkmax, -ntk);
                    float e0, e1;
                    iou_screen_d2(bk, ntk, cp, c1_2, e0, e1);
                    dmax0 = fmaxf(dmax0, e0);
                    dmax1 = fmaxf(dmax1, e1);
                }
            } else {
                screen(cnt_known, cnt, nk, true);
            }
            nk_known = nk;
            const float m0 = __fmul_rn(__fadd_rn(tkmax, tc0), kScreenBand), m1 = __fmul_rn(__fadd_rn(tkmax, tc1), kScreenBand);
            bool r0 = dmax0 > m0, r1 = dmax1 > m1;
            const bool unsure = (!r0 && dmax0 >= -m0) || (!r1 && dmax1 >= -m1);
            if (__any_sync(0xffffffffu, unsure)) {  // a pair within 2^-20 of the threshold: exact division
                PROFW_COUNT;
                PROF_FALLBACK;
                r0 = false; r1 = false;
                for (int k = k0; k < nk; k += nsrc) {
                    const int ki = sel[k];
                    r0 |= iou_gt(sb[ki], sa[ki], b0, a0, thr);
                    r1 |= iou_gt(sb[ki], sa[ki], b1, a1, thr);
                }
            }
            const uint64_t hit = ballot64(r0, r1);
            PROFW_MARK(2);
            if (lane < csize) st_async_u64(mapa_u32(my_far, (uint32_t)lane), hit, mapa_u32(bar_u, (uint32_t)lane));
            PROFW_MARK(3);
            if (fw == 0) PROF_TL(u, 5);
        }
        PROFW_DUMP;
    } else {
        // ================= row warps: diag(v) + cross(v), one tile ahead of the far warps =================
        // Rows do not depend on the kept list; the release barrier only provides flow control: rows(u+1) goes into ring
        // slot (u+1) % kRing of every CTA and is sent once resolve(u-kDepth) has released tile u (ring depth: nms_dev.cuh).
        const int rw = ri - nfar;  // 0..nrow-1
        const float cthr = __fadd_rn(1.0f, thr);
        const unsigned long long c1_2 = pack_f2(cthr, cthr);
        auto send_rows = [&](int v) {
            const int vbase = v * kTile;
            const int c0 = vbase + lane, c1 = c0 + 32;
            const float4 b0 = (c0 < n) ? sb[c0] : kNone, b1 = (c1 < n) ? sb[c1] : kNone;
            const float a0 = (c0 < n) ? sa[c0] : 1.0f, a1 = (c1 < n) ? sa[c1] : 1.0f;
            const uint32_t bar_v = bar_base + 8u * (uint32_t)(v & (kRing - 1));
            const float tc0 = __fmul_rn(thr, a0), tc1 = __fmul_rn(thr, a1);
            CandPair cp;
            cp.y1a = b0.x; cp.x1a = b0.y; cp.y2a = b0.z; cp.x2a = b0.w;
            cp.y1b = b1.x; cp.x1b = b1.y; cp.y2b = b1.z; cp.x2b = b1.w;
            cp.ntac = pack_f2(-tc0, -tc1);
            // this CTA owns the contiguous rows [crank * rpc, (crank + 1) * rpc) of the tile (0..63 diag, 64 d + i:
            // candidate i of tile v-d against tile v); they are staged locally and travel as ONE bulk copy per peer:
            // the receivers' mbarriers see csize transactions per tile instead of one per row
            const int rpc = kRows / csize;
            unsigned long long* stage = &s_stage[v & (kRing - 1)][0];
            for (int j = rw; j < rpc; j += nrow) {  // (four rows per pass with one vote was measured: slower, 135 -> 147 us)
                const int r = crank * rpc + j;
                const int i = r & (kTile - 1);
                const int ci = vbase - (r / kTile) * kTile + i;  // negative: that earlier tile does not exist
                const float4 bi = (ci >= 0 && ci < n) ? sb[ci] : kNone;
                const float ai = (ci >= 0 && ci < n) ? sa[ci] : 1.0f;
                const float ti = __fmul_rn(thr, ai);  // screen first; the exact division only inside the 2^-20 band
                float e0, e1;
                iou_screen_d2(bi, -ti, cp, c1_2, e0, e1);
                const float m0 = __fmul_rn(__fadd_rn(ti, tc0), kScreenBand), m1 = __fmul_rn(__fadd_rn(ti, tc1), kScreenBand);
                bool h0 = e0 > m0, h1 = e1 > m1;
                if (__any_sync(0xffffffffu, fabsf(e0) <= m0 || fabsf(e1) <= m1)) {
                    PROF_FALLBACK;
                    h0 = iou_gt(bi, ai, b0, a0, thr);
                    h1 = iou_gt(bi, ai, b1, a1, thr);
                }
                uint64_t row = ballot64(h0, h1);
                if (r < kTile) row &= ~(1ull << i);
                if (lane == 0) stage[j] = row;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy reads
            named_barrier(1, nrow * 32);
            if (rw == 0 && lane < csize)
                bulk_copy_to_peer(mapa_u32(smem_u32(&s_rows[v & (kRing - 1)][crank * rpc]), (uint32_t)lane), smem_u32(stage),
                                  (uint32_t)rpc * 8u, mapa_u32(bar_v, (uint32_t)lane));
            if (rw == 0) PROF_TL(v, 6);
        };
        if (tiles > 1) send_rows(1);
        for (int u = 1; u < tiles; ++u) {
            if (u >= kDepth) {
                asm volatile("bar.sync %0, %1;" ::"r"(3 + (u % kDepth)), "r"(rel_threads) : "memory");
                if (s_nk[(u - kDepth) & (kRing - 1)] >= max_out) break;  // same test as the far warps: both leave at the same tile
            }
            if (u + 1 < tiles) send_rows(u + 1);
        }
    }
    if (tid == 0) { s_final[0] = t; s_final[1] = nkept; }  // the resolver's loop state, for everybody
    PROF_TL(100, 5);
    __syncthreads();
    PROF_TL(100, 6);
    t = s_final[0];
    nkept = s_final[1];
    PROF_DUMP;
    // The sets of tiles t .. t+kDepth-1 (far or empty far partials + rows) were sent but never consumed: drain
    // them, so that no st.async is in flight towards this CTA when it exits; the cluster barrier then keeps every CTA
    // alive until its peers have drained
    if (warp == 0) {
        for (int d = t; d < t + kDepth; ++d)
            if (d >= 1 && d < tiles)
                mbar_wait(bar_base + 8u * (uint32_t)(d & (kRing - 1)),
                          (uint32_t)(((d >> 3) - ((d & (kRing - 1)) == 0 ? 1 : 0)) & 1));
        if (grows)   // row copies issued for tiles t .. t + kRing - 1 (resolve(d - kRing) fetched tile d) and never consumed
            for (int d = max(t, 1); d < t + kRing && d < tiles; ++d)
                mbar_wait(rbar_base + 8u * (uint32_t)(d & (kRing - 1)),
                          (uint32_t)(((d >> 3) - ((d & (kRing - 1)) == 0 ? 1 : 0)) & 1));
    }
    cluster.sync();
    if (crank != 0) return;
    const int total = nkept;
    nms_write_outputs(epi, bx, b, M, max_out, total, sel, FUSED ? s_orig : nullptr, eidx, rel_threads);
    PROF_TL(100, 7);
}

template <bool COMPACT, bool FUSED>
static const void* nms_kernel_ptr() { return (const void*)nms_lazy_kernel<COMPACT, FUSED>; }

// Launch.  in.scores / in.keys != NULL selects the fused-ordering kernel (the caller has checked that it applies:
// nms_fused_applies).  Cluster size: spread one image over as many SMs as the batch leaves free (1 CTA per SM), up to
// the portable maximum of 8 (16-CTA clusters do not co-schedule for 8 images on this part: measured); small candidate
// sets do not amortise the exchange.
static bool nms_compact_fits(int M, int max_out, bool fused, size_t* smem_out) {
    const void* k = fused ? nms_kernel_ptr<true, true>() : nms_kernel_ptr<true, false>();
    const size_t max_dyn = (size_t)device_props().smem_optin - static_smem_bytes(k) - 256;
    const size_t c = fused ? nms_fused_smem_bytes(M, max_out, true) : nms_smem_bytes(M, max_out, true);
    const bool compact = c <= max_dyn;
    *smem_out = compact ? c : (fused ? nms_fused_smem_bytes(M, max_out, false) : nms_smem_bytes(M, max_out, false));
    return compact;
}

bool nms_fused_applies(int M, int max_out) {
    if (M > 2048) return false;  // cluster size 1 only: every CTA of a cluster would repeat the ordering
    size_t smem;
    (void)nms_compact_fits(M, max_out, true, &smem);
    const size_t max_dyn = (size_t)device_props().smem_optin - static_smem_bytes(nms_kernel_ptr<false, true>()) - 256;
    return smem <= max_dyn;
}

int launch_nms(const NmsInput& in, int B, int M, int max_out, float thr, const NmsEpilogue& epi, cudaStream_t stream) {
    const bool fused = in.scores != nullptr || in.keys != nullptr;
    if (!fused && M > 2048 && in.rows == nullptr && tuning_knob("MRCNN_NMS_SWEEP", 1)) {   // cluster problems: nms_sweep.cu
        const int rc = launch_nms_sweep(in.boxes, in.valid, B, M, max_out, thr, epi, stream, in.unit);
        if (rc != -1) return rc;
    }
    size_t smem;
    const bool compact = nms_compact_fits(M, max_out, fused, &smem);
    const void* kernel = fused ? (compact ? nms_kernel_ptr<true, true>() : nms_kernel_ptr<false, true>())
                               : (compact ? nms_kernel_ptr<true, false>() : nms_kernel_ptr<false, false>());
    int cs = 1;
    if (!fused && M > 2048)
        cs = pick_cluster_size(kernel, kNmsThreads, B, tuning_knob("MRCNN_NMS_MAX_CLUSTER", 8), [smem](int) { return smem; });
    {   // per launch: the occupancy cache above may have set another problem's (smaller) limit last
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kNmsThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1, stream);
    // warp roles: 1 resolver, nfar far warps (the kept list x tile tests), nrow row warps (diag + cross blocks: 128 rows
    // per tile dealt over the cluster); the remaining warps only help with the staging and exit.  Few, fat workers:
    // the per-tile work of a warp is a latency-bound chain (~5 cycles per instruction), so fixed overhead per warp
    // costs more than it buys.  Measured at config 2 (ProposalLayer, B=8 / B=16): (8,4) 134.5 / 200 us, (6,8) 129.2 /
    // 194, (6,12) 130.4 / 192, (8,12) 132.4 / 198, (10,14) 136.6 / 202; the single-CTA detection NMS does not react.
    int nfar = cs >= 16 ? 6 : (cs >= 4 ? 8 : 12), nrow = cs >= 4 ? 8 : (cs == 2 ? 8 : 12), roles = 1;
    {   // measurement knobs (no cached state): MRCNN_NMS_NFAR / _NROW / _ROLES
        const int kf = tuning_knob("MRCNN_NMS_NFAR", 0), kr = tuning_knob("MRCNN_NMS_NROW", 0);
        roles = tuning_knob("MRCNN_NMS_ROLES", 1) ? 1 : 0;
        const int max_workers = roles ? 31 : 24;
        if (kf > 0 && kr > 0 && kf + kr <= max_workers && kf * cs <= kMaxFarSrc && kf <= 12) { nfar = kf; nrow = kr; }
    }
    void* args[] = {(void*)&in, (void*)&M, (void*)&max_out, (void*)&thr, (void*)&nfar, (void*)&nrow, (void*)&roles,
                    (void*)&epi};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, args);
    if (e != cudaSuccess) return (int)e;
    return last_error();
}

size_t nms_rows_ws_bytes(int B, int M) {
    return align_up((size_t)B * ((M + kTile - 1) / kTile) * kRows * sizeof(unsigned long long), 256);
}

// rows_ws: nms_rows_ws_bytes(B, M) of scratch for the precomputed rows (cluster problems, M > 2048), or NULL
int launch_nms_sorted(const float4* boxes_sorted, const int32_t* valid, int B, int M, int max_out, float thr,
                      const NmsEpilogue& epi, void* rows_ws, cudaStream_t stream, bool unit_boxes) {
    NmsInput in{};
    in.boxes = boxes_sorted;
    in.valid = valid;
    in.unit = unit_boxes;
    if (rows_ws != nullptr && M > 2048 && tuning_knob("MRCNN_NMS_GLOBAL_ROWS", 0)) {
        const int tiles_max = (M + kTile - 1) / kTile;
        cudaError_t e = launch_pdl(nms_rows_kernel, dim3(tiles_max, B), dim3(128), 0, stream, boxes_sorted, valid, M,
                                   tiles_max, thr, (unsigned long long*)rows_ws);
        if (e != cudaSuccess) return (int)e;
        in.rows = (const unsigned long long*)rows_ws;
    }
    return launch_nms(in, B, M, max_out, thr, epi, stream);
}

int launch_nms_unsorted(const float4* boxes, const float* scores, const uint32_t* keys, const int32_t* valid, int B, int M,
                        int max_out, float thr, const NmsEpilogue& epi, cudaStream_t stream) {
    NmsInput in{};
    in.boxes = boxes;
    in.valid = valid;
    in.scores = scores;
    in.keys = keys;
    return launch_nms(in, B, M, max_out, thr, epi, stream);
}

// generic entry: sort candidates (score > -inf) by (score desc, index asc); one CTA per image
__global__ void __launch_bounds__(1024)
nms_sort_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ valid,
                int M, float4* __restrict__ boxes_sorted, int32_t* __restrict__ orig_idx, int32_t* __restrict__ ncand) {
    extern __shared__ __align__(16) uint64_t s[];
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
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
           align_up((size_t)B * sizeof(int32_t), 256) + nms_rows_ws_bytes(B, M);
}

}  // namespace mrcnn

using namespace mrcnn;

#ifdef MRCNN_NMS_PROFILE
MRCNN_EXPORT int mrcnn_debug_nms_profile(long long* host_out8) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out8, g_nms_prof, sizeof(long long) * 8);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(host_out8 + 8, g_nms_profw, sizeof(long long) * 8);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(host_out8 + 16, g_nms_tl, sizeof(long long) * 16);
    return (int)e;
}
MRCNN_EXPORT int mrcnn_debug_nms_timeline(long long* host_out_128x8) {
    return (int)cudaMemcpyFromSymbol(host_out_128x8, g_nms_timeline, sizeof(long long) * 128 * 8);
}
// number of exact-division fallbacks of the threshold screen since the last call (far and row warps, all CTAs)
MRCNN_EXPORT int mrcnn_debug_nms_fallbacks(unsigned long long* host_out, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out, g_nms_fallbacks, sizeof(unsigned long long));
    const unsigned long long zero = 0;
    if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(g_nms_fallbacks, &zero, sizeof(zero));
    return (int)e;
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
    NmsEpilogue epi{};
    epi.mode = 0;
    epi.keep = keep;
    epi.count = count;
    // single-CTA problems: the NMS kernel orders its candidates itself (one launch, no workspace traffic)
    if (nms_fused_applies(M, max_out))
        return launch_nms_unsorted((const float4*)boxes, scores, nullptr, valid, B, M, max_out, thr, epi, st);
    NmsWs w;
    char* p = (char*)ws;
    w.boxes_sorted = (float4*)p; p += align_up((size_t)B * M * sizeof(float4), 256);
    w.orig_idx = (int32_t*)p;    p += align_up((size_t)B * M * sizeof(int32_t), 256);
    w.ncand = (int32_t*)p;       p += align_up((size_t)B * sizeof(int32_t), 256);
    void* rows_ws = p;
    const int sort_n = next_pow2(M < 32 ? 32 : M);
    const size_t smem = (size_t)sort_n * sizeof(uint64_t) + (sort_n >= 1024 ? block_sort_xch_bytes(sort_n / 1024) : 0);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaError_t e = launch_pdl(nms_sort_kernel, dim3(B), dim3(1024), smem, st, (const float4*)boxes, scores, valid, M,
                               w.boxes_sorted, w.orig_idx, w.ncand);
    if (e != cudaSuccess) return (int)e;
    epi.orig_idx = w.orig_idx;
    return launch_nms_sorted(w.boxes_sorted, w.ncand, B, M, max_out, thr, epi, rows_ws, st);
}
