// nms_sweep.cu -- the cluster NMS for ORDERED candidate lists (ProposalLayer: tf.image.non_max_suppression at
// mrcnn_layers.py:225 behind the top-k of :246), TF NonMaxSuppressionV3 semantics, bit-exact keep lists.
//
// Round-2 redesign of the cluster path of nms.cu (nms_lazy_kernel stays for single-CTA / fused-ordering problems and as
// the fallback when this kernel's shared memory does not fit).  Same idea -- walk the candidates in 64-box tiles, test a
// tile only against what is KEPT, never build the M x M matrix -- but every per-tile fixed cost of the old kernel is
// gone.  A cluster of 1..16 CTAs per image (10 at batch 8 on a B200; sizes need not be powers of two), 20 warps per CTA,
// roles ordered by the priority the issue arbiter gives them (highest warp id first):
//   resolver (last warp, alone on its SM sub-partition; every CTA for itself, identically): per tile ONE mbarrier wait,
//     eight shared-memory loads, two warp-wide ORs, the ballot fixed point over the tile's own block, and the append to
//     the CTA's copy of the kept list.  ~900 cycles per tile; it also writes the outputs when the sweep is over.
//   far warps: tile u belongs to the far warps of CTA (u mod cluster size), which keep the tile's candidates in registers
//     for the tile's whole life.  Two groups: the TAIL group (3 warps, highest priority after the resolver) screens the
//     boxes kept in the last `cluster size` tiles in instalments -- after the releases of tiles u-8, u-6, u-4 and then
//     the boxes kept in tile u-3, the only work between a release and the tile that needs the result; the BULK group
//     (9 warps, lowest priority) screens everything kept before that and has cluster-size - 1 resolver periods for it.
//   row warps (3, between the two): what does not depend on the kept list -- the tile's own 64 x 64 block and the cross
//     blocks tile u-1 / u-2 (rows) x tile u (columns) -- as 32-row jobs dealt over the cluster.  A lane keeps the masks
//     of ITS two candidates (bit i = row i overlaps my candidate: a funnel shift of the sign of band - d, no ballots),
//     the warp stores them into this CTA's slot and sends the slot to every peer as one bulk copy; the resolver's lane
//     reads them back as its own columns, so "removed by a box kept in the previous tiles" is four ANDs, not a reduction.
// Every per-tile object (column masks, far partials, the data mbarrier, the release mbarrier, the kept count) has its
// OWN shared-memory slot -- 1.7 KB per tile, 94 tiles at M = 6000 -- so there is no ring, no phase reuse and no
// flow-control argument: every mbarrier completes exactly one phase.  After the sweep stops (max_out reached) the
// workers abandon what they are doing, complete the transaction counts of the unvisited tiles without moving data
// (mbarrier.complete_tx on the peers' barriers) and every CTA drains all of its tile barriers before the final cluster
// barrier, so nothing is ever in flight towards an exited CTA and no bulk copy reads from one.
// What bounds it (ncu): the alu pipes of the cluster's SMs (FMNMX of the IoU screens: 53 % busy
// on average, saturated late in the sweep) and the resolver's chain.
#include <cstdlib>

#include "nms_dev.cuh"

namespace cg = cooperative_groups;

namespace mrcnn {

constexpr int kSwThreads = 640;
constexpr int kSwWarps = kSwThreads / 32;
constexpr int kSwScratch = 768;    // per warp: 32 x (box 16 B, -thr * area 4 B, area 4 B)

#ifdef MRCNN_NMS_PROFILE
static __device__ long long g_sw_timeline[128 * 8];
#define SW_TL(tile, ev) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile) < 128) g_sw_timeline[(tile) * 8 + (ev)] = clock64(); } while (0)
static __device__ long long g_sw_rowtl[128 * 4];
#define SW_RTL(on, tile, ev) do { if ((on) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile) < 128) g_sw_rowtl[(tile) * 4 + (ev)] = clock64(); } while (0)
#else
#define SW_TL(tile, ev)
#define SW_RTL(on, tile, ev)
#endif

struct SwLayout {
    uint32_t cols, far, bar, rel, kb, kt, ka, sel, nk, scr, total;
};
// DEPTH: far(u) covers the boxes kept in tiles <= u - DEPTH; the DEPTH - 1 tiles in between are covered by cross blocks
// (tile u - d rows x tile u columns, independent of what is kept).  DEPTH = 3 gives the far warps two resolver periods
// between the release they wait for and the tile that needs their partial (measured: at DEPTH = 2 every third tile
// waited for it); DEPTH = 2 is the smaller-footprint instance.
__host__ __device__ inline SwLayout sw_layout(int tiles_max, int cap, int nfar, int depth) {
    SwLayout L;
    uint32_t o = 0;
    L.cols = o; o += (uint32_t)tiles_max * (uint32_t)depth * 512u;   // [tile][job 0 .. 2 depth - 1][lane] u64
    L.far = o;  o += (uint32_t)tiles_max * (uint32_t)nfar * 8u;      // [tile][far warp of the owning CTA] u64
    L.bar = o;  o += (uint32_t)tiles_max * 8u;               // data barrier of the tile: rows + far partials
    L.rel = o;  o += (uint32_t)tiles_max * 8u;               // release barrier: the tile is resolved
    o = (o + 15u) & ~15u;
    L.kb = o;   o += (uint32_t)cap * 16u;                    // kept boxes, normalised
    L.kt = o;   o += (uint32_t)cap * 4u;                     // -thr * area
    L.ka = o;   o += (uint32_t)cap * 4u;                     // area (exact-division path only)
    L.sel = o;  o += (uint32_t)cap * 4u;                     // candidate position
    L.nk = o;   o += (uint32_t)tiles_max * 4u;               // kept count after the tile
    o = (o + 15u) & ~15u;
    L.scr = o;  o += (uint32_t)kSwWarps * (uint32_t)kSwScratch;
    L.total = o;
    return L;
}

__device__ __forceinline__ void mbar_arrive_release(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct SwCands {   // the two candidates of a lane: c0 = 64 v + lane, c1 = c0 + 32
    CandPair cp;
    float4 b0, b1;
    float a0, a1, tc0, tc1;
};
__device__ __forceinline__ float4 sw_none() { return make_float4(1.0e18f, 1.0e18f, -1.0e18f, -1.0e18f); }
// TF: a box with area <= 0 has IoU 0 with everything -> replaced by a far-away box of area 1
__device__ __forceinline__ float4 sw_load_box(const float4* __restrict__ bx, int n, int c, float& a) {
    float4 t = sw_none();
    a = 1.0f;
    if (c >= 0 && c < n) {
        t = normalise_box(__ldg(bx + c), a);
        if (!(a > 0.0f)) { t = sw_none(); a = 1.0f; }
    }
    return t;
}
__device__ __forceinline__ void sw_load_cands(const float4* __restrict__ bx, int n, int v, int lane, float thr, SwCands& c) {
    c.b0 = sw_load_box(bx, n, v * kTile + lane, c.a0);
    c.b1 = sw_load_box(bx, n, v * kTile + lane + 32, c.a1);
    c.tc0 = __fmul_rn(thr, c.a0);
    c.tc1 = __fmul_rn(thr, c.a1);
    c.cp.y1a = c.b0.x; c.cp.x1a = c.b0.y; c.cp.y2a = c.b0.z; c.cp.x2a = c.b0.w;
    c.cp.y1b = c.b1.x; c.cp.x1b = c.b1.y; c.cp.y2b = c.b1.z; c.cp.x2b = c.b1.w;
    c.cp.ntac = pack_f2(-c.tc0, -c.tc1);
}

// One row job: the NR candidates from position `row0` on (rows; positions outside [0, n) overlap nothing) against the
// lane's two candidates of tile v.  m0 / m1: bit i = row i overlaps candidate c0 / c1.  ONLY1: c0's mask is not needed
// (rows 32..63 of a tile's own block are later than its candidates 0..31) -- the scalar screen on c1 alone.
// alu-pipe budget per row: 8 FMNMX (4 ONLY1) + 2 clamps (none for unit boxes) + one funnel shift per candidate (the sign
// bit of margin - d IS the hit) + one FMNMX3 for "some pair within the band".
// `stop_flag` (or NULL): polled every eight rows; once set the job is abandoned (returns false, masks undefined).
template <int NR, bool UNIT, bool ONLY1>
__device__ __forceinline__ bool sw_row_job(const float4* __restrict__ bx, int n, int v, int row0, int lane, float thr,
                                           float cthr, unsigned char* scr, uint32_t& m0, uint32_t& m1,
                                           const int* stop_flag = nullptr, bool prof = false) {
    SwCands c;
    sw_load_cands(bx, n, v, lane, thr, c);
    float ra;
    const float4 rb = sw_load_box(bx, n, (lane < NR) ? row0 + lane : -1, ra);
    float4* s_b = reinterpret_cast<float4*>(scr);
    float* s_nt = reinterpret_cast<float*>(scr + 512);
    float* s_a = reinterpret_cast<float*>(scr + 640);
    __syncwarp();   // the previous job's reads of the scratch are done
    s_b[lane] = rb;
    s_nt[lane] = -__fmul_rn(thr, ra);
    s_a[lane] = ra;
    __syncwarp();
    SW_RTL(prof, v, 1);
    const unsigned long long c1_2 = pack_f2(cthr, cthr);
    m0 = 0u; m1 = 0u;
    float amin = 1.0f;   // min over the pairs of |d| - margin: <= 0 = some pair within the band
    for (int i8 = NR - 8; i8 >= 0; i8 -= 8) {
    if (stop_flag != nullptr && *(volatile const int*)stop_flag) return false;
#pragma unroll
    for (int i = i8 + 7; i >= i8; --i) {   // downwards: every row shifts its bit in from the right
        const float4 bi = s_b[i];
        const float nti = s_nt[i];
        if (ONLY1) {
            float e1;
            if (UNIT) {
                const float dh = sub_sat(fminf(bi.z, c.b1.z), fmaxf(bi.x, c.b1.x));
                const float dw = __fsub_rn(fminf(bi.w, c.b1.w), fmaxf(bi.y, c.b1.y));
                e1 = __fmaf_rn(__fmul_rn(dh, dw), cthr, __fadd_rn(nti, -c.tc1));
            } else {
                e1 = iou_screen_d(bi, -nti, c.b1, c.tc1, cthr);
            }
            const float g1 = __fmul_rn(__fsub_rn(c.tc1, nti), kScreenBand);
            m1 = __funnelshift_l(__float_as_uint(__fsub_rn(g1, e1)), m1, 1);   // sign(g - e) = (e > g)
            amin = fminf(amin, __fsub_rn(fabsf(e1), g1));
        } else {
            float e0, e1;
            iou_screen_d2t<UNIT>(bi, nti, c.cp, c1_2, e0, e1);
            const float g0 = __fmul_rn(__fsub_rn(c.tc0, nti), kScreenBand), g1 = __fmul_rn(__fsub_rn(c.tc1, nti), kScreenBand);
            m0 = __funnelshift_l(__float_as_uint(__fsub_rn(g0, e0)), m0, 1);
            m1 = __funnelshift_l(__float_as_uint(__fsub_rn(g1, e1)), m1, 1);
            amin = fmin3(amin, __fsub_rn(fabsf(e0), g0), __fsub_rn(fabsf(e1), g1));
        }
    }
    }
    if (__any_sync(0xffffffffu, amin <= 0.0f)) {   // a pair within 2^-20 of the threshold: the whole job by exact division
        PROF_FALLBACK;
        m0 = 0u; m1 = 0u;
        for (int i = 0; i < NR; ++i) {
            const float4 bi = s_b[i];
            const float ai = s_a[i];
            m0 |= iou_gt(bi, ai, c.b0, c.a0, thr) ? (1u << i) : 0u;
            m1 |= iou_gt(bi, ai, c.b1, c.a1, thr) ? (1u << i) : 0u;
        }
    }
    return true;
}

template <int DEPTH, bool UNIT>
__global__ void __launch_bounds__(kSwThreads, 1)
nms_sweep_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ valid, int M, int max_out, float thr,
                 int nfar, int ntail, int nrow, int layout, int look, int sched, NmsEpilogue epi) {
    constexpr int kJobs = 2 * DEPTH;           // 32-row jobs per tile: the tile's own block + DEPTH - 1 cross blocks
    constexpr int kColWords = kJobs * 32;      // u64 words of column masks per tile
    extern __shared__ __align__(16) unsigned char sw_smem[];
    __shared__ int s_stop;       // the resolver has left its loop: whatever is still to be sent may be empty
    __shared__ float s_tk[32];   // per resolver lane: largest thr * area among the boxes it has appended to the kept list
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
    const int b = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_max = (M + kTile - 1) / kTile, cap = min(max_out, M);
    const SwLayout L = sw_layout(tiles_max, cap, nfar, DEPTH);
    unsigned long long* s_cols = reinterpret_cast<unsigned long long*>(sw_smem + L.cols);
    unsigned long long* s_far = reinterpret_cast<unsigned long long*>(sw_smem + L.far);
    float4* kb = reinterpret_cast<float4*>(sw_smem + L.kb);
    float* kt = reinterpret_cast<float*>(sw_smem + L.kt);
    float* ka = reinterpret_cast<float*>(sw_smem + L.ka);
    int32_t* sel = reinterpret_cast<int32_t*>(sw_smem + L.sel);
    int* s_nk = reinterpret_cast<int*>(sw_smem + L.nk);
    unsigned char* scr = sw_smem + L.scr + warp * kSwScratch;
    const uint32_t bar_base = smem_u32(sw_smem + L.bar), rel_base = smem_u32(sw_smem + L.rel);
    const uint32_t cols_base = smem_u32(s_cols), far_base = smem_u32(s_far);
    SW_TL(120, 0);
    pdl_launch_dependents();
    // every tile has its own barriers; each completes exactly one phase (parity 0).  Tiles 0 and 1 get their rows from
    // this CTA's own prologue (plain stores), tile 0 has no far set.
    for (int i = tid; i < tiles_max; i += kSwThreads) {
        mbar_init(bar_base + 8u * i, 1);
        mbar_init(rel_base + 8u * i, 1);
    }
    if (tid == 0) s_stop = 0;
    if (tid < 32) s_tk[tid] = 0.0f;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = tid; i < tiles_max; i += kSwThreads)
        mbar_arm(bar_base + 8u * i, (i >= 2 ? (uint32_t)kJobs * 256u : 0u) + 8u * (uint32_t)nfar);
    // every CTA's barriers are initialised once all threads of the cluster have ARRIVED here; the wait sits behind the
    // prologue jobs, which touch only this CTA's memory
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    pdl_wait();  // everything above overlaps the tail of the producing kernel; global memory is touched from here on
    SW_TL(120, 1);
    const int n = valid ? min(max(valid[b], 0), M) : M;
    const float4* bx = boxes + (size_t)b * M;
    const int tiles = (n + kTile - 1) / kTile;
    const float cthr = __fadd_rn(1.0f, thr);
    // prologue, every CTA for itself: the rows of tiles 0 and 1 as twelve 16-row jobs (warps 0..11): tile 0's own block
    // (4), tile 1's own block (4), tile 0 x tile 1 (4); tile 1's second cross block does not exist
    if (warp < 12 && tiles > 0) {
        const int pj = warp;
        const int v = pj < 4 ? 0 : 1, blk = pj < 8 ? 0 : 1, h = pj & 3;     // h: rows 16 h .. 16 h + 15 of the row tile
        uint32_t m0, m1;
        sw_row_job<16, UNIT, false>(bx, n, v, (v - blk) * kTile + 16 * h, lane, thr, cthr, scr, m0, m1);
        // column word (tile v, job 2 blk + (h >> 1), lane) = m0 | m1 << 32; this job owns bits 16 (h & 1) .. + 15 of both
        uint16_t* w = reinterpret_cast<uint16_t*>(s_cols + (size_t)v * kColWords + (2 * blk + (h >> 1)) * 32 + lane);
        w[h & 1] = (uint16_t)m0;
        w[2 + (h & 1)] = (uint16_t)m1;
    }
    if (DEPTH > 2 && tid < 64) s_cols[(size_t)1 * kColWords + 4 * 32 + tid] = 0ull;   // tile 1, cross block 2: empty
    SW_TL(120, 2);
    __syncthreads();   // the prologue rows are in place
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // every CTA is resident, its mbarriers initialised
    SW_TL(120, 3);

    // Roles by priority: the issue arbiter of an SM sub-partition prefers the HIGHEST warp id, so the resolver (the serial
    // chain) is the last warp, the far warps (whose last instalment sits between a release and a tile that needs it) come
    // next, and the row warps, which only have to stay ahead, are the lowest ids.
    // The resolver also has its sub-partition (warp id mod 4 == 3) to itself -- the alu pipe and the issue port are per
    // sub-partition, and a far group in full swing next to it stretched its tile from 900 to 1500 cycles (measured); the
    // other fifteen warps (three sub-partitions x five) are the workers.
    int role;   // 0 = resolver, 1 .. ntail = far (tail), then nrow row warps, then nfar - ntail far (bulk), beyond = idle
    if (warp == kSwWarps - 1) role = 0;
    else if (!layout) role = kSwWarps - 1 - warp;            // measurement knob: the resolver shares its sub-partition
    else if ((warp & 3) == 3) role = kSwWarps;
    else role = 1 + (14 - ((warp >> 2) * 3 + (warp & 3)));
    if (role == 0) {
        // ================= resolver: tile by tile =================
        int nkept = 0, t = 0;
        uint32_t kp_lo[DEPTH - 1], kp_hi[DEPTH - 1];   // kept masks of the previous DEPTH - 1 tiles
#pragma unroll
        for (int d = 0; d < DEPTH - 1; ++d) { kp_lo[d] = 0u; kp_hi[d] = 0u; }
        const uint32_t lt = (1u << lane) - 1u;
        float mytk = 0.0f;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 r0 = (lane < n) ? __ldg(bx + lane) : zero4, r1 = (lane + 32 < n) ? __ldg(bx + lane + 32) : zero4;
        for (; t < tiles && nkept < max_out; ++t) {
            SW_TL(t, 0);
            const int c0 = t * kTile + lane, c1 = c0 + 32;
            float a0, a1;   // this lane's two candidates, for the kept list (the next tile's are fetched meanwhile)
            float4 m0 = normalise_box(r0, a0), m1 = normalise_box(r1, a1);
            if (!(a0 > 0.0f)) { m0 = sw_none(); a0 = 1.0f; }
            if (!(a1 > 0.0f)) { m1 = sw_none(); a1 = 1.0f; }
            const float nt0 = -__fmul_rn(thr, a0), nt1 = -__fmul_rn(thr, a1);
            r0 = (c0 + kTile < n) ? __ldg(bx + c0 + kTile) : zero4;
            r1 = (c1 + kTile < n) ? __ldg(bx + c1 + kTile) : zero4;
            const unsigned long long* ct = s_cols + (size_t)t * kColWords;
            uint32_t rem0 = 0u, rem1 = 0u;
            if (t >= 1) {
                mbar_wait_parked(bar_base + 8u * (uint32_t)t, 0u);
                SW_TL(t, 1);
                const unsigned long long fv = (lane < nfar) ? s_far[(size_t)t * nfar + lane] : 0ull;
#pragma unroll
                for (int d = 1; d < DEPTH; ++d) {   // removed by a box kept in tile t - d: my column of that cross block
                    const unsigned long long qa = ct[(2 * d) * 32 + lane], qb = ct[(2 * d + 1) * 32 + lane];
                    rem0 |= ((uint32_t)qa & kp_lo[d - 1]) | ((uint32_t)qb & kp_hi[d - 1]);
                    rem1 |= ((uint32_t)(qa >> 32) & kp_lo[d - 1]) | ((uint32_t)(qb >> 32) & kp_hi[d - 1]);
                }
                const uint32_t far_lo = __reduce_or_sync(0xffffffffu, (unsigned)fv);
                const uint32_t far_hi = __reduce_or_sync(0xffffffffu, (unsigned)(fv >> 32));
                rem0 |= (far_lo >> lane) & 1u;
                rem1 |= (far_hi >> lane) & 1u;
            }
            const unsigned long long q0 = ct[lane], q1 = ct[32 + lane];
            // earlier candidates of the tile that overlap mine
            const uint32_t blk0 = (uint32_t)q0 & lt, blk1l = (uint32_t)(q0 >> 32), blk1h = (uint32_t)(q1 >> 32) & lt;
            uint32_t und_lo = __ballot_sync(0xffffffffu, c0 < n && !rem0), und_hi = __ballot_sync(0xffffffffu, c1 < n && !rem1);
            uint32_t kept_lo = 0u, kept_hi = 0u;
            while (und_lo | und_hi) {   // fixed point: kept once every earlier overlapping candidate is decided-removed,
                const bool u0 = (und_lo >> lane) & 1u, u1 = (und_hi >> lane) & 1u;   // removed once one is decided-kept
                const bool d0 = u0 && (blk0 & kept_lo), d1 = u1 && ((blk1l & kept_lo) | (blk1h & kept_hi));
                const bool k0 = u0 && !d0 && !(blk0 & und_lo), k1 = u1 && !d1 && !((blk1l & und_lo) | (blk1h & und_hi));
                const uint32_t nk_lo = __ballot_sync(0xffffffffu, k0), nk_hi = __ballot_sync(0xffffffffu, k1);
                const uint32_t nd_lo = __ballot_sync(0xffffffffu, d0), nd_hi = __ballot_sync(0xffffffffu, d1);
                kept_lo |= nk_lo; kept_hi |= nk_hi;
                und_lo &= ~(nk_lo | nd_lo); und_hi &= ~(nk_hi | nd_hi);
            }
            const int room = max_out - nkept;
            if (__popc(kept_lo) + __popc(kept_hi) > room) {   // last tile only: the first `room` of them
                const bool f0 = ((kept_lo >> lane) & 1u) && __popc(kept_lo & lt) < room;
                const bool f1 = ((kept_hi >> lane) & 1u) && __popc(kept_lo) + __popc(kept_hi & lt) < room;
                kept_lo = __ballot_sync(0xffffffffu, f0);
                kept_hi = __ballot_sync(0xffffffffu, f1);
            }
            SW_TL(t, 2);
            if ((kept_lo >> lane) & 1u) {
                const int pos = nkept + __popc(kept_lo & lt);
                kb[pos] = m0; kt[pos] = nt0; ka[pos] = a0; sel[pos] = c0;
                mytk = fmaxf(mytk, -nt0);
            }
            if ((kept_hi >> lane) & 1u) {
                const int pos = nkept + __popc(kept_lo) + __popc(kept_hi & lt);
                kb[pos] = m1; kt[pos] = nt1; ka[pos] = a1; sel[pos] = c1;
                mytk = fmaxf(mytk, -nt1);
            }
            s_tk[lane] = mytk;
            nkept += __popc(kept_lo) + __popc(kept_hi);
#pragma unroll
            for (int d = DEPTH - 2; d > 0; --d) { kp_lo[d] = kp_lo[d - 1]; kp_hi[d] = kp_hi[d - 1]; }
            kp_lo[0] = kept_lo; kp_hi[0] = kept_hi;
            __syncwarp();
            if (lane == 0) {
                s_nk[t] = nkept;
                mbar_arrive_release(rel_base + 8u * (uint32_t)t);
            }
            SW_TL(t, 3);
        }
        SW_TL(120, 4);
        if (lane == 0) *(volatile int*)&s_stop = 1;
        __syncwarp();
        // stopped before the last tile: wake whoever waits for a later release (they read nk >= max_out or the stop flag
        // and finish with empty sends) -- first, so that the workers complete the unvisited tiles while the outputs are written
        for (int u = t + lane; u < tiles; u += 32) {
            s_nk[u] = max(nkept, max_out);
            mbar_arrive_release(rel_base + 8u * (uint32_t)u);
        }
        // every CTA holds the whole kept list: the resolver warps of the cluster share the output rows among themselves
        // (row r goes to lane r mod 32 csize)
        nms_write_outputs(epi, bx, b, M, max_out, nkept, sel, nullptr, crank * 32 + lane, csize * 32);
        SW_TL(120, 6);
        // then drain every tile barrier of this CTA
        for (int u = max(t, 1) + lane; u < tiles; u += 32) mbar_wait_parked(bar_base + 8u * (uint32_t)u, 0u);
#ifdef MRCNN_NMS_PROFILE
        if (blockIdx.x == 0) g_sw_timeline[(64 + lane) * 8 + 7] = clock64();   // when this lane's barriers were complete
#endif
        __syncwarp();
        SW_TL(120, 5);
    } else if (role <= ntail || (role > ntail + nrow && role <= nfar + nrow)) {
        // ================= far warps: tiles u = crank (mod csize), u >= 1 =================
        // Two groups.  The BULK group (lowest priority of all) screens tile u against everything kept through tile u - X
        // (X = the cluster size: the CTA's previous tile is u - X, so the bulk of one tile has X - 1 resolver periods);
        // the TAIL group (highest priority after the resolver) takes the boxes kept in tiles u - X + 1 .. u - DEPTH, in two
        // instalments -- through tile u - DEPTH - 1, then the boxes kept in tile u - DEPTH, the only work between a
        // release and the tile that needs the result.  Every far warp of either group sends its own 64-bit partial.
        const bool tail = role <= ntail;
        const int fi = tail ? role - 1 : role - 1 - ntail - nrow;   // index inside the group
        const int fn = tail ? ntail : nfar - ntail;                 // warps of the group: box k belongs to warp k mod fn
        const int slot = tail ? fi : ntail + fi;                    // far partial slot of the tile
        const int X = max(csize, DEPTH + 1);
        const unsigned long long c1_2 = pack_f2(cthr, cthr);
        const float kcap = (thr > 0.0f) ? __fdiv_ru(1.05f, thr) : 3.0e38f;
        float tkmax = 0.0f;   // an upper bound of thr * area over the kept list as far as this warp has read it (s_tk)
        bool stop = false;
        for (int u = (crank == 0) ? csize : crank; u < tiles; u += csize) {
            const uint32_t my_far = far_base + 8u * (uint32_t)(u * nfar + slot), bar_u = bar_base + 8u * (uint32_t)u;
            uint64_t hit = 0ull;
            stop = stop || *(volatile int*)&s_stop;
            if (!stop) {
                SwCands c;
                sw_load_cands(bx, n, u, lane, thr, c);
                float dmax0 = -3.0e38f, dmax1 = -3.0e38f;
                const int tgt = u - DEPTH;        // far(u) covers the boxes kept in tiles <= u - DEPTH
                const int split = min(u - X, tgt);   // bulk: tiles <= split; tail: split < tile <= tgt
                int k = fi, k_first = fi;   // this warp's boxes: kept list entries k_first, k_first + fn, ... below nk
                int nk = 0;
                auto advance = [&](int tile) {   // screen the share of everything kept through `tile` (released or awaited)
                    mbar_wait_parked(rel_base + 8u * (uint32_t)tile, 0u);
                    nk = s_nk[tile];
                    if (nk >= max_out) { stop = true; return; }
                    tkmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(s_tk[lane])));   // values >= 0
                    for (; k + 3 * fn < nk; k += 4 * fn) {   // four kept boxes per round: independent chains
                        if (*(volatile int*)&s_stop) { stop = true; return; }   // the sweep is over: abandon
                        float d0[4], d1[4];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const float4 bk = kb[k + qq * fn];
                            const float ntk = kt[k + qq * fn];
                            iou_screen_d2t<UNIT>(bk, ntk, c.cp, c1_2, d0[qq], d1[qq]);
                        }
                        dmax0 = fmax3(fmax3(dmax0, d0[0], d0[1]), d0[2], d0[3]);
                        dmax1 = fmax3(fmax3(dmax1, d1[0], d1[1]), d1[2], d1[3]);
                    }
                    for (; k < nk; k += fn) {
                        const float4 bk = kb[k];
                        const float ntk = kt[k];
                        float e0, e1;
                        iou_screen_d2t<UNIT>(bk, ntk, c.cp, c1_2, e0, e1);
                        dmax0 = fmaxf(dmax0, e0);
                        dmax1 = fmaxf(dmax1, e1);
                    }
                };
                if (!tail) {
                    if (split >= 0) advance(split);
                } else if (tgt >= 0) {
                    if (split >= 0) {   // the tail starts where the bulk ends
                        mbar_wait_parked(rel_base + 8u * (uint32_t)split, 0u);
                        const int nb = s_nk[split];
                        if (nb >= max_out) stop = true;
                        const int r = (fi - nb % fn + fn) % fn;
                        k = k_first = nb + r;
                    }
                    // instalments: bit k of `sched` = one after the release of tile tgt - k (most of the tail's boxes
                    // early), the last one after the release of tile tgt itself
#pragma unroll 1
                    for (int k = 7; k >= 1; --k)
                        if (!stop && ((sched >> k) & 1) && tgt >= k && tgt - k > split) advance(tgt - k);
                    if (!stop && tgt > split) advance(tgt);
                }
                if (!stop) {
                    // Band of the OR over the whole share: the exact band of a pair is (tk + tc) * 2^-20.  A pair whose tk
                    // exceeds 1.05 * tc / thr cannot matter: IoU <= area ratio < thr / 1.05, its d is below -0.07 * tc, far
                    // outside any band.  So the band of the pairs that can matter is bounded per candidate, by
                    // (min(largest tk, 1.05 * tc / thr) + tc) * 2^-20 -- without the min the few large kept boxes set the band
                    // of every small candidate and the exact path fires a few times per image (measured: 20 k cycles each).
                    const float g0 = __fmul_rn(__fadd_rn(fminf(tkmax, __fmul_rn(c.tc0, kcap)), c.tc0), kScreenBand);
                    const float g1 = __fmul_rn(__fadd_rn(fminf(tkmax, __fmul_rn(c.tc1, kcap)), c.tc1), kScreenBand);
                    bool h0 = dmax0 > g0, h1 = dmax1 > g1;
                    const bool unsure = (!h0 && dmax0 >= -g0) || (!h1 && dmax1 >= -g1);
                    if (__any_sync(0xffffffffu, unsure)) {   // a pair within 2^-20 of the threshold: exact division
                        PROF_FALLBACK;
                        h0 = false; h1 = false;
                        for (int kk = k_first; kk < nk; kk += fn) {
                            const float4 bk = kb[kk];
                            const float ak = ka[kk];
                            h0 |= iou_gt(bk, ak, c.b0, c.a0, thr);
                            h1 |= iou_gt(bk, ak, c.b1, c.a1, thr);
                        }
                    }
                    hit = ballot64(h0, h1);
                }
            }
            if (stop) {
                // the sweep is over: this warp's partial of every remaining tile of this CTA is completed without data, all
                // (tile, peer) pairs spread over the lanes -- a remote mbarrier.complete_tx costs a round trip through the
                // cluster network, and one per tile in sequence made the drain 4 k cycles long (measured)
                const int cnt = (tiles - 1 - u) / csize + 1;
                for (int idx = lane; idx < cnt * csize; idx += 32) {
                    const int j = idx / csize, p = idx - j * csize;
                    mbar_complete_tx_cluster(mapa_u32(bar_base + 8u * (uint32_t)(u + j * csize), (uint32_t)p), 8u);
                }
                break;
            }
            if (lane < csize) st_async_u64(mapa_u32(my_far, (uint32_t)lane), hit, mapa_u32(bar_u, (uint32_t)lane));
            if (tail && fi == 0) SW_TL(u, 5);
            if (!tail && fi == 0) SW_TL(u, 4);
        }
        if (tail && fi == 0) SW_TL(121, 0);
        if (!tail && fi == 0) SW_TL(121, 1);
    } else if (role > ntail && role <= ntail + nrow) {
        // ================= row warps: 32-row jobs (tile v >= 2, job q), dealt over the cluster =================
        const int rw = role - 1 - ntail;
        const int njobs = (tiles - 2) * kJobs;
        bool stop = false;
        for (int jb = rw * csize + crank; jb < njobs; jb += nrow * csize) {
            const int v = 2 + jb / kJobs, q = jb % kJobs;
            uint32_t m0 = 0u, m1 = 0u;
            if (!stop && v >= look) {   // rows run at most `look` tiles ahead of the resolver: bounds the work past the stop
                mbar_wait_parked(rel_base + 8u * (uint32_t)(v - look), 0u);
                stop = s_nk[v - look] >= max_out;
            }
            stop = stop || *(volatile int*)&s_stop;
            SW_RTL(rw == 0, v, 0);
            const uint32_t slot = cols_base + 8u * (uint32_t)(v * kColWords + q * 32), bar_v = bar_base + 8u * (uint32_t)v;
            if (!stop) {
                const int row0 = (v - (q >> 1)) * kTile + (q & 1) * 32;
                bool done;
                if (q == 1) done = sw_row_job<32, UNIT, true>(bx, n, v, row0, lane, thr, cthr, scr, m0, m1, &s_stop, rw == 0);
                else done = sw_row_job<32, UNIT, false>(bx, n, v, row0, lane, thr, cthr, scr, m0, m1, &s_stop, rw == 0);
                stop = !done;
                SW_RTL(rw == 0, v, 2);
            }
            if (!stop) {
                // the job's 256 bytes go into THIS CTA's slot with plain stores and from there to every peer's slot as one
                // bulk copy each (one transaction per peer instead of 32); the slot is never written again, and the final
                // cluster barrier keeps this CTA alive until every peer has received its copy
                s_cols[(size_t)v * kColWords + q * 32 + lane] = (uint64_t)m0 | ((uint64_t)m1 << 32);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy reads
                __syncwarp();
                if (lane < csize) {
                    if (lane != crank) {
                        bulk_copy_to_peer(mapa_u32(slot, (uint32_t)lane), slot, 256u, mapa_u32(bar_v, (uint32_t)lane));
                    } else {
                        __threadfence_block();   // the warp's stores (ordered by the __syncwarp) before the completion
                        mbar_complete_tx_cluster(mapa_u32(bar_v, (uint32_t)lane), 256u);
                    }
                }
            } else {
                // the sweep is over: every remaining job of this warp is completed without data, (job, peer) pairs spread
                // over the lanes (see the far warps)
                const int stride = nrow * csize, cnt = (njobs - 1 - jb) / stride + 1;
                for (int idx = lane; idx < cnt * csize; idx += 32) {
                    const int j = idx / csize, p = idx - j * csize;
                    const int vv = 2 + (jb + j * stride) / kJobs;
                    mbar_complete_tx_cluster(mapa_u32(bar_base + 8u * (uint32_t)vv, (uint32_t)p), 256u);
                }
                break;
            }
            if (rw == 0) SW_TL(v, 6);
            SW_RTL(rw == 0, v, 3);
        }
        if (rw == 0) SW_TL(121, 2);
    }
    // every CTA has drained its own tile barriers; the cluster barrier then says that every bulk copy this CTA sourced has
    // been received
    cluster.sync();
}

template <int DEPTH>
static const void* sweep_kernel_ptr(bool unit) {
    return unit ? (const void*)nms_sweep_kernel<DEPTH, true> : (const void*)nms_sweep_kernel<DEPTH, false>;
}

// shared memory of one CTA; 0 = does not fit
static size_t sweep_smem_bytes(const void* kernel, int M, int max_out, int nfar, int depth) {
    const SwLayout L = sw_layout((M + kTile - 1) / kTile, max_out < M ? max_out : M, nfar, depth);
    const size_t max_dyn = (size_t)device_props().smem_optin - static_smem_bytes(kernel) - 256;
    return L.total <= max_dyn ? (size_t)L.total : 0;
}

// returns -1 when the problem does not fit this kernel (the caller keeps nms_lazy_kernel)
int launch_nms_sweep(const float4* boxes_sorted, const int32_t* valid, int B, int M, int max_out, float thr,
                     const NmsEpilogue& epi, cudaStream_t stream, bool unit_boxes) {
    unit_boxes = unit_boxes && tuning_knob("MRCNN_SWEEP_UNIT", 1);
    // far warps (tail + bulk) and row warps of a CTA; nfar + nrow <= 15 workers
    int nfar = tuning_knob("MRCNN_SWEEP_NFAR", 12), nrow = tuning_knob("MRCNN_SWEEP_NROW", 3);
    int ntail = tuning_knob("MRCNN_SWEEP_NTAIL", 3);
    if (nfar < 2 || nrow < 1 || nfar + nrow > 15 || ntail < 1 || ntail >= nfar) { nfar = 12; nrow = 3; ntail = 3; }
    const int want = tuning_knob("MRCNN_SWEEP_DEPTH", 3);
    const void* kernel = sweep_kernel_ptr<3>(unit_boxes);
    size_t smem = want >= 3 ? sweep_smem_bytes(kernel, M, max_out, nfar, 3) : 0;
    if (smem == 0) {
        kernel = sweep_kernel_ptr<2>(unit_boxes);
        smem = sweep_smem_bytes(kernel, M, max_out, nfar, 2);
    }
    if (smem == 0) return -1;
    // Cluster size: the sweep is bound by the alu pipes of the cluster's SMs, so as many SMs per image as can be
    // co-scheduled for the whole batch (one CTA per SM); sizes need not be powers of two, above 8 is the non-portable range
    const int cs = pick_cluster_size_any(kernel, kSwThreads, B, tuning_knob("MRCNN_NMS_MAX_CLUSTER", 16), [smem](int) { return smem; });
    const int layout = tuning_knob("MRCNN_SWEEP_LAYOUT", 1);
    int look = tuning_knob("MRCNN_SWEEP_LOOK", 12);
    if (look < 4) look = 4;
    const int sched = tuning_knob("MRCNN_SWEEP_TAILSCHED", 42) & 0xfe;   // tail instalments after tiles tgt-5, tgt-3, tgt-1 (and tgt)
    {   // per launch: the occupancy cache may have set another problem's (smaller) limit last
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cs));
    cfg.blockDim = dim3(kSwThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1 + (unsigned)pdl_attr(attr + 1, stream);
    void* args[] = {(void*)&boxes_sorted, (void*)&valid, (void*)&M, (void*)&max_out, (void*)&thr, (void*)&nfar,
                    (void*)&ntail, (void*)&nrow, (void*)&layout, (void*)&look, (void*)&sched, (void*)&epi};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, args);
    if (e != cudaSuccess) return (int)e;
    return last_error();
}

}  // namespace mrcnn

#ifdef MRCNN_NMS_PROFILE
MRCNN_EXPORT int mrcnn_debug_nms_sweep_timeline(long long* host_out_128x8) {
    return (int)cudaMemcpyFromSymbol(host_out_128x8, mrcnn::g_sw_timeline, sizeof(long long) * 128 * 8);
}
MRCNN_EXPORT int mrcnn_debug_nms_sweep_rows(long long* host_out_128x4) {
    return (int)cudaMemcpyFromSymbol(host_out_128x4, mrcnn::g_sw_rowtl, sizeof(long long) * 128 * 4);
}
#endif
