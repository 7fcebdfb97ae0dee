// roialign.cu -- PyramidROIAlign.call (mrcnn_layers.py:583-664) forward and its feature-map gradient.
//
// forward = roialign_prep_kernel: FPN level per ROI (L:596-607, utils.py:825-827) and, per level, the first
//           flattened index at which it appears (tf.unique order over the whole batch, L:613-615, quirk Q2);
//         + roialign_fwd_kernel (one warp per output row of a ROI): each warp turns the four first-appearance
//           indices into its ROI's map index and walks the bins of its row with TF CropAndResize bilinear sampling
//           (crop_and_resize_op.cc; called at L:641) with 128-bit channel-vectorised NHWC loads, written straight
//           into [B,N,ph,pw,C] in input ROI order -- the reference's concat / top_k re-sort / gather passes
//           (L:644-659) have no counterpart here because nothing is ever out of order.
// backward (with a workspace) = deterministic pixel-centric gather: count the corner samples that land on every
//           gradient-map pixel, give each touched pixel a segment of a key list, fill it, then ONE pass over all
//           pixels writes each of them exactly once -- zeros, or the sum of its samples in TF CropAndResizeGradImage's
//           own (box, y, x, corner) order, so the result is bit-identical to the sequential CPU kernel and the
//           713 MB zero-fill and the accumulation are the same HBM write.  Pixels that collect more than 1024
//           samples (zero-padded ROIs pile thousands on pixel (0,0), quirk Q5) fall back to vector atomics.
// backward (no workspace) = memset of the four gradient maps + roialign_bwd_kernel: scatter with
//           red.global.add.v4.f32 (one 16-byte reduction per lane and corner).
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace mrcnn {

struct MapTable {
    const float* ptr[4];
    int H[4];
    int W[4];
};
struct GradTable {
    float* ptr[4];
    int H[4];
    int W[4];
};

constexpr int kPrepThreads = 1024; // ROIs per prep CTA = ROIs per entry of the first-appearance partial list
constexpr int kRoiThreads = 256;   // 8 warps per CTA

__device__ __forceinline__ int roi_level_of(float4 b, float denom) {
    const float h = __fsub_rn(b.z, b.x);
    const float w = __fsub_rn(b.w, b.y);
    const float x = __fdiv_rn(__fsqrt_rn(__fmul_rn(h, w)), denom);                 // L:605
    const float lv = __fdiv_rn(det_logf(x), 0.693147182464599609375f);              // utils.log2_graph
    const int r = cast_i32_x86(rintf(lv));                                           // tf.round + int32 cast
    int level = (r < 0) ? max(4 + r, 2) : min(4 + r, 5);                             // L:607, overflow-free
    return min(max(level, 2), 5);
}

// level per ROI + first flattened index at which each level appears (tf.unique order, L:613).  Every CTA writes the
// minima of its own kPrepThreads ROIs to partial[blockIdx.x] (INT_MAX = level absent); the consumers reduce the
// partial list themselves (first_appearance below: one int4 per lane at B*N <= 32768) -- no atomics, so no memset node
// in front of the kernel, and the launch chain stays eligible for programmatic dependent launch.
__global__ void __launch_bounds__(kPrepThreads)
roialign_prep_kernel(const float4* __restrict__ boxes, const float* __restrict__ image_meta, int BN,
                     float denominator, int32_t* __restrict__ level_ws, int4* __restrict__ partial,
                     int32_t* __restrict__ roi_level) {
    __shared__ int s_min[kPrepThreads / 32][4];
    pdl_launch_dependents();
    pdl_wait();
    const int f = blockIdx.x * kPrepThreads + threadIdx.x;
    const float image_area = __fmul_rn(image_meta[4], image_meta[5]);               // L:600,604 (image 0)
    const float denom = __fdiv_rn(denominator, __fsqrt_rn(image_area));
    int level = 0;
    if (f < BN) {
        level = roi_level_of(__ldg(boxes + f), denom);
        level_ws[f] = level;
        if (roi_level) roi_level[f] = level;
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const int m = __reduce_min_sync(0xffffffffu, (level == l + 2) ? f : INT_MAX);
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5][l] = m;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        int m = INT_MAX;
#pragma unroll
        for (int w = 0; w < kPrepThreads / 32; ++w) m = min(m, s_min[w][threadIdx.x]);
        reinterpret_cast<int*>(partial + blockIdx.x)[threadIdx.x] = m;
    }
}

// the four first-appearance indices, reduced from the prep kernel's per-CTA minima; every lane of the warp must call
__device__ __forceinline__ int4 first_appearance(const int4* __restrict__ partial, int nparts) {
    int4 m = make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
    for (int i = threadIdx.x & 31; i < nparts; i += 32) {
        const int4 p = __ldg(partial + i);
        m.x = min(m.x, p.x); m.y = min(m.y, p.y); m.z = min(m.z, p.z); m.w = min(m.w, p.w);
    }
    m.x = __reduce_min_sync(0xffffffffu, m.x);
    m.y = __reduce_min_sync(0xffffffffu, m.y);
    m.z = __reduce_min_sync(0xffffffffu, m.z);
    m.w = __reduce_min_sync(0xffffffffu, m.w);
    return m;
}

// Sampling geometry of one ROI on its feature map, computed once per warp (all lanes redundantly, no divergence):
// TF crop_and_resize_op.cc evaluates in_y = y1*(H-1) + y*height_scale per output row (and the box centre when the
// crop has a single row); y0 / x0 hold the y- and x-independent terms so a bin costs one multiply and one add.
struct RoiGeom {
    float y0, x0, hs, ws;  // in_y = y0 + y * hs, in_x = x0 + x * ws (individually rounded, as TF)
    int H, W, m;
};

__device__ __forceinline__ RoiGeom roi_geom(float4 b, int m, const int (&Hs)[4], const int (&Ws)[4], int ph, int pw) {
    RoiGeom g;
    g.m = m;
    g.H = (m == 0) ? Hs[0] : (m == 1) ? Hs[1] : (m == 2) ? Hs[2] : Hs[3];
    g.W = (m == 0) ? Ws[0] : (m == 1) ? Ws[1] : (m == 2) ? Ws[2] : Ws[3];
    if (ph > 1) { g.hs = crop_scale(b.x, b.z, g.H, ph); g.y0 = __fmul_rn(b.x, (float)(g.H - 1)); }
    else { g.hs = 0.0f; g.y0 = (float)(0.5 * (double)__fadd_rn(b.x, b.z) * (double)(g.H - 1)); }
    if (pw > 1) { g.ws = crop_scale(b.y, b.w, g.W, pw); g.x0 = __fmul_rn(b.y, (float)(g.W - 1)); }
    else { g.ws = 0.0f; g.x0 = (float)(0.5 * (double)__fadd_rn(b.y, b.w) * (double)(g.W - 1)); }
    return g;
}

struct AxisTap {
    int lo, hi;
    float lerp;
    bool valid;
};
__device__ __forceinline__ AxisTap axis_tap(float c0, float scale, int t, int size) {
    AxisTap r;
    const float in = __fadd_rn(c0, __fmul_rn((float)t, scale));
    r.valid = (in >= 0.0f && in <= (float)(size - 1));
    const float fl = floorf(in);
    r.lo = r.valid ? (int)fl : 0;
    r.hi = r.valid ? (int)ceilf(in) : 0;
    r.lerp = __fsub_rn(in, fl);
    return r;
}

// map index of a ROI from its level and the four first-appearance indices (L:613-619), or level-2 (map_mode 1)
__device__ __forceinline__ int roi_map_index(int level, const int4& fa, int map_mode) {
    if (map_mode != 0) return level - 2;
    const int mine = (level == 2) ? fa.x : (level == 3) ? fa.y : (level == 4) ? fa.z : fa.w;
    return (fa.x < mine) + (fa.y < mine) + (fa.z < mine) + (fa.w < mine);
}

// One warp per output ROW (roi f, row y): lanes span the channels with 128-bit accesses, the warp walks the pw
// bins of its row.  VPL = float4 vectors per lane (C == VPL * 128); VPL == 0 -> any C that is a multiple of 4.
// XSPLIT: warps per output row (compile-time: the one-warp-per-row code of narrow crops must not change).
template <int VPL, int XSPLIT>
__global__ void __launch_bounds__(kRoiThreads, 5)
roialign_fwd_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                    const int4* __restrict__ partial, int nparts, int map_mode, MapTable tbl, int C, int N, int ph, int pw,
                    int total_rows, float* __restrict__ out, int32_t* __restrict__ roi_map) {
    constexpr int xsplit = XSPLIT;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    const int4 fa = first_appearance(partial, nparts);  // every warp, before the early exit (warp-uniform anyway)
    // a warp owns one output row, or 1/xsplit of it (wide crops: more, shorter warps fill the last wave better)
    const int unit = blockIdx.x * (kRoiThreads / 32) + (threadIdx.x >> 5);
    if (unit >= total_rows * xsplit) return;
    const int row = unit / xsplit, part = unit - row * xsplit;
    const int x_begin = part * pw / xsplit, x_end = (part + 1) * pw / xsplit;
    const int f = row / ph, y = row - f * ph;
    const int m = roi_map_index(level_ws[f], fa, map_mode);
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    if (y == 0 && part == 0 && lane == 0) roi_map[f] = m;
    const int c4 = C >> 2;
    const float4* img = reinterpret_cast<const float4*>(
        ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
        (size_t)(f / N) * g.H * g.W * C);
    float4* o = reinterpret_cast<float4*>(out) + ((size_t)row * pw + x_begin) * c4;
    const AxisTap ty = axis_tap(g.y0, g.hs, y, g.H);
    const int top = ty.lo * g.W, bot = ty.hi * g.W;  // pixel index of the two sampled rows
    const float ly = ty.lerp;
    for (int x = x_begin; x < x_end; ++x, o += c4) {
        const AxisTap tx = axis_tap(g.x0, g.ws, x, g.W);
        if (!(ty.valid && tx.valid)) {  // extrapolation_value = 0
            for (int v = lane; v < c4; v += 32) __stcs(o + v, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const float4* tl = img + (size_t)(top + tx.lo) * c4;
        const float4* tr = img + (size_t)(top + tx.hi) * c4;
        const float4* bl = img + (size_t)(bot + tx.lo) * c4;
        const float4* br = img + (size_t)(bot + tx.hi) * c4;
        const float lx = tx.lerp;
        auto lerp1 = [&](float a, float b, float c, float d) {
            const float t = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), lx));
            const float u = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), lx));
            return __fadd_rn(t, __fmul_rn(__fsub_rn(u, t), ly));
        };
        auto lerp4 = [&](const float4& a, const float4& b, const float4& c, const float4& d) {
            return make_float4(lerp1(a.x, b.x, c.x, d.x), lerp1(a.y, b.y, c.y, d.y), lerp1(a.z, b.z, c.z, d.z),
                               lerp1(a.w, b.w, c.w, d.w));
        };
        if (VPL > 0) {
            float4 a[VPL > 0 ? VPL : 1], b[VPL > 0 ? VPL : 1], c[VPL > 0 ? VPL : 1], d[VPL > 0 ? VPL : 1];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int i = lane + 32 * v;
                a[v] = __ldg(tl + i); b[v] = __ldg(tr + i); c[v] = __ldg(bl + i); d[v] = __ldg(br + i);
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) __stcs(o + lane + 32 * v, lerp4(a[v], b[v], c[v], d[v]));
        } else {
            for (int i = lane; i < c4; i += 32) __stcs(o + i, lerp4(__ldg(tl + i), __ldg(tr + i), __ldg(bl + i), __ldg(br + i)));
        }
    }
}

// ---- TMA-staged forward (C == 256: one feature-map pixel = one 1 KB bulk copy) ----------------------------------
// Persistent warps; each warp owns a ring of STAGES x 4 KB in shared memory.  For every output bin the warp's lane 0
// arms an mbarrier with 4096 expected bytes and issues four `cp.async.bulk.shared.global` copies (the four bilinear
// corners; SASS: UBLKCP) -- the copy engine, not the LSU, moves the pixels, the warp never holds them in registers
// while they are in flight, and the ring runs STAGES-1 bins ahead ACROSS row boundaries (a second geometry context
// follows the producer cursor), so there is no exposed latency per row.  Consumption: 8 x LDS.128 per lane, the same
// individually rounded lerps as roialign_fwd_kernel, 2 x streaming STG.128.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_mbar_arm(uint32_t bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct RowCtx {        // everything about one output row (roi f, row y) that does not depend on x
    const float4* img; // feature map of the ROI's image
    float x0, ws, ly;
    int top, bot, W;   // pixel index of the two sampled map rows
    bool yvalid;
};

__device__ __forceinline__ RowCtx make_row_ctx(int row, const float4* __restrict__ boxes,
                                               const int32_t* __restrict__ level_ws, const int4& fa,
                                               int map_mode, const MapTable& tbl, int N, int ph, int pw, int& m_out) {
    const int f = row / ph, y = row - f * ph;
    const int m = roi_map_index(level_ws[f], fa, map_mode);
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const AxisTap ty = axis_tap(g.y0, g.hs, y, g.H);
    RowCtx c;
    c.img = reinterpret_cast<const float4*>(((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2]
                                                                                              : tbl.ptr[3]) +
                                            (size_t)(f / N) * g.H * g.W * 256);
    c.x0 = g.x0; c.ws = g.ws; c.ly = ty.lerp; c.top = ty.lo * g.W; c.bot = ty.hi * g.W; c.W = g.W; c.yvalid = ty.valid;
    m_out = m;
    return c;
}

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32)
roialign_fwd_tma_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                        const int4* __restrict__ partial, int nparts, int map_mode, MapTable tbl, int N, int ph, int pw,
                        int total_rows, float* __restrict__ out, int32_t* __restrict__ roi_map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int4 fa = first_appearance(partial, nparts);
    float4* ring = reinterpret_cast<float4*>(smem_raw) + (size_t)warp * STAGES * 256;        // [STAGES][4 corners][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)WARPS * STAGES * 4096) + warp * STAGES;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) tma_mbar_init(smem_addr(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int nwarps = gridDim.x * WARPS;
    const int w0 = blockIdx.x * WARPS + warp;
    if (w0 >= total_rows) return;
    const int my_rows = (total_rows - w0 + nwarps - 1) / nwarps;  // rows w0, w0 + nwarps, ...
    const int my_bins = my_rows * pw;

    int m_dummy;
    RowCtx pc = make_row_ctx(w0, boxes, level_ws, fa, map_mode, tbl, N, ph, pw, m_dummy);  // producer cursor
    int p_row = 0, p_x = 0, issued = 0;
    auto produce = [&]() {  // bin `issued` of this warp's stream -> stage issued % STAGES
        const AxisTap tx = axis_tap(pc.x0, pc.ws, p_x, pc.W);
        const bool valid = pc.yvalid && tx.valid;
        if (lane == 0) {
            const int s = issued % STAGES;
            const uint32_t bar = smem_addr(bars + s), dst = smem_addr(ring + s * 256);
            tma_mbar_arm(bar, valid ? 4096u : 0u);
            if (valid && tx.hi == tx.lo + 1) {  // left and right pixels are neighbours in memory: one 2 KB copy per row
                tma_load_1d(dst, pc.img + (size_t)(pc.top + tx.lo) * 64, 2048, bar);
                tma_load_1d(dst + 2048, pc.img + (size_t)(pc.bot + tx.lo) * 64, 2048, bar);
            } else if (valid) {                 // integer sample position: lo == hi
                tma_load_1d(dst, pc.img + (size_t)(pc.top + tx.lo) * 64, 1024, bar);
                tma_load_1d(dst + 1024, pc.img + (size_t)(pc.top + tx.hi) * 64, 1024, bar);
                tma_load_1d(dst + 2048, pc.img + (size_t)(pc.bot + tx.lo) * 64, 1024, bar);
                tma_load_1d(dst + 3072, pc.img + (size_t)(pc.bot + tx.hi) * 64, 1024, bar);
            }
        }
        ++issued;
        if (++p_x == pw) {
            p_x = 0;
            if (++p_row < my_rows)
                pc = make_row_ctx(w0 + p_row * nwarps, boxes, level_ws, fa, map_mode, tbl, N, ph, pw, m_dummy);
        }
    };
    for (int k = 0; k < STAGES - 1 && issued < my_bins; ++k) produce();

    int consumed = 0;
#pragma unroll 1
    for (int r = 0; r < my_rows; ++r) {
        const int row = w0 + r * nwarps;
        int m;
        const RowCtx cc = make_row_ctx(row, boxes, level_ws, fa, map_mode, tbl, N, ph, pw, m);
        if (lane == 0 && row % ph == 0) roi_map[row / ph] = m;
        float4* o = reinterpret_cast<float4*>(out) + (size_t)row * pw * 64;
        const float ly = cc.ly;
#pragma unroll 1
        for (int x = 0; x < pw; ++x, o += 64, ++consumed) {
            __syncwarp();                       // every lane is done reading the stage the next copy will overwrite
            if (issued < my_bins) produce();
            const AxisTap tx = axis_tap(cc.x0, cc.ws, x, cc.W);
            const int s = consumed % STAGES;
            tma_mbar_wait(smem_addr(bars + s), (uint32_t)(consumed / STAGES) & 1u);
            if (!(cc.yvalid && tx.valid)) {     // extrapolation_value = 0
                __stcs(o + lane, make_float4(0.f, 0.f, 0.f, 0.f));
                __stcs(o + lane + 32, make_float4(0.f, 0.f, 0.f, 0.f));
                continue;
            }
            const float4* st = ring + s * 256;
            const float lx = tx.lerp;
            auto lerp1 = [&](float a, float b, float c, float d) {
                const float t = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), lx));
                const float u = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), lx));
                return __fadd_rn(t, __fmul_rn(__fsub_rn(u, t), ly));
            };
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const float4 a = st[lane + 32 * v], b = st[64 + lane + 32 * v], c = st[128 + lane + 32 * v],
                             d = st[192 + lane + 32 * v];
                __stcs(o + lane + 32 * v, make_float4(lerp1(a.x, b.x, c.x, d.x), lerp1(a.y, b.y, c.y, d.y),
                                                       lerp1(a.z, b.z, c.z, d.z), lerp1(a.w, b.w, c.w, d.w)));
            }
        }
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// TF CropAndResizeGradImage: dtop = (1-ly) g; tl += (1-lx) dtop; tr += lx dtop; dbot = ly g; bl += (1-lx) dbot;
// br += lx dbot -- skipped exactly where the forward pass extrapolated.
//
// Atomic mode (no workspace): memset of the maps, then roialign_bwd_kernel scatters with 16-byte vector reductions.
// One CTA per (ROI, group of <= 8 output rows), one warp per row, lanes over channels.  Contention control:
//   * a corner whose weight is exactly 0 is not touched (finite gradients: adding 0 changes nothing);
//   * bins of a row that hit the same pixels (zero width scale) are summed in registers before one reduction;
//   * ROIs whose taps are constant in both axes (zero-padded ROIs: every bin samples pixel (0,0), quirk Q5) are
//     summed across the CTA's rows in shared memory and issue ONE reduction per CTA and corner, which keeps the
//     thousands of padded rows of a training batch from serialising on a single L2 line.
//
// Deterministic mode (workspace): tile-owner accumulation, see roialign_bwd_tile_kernel below.
struct PixelSpace {   // global pixel id = base[m] + (b * H[m] + y) * W[m] + x ; base[4] = number of pixels
    int base[5];
};

constexpr int kTileH = 4, kTileW = 8; // gradient-map tile of the deterministic mode: 4 rows x 8 columns of pixels
constexpr int kTilePix = kTileH * kTileW;
constexpr int kTileCap = 1024;       // samples one tile orders in shared memory per round
constexpr int kTileThreads = 64;     // one thread per four channels (256 channels per pass)
constexpr int kTileChunk = 64;       // float4 channel vectors per pass
constexpr int kTileQueue = 128;      // samples decoded per consume round
constexpr int kTileBuckets = 2;      // work lists by sample count (>= 128 / fewer): the heavy tiles are started first
constexpr int kTileCtasPerSm = 5;    // persistent grid of the tile kernel (42 KB of shared memory per CTA)
constexpr int kTileFBuckets = 256;   // a tile with more than kTileCap samples is taken in rounds of ROI-index ranges
// sample key: ROI index << 14 | output row << 7 | output column -- ascending key = TF's accumulation order
constexpr int kKeyXBits = 7, kKeyYBits = 7, kKeyFBits = 18;

struct TileSpace {    // tile id = base[m] + (b * ty[m] + tile_y) * tx[m] + tile_x ; base[4] = number of tiles
    int base[5];
    int ty[4], tx[4];
};
// misc words of the deterministic workspace
enum { kMiscBump = 0, kMiscOverflow = 1, kMiscBucket0 = 2 /* .. 5 */, kMiscWords = 16 };

__device__ __forceinline__ int tile_of_pixel(const TileSpace& ts, int m, int b, int y, int x) {
    const int ty = (m == 0) ? ts.ty[0] : (m == 1) ? ts.ty[1] : (m == 2) ? ts.ty[2] : ts.ty[3];
    const int tx = (m == 0) ? ts.tx[0] : (m == 1) ? ts.tx[1] : (m == 2) ? ts.tx[2] : ts.tx[3];
    const int base = (m == 0) ? ts.base[0] : (m == 1) ? ts.base[1] : (m == 2) ? ts.base[2] : ts.base[3];
    return base + (b * ty + y / kTileH) * tx + x / kTileW;
}

// `tile_start` == nullptr: every corner is scattered (atomic mode).  Otherwise only corners whose pixel lies in a tile
// that overflowed kTileRoiMax (tile_start < 0) are; the tile kernel owns the rest.
struct ScatterFilter {
    const int* tile_start;
    TileSpace ts;
    int m, b, W;
    __device__ __forceinline__ bool ours(int p) const {   // p = pixel index inside the (image, map) plane
        if (tile_start == nullptr) return true;
        return __ldg(tile_start + tile_of_pixel(ts, m, b, p / W, p - (p / W) * W)) < 0;
    }
};

__device__ __forceinline__ void scatter_corners(float* gimg, int C, const ScatterFilter& flt, int ptl, int ptr_, int pbl,
                                                int pbr, int i, const float4& v, float wy0, float ly, float wx0, float lx) {
    const float4 dt = make_float4(__fmul_rn(wy0, v.x), __fmul_rn(wy0, v.y), __fmul_rn(wy0, v.z), __fmul_rn(wy0, v.w));
    const float4 db = make_float4(__fmul_rn(ly, v.x), __fmul_rn(ly, v.y), __fmul_rn(ly, v.z), __fmul_rn(ly, v.w));
    if (wy0 != 0.0f && wx0 != 0.0f && flt.ours(ptl))
        red_add_v4(gimg + (size_t)ptl * C + 4 * i, __fmul_rn(wx0, dt.x), __fmul_rn(wx0, dt.y), __fmul_rn(wx0, dt.z), __fmul_rn(wx0, dt.w));
    if (wy0 != 0.0f && lx != 0.0f && flt.ours(ptr_))
        red_add_v4(gimg + (size_t)ptr_ * C + 4 * i, __fmul_rn(lx, dt.x), __fmul_rn(lx, dt.y), __fmul_rn(lx, dt.z), __fmul_rn(lx, dt.w));
    if (ly != 0.0f && wx0 != 0.0f && flt.ours(pbl))
        red_add_v4(gimg + (size_t)pbl * C + 4 * i, __fmul_rn(wx0, db.x), __fmul_rn(wx0, db.y), __fmul_rn(wx0, db.z), __fmul_rn(wx0, db.w));
    if (ly != 0.0f && lx != 0.0f && flt.ours(pbr))
        red_add_v4(gimg + (size_t)pbr * C + 4 * i, __fmul_rn(lx, db.x), __fmul_rn(lx, db.y), __fmul_rn(lx, db.z), __fmul_rn(lx, db.w));
}

__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ boxes,
                    const int32_t* __restrict__ roi_map, GradTable tbl, int C, int N, int ph, int pw, int groups,
                    int rows_per_group, TileSpace ts, const int* __restrict__ tile_start,
                    const int* __restrict__ overflow) {
    extern __shared__ __align__(16) float4 s_acc[];  // [8 warps][C/4], constant-tap ROIs only
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // deterministic mode: nothing to do unless some tile overflowed kTileRoiMax
    if (tile_start != nullptr && *overflow == 0) return;
    const int f = blockIdx.x / groups, grp = blockIdx.x - f * groups;
    const int y = grp * rows_per_group + warp;
    const bool has_row = (warp < rows_per_group) && (y < ph);
    const int m = roi_map[f];
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const int c4 = C >> 2;
    const int b = f / N;
    float* gimg = ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
                  (size_t)b * g.H * g.W * C;
    ScatterFilter flt;
    flt.tile_start = tile_start; flt.ts = ts; flt.m = m; flt.b = b; flt.W = g.W;
    const bool constant = (g.hs == 0.0f && g.ws == 0.0f);  // CTA-uniform
    const AxisTap ty = axis_tap(g.y0, g.hs, has_row ? y : 0, g.H);
    const float4* gr = grad_out + ((size_t)f * ph + (has_row ? y : 0)) * pw * c4;
    const int top = ty.lo * g.W, bot = ty.hi * g.W;
    const float ly = ty.lerp, wy0 = __fsub_rn(1.0f, ty.lerp);
    if (constant) {
        const AxisTap tx = axis_tap(g.x0, g.ws, 0, g.W);
        for (int i = lane; i < c4; i += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_row)
                for (int x = 0; x < pw; ++x) {
                    const float4 v = __ldcs(gr + (size_t)x * c4 + i);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            s_acc[warp * c4 + i] = acc;
        }
        __syncthreads();
        if (!(ty.valid && tx.valid)) return;  // uniform: the same tap for every bin of the ROI
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = threadIdx.x; i < c4; i += kRoiThreads) {
            float4 acc = s_acc[i];
            for (int w = 1; w < kRoiThreads / 32; ++w) {
                const float4 v = s_acc[w * c4 + i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            scatter_corners(gimg, C, flt, top + tx.lo, top + tx.hi, bot + tx.lo, bot + tx.hi, i, acc, wy0, ly, wx0, lx);
        }
        return;
    }
    if (!has_row || !ty.valid) return;
    if (g.ws == 0.0f && tile_start == nullptr) {  // every bin of the row hits the same pixels: sum the row first
        const AxisTap tx = axis_tap(g.x0, g.ws, 0, g.W);
        if (!tx.valid) return;
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int x = 0; x < pw; ++x) {
                const float4 v = __ldcs(gr + (size_t)x * c4 + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            scatter_corners(gimg, C, flt, top + tx.lo, top + tx.hi, bot + tx.lo, bot + tx.hi, i, acc, wy0, ly, wx0, lx);
        }
        return;
    }
    for (int x = 0; x < pw; ++x, gr += c4) {
        const AxisTap tx = axis_tap(g.x0, g.ws, x, g.W);
        if (!tx.valid) continue;
        const int ptl = top + tx.lo, ptr_ = top + tx.hi, pbl = bot + tx.lo, pbr = bot + tx.hi;
        if (tile_start != nullptr && !(flt.ours(ptl) || flt.ours(ptr_) || flt.ours(pbl) || flt.ours(pbr)))
            continue;  // warp-uniform: no gradient row is read for bins the tile kernel owns entirely
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32)
            scatter_corners(gimg, C, flt, ptl, ptr_, pbl, pbr, i, __ldcs(gr + i), wy0, ly, wx0, lx);
    }
}

// ---- deterministic backward: tile-owner accumulation ------------------------------------------------------------
// Every gradient map is cut into tiles of 4 x 8 pixels and every tile is written exactly once, by ONE CTA that adds up
// all the samples landing on it in TF CropAndResizeGradImage's own order -- ROI index ascending, then output row, column,
// corner -- with every product and sum individually rounded: the result is bit-identical to the sequential CPU kernel
// and reproducible, and the zero-fill of the maps (713 MB per 8 images at 1024^2) and the accumulation are the same
// HBM write.  Steps:
//   roialign_bwd_taps_kernel    thread per (ROI, output row or column): the TF sampling tap (lo, hi, lerp, valid) -- the
//                               only place the ROI geometry (two fp32 divisions per axis) is evaluated;
//   roialign_bwd_const_kernel   zero-size ROIs (zero-padded target rows, quirk Q5: all ph*pw bins on the same four taps)
//                               are pre-reduced to one gradient row each (fixed order) and enter as ONE sample;
//   roialign_bwd_bin_kernel<0>  thread per sample: one count per tile its (up to four) corners land on;
//   roialign_bwd_alloc_kernel   a segment of the (tile, sample) list per tile (block scan, one atomic per CTA), and the
//                               tile goes on one of four work lists by sample count (the heavy tiles are started first);
//   roialign_bwd_bin_kernel<1>  the same walk, now storing the sample key into the segments;
//   roialign_bwd_tile_kernel    CTAs [0, NT): the ranked non-empty tiles; CTAs [NT, 2 NT): zero-fill of the empty ones.
//                               A tile CTA (64 threads, a thread owns four channels of every pixel of the tile in a
//                               32 KB shared-memory accumulator) sorts its sample keys, then alternates: 64 threads
//                               decode 192 samples (taps -> which corner lands on which tile pixel with which weight),
//                               then all of them walk those samples in order, gradient rows fetched eight samples
//                               ahead.  A gradient row is read once per tile it touches (1.4 x in all at 1024^2, from
//                               L2) instead of once per corner (4 x).  More than kTileCap samples on a tile (thousands of
//                               ROIs per image): rounds over ROI-index ranges chosen from a 256-bucket histogram.
// Not bit-identical to the sequential order, but still deterministic: pixels under zero-size ROIs (their pre-reduced
// row is added as one sample) -- thousands of samples on one pixel cannot be added one after the other at any speed;
// non-deterministic: a tile where ONE histogram bucket of ROI indices holds more than kTileCap samples (atomic scatter
// fallback; needs > 5000 ROIs per image on the same 32 pixels).
// Maps up to 512 x 512 pixels, ph, pw <= 128, B * N < 2^18.
struct TapWord {   // .x = lo | (hi - lo) << 16 | valid << 17 | zero-size ROI << 18 ; .y = lerp bits
    uint32_t x, y;
};
__device__ __forceinline__ int tap_lo(const TapWord& t) { return (int)(t.x & 0xffffu); }
__device__ __forceinline__ int tap_hi(const TapWord& t) { return (int)(t.x & 0xffffu) + (int)((t.x >> 16) & 1u); }
__device__ __forceinline__ bool tap_valid(const TapWord& t) { return (t.x >> 17) & 1u; }
__device__ __forceinline__ bool tap_const(const TapWord& t) { return (t.x >> 18) & 1u; }

__global__ void __launch_bounds__(256)
roialign_bwd_taps_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ roi_map, GradTable tbl, int BN,
                         int ph, int pw, TapWord* __restrict__ taps /*[BN][ph + pw]*/) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= BN * (ph + pw)) return;
    const int f = i / (ph + pw), t = i - f * (ph + pw);
    const RoiGeom g = roi_geom(__ldg(boxes + f), roi_map[f], tbl.H, tbl.W, ph, pw);
    const AxisTap a = (t < ph) ? axis_tap(g.y0, g.hs, t, g.H) : axis_tap(g.x0, g.ws, t - ph, g.W);
    TapWord w;
    w.x = (uint32_t)a.lo | ((uint32_t)(a.hi - a.lo) << 16) | (a.valid ? 1u << 17 : 0u) |
          ((g.hs == 0.0f && g.ws == 0.0f) ? 1u << 18 : 0u);
    w.y = __float_as_uint(a.lerp);
    taps[i] = w;
}

__global__ void __launch_bounds__(64)
roialign_bwd_const_kernel(const float4* __restrict__ grad_out, const TapWord* __restrict__ taps, int C, int ph, int pw,
                          float4* __restrict__ partial /*[BN][C/4]*/) {
    const int f = blockIdx.x;
    if (!tap_const(taps[(size_t)f * (ph + pw)])) return;
    const int c4 = C >> 2, bins = ph * pw;
    const float4* gr = grad_out + (size_t)f * bins * c4;
    for (int i = threadIdx.x; i < c4; i += 64) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0;
        for (; s + 8 <= bins; s += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(gr + (size_t)(s + u) * c4 + i);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y);
                acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
            }
        }
        for (; s < bins; ++s) {
            const float4 v = __ldcs(gr + (size_t)s * c4 + i);
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        partial[(size_t)f * c4 + i] = acc;
    }
}

// the (up to four) distinct tiles the corners of one sample land on with non-zero weight; returns their number
__device__ __forceinline__ int sample_tiles(const TapWord& ty, const TapWord& tx, int tbase, int TX, int (&tiles)[4]) {
    if (!(tap_valid(ty) && tap_valid(tx))) return 0;
    const float ly = __uint_as_float(ty.y), lx = __uint_as_float(tx.y);
    const bool t_on = __fsub_rn(1.0f, ly) != 0.0f, b_on = ly != 0.0f, l_on = __fsub_rn(1.0f, lx) != 0.0f, r_on = lx != 0.0f;
    const int r0 = tap_lo(ty) / kTileH, r1 = tap_hi(ty) / kTileH, c0 = tap_lo(tx) / kTileW, c1 = tap_hi(tx) / kTileW;
    int n = 0;
    auto add = [&](int r, int c) {
        const int t = tbase + r * TX + c;
        for (int i = 0; i < n; ++i)
            if (tiles[i] == t) return;
        tiles[n++] = t;
    };
    if (t_on && l_on) add(r0, c0);
    if (t_on && r_on) add(r0, c1);
    if (b_on && l_on) add(r1, c0);
    if (b_on && r_on) add(r1, c1);
    return n;
}

template <int PASS>
__global__ void __launch_bounds__(256)
roialign_bwd_bin_kernel(const TapWord* __restrict__ taps, const int32_t* __restrict__ roi_map, TileSpace ts, int N,
                        int ph, int pw, int total_bins, int* __restrict__ count, const int* __restrict__ start,
                        int* __restrict__ cursor, uint32_t* __restrict__ entries) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= total_bins) return;
    const int bins = ph * pw;
    const int f = s / bins, r = s - f * bins, y = r / pw, x = r - y * pw;
    const TapWord* tf = taps + (size_t)f * (ph + pw);
    const TapWord ty = tf[y], tx = tf[ph + x];
    if (tap_const(ty) && r != 0) return;           // a zero-size ROI is ONE sample (its pre-reduced row)
    const int m = roi_map[f];
    const int TX = (m == 0) ? ts.tx[0] : (m == 1) ? ts.tx[1] : (m == 2) ? ts.tx[2] : ts.tx[3];
    int tiles[4];
    const int n = sample_tiles(ty, tx, tile_of_pixel(ts, m, f / N, 0, 0), TX, tiles);
    const uint32_t key = ((uint32_t)f << (kKeyXBits + kKeyYBits)) | ((uint32_t)y << kKeyXBits) | (uint32_t)x;
    for (int i = 0; i < n; ++i) {
        if (PASS == 0) atomicAdd(count + tiles[i], 1);
        else {
            const int seg = start[tiles[i]];
            if (seg >= 0) entries[seg + atomicAdd(cursor + tiles[i], 1)] = key;
        }
    }
}

__global__ void __launch_bounds__(256)
roialign_bwd_alloc_kernel(const int* __restrict__ count, int NT, int* __restrict__ start, int* __restrict__ misc,
                          int4* __restrict__ lists /*[kTileBuckets][NT]: (tile, segment start, samples, -)*/) {
    __shared__ int warp_sums[32];
    __shared__ int block_total, block_base;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int c = (t < NT) ? count[t] : 0;
    int off = block_exclusive_scan(c, warp_sums, &block_total);
    if (threadIdx.x == 0) block_base = block_total > 0 ? atomicAdd(&misc[kMiscBump], block_total) : 0;
    __syncthreads();
    if (t >= NT) return;
    start[t] = block_base + off;
    if (c > 0) {   // one atomic per bucket and warp
        const int bucket = (c >= 128) ? 0 : 1;
        const unsigned peers = __match_any_sync(__activemask(), bucket);
        const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
        int base = 0;
        if (lane == leader) base = atomicAdd(&misc[kMiscBucket0 + bucket], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        lists[(size_t)bucket * NT + base + __popc(peers & ((1u << lane) - 1u))] = make_int4(t, block_base + off, c, 0);
    }
}

__device__ __forceinline__ void acc_corner(float4* acc, const float4& d, float w) {
    float4 a = *acc;   // acc += w * d with the product and the sum individually rounded (TF's two statements)
    a.x = __fadd_rn(a.x, __fmul_rn(w, d.x));
    a.y = __fadd_rn(a.y, __fmul_rn(w, d.y));
    a.z = __fadd_rn(a.z, __fmul_rn(w, d.z));
    a.w = __fadd_rn(a.w, __fmul_rn(w, d.w));
    *acc = a;
}

// One decoded sample of the tile's queue: where its gradient row lies and which of its four corners land on which tile
// pixel.  pk: bits 0-4 top row's first pixel (row * 8), 5-7 column lo, 8-10 column hi, 11-15 bottom row's first pixel,
// 16 tl, 17 tr, 18 bl, 19 br (corner lands in the tile with non-zero weight), 20 pre-reduced row (zero-size ROI)
struct TileQueue {
    int off[kTileQueue];
    uint32_t pk[kTileQueue];
    float4 w[kTileQueue];   // (1 - ly, ly, 1 - lx, lx)
};

// geometry of a tile: which map, which image, first pixel, extent
struct TileGeom {
    float4* gmap;
    int W, y0, x0, rows_in, cols_in;
};
__device__ __forceinline__ TileGeom tile_geom(const GradTable& tbl, const TileSpace& ts, int tile, int c4) {
    const int m = (tile >= ts.base[3]) ? 3 : (tile >= ts.base[2]) ? 2 : (tile >= ts.base[1]) ? 1 : 0;
    const int H = tbl.H[m], W = tbl.W[m], TY = ts.ty[m], TX = ts.tx[m];
    const int local = tile - ts.base[m];
    const int b = local / (TY * TX), rem = local - b * (TY * TX);
    TileGeom g;
    g.W = W;
    g.y0 = (rem / TX) * kTileH;
    g.x0 = (rem - (rem / TX) * TX) * kTileW;
    g.rows_in = min(kTileH, H - g.y0);
    g.cols_in = min(kTileW, W - g.x0);
    g.gmap = reinterpret_cast<float4*>(tbl.ptr[m]) + (size_t)b * H * W * c4;
    return g;
}
__device__ __forceinline__ void tile_zero(const TileGeom& g, int c4, int tid) {
    for (int r = 0; r < g.rows_in; ++r) {
        float4* dst = g.gmap + ((size_t)(g.y0 + r) * g.W + g.x0) * c4;
        for (int v = tid; v < g.cols_in * c4; v += kTileThreads) __stcs(dst + v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
}

// Persistent grid (kTileCtasPerSm CTAs per SM): every CTA first takes non-empty tiles off the two work lists (heavy
// tiles first, static round robin), then zero-fills its share of the tiles nobody samples.
__global__ void __launch_bounds__(kTileThreads)
roialign_bwd_tile_kernel(const float4* __restrict__ grad_out, const TapWord* __restrict__ taps,
                         const float4* __restrict__ const_partial, GradTable tbl, TileSpace ts, int C, int ph, int pw,
                         const int* __restrict__ count, int* __restrict__ start, const uint32_t* __restrict__ entries,
                         int* __restrict__ misc, const int4* __restrict__ lists) {
    extern __shared__ __align__(16) float4 acc[];                 // [pixel][channel vector]: 32 KB (dynamic)
    __shared__ uint32_t s_key[kTileCap], s_tmp[kTileCap];
    __shared__ TileQueue q;
    __shared__ int s_hist[kTileFBuckets];
    __shared__ int s_n;
    const int NT = ts.base[4], tid = threadIdx.x;
    const int c4 = C >> 2;
    const int n_heavy = misc[kMiscBucket0], n_items = n_heavy + misc[kMiscBucket0 + 1];
    auto item_of = [&](int w) { return __ldg(lists + (w < n_heavy ? (size_t)w : (size_t)NT + (w - n_heavy))); };
    int4 next = make_int4(0, 0, 0, 0);
    if ((int)blockIdx.x < n_items) next = item_of(blockIdx.x);
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int4 item = next;
        if (w + (int)gridDim.x < n_items) next = item_of(w + gridDim.x);     // the next header is in flight during this tile
        const int tile = item.x, n = item.z;
        const TileGeom tg = tile_geom(tbl, ts, tile, c4);
        const int y0 = tg.y0, x0 = tg.x0;
        const uint32_t* seg = entries + item.y;
        // ---- rounds: all samples at once when they fit, else ranges of ROI indices of at most kTileCap samples ----
        uint32_t f_lo = 0u, f_width = 1u;   // histogram bucket of a key: ((key >> 14) - f_lo) / f_width
        bool overflow = false;
        __syncthreads();                    // the previous tile is done with the shared arrays
        if (n > kTileCap) {
            uint32_t mn = 0xffffffffu, mx = 0u;
            for (int i = tid; i < n; i += kTileThreads) { const uint32_t f = seg[i] >> (kKeyXBits + kKeyYBits); mn = min(mn, f); mx = max(mx, f); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
            for (int i = tid; i < kTileFBuckets; i += kTileThreads) s_hist[i] = 0;
            if (tid == 0) { s_key[0] = 0xffffffffu; s_key[1] = 0u; }
            __syncthreads();
            if ((tid & 31) == 0) { atomicMin(&s_key[0], mn); atomicMax(&s_key[1], mx); }
            __syncthreads();
            f_lo = s_key[0];
            f_width = (s_key[1] - f_lo) / kTileFBuckets + 1u;
            __syncthreads();
            for (int i = tid; i < n; i += kTileThreads) atomicAdd(&s_hist[((seg[i] >> (kKeyXBits + kKeyYBits)) - f_lo) / f_width], 1);
            __syncthreads();
            for (int i = tid; i < kTileFBuckets; i += kTileThreads) overflow |= s_hist[i] > kTileCap;
            overflow = __syncthreads_or(overflow);
        }
        if (overflow) {   // (pathological) zero the tile, flag it: the scatter kernel adds this tile's samples atomically
            if (tid == 0) { start[tile] = -1; misc[kMiscOverflow] = 1; }
            tile_zero(tg, c4, tid);
            continue;
        }
        for (int cc = 0; cc < c4; cc += kTileChunk) {   // 256 channels per pass
            const int v = cc + tid;
            const bool have_v = v < c4;
            for (int i = tid; i < kTilePix * kTileChunk; i += kTileThreads) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            int bkt = 0;
            while (bkt < kTileFBuckets) {
                // ---- this round's samples -> s_key, ascending ----
                int nr;
                if (n <= kTileCap) {
                    for (int i = tid; i < n; i += kTileThreads) s_tmp[i] = __ldg(seg + i);
                    nr = n;
                    bkt = kTileFBuckets;
                } else {
                    int b1 = bkt, sum = 0;
                    while (b1 < kTileFBuckets && sum + s_hist[b1] <= kTileCap) sum += s_hist[b1++];   // uniform
                    if (tid == 0) s_n = 0;
                    __syncthreads();
                    for (int i = tid; i < n; i += kTileThreads) {
                        const uint32_t key = __ldg(seg + i);
                        const int kb = (int)(((key >> (kKeyXBits + kKeyYBits)) - f_lo) / f_width);
                        if (kb >= bkt && kb < b1) s_tmp[atomicAdd(&s_n, 1)] = key;
                    }
                    __syncthreads();
                    nr = s_n;
                    bkt = b1;
                }
                __syncthreads();
                if (nr <= 128) {   // the usual case: rank by counting (broadcast loads), one barrier
                    for (int i = tid; i < nr; i += kTileThreads) {
                        const uint32_t x = s_tmp[i];
                        int rank = 0;
                        for (int j = 0; j < nr; ++j) rank += (s_tmp[j] < x);
                        s_key[rank] = x;
                    }
                    __syncthreads();
                } else {
                    const int np2 = 1 << (32 - __clz(nr - 1));
                    for (int i = tid; i < np2; i += kTileThreads) s_key[i] = (i < nr) ? s_tmp[i] : 0xffffffffu;
                    __syncthreads();
                    for (int k = 2; k <= np2; k <<= 1)
                        for (int j = k >> 1; j > 0; j >>= 1) {
                            for (int t = tid; t < (np2 >> 1); t += kTileThreads) {
                                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i | j;
                                const bool up = ((i & k) == 0);
                                const uint32_t a = s_key[i], c = s_key[p];
                                if ((a > c) == up) { s_key[i] = c; s_key[p] = a; }
                            }
                            __syncthreads();
                        }
                }
                // ---- decode kTileQueue samples, then walk them in order ----
                for (int base = 0; base < nr; base += kTileQueue) {
                    const int nq = min(kTileQueue, nr - base);
                    for (int k = tid; k < nq; k += kTileThreads) {
                        const uint32_t key = s_key[base + k];
                        const int f = (int)(key >> (kKeyXBits + kKeyYBits)), y = (int)((key >> kKeyXBits) & ((1u << kKeyYBits) - 1u)),
                                  x = (int)(key & ((1u << kKeyXBits) - 1u));
                        const TapWord* tf = taps + (size_t)f * (ph + pw);
                        const TapWord ty = tf[y], tx = tf[ph + x];
                        const float ly = __uint_as_float(ty.y), lx = __uint_as_float(tx.y);
                        const float wy0 = __fsub_rn(1.0f, ly), wx0 = __fsub_rn(1.0f, lx);
                        const int rt = tap_lo(ty) - y0, rb = tap_hi(ty) - y0, cl = tap_lo(tx) - x0, cr = tap_hi(tx) - x0;
                        const bool t_in = wy0 != 0.0f && (unsigned)rt < (unsigned)kTileH;
                        const bool b_in = ly != 0.0f && (unsigned)rb < (unsigned)kTileH;
                        const bool l_in = wx0 != 0.0f && (unsigned)cl < (unsigned)kTileW;
                        const bool r_in = lx != 0.0f && (unsigned)cr < (unsigned)kTileW;
                        const uint32_t fl = (t_in && l_in ? 1u : 0u) | (t_in && r_in ? 2u : 0u) | (b_in && l_in ? 4u : 0u) |
                                            (b_in && r_in ? 8u : 0u);
                        const bool constant = tap_const(ty);
                        q.off[k] = constant ? f : (f * ph + y) * pw + x;
                        q.pk[k] = (uint32_t)((rt & 3) * kTileW) | ((uint32_t)(cl & 7) << 5) | ((uint32_t)(cr & 7) << 8) |
                                  ((uint32_t)((rb & 3) * kTileW) << 11) | (fl << 16) | (constant ? (1u << 20) : 0u);
                        q.w[k] = make_float4(wy0, ly, wx0, lx);
                    }
                    __syncthreads();
                    auto row_ptr = [&](int i) {
                        const int o = q.off[i];
                        return ((q.pk[i] >> 20) & 1u) ? const_partial + (size_t)o * c4 + v : grad_out + (size_t)o * c4 + v;
                    };
                    float4 val[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (u < nq && have_v) val[u] = __ldg(row_ptr(u));
                    for (int i0 = 0; i0 < nq; i0 += 8) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int i = i0 + u;
                            if (i < nq && have_v) {
                                const uint32_t pk = q.pk[i];
                                const float4 wt = q.w[i];
                                const float4 gv = val[u];
                                float4* at = acc + (size_t)(pk & 31u) * kTileChunk + tid;
                                float4* ab = acc + (size_t)((pk >> 11) & 31u) * kTileChunk + tid;
                                const int cl = (pk >> 5) & 7u, cr = (pk >> 8) & 7u;
                                if (pk & (3u << 16)) {
                                    const float4 d = make_float4(__fmul_rn(wt.x, gv.x), __fmul_rn(wt.x, gv.y),
                                                                 __fmul_rn(wt.x, gv.z), __fmul_rn(wt.x, gv.w));
                                    if (pk & (1u << 16)) acc_corner(at + cl * kTileChunk, d, wt.z);
                                    if (pk & (2u << 16)) acc_corner(at + cr * kTileChunk, d, wt.w);
                                }
                                if (pk & (12u << 16)) {
                                    const float4 d = make_float4(__fmul_rn(wt.y, gv.x), __fmul_rn(wt.y, gv.y),
                                                                 __fmul_rn(wt.y, gv.z), __fmul_rn(wt.y, gv.w));
                                    if (pk & (4u << 16)) acc_corner(ab + cl * kTileChunk, d, wt.z);
                                    if (pk & (8u << 16)) acc_corner(ab + cr * kTileChunk, d, wt.w);
                                }
                            }
                            const int j = i + 8;
                            if (j < nq && have_v) val[u] = __ldg(row_ptr(j));
                        }
                    }
                    __syncthreads();   // the queue is rewritten next
                }
            }
            // every pixel of the tile exactly once
            if (have_v)
                for (int r = 0; r < tg.rows_in; ++r)
                    for (int px = 0; px < tg.cols_in; ++px)
                        __stcs(tg.gmap + ((size_t)(y0 + r) * tg.W + x0 + px) * c4 + v,
                               acc[(size_t)(r * kTileW + px) * kTileChunk + tid]);
            __syncthreads();
        }
    }
    // ---- zero-fill: the tiles nobody samples, 64 candidates per step (one count per thread) ----
    for (int base = blockIdx.x * kTileThreads; base < NT; base += gridDim.x * kTileThreads) {
        const int t = base + tid;
        const bool empty = (t < NT) && count[t] == 0;
        const unsigned vote = __ballot_sync(0xffffffffu, empty);
        __syncthreads();                                   // s_hist is free (previous step / the tile loop)
        if ((tid & 31) == 0) s_hist[tid >> 5] = (int)vote;
        __syncthreads();
        const unsigned long long all = (unsigned long long)(unsigned)s_hist[0] | ((unsigned long long)(unsigned)s_hist[1] << 32);
        for (unsigned long long mm = all; mm; mm &= mm - 1)
            tile_zero(tile_geom(tbl, ts, base + __ffsll((long long)mm) - 1, c4), c4, tid);
    }
}

// workspace: [first-appearance partial list: one int4 per kPrepThreads ROIs][level per ROI]
static int roialign_nparts(int BN) { return (BN + kPrepThreads - 1) / kPrepThreads; }
static size_t roialign_ws_bytes(int B, int N) {
    return align_up((size_t)roialign_nparts(B * N) * sizeof(int4), 256) + align_up((size_t)B * N * sizeof(int32_t), 256);
}

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_roialign_workspace_bytes(int B, int N, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1) return MRCNN_ERR_RANGE;
    *bytes = roialign_ws_bytes(B, N);  // per-ROI level + the four first-appearance indices
    return MRCNN_OK;
}

// MRCNN_ROIALIGN_FWD (read once per process): 0 / unset = LDG kernel (default, the faster one);
// 1 = TMA-staged kernel, 4 warps x 2 stages x 6 CTAs per SM; 2, 3 = other ring shapes kept for the measurement script
static int fwd_variant() {
    static const int v = [] { const char* e = getenv("MRCNN_ROIALIGN_FWD"); return e ? atoi(e) : 0; }();
    return v;
}

static int check_maps(const void* const* maps, const int* H, const int* W, int C) {
    if (!maps || !H || !W) return MRCNN_ERR_NULL;
    if (C < 4 || (C & 3)) return MRCNN_ERR_RANGE;
    for (int l = 0; l < 4; ++l) {
        if (!maps[l]) return MRCNN_ERR_NULL;
        if (H[l] < 1 || W[l] < 1) return MRCNN_ERR_RANGE;
        if (!aligned16(maps[l])) return MRCNN_ERR_ALIGN;
    }
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_roialign_forward(const float* boxes, const float* image_meta, int meta_len,
                                        const float* const* fmaps, const int* H, const int* W, int C, int B, int N,
                                        int ph, int pw, float denominator, int map_mode, float* out, int32_t* roi_map,
                                        int32_t* roi_level, void* ws, size_t ws_bytes, void* stream) {
    if (!boxes || !image_meta || !out || !roi_map || !ws) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || meta_len < 6 || (map_mode != 0 && map_mode != 1) ||
        (long long)B * N * ph > INT_MAX / 2 || !(denominator > 0.0f))
        return MRCNN_ERR_RANGE;
    if (ws_bytes < roialign_ws_bytes(B, N)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(boxes) || !aligned16(out) || !aligned16(ws)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    MapTable tbl;
    for (int l = 0; l < 4; ++l) { tbl.ptr[l] = fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l]; }
    const int BN = B * N;
    const int nparts = roialign_nparts(BN);
    int4* partial = (int4*)ws;
    int32_t* level_ws = (int32_t*)((char*)ws + align_up((size_t)nparts * sizeof(int4), 256));
    cudaError_t e = launch_pdl(roialign_prep_kernel, dim3(nparts), dim3(kPrepThreads), 0, st, (const float4*)boxes,
                               image_meta, BN, denominator, level_ws, partial, roi_level);
    if (e != cudaSuccess) return (int)e;
    const int total_rows = BN * ph;
    // two warps per output row for wide crops on large maps (memory-latency-bound: 14x14 at S=1024 62 -> 58 us at B=8,
    // 236 -> 232 at B=32); 7x7 and the small, cache-resident maps of config 4 (S=512 / 256: 189 -> 195, 150 -> 162 us)
    // are issue-bound and lose with the split
    const int xsplit = (pw >= 12 && (long long)H[0] * W[0] >= 192LL * 192LL) ? 2 : 1;
    const int grid = (total_rows * xsplit + kRoiThreads / 32 - 1) / (kRoiThreads / 32);
#define MRCNN_FWD(V)                                                                                              \
    do {                                                                                                          \
        if (xsplit == 2)                                                                                          \
            e = launch_pdl(roialign_fwd_kernel<V, 2>, dim3(grid), dim3(kRoiThreads), 0, st, (const float4*)boxes,  \
                           level_ws, partial, nparts, map_mode, tbl, C, N, ph, pw, total_rows, out, roi_map);     \
        else                                                                                                      \
            e = launch_pdl(roialign_fwd_kernel<V, 1>, dim3(grid), dim3(kRoiThreads), 0, st, (const float4*)boxes,  \
                           level_ws, partial, nparts, map_mode, tbl, C, N, ph, pw, total_rows, out, roi_map);     \
    } while (0)
    const int variant = fwd_variant();
    if (C == 256 && variant != 0) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define MRCNN_TMA(WARPS, STAGES, CTAS)                                                                             \
        {                                                                                                          \
            const size_t smem = (size_t)WARPS * STAGES * 4096 + (size_t)WARPS * STAGES * 8;                        \
            cudaError_t e2 = cudaFuncSetAttribute(roialign_fwd_tma_kernel<WARPS, STAGES>,                         \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
            if (e2 != cudaSuccess) return (int)e2;                                                                 \
            const int g2 = min((total_rows + WARPS - 1) / WARPS, CTAS * sms);                                      \
            e = launch_pdl(roialign_fwd_tma_kernel<WARPS, STAGES>, dim3(g2), dim3(WARPS * 32), smem, st,           \
                           (const float4*)boxes, level_ws, partial, nparts, map_mode, tbl, N, ph, pw, total_rows,  \
                           out, roi_map);                                                                          \
        }
        // measured on B200, config 2 (profiles/r1_roialign_fwd_tma.md): 4 warps x 2 stages x 6 CTAs/SM is the best
        // of the shapes tried and ties the LDG kernel at 7x7 (138.8 vs 140.0 us) but loses at 14x14 (103 vs 59 us)
        if (variant == 2) MRCNN_TMA(8, 3, 2)
        else if (variant == 3) MRCNN_TMA(4, 3, 4)
        else MRCNN_TMA(4, 2, 6)
#undef MRCNN_TMA
        return e != cudaSuccess ? (int)e : last_error();
    }
    if (C == 128) MRCNN_FWD(1);
    else if (C == 256) MRCNN_FWD(2);
    else if (C == 512) MRCNN_FWD(4);
    else MRCNN_FWD(0);
#undef MRCNN_FWD
    return e != cudaSuccess ? (int)e : last_error();
}

static int pixel_space(const int* H, const int* W, int B, PixelSpace* ps) {
    long long acc = 0;
    for (int l = 0; l < 4; ++l) { ps->base[l] = (int)acc; acc += (long long)B * H[l] * W[l]; }
    if (acc > INT_MAX / 2) return MRCNN_ERR_RANGE;
    ps->base[4] = (int)acc;
    return MRCNN_OK;
}

static int tile_space(const int* H, const int* W, int B, TileSpace* ts) {
    long long acc = 0;
    for (int l = 0; l < 4; ++l) {
        ts->ty[l] = (H[l] + kTileH - 1) / kTileH;
        ts->tx[l] = (W[l] + kTileW - 1) / kTileW;
        if (ts->ty[l] > 128 || ts->tx[l] > 64) return MRCNN_ERR_RANGE;   // the tile masks: 128 rows, 64 columns (512 x 512 maps)
        ts->base[l] = (int)acc;
        acc += (long long)B * ts->ty[l] * ts->tx[l];
    }
    if (acc > INT_MAX / 8) return MRCNN_ERR_RANGE;
    ts->base[4] = (int)acc;
    return MRCNN_OK;
}

// workspace of the deterministic backward:
// [count NT | cursor NT | misc] (zeroed per call) [start NT] [work lists 4 NT] [taps BN (ph + pw)] [(tile, sample) keys:
// 4 per sample] [pre-reduced rows of the zero-size ROIs: BN x C floats]
struct BwdWsLayout {
    size_t zeroed, start, lists, taps, entries, partial, total;
};
static BwdWsLayout roialign_bwd_ws_layout(const TileSpace& ts, int B, int N, int ph, int pw, int C) {
    const size_t NT = (size_t)ts.base[4], BN = (size_t)B * N;
    BwdWsLayout w;
    w.zeroed = align_up((2 * NT + kMiscWords) * sizeof(int), 256);
    w.start = w.zeroed;
    w.lists = w.start + align_up(NT * sizeof(int), 256);
    w.taps = w.lists + align_up((size_t)kTileBuckets * NT * sizeof(int4), 256);
    w.entries = w.taps + align_up(BN * (size_t)(ph + pw) * sizeof(TapWord), 256);
    w.partial = w.entries + align_up(BN * (size_t)ph * pw * 4 * sizeof(uint32_t), 256);
    w.total = w.partial + align_up(BN * (size_t)C * sizeof(float), 256);
    return w;
}

MRCNN_EXPORT int mrcnn_roialign_backward_workspace_bytes(int B, int N, int ph, int pw, const int* H, const int* W,
                                                         int C, size_t* bytes) {
    if (!bytes || !H || !W) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || C < 4 || (C & 3) || C > 8192) return MRCNN_ERR_RANGE;
    for (int l = 0; l < 4; ++l)
        if (H[l] < 1 || W[l] < 1) return MRCNN_ERR_RANGE;
    TileSpace ts;
    const int rc = tile_space(H, W, B, &ts);
    if (rc != MRCNN_OK) return rc;
    if ((long long)B * N * ph * pw >= (1LL << 29) || ph > (1 << kKeyYBits) || pw > (1 << kKeyXBits) ||
        (long long)B * N >= (1LL << kKeyFBits))
        return MRCNN_ERR_RANGE;
    *bytes = roialign_bwd_ws_layout(ts, B, N, ph, pw, C).total;
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_roialign_backward(const float* grad_out, const float* boxes, const int32_t* roi_map,
                                         float* const* grad_fmaps, const int* H, const int* W, int C, int B, int N,
                                         int ph, int pw, void* ws, size_t ws_bytes, void* stream) {
    if (!grad_out || !boxes || !roi_map) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)grad_fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || (long long)B * N * ph > INT_MAX / 2 || C > 8192) return MRCNN_ERR_RANGE;
    if (!aligned16(boxes) || !aligned16(grad_out)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    GradTable tbl;
    for (int l = 0; l < 4; ++l) { tbl.ptr[l] = grad_fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l]; }
    const int groups = (ph + kRoiThreads / 32 - 1) / (kRoiThreads / 32);
    const int rows_per_group = (ph + groups - 1) / groups;  // 7x7 -> 1 x 7 rows, 14x14 -> 2 x 7, 28x28 -> 4 x 7
    const size_t smem = (size_t)(kRoiThreads / 32) * C * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(roialign_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    TileSpace ts{};
    if (ws == nullptr) {  // atomic mode: zero-fill, then scatter every sample with vector reductions
        for (int l = 0; l < 4; ++l) {
            cudaError_t e = cudaMemsetAsync(grad_fmaps[l], 0, (size_t)B * H[l] * W[l] * C * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
        }
        roialign_bwd_kernel<<<B * N * groups, kRoiThreads, smem, st>>>((const float4*)grad_out, (const float4*)boxes,
                                                                      roi_map, tbl, C, N, ph, pw, groups, rows_per_group,
                                                                      ts, nullptr, nullptr);
        return last_error();
    }
    // deterministic mode
    if ((long long)B * N * ph * pw >= (1LL << 29) || ph > (1 << kKeyYBits) || pw > (1 << kKeyXBits) ||
        (long long)B * N >= (1LL << kKeyFBits))
        return MRCNN_ERR_RANGE;
    rc = tile_space(H, W, B, &ts);
    if (rc != MRCNN_OK) return rc;
    if (!aligned16(ws)) return MRCNN_ERR_ALIGN;
    const BwdWsLayout lay = roialign_bwd_ws_layout(ts, B, N, ph, pw, C);
    if (ws_bytes < lay.total) return MRCNN_ERR_WORKSPACE;
    const int NT = ts.base[4], BN = B * N, bins = BN * ph * pw;
    int* count = (int*)ws;
    int* cursor = count + NT;
    int* misc = count + 2 * (size_t)NT;
    int* start = (int*)((char*)ws + lay.start);
    int4* lists = (int4*)((char*)ws + lay.lists);
    TapWord* taps = (TapWord*)((char*)ws + lay.taps);
    uint32_t* entries = (uint32_t*)((char*)ws + lay.entries);
    float4* partial = (float4*)((char*)ws + lay.partial);
    cudaError_t e = cudaMemsetAsync(ws, 0, lay.zeroed, st);
    if (e != cudaSuccess) return (int)e;
    roialign_bwd_taps_kernel<<<(BN * (ph + pw) + 255) / 256, 256, 0, st>>>((const float4*)boxes, roi_map, tbl, BN, ph, pw,
                                                                           taps);
    roialign_bwd_const_kernel<<<BN, 64, 0, st>>>((const float4*)grad_out, taps, C, ph, pw, partial);
    roialign_bwd_bin_kernel<0><<<(bins + 255) / 256, 256, 0, st>>>(taps, roi_map, ts, N, ph, pw, bins, count, start,
                                                                  cursor, entries);
    roialign_bwd_alloc_kernel<<<(NT + 255) / 256, 256, 0, st>>>(count, NT, start, misc, lists);
    roialign_bwd_bin_kernel<1><<<(bins + 255) / 256, 256, 0, st>>>(taps, roi_map, ts, N, ph, pw, bins, count, start,
                                                                  cursor, entries);
    const size_t tile_smem = (size_t)kTilePix * kTileChunk * sizeof(float4);
    e = cudaFuncSetAttribute(roialign_bwd_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem);
    if (e != cudaSuccess) return (int)e;
    const int tile_grid = min(2 * NT, device_props().sms * kTileCtasPerSm);
    roialign_bwd_tile_kernel<<<tile_grid, kTileThreads, tile_smem, st>>>((const float4*)grad_out, taps, partial, tbl, ts, C, ph, pw,
                                                              count, start, entries, misc, lists);
    roialign_bwd_kernel<<<B * N * groups, kRoiThreads, smem, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map,
                                                                  tbl, C, N, ph, pw, groups, rows_per_group, ts, start,
                                                                  misc + kMiscOverflow);
    return last_error();
}


// ---------------------------------------------------------------------------------------------------------------
// Demand-driven host -> device staging of the feature maps (maps resident in PINNED HOST memory).
// PyramidROIAlign touches only the pixels its ROIs sample (about half of the 89 MB per image at 1024^2 for 1000
// proposals); when the maps start in host memory, copying all of them is the whole cost of the stage (PCIe).  Instead:
//   roialign_mark_kernel   thread per output bin: the same taps as the forward kernel -> one bit per sampled map pixel
//                          (skipping pixels that are already resident);
//   roialign_fetch_kernel  warp per 32-pixel bitmap word: every marked pixel (C floats, 1 KB at C = 256) is read ONCE
//                          straight out of the pinned host map (zero-copy loads over PCIe / C2C, 128-bit per lane,
//                          several pixels in flight per warp) and written to the device staging map; the word is then
//                          merged into the `resident` bitmap, so a later call (the mask branch's ROIAlign on the
//                          detections) only fetches what is still missing.
// The ordinary forward kernel then runs on the staging maps, bit-identical to a full copy.
// ---------------------------------------------------------------------------------------------------------------
namespace mrcnn {

__global__ void __launch_bounds__(256)
roialign_mark_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                     const int4* __restrict__ partial, int nparts, int map_mode, MapTable tbl, PixelSpace ps, int N, int ph,
                     int pw, int total_bins, const uint32_t* __restrict__ resident, uint32_t* __restrict__ need) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    const int4 fa = first_appearance(partial, nparts);  // all lanes, before the exit
    if (s >= total_bins) return;
    const int f = s / (ph * pw), r = s - f * (ph * pw), y = r / pw, x = r - y * pw;
    const int m = roi_map_index(level_ws[f], fa, map_mode);
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const AxisTap ty = axis_tap(g.y0, g.hs, y, g.H), tx = axis_tap(g.x0, g.ws, x, g.W);
    if (!(ty.valid && tx.valid)) return;  // the forward kernel writes zeros without reading
    const int base = ((m == 0) ? ps.base[0] : (m == 1) ? ps.base[1] : (m == 2) ? ps.base[2] : ps.base[3]) +
                     (f / N) * g.H * g.W;
    const int q[4] = {base + ty.lo * g.W + tx.lo, base + ty.lo * g.W + tx.hi, base + ty.hi * g.W + tx.lo,
                      base + ty.hi * g.W + tx.hi};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c == 1 && q[1] == q[0]) continue;
        if (c == 2 && q[2] == q[0]) continue;
        if (c == 3 && (q[3] == q[1] || q[3] == q[2])) continue;
        const uint32_t bit = 1u << (q[c] & 31);
        const int w = q[c] >> 5;
        if ((__ldg(resident + w) | need[w]) & bit) continue;  // racy pre-check: saves most of the atomics
        atomicOr(need + w, bit);
    }
}

// pixel q of the concatenated pixel space -> map index and offset (in pixels) inside that map
__device__ __forceinline__ int pixel_map_of(const PixelSpace& ps, int q) {
    return (q >= ps.base[3]) ? 3 : (q >= ps.base[2]) ? 2 : (q >= ps.base[1]) ? 1 : 0;
}

template <int VPL>
__global__ void __launch_bounds__(256)
roialign_fetch_kernel(MapTable host, GradTable dev, PixelSpace ps, int C, int words, uint32_t* __restrict__ need,
                      uint32_t* __restrict__ resident, unsigned long long* __restrict__ fetched) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = (gridDim.x * 256) >> 5;
    const int c4 = C >> 2;
    unsigned long long mine = 0;
    for (int w = warp; w < words; w += nwarps) {
        uint32_t bits = need[w] & ~resident[w];
        if (bits == 0u) continue;
        mine += __popc(bits);
        const uint32_t all = bits;
        while (bits) {  // two pixels per round: 2 * VPL independent 16-byte host loads per lane in flight
            const int i0 = __ffs(bits) - 1;
            bits &= bits - 1;
            const int i1 = bits ? __ffs(bits) - 1 : -1;
            if (bits) bits &= bits - 1;
            const int q0 = w * 32 + i0, q1 = w * 32 + (i1 < 0 ? i0 : i1);
            const int m0 = pixel_map_of(ps, q0), m1 = pixel_map_of(ps, q1);
            const size_t o0 = (size_t)(q0 - ps.base[m0]) * c4, o1 = (size_t)(q1 - ps.base[m1]) * c4;
            const float4* h0 = reinterpret_cast<const float4*>(host.ptr[m0]) + o0;
            const float4* h1 = reinterpret_cast<const float4*>(host.ptr[m1]) + o1;
            float4* d0 = reinterpret_cast<float4*>(dev.ptr[m0]) + o0;
            float4* d1 = reinterpret_cast<float4*>(dev.ptr[m1]) + o1;
            if (VPL > 0) {
                float4 a[VPL > 0 ? VPL : 1], b[VPL > 0 ? VPL : 1];
#pragma unroll
                for (int v = 0; v < VPL; ++v) { a[v] = __ldcs(h0 + lane + 32 * v); b[v] = __ldcs(h1 + lane + 32 * v); }
#pragma unroll
                for (int v = 0; v < VPL; ++v) { d0[lane + 32 * v] = a[v]; if (i1 >= 0) d1[lane + 32 * v] = b[v]; }
            } else {
                for (int i = lane; i < c4; i += 32) { d0[i] = __ldcs(h0 + i); if (i1 >= 0) d1[i] = __ldcs(h1 + i); }
            }
        }
        if (lane == 0) { resident[w] |= all; need[w] = 0u; }  // one warp owns a word: no atomics needed
    }
    if (fetched && lane == 0 && mine) atomicAdd(fetched, mine);
}

static size_t fetch_ws_bytes(int B, int N, int NP) {
    return roialign_ws_bytes(B, N) + align_up((size_t)((NP + 31) / 32) * sizeof(uint32_t), 256);
}

}  // namespace mrcnn

MRCNN_EXPORT int mrcnn_roialign_resident_words(int B, const int* H, const int* W, size_t* words) {
    if (!H || !W || !words) return MRCNN_ERR_NULL;
    PixelSpace ps;
    if (B < 1) return MRCNN_ERR_RANGE;
    const int rc = pixel_space(H, W, B, &ps);
    if (rc != MRCNN_OK) return rc;
    *words = (size_t)((ps.base[4] + 31) / 32);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_roialign_fetch_workspace_bytes(int B, int N, const int* H, const int* W, size_t* bytes) {
    if (!H || !W || !bytes) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1) return MRCNN_ERR_RANGE;
    PixelSpace ps;
    const int rc = pixel_space(H, W, B, &ps);
    if (rc != MRCNN_OK) return rc;
    *bytes = fetch_ws_bytes(B, N, ps.base[4]);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_roialign_fetch_hostmaps(const float* boxes, const float* image_meta, int meta_len,
                                               const float* const* host_fmaps, float* const* dev_fmaps, const int* H,
                                               const int* W, int C, int B, int N, int ph, int pw, float denominator,
                                               int map_mode, uint32_t* resident, int reset,
                                               unsigned long long* fetched, void* ws, size_t ws_bytes, void* stream) {
    if (!boxes || !image_meta || !resident || !ws) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)host_fmaps, H, W, C);
    if (rc == MRCNN_OK) rc = check_maps((const void* const*)dev_fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || meta_len < 6 || (map_mode != 0 && map_mode != 1) ||
        (long long)B * N * ph * pw > INT_MAX / 2 || !(denominator > 0.0f))
        return MRCNN_ERR_RANGE;
    PixelSpace ps;
    rc = pixel_space(H, W, B, &ps);
    if (rc != MRCNN_OK) return rc;
    if (ws_bytes < fetch_ws_bytes(B, N, ps.base[4])) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(boxes) || !aligned16(ws)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    MapTable host;
    GradTable dev;
    for (int l = 0; l < 4; ++l) {
        host.ptr[l] = host_fmaps[l]; host.H[l] = H[l]; host.W[l] = W[l];
        dev.ptr[l] = dev_fmaps[l]; dev.H[l] = H[l]; dev.W[l] = W[l];
    }
    const int BN = B * N, words = (ps.base[4] + 31) / 32;
    const int nparts = roialign_nparts(BN);
    int4* partial = (int4*)ws;
    int32_t* level_ws = (int32_t*)((char*)ws + align_up((size_t)nparts * sizeof(int4), 256));
    uint32_t* need = (uint32_t*)((char*)ws + roialign_ws_bytes(B, N));
    cudaError_t e = cudaMemsetAsync(need, 0, (size_t)words * sizeof(uint32_t), st);
    if (e == cudaSuccess && reset) e = cudaMemsetAsync(resident, 0, (size_t)words * sizeof(uint32_t), st);
    if (e != cudaSuccess) return (int)e;
    roialign_prep_kernel<<<nparts, kPrepThreads, 0, st>>>((const float4*)boxes, image_meta, BN, denominator, level_ws,
                                                         partial, nullptr);
    const int bins = BN * ph * pw;
    roialign_mark_kernel<<<(bins + 255) / 256, 256, 0, st>>>((const float4*)boxes, level_ws, partial, nparts, map_mode,
                                                            host, ps, N, ph, pw, bins, resident, need);
    int devid = 0, sms = 148;
    if (cudaGetDevice(&devid) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid);
    const int grid = min((words + 7) / 8, 8 * sms);
#define MRCNN_FETCH(V) roialign_fetch_kernel<V><<<grid, 256, 0, st>>>(host, dev, ps, C, words, need, resident, fetched)
    if (C == 128) MRCNN_FETCH(1);
    else if (C == 256) MRCNN_FETCH(2);
    else if (C == 512) MRCNN_FETCH(4);
    else MRCNN_FETCH(0);
#undef MRCNN_FETCH
    return last_error();
}
