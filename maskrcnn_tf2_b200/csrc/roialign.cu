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
// Occupancy: 5 CTAs/SM (48 registers) for narrow crops, 6 (40 registers, 16 bytes of spill) for wide ones -- the 14x14
// kernel is bound by the serial load -> lerp -> store chain of a warp's bins, so more resident warps pay (58.4 -> 56.4 us
// per layer call; 8 CTAs/SM with 32 registers: 66.5 us; the L2-bound 7x7 kernel loses at 6: 140.3 -> 142.2 us).
template <int VPL, int XSPLIT>
__global__ void __launch_bounds__(kRoiThreads, XSPLIT > 1 ? 6 : 5)
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
        // TF's three lerps per element, every subtraction, product and sum individually rounded.
        auto lerp1 = [&](float a, float b, float c, float d) {
            const float t = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), lx));
            const float u = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), lx));
            return __fadd_rn(t, __fmul_rn(__fsub_rn(u, t), ly));
        };
        // Wide crops (XSPLIT == 2: 14x14 and up, L1 hit rate 66 %, issue slots 56 % busy) run the subtractions and sums on
        // the f32x2 pipe (FADD2: two elements per instruction, same IEEE rounding); the products stay scalar -- ptxas
        // contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which would change the rounding.  Measured on B200: 14x14
        // 60.4 -> 58.4 us; the memory-bound 7x7 kernel loses with it (140.3 -> 144.4 us) and keeps the scalar form.
        auto lerp2 = [&](unsigned long long a, unsigned long long b, float w) {   // a + (b - a) * w on two elements
            float d0, d1;
            unpack_f2(sub_f2(b, a), d0, d1);
            return add_f2(a, pack_f2(__fmul_rn(d0, w), __fmul_rn(d1, w)));
        };
        auto lerp4 = [&](const float4& a, const float4& b, const float4& c, const float4& d) {
            if (XSPLIT == 1)
                return make_float4(lerp1(a.x, b.x, c.x, d.x), lerp1(a.y, b.y, c.y, d.y), lerp1(a.z, b.z, c.z, d.z),
                                   lerp1(a.w, b.w, c.w, d.w));
            const unsigned long long t0 = lerp2(pack_f2(a.x, a.y), pack_f2(b.x, b.y), lx);
            const unsigned long long t1 = lerp2(pack_f2(a.z, a.w), pack_f2(b.z, b.w), lx);
            const unsigned long long u0 = lerp2(pack_f2(c.x, c.y), pack_f2(d.x, d.y), lx);
            const unsigned long long u1 = lerp2(pack_f2(c.z, c.w), pack_f2(d.z, d.w), lx);
            float4 o;
            unpack_f2(lerp2(t0, u0, ly), o.x, o.y);
            unpack_f2(lerp2(t1, u1, ly), o.z, o.w);
            return o;
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
// One CTA per (ROI, group of <= 8 output rows), one warp per row, lanes over channels.  Contention control:
//   * a corner whose weight is exactly 0 is not touched (finite gradients: adding 0 changes nothing);
//   * bins of a row that hit the same pixels (zero width scale) are summed in registers before one reduction;
//   * ROIs whose taps are constant in both axes (zero-padded ROIs: every bin samples pixel (0,0), quirk Q5) are
//     summed across the CTA's rows in shared memory and issue ONE reduction per CTA and corner, which keeps the
//     thousands of padded rows of a training batch from serialising on a single L2 line.
constexpr int kLightMax = 32;       // samples a pixel may collect and still be summed, in TF's order, by one warp
constexpr int kMediumMax = 1024;    // ... or, thread per channel, by one CTA, in ONE round; more samples: several rounds
constexpr int kHeavyMax = 1 << 20;  // pixels above this (never seen) go to the atomic fallback
constexpr int kHeavyBuckets = 1024; // sample-index histogram of a pixel with more than kMediumMax samples
constexpr int kMediumCtas = 296;    // CTAs at the head of the gather grid that walk the list of such pixels

struct PixelSpace {   // global pixel id = base[m] + (b * H[m] + y) * W[m] + x ; base[4] = number of pixels
    int base[5];
};

// `count` == nullptr: every corner is scattered (atomic mode).  Otherwise only corners whose pixel was left to the
// atomic fallback (more than kLightMax samples, or a constant-tap ROI) are; the gather kernel owns the rest.
// (cursor[q] < 0: the gather kernel handed the pixel over -- one bucket of its sample histogram overflowed a round)
__device__ __forceinline__ bool corner_is_ours(const uint32_t* __restrict__ count, int q) {
    return count == nullptr || __ldg(count + q) > (uint32_t)kHeavyMax || (__ldg(count + q) & 0x40000000u) != 0u;
}

__device__ __forceinline__ void scatter_corners(float* gimg, int C, const uint32_t* __restrict__ count, int qbase,
                                                int ptl, int ptr_, int pbl, int pbr, int i, const float4& v,
                                                float wy0, float ly, float wx0, float lx) {
    const float4 dt = make_float4(__fmul_rn(wy0, v.x), __fmul_rn(wy0, v.y), __fmul_rn(wy0, v.z), __fmul_rn(wy0, v.w));
    const float4 db = make_float4(__fmul_rn(ly, v.x), __fmul_rn(ly, v.y), __fmul_rn(ly, v.z), __fmul_rn(ly, v.w));
    if (wy0 != 0.0f && wx0 != 0.0f && corner_is_ours(count, qbase + ptl))
        red_add_v4(gimg + (size_t)ptl * C + 4 * i, __fmul_rn(wx0, dt.x), __fmul_rn(wx0, dt.y), __fmul_rn(wx0, dt.z), __fmul_rn(wx0, dt.w));
    if (wy0 != 0.0f && lx != 0.0f && corner_is_ours(count, qbase + ptr_))
        red_add_v4(gimg + (size_t)ptr_ * C + 4 * i, __fmul_rn(lx, dt.x), __fmul_rn(lx, dt.y), __fmul_rn(lx, dt.z), __fmul_rn(lx, dt.w));
    if (ly != 0.0f && wx0 != 0.0f && corner_is_ours(count, qbase + pbl))
        red_add_v4(gimg + (size_t)pbl * C + 4 * i, __fmul_rn(wx0, db.x), __fmul_rn(wx0, db.y), __fmul_rn(wx0, db.z), __fmul_rn(wx0, db.w));
    if (ly != 0.0f && lx != 0.0f && corner_is_ours(count, qbase + pbr))
        red_add_v4(gimg + (size_t)pbr * C + 4 * i, __fmul_rn(lx, db.x), __fmul_rn(lx, db.y), __fmul_rn(lx, db.z), __fmul_rn(lx, db.w));
}

__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ boxes,
                    const int32_t* __restrict__ roi_map, GradTable tbl, int C, int N, int ph, int pw, int groups,
                    int rows_per_group, PixelSpace ps, const uint32_t* __restrict__ count,
                    const int* __restrict__ heavy_normal) {
    extern __shared__ __align__(16) float4 s_acc[];  // [8 warps][C/4], constant-tap ROIs only
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // gather mode: nothing to do unless some pixel was left to the fallback (more than kHeavyMax samples, or more than
    // kMediumMax of them inside one bucket of its sample-index histogram)
    if (count != nullptr && *heavy_normal == 0) return;
    const int f = blockIdx.x / groups, grp = blockIdx.x - f * groups;
    const int y = grp * rows_per_group + warp;
    const bool has_row = (warp < rows_per_group) && (y < ph);
    const int m = roi_map[f];
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const int c4 = C >> 2;
    const int b = f / N;
    float* gimg = ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
                  (size_t)b * g.H * g.W * C;
    const int qbase = ps.base[m] + b * g.H * g.W;
    const bool constant = (g.hs == 0.0f && g.ws == 0.0f);  // CTA-uniform
    const AxisTap ty = axis_tap(g.y0, g.hs, has_row ? y : 0, g.H);
    const float4* gr = grad_out + ((size_t)f * ph + (has_row ? y : 0)) * pw * c4;
    const int top = ty.lo * g.W, bot = ty.hi * g.W;
    const float ly = ty.lerp, wy0 = __fsub_rn(1.0f, ty.lerp);
    if (constant) {  // its pixels carry kConstFlag, so in gather mode they are always ours
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
            scatter_corners(gimg, C, count, qbase, top + tx.lo, top + tx.hi, bot + tx.lo, bot + tx.hi, i, acc, wy0, ly,
                            wx0, lx);
        }
        return;
    }
    if (!has_row || !ty.valid) return;
    if (g.ws == 0.0f && count == nullptr) {  // every bin of the row hits the same pixels: sum the row first
        const AxisTap tx = axis_tap(g.x0, g.ws, 0, g.W);
        if (!tx.valid) return;
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int x = 0; x < pw; ++x) {
                const float4 v = __ldcs(gr + (size_t)x * c4 + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            scatter_corners(gimg, C, nullptr, qbase, top + tx.lo, top + tx.hi, bot + tx.lo, bot + tx.hi, i, acc, wy0, ly,
                            wx0, lx);
        }
        return;
    }
    for (int x = 0; x < pw; ++x, gr += c4) {
        const AxisTap tx = axis_tap(g.x0, g.ws, x, g.W);
        if (!tx.valid) continue;
        const int ptl = top + tx.lo, ptr_ = top + tx.hi, pbl = bot + tx.lo, pbr = bot + tx.hi;
        if (count != nullptr && !(corner_is_ours(count, qbase + ptl) || corner_is_ours(count, qbase + ptr_) ||
                                  corner_is_ours(count, qbase + pbl) || corner_is_ours(count, qbase + pbr)))
            continue;  // warp-uniform: no gradient row is read for bins the gather kernel owns entirely
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32)
            scatter_corners(gimg, C, count, qbase, ptl, ptr_, pbl, pbr, i, __ldcs(gr + i), wy0, ly, wx0, lx);
    }
}

// ---- deterministic gather backward -------------------------------------------------------------------------
// One thread per output bin (f, y, x): the up-to-four gradient-map pixels its gradient lands on.
struct BinTaps {
    int q[4];      // global pixel ids: tl, tr, bl, br
    bool on[4];    // weight != 0 and the forward pass did not extrapolate
    bool constant; // zero-size ROI: every bin of the ROI samples the same pixels
    float wy[2], wx[2];  // (1 - ly, ly), (1 - lx, lx): corner c weighs wx[c & 1] * (wy[c >> 1] * g)
};

__device__ __forceinline__ BinTaps bin_taps(const float4* __restrict__ boxes, const int32_t* __restrict__ roi_map,
                                            const GradTable& tbl, const PixelSpace& ps, int N, int ph, int pw, int s) {
    const int bins = ph * pw;
    const int f = s / bins, r = s - f * bins, y = r / pw, x = r - y * pw;
    const int m = roi_map[f];
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const AxisTap ty = axis_tap(g.y0, g.hs, y, g.H), tx = axis_tap(g.x0, g.ws, x, g.W);
    const int qbase = ps.base[m] + (f / N) * g.H * g.W;
    const float wy0 = __fsub_rn(1.0f, ty.lerp), wx0 = __fsub_rn(1.0f, tx.lerp);
    const bool valid = ty.valid && tx.valid;
    BinTaps t;
    t.wy[0] = wy0; t.wy[1] = ty.lerp; t.wx[0] = wx0; t.wx[1] = tx.lerp;
    t.constant = (g.hs == 0.0f && g.ws == 0.0f);
    t.q[0] = qbase + ty.lo * g.W + tx.lo; t.on[0] = valid && wy0 != 0.0f && wx0 != 0.0f;
    t.q[1] = qbase + ty.lo * g.W + tx.hi; t.on[1] = valid && wy0 != 0.0f && tx.lerp != 0.0f;
    t.q[2] = qbase + ty.hi * g.W + tx.lo; t.on[2] = valid && ty.lerp != 0.0f && wx0 != 0.0f;
    t.q[3] = qbase + ty.hi * g.W + tx.hi; t.on[3] = valid && ty.lerp != 0.0f && tx.lerp != 0.0f;
    return t;
}

__global__ void __launch_bounds__(256)
roialign_bwd_count_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ roi_map, GradTable tbl,
                          PixelSpace ps, int N, int ph, int pw, int total_bins, uint32_t* __restrict__ count,
                          int* __restrict__ misc) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= total_bins) return;
    const BinTaps t = bin_taps(boxes, roi_map, tbl, ps, N, ph, pw, s);
    if (t.constant && s % (ph * pw) != 0) return;  // a zero-size ROI is ONE sample: its pre-reduced gradient row
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (t.on[c]) atomicAdd(count + t.q[c], 1u);
}

// Four pixels per thread: a segment of the key list for every pixel the gather kernel will sum (1..kMediumMax samples);
// pixels above kLightMax are also appended to the medium list.  Segment order is whatever the atomics give; only the
// order INSIDE a segment matters and that is sorted later.  misc: [0] key-list bump pointer, [1] "the atomic fallback
// has work" (a constant-tap ROI, or ordinary samples on a pixel above kMediumMax), [2] length of the medium list.
__global__ void __launch_bounds__(256)
roialign_bwd_alloc_kernel(const uint32_t* __restrict__ count, int NP, int* __restrict__ start, int* __restrict__ misc,
                          int* __restrict__ medium) {
    __shared__ int warp_sums[32];
    __shared__ int block_total, block_base;
    const int q0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    uint32_t c[4];
    int n[4], mine = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c[i] = (q0 + i < NP) ? count[q0 + i] : 0u;
        n[i] = (c[i] <= (uint32_t)kHeavyMax) ? (int)c[i] : 0;
        mine += n[i];
        if (c[i] > (uint32_t)kHeavyMax) misc[1] = 1;
        if (n[i] > kLightMax) medium[atomicAdd(&misc[2], 1)] = q0 + i;
    }
    int off = block_exclusive_scan(mine, warp_sums, &block_total);
    if (threadIdx.x == 0) block_base = (block_total > 0) ? atomicAdd(&misc[0], block_total) : 0;
    __syncthreads();
    off += block_base;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (q0 + i < NP) start[q0 + i] = off;
        off += n[i];
    }
}

__global__ void __launch_bounds__(256)
roialign_bwd_fill_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ roi_map, GradTable tbl,
                         PixelSpace ps, int N, int ph, int pw, int total_bins, const uint32_t* __restrict__ count,
                         const int* __restrict__ start, int* __restrict__ cursor, int4* __restrict__ entries) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= total_bins) return;
    const BinTaps t = bin_taps(boxes, roi_map, tbl, ps, N, ph, pw, s);
    if (t.constant && s % (ph * pw) != 0) return;
    const int fetch = t.constant ? total_bins + s / (ph * pw) : s;   // rows >= total_bins: the pre-reduced rows, by ROI
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (t.on[c] && count[t.q[c]] <= (uint32_t)kHeavyMax)
            entries[start[t.q[c]] + atomicAdd(cursor + t.q[c], 1)] =  // key order = TF's accumulation order
                make_int4((s << 2) | c, __float_as_int(t.wy[c >> 1]), __float_as_int(t.wx[c & 1]), fetch);
}

// zero-size ROIs (zero-padded target rows, quirk Q5): all ph*pw bins sample the same four taps.  Thousands of them can
// sit on one pixel, and that many rows cannot be added one after the other at any speed, so each such ROI is reduced
// to ONE row first (its bins in order, fixed) and enters the pixel sums as one sample at the position of its first bin:
// deterministic, and bit-identical to the sequential order everywhere except on the pixels under such ROIs.
__global__ void __launch_bounds__(64)
roialign_bwd_const_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ boxes,
                          const int32_t* __restrict__ roi_map, GradTable tbl, int C, int ph, int pw,
                          float4* __restrict__ partial /*[BN][C/4]*/) {
    const int f = blockIdx.x;
    const RoiGeom g = roi_geom(__ldg(boxes + f), roi_map[f], tbl.H, tbl.W, ph, pw);
    if (!(g.hs == 0.0f && g.ws == 0.0f)) return;
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

__device__ __forceinline__ int warp_sort_asc(int v, int lane) {  // bitonic network over the 32 lanes
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, v, j);
            const bool up = ((lane & k) == 0), lower = ((lane & j) == 0);
            v = (lower == up) ? min(v, o) : max(v, o);
        }
    return v;
}

__device__ __forceinline__ void fma_in_order(float4& acc, const float4& v, float a, float b) {
    acc.x = __fadd_rn(acc.x, __fmul_rn(b, __fmul_rn(a, v.x)));  // acc += wx * (wy * g): TF's two products, then the add
    acc.y = __fadd_rn(acc.y, __fmul_rn(b, __fmul_rn(a, v.y)));
    acc.z = __fadd_rn(acc.z, __fmul_rn(b, __fmul_rn(a, v.z)));
    acc.w = __fadd_rn(acc.w, __fmul_rn(b, __fmul_rn(a, v.w)));
}

__device__ __forceinline__ int c4_of(int C) { return C >> 2; }
__device__ __forceinline__ float4* pixel_ptr(const GradTable& tbl, const PixelSpace& ps, int q, int c4) {
    const int m = (q >= ps.base[3]) ? 3 : (q >= ps.base[2]) ? 2 : (q >= ps.base[1]) ? 1 : 0;
    return reinterpret_cast<float4*>((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
           (size_t)(q - ps.base[m]) * c4;
}

// The gather pass.  Every pixel of every gradient map is written exactly once.  An entry is (key, wy, wx): sample
// s = key >> 2 (= its row of grad_out), corner = key & 3; ascending key = TF's accumulation order.
//   * CTAs [kMediumCtas, grid): persistent warps stream over the pixels -- zeros, or (<= kLightMax samples) the sum
//     in TF's order: each lane ranks its entry among the pixel's (n shuffles), then the rows are fetched two at a
//     time and added strictly in rank order;
//   * CTAs [0, kMediumCtas): walk the medium list, one CTA per pixel: (key, slot) pairs sorted in shared memory, then
//     one thread per channel adds the rows strictly in order, sixteen loads in flight.  They sit at the head of the
//     grid so these long sums start first and overlap the streaming pass.
template <int VPL>
__global__ void __launch_bounds__(256, 4)
roialign_bwd_gather_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ const_partial, int total_bins,
                           GradTable tbl, PixelSpace ps, int C, uint32_t* __restrict__ count,
                           const int* __restrict__ start, const int4* __restrict__ entries, int* __restrict__ misc,
                           const int* __restrict__ medium) {
    constexpr int V = VPL > 0 ? VPL : 1;
    __shared__ uint64_t s_sort[kMediumMax];
    __shared__ int s_row[kMediumMax];
    __shared__ float s_wy[kMediumMax], s_wx[kMediumMax];
    __shared__ int s_hist[kHeavyBuckets];
    __shared__ int s_n, s_lohi[2];
    // a gradient row: rows < total_bins are rows of grad_out, the others the pre-reduced rows of the zero-size ROIs
    auto row_ptr = [&](int r) { return r < total_bins ? grad_out + (size_t)r * c4_of(C) : const_partial + (size_t)(r - total_bins) * c4_of(C); };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = C >> 2;
    if (blockIdx.x < kMediumCtas) {
        const int n_medium = misc[2];
        for (int mi = blockIdx.x; mi < n_medium; mi += kMediumCtas) {
            const int q = medium[mi];
            const int n = (int)count[q], seg = start[q];
            float* dst = reinterpret_cast<float*>(pixel_ptr(tbl, ps, q, c4));
            // More than kMediumMax samples: rounds over ranges of the sample index, chosen from a histogram so that each
            // round holds at most kMediumMax entries; the per-channel sums run on across the rounds (C <= 1024).
            int s_lo = 0, s_width = 1;
            bool hand_over = false;
            if (n > kMediumMax) {
                int mn = INT_MAX, mx = 0;
                for (int e = threadIdx.x; e < n; e += 256) { const int smp = __ldg(&entries[seg + e].x) >> 2; mn = min(mn, smp); mx = max(mx, smp); }
                mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
                for (int i = threadIdx.x; i < kHeavyBuckets; i += 256) s_hist[i] = 0;
                if (threadIdx.x == 0) { s_lohi[0] = INT_MAX; s_lohi[1] = 0; }
                __syncthreads();
                if (lane == 0) { atomicMin(&s_lohi[0], mn); atomicMax(&s_lohi[1], mx); }
                __syncthreads();
                s_lo = s_lohi[0];
                s_width = (s_lohi[1] - s_lo) / kHeavyBuckets + 1;
                for (int e = threadIdx.x; e < n; e += 256) atomicAdd(&s_hist[((__ldg(&entries[seg + e].x) >> 2) - s_lo) / s_width], 1);
                __syncthreads();
                bool over = false;
                for (int i = threadIdx.x; i < kHeavyBuckets; i += 256) over |= s_hist[i] > kMediumMax;
                hand_over = __syncthreads_or(over) || C > 1024;
            }
            if (hand_over) {   // (never seen) the atomic fallback adds this pixel's samples: zero it and flag it
                if (threadIdx.x == 0) { count[q] |= 0x40000000u; misc[1] = 1; }
                for (int ch = threadIdx.x; ch < C; ch += 256) __stcs(dst + ch, 0.0f);
                __syncthreads();
                continue;
            }
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};   // channels threadIdx.x + 256 k
            int bkt = 0;
            while (bkt < kHeavyBuckets) {
                int nr;
                if (n <= kMediumMax) {
                    for (int e = threadIdx.x; e < n; e += 256)
                        s_sort[e] = (((uint64_t)(uint32_t)__ldg(&entries[seg + e].x) << 21) | (uint32_t)e) + 1u;
                    nr = n;
                    bkt = kHeavyBuckets;
                } else {
                    int b1 = bkt, sum = 0;
                    while (b1 < kHeavyBuckets && sum + s_hist[b1] <= kMediumMax) sum += s_hist[b1++];   // uniform
                    if (threadIdx.x == 0) s_n = 0;
                    __syncthreads();
                    for (int e = threadIdx.x; e < n; e += 256) {
                        const uint32_t key = (uint32_t)__ldg(&entries[seg + e].x);
                        const int kb = ((int)(key >> 2) - s_lo) / s_width;
                        if (kb >= bkt && kb < b1) s_sort[atomicAdd(&s_n, 1)] = (((uint64_t)key << 21) | (uint32_t)e) + 1u;
                    }
                    __syncthreads();
                    nr = s_n;
                    bkt = b1;
                }
                const int np2 = max(64, 1 << (32 - __clz(max(nr, 1) - 1)));
                for (int e = nr + threadIdx.x; e < np2; e += 256) s_sort[e] = 0ull;
                __syncthreads();
                block_bitonic_sort_desc(s_sort, np2);  // descending, zero padding last: ascending rank e sits at nr - 1 - e
                for (int e = threadIdx.x; e < nr; e += 256) {
                    const int4 ent = __ldg(entries + seg + (int)((s_sort[nr - 1 - e] - 1u) & 0x1fffffu));
                    s_row[e] = ent.w;
                    s_wy[e] = __int_as_float(ent.y);
                    s_wx[e] = __int_as_float(ent.z);
                }
                __syncthreads();
                // one thread per channel (a warp covers 128 B of every gradient row), sixteen rows in flight, strict order
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ch = threadIdx.x + 256 * k;
                    if (ch >= C) break;
                    float a = acc[k];
                    int e = 0;
                    for (; e + 16 <= nr; e += 16) {
                        float val[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) val[u] = __ldg(reinterpret_cast<const float*>(row_ptr(s_row[e + u])) + ch);
#pragma unroll
                        for (int u = 0; u < 16; ++u) a = __fadd_rn(a, __fmul_rn(s_wx[e + u], __fmul_rn(s_wy[e + u], val[u])));
                    }
                    for (; e < nr; ++e)
                        a = __fadd_rn(a, __fmul_rn(s_wx[e], __fmul_rn(s_wy[e], __ldg(reinterpret_cast<const float*>(row_ptr(s_row[e])) + ch))));
                    acc[k] = a;
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (threadIdx.x + 256 * k < C) __stcs(dst + threadIdx.x + 256 * k, acc[k]);
        }
        return;
    }
    // persistent streaming pass: a CTA walks 32-pixel chunks (32 KB of contiguous gradient map at C = 256, 4 pixels per
    // warp); the next chunk's headers are fetched one iteration ahead and the entry lists of a warp's four pixels are
    // requested together, so a touched pixel costs one gradient-row latency, not three dependent ones.
    const int NP = ps.base[4];
    const int nctas = gridDim.x - kMediumCtas, chunks = (NP + 31) >> 5;
    int chunk = blockIdx.x - kMediumCtas;
    auto header = [&](int ch, uint32_t& cnt, int& st) {
        const int q = ch * 32 + warp * 4 + lane;
        const bool ok = (lane < 4) && (ch < chunks) && (q < NP);
        cnt = ok ? __ldg(count + q) : 0u;
        st = ok ? __ldg(start + q) : 0;
    };
    uint32_t c_next;
    int seg_next;
    header(chunk, c_next, seg_next);
#pragma unroll 1
    for (; chunk < chunks; chunk += nctas) {
        const uint32_t my_count = c_next;
        const int my_start = seg_next;
        header(chunk + nctas, c_next, seg_next);
        const int q0 = chunk * 32 + warp * 4;
        // fast path (most of the fine maps): none of the warp's four pixels has a sample and they are contiguous in
        // one map -> eight stores
        if (__all_sync(0xffffffffu, my_count == 0u) && q0 + 3 < NP &&
            (q0 + 3 < ps.base[1] || (q0 >= ps.base[1] && q0 + 3 < ps.base[2]) || (q0 >= ps.base[2] && q0 + 3 < ps.base[3]) ||
             q0 >= ps.base[3])) {
            float4* dst = pixel_ptr(tbl, ps, q0, c4);
            for (int v = lane; v < 4 * c4; v += 32) __stcs(dst + v, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        // -1: a medium CTA writes this pixel; 0: zeros (untouched, or the atomic fallback adds to it later)
        auto samples_of = [&](int i) {
            const uint32_t c = __shfl_sync(0xffffffffu, my_count, i);
            return (c <= (uint32_t)kLightMax) ? (int)c : (c <= (uint32_t)kHeavyMax) ? -1 : 0;
        };
        auto entry_of = [&](int i, int n) {
            const int seg = __shfl_sync(0xffffffffu, my_start, i);
            return (lane < n) ? __ldg(entries + seg + lane) : make_int4(INT_MAX, 0, 0, 0);
        };
        int n_next = samples_of(0);
        int4 ent_next = entry_of(0, n_next);
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + i, n = n_next;
            const int4 ent = ent_next;
            if (i < 3) { n_next = samples_of(i + 1); ent_next = entry_of(i + 1, n_next); }  // one pixel ahead
            if (q >= NP || n < 0) continue;
            float4* dst = pixel_ptr(tbl, ps, q, c4);
            if (n == 0) {
                if (VPL > 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) __stcs(dst + lane + 32 * v, make_float4(0.f, 0.f, 0.f, 0.f));
                } else {
                    for (int v = lane; v < c4; v += 32) __stcs(dst + v, make_float4(0.f, 0.f, 0.f, 0.f));
                }
                continue;
            }
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += (__shfl_sync(0xffffffffu, ent.x, j) < ent.x);
            if (lane >= n) rank = -1;
            const int row = ent.w;   // the gradient row to fetch (a pre-reduced row for a zero-size ROI)
            const float wy = __int_as_float(ent.y), wx = __int_as_float(ent.z);
            auto lane_of_rank = [&](int j) { return __ffs(__ballot_sync(0xffffffffu, rank == j)) - 1; };
            if (VPL > 0) {
                float4 acc[V];
#pragma unroll
                for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                int j = 0;
                for (; j + 2 <= n; j += 2) {  // two gradient rows in flight, added strictly in order
                    const int s0 = lane_of_rank(j), s1 = lane_of_rank(j + 1);
                    const float4* g0 = row_ptr(__shfl_sync(0xffffffffu, row, s0));
                    const float4* g1 = row_ptr(__shfl_sync(0xffffffffu, row, s1));
                    const float a0 = __shfl_sync(0xffffffffu, wy, s0), b0 = __shfl_sync(0xffffffffu, wx, s0);
                    const float a1 = __shfl_sync(0xffffffffu, wy, s1), b1 = __shfl_sync(0xffffffffu, wx, s1);
                    float4 v0[V], v1[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) { v0[v] = __ldg(g0 + lane + 32 * v); v1[v] = __ldg(g1 + lane + 32 * v); }
#pragma unroll
                    for (int v = 0; v < V; ++v) { fma_in_order(acc[v], v0[v], a0, b0); fma_in_order(acc[v], v1[v], a1, b1); }
                }
                if (j < n) {
                    const int s0 = lane_of_rank(j);
                    const float4* g0 = row_ptr(__shfl_sync(0xffffffffu, row, s0));
                    const float a0 = __shfl_sync(0xffffffffu, wy, s0), b0 = __shfl_sync(0xffffffffu, wx, s0);
#pragma unroll
                    for (int v = 0; v < V; ++v) fma_in_order(acc[v], __ldg(g0 + lane + 32 * v), a0, b0);
                }
#pragma unroll
                for (int v = 0; v < V; ++v) __stcs(dst + lane + 32 * v, acc[v]);
            } else {
                for (int v0 = 0; v0 < c4; v0 += 32) {  // warp-uniform trip count: the shuffles need every lane
                    const int v = v0 + lane;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int j = 0; j < n; ++j) {
                        const int s0 = lane_of_rank(j);
                        const float4* gr = row_ptr(__shfl_sync(0xffffffffu, row, s0));
                        const float a0 = __shfl_sync(0xffffffffu, wy, s0), b0 = __shfl_sync(0xffffffffu, wx, s0);
                        if (v < c4) fma_in_order(acc, __ldg(gr + v), a0, b0);
                    }
                    if (v < c4) __stcs(dst + v, acc);
                }
            }
        }
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

// workspace of the deterministic backward:
// [count NP | cursor NP | misc 64] (zeroed per call) [start NP] [medium list] [entries 4*bins x 16 B] [pre-reduced rows of
// the zero-size ROIs: BN x C floats]
static size_t roialign_bwd_ws_bytes(int NP, long long bins, int C, long long BN) {
    return align_up((2 * (size_t)NP + 64) * sizeof(int), 256) + align_up((size_t)NP * sizeof(int), 256) +
           align_up(((size_t)NP / 32 + 1 + 4 * (size_t)bins / 32) * sizeof(int), 256) +
           align_up(4 * (size_t)bins * sizeof(int4), 256) + align_up((size_t)BN * C * sizeof(float), 256);
}

MRCNN_EXPORT int mrcnn_roialign_backward_workspace_bytes(int B, int N, int ph, int pw, const int* H, const int* W,
                                                         int C, size_t* bytes) {
    if (!bytes || !H || !W) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || C < 4 || (C & 3) || C > 8192) return MRCNN_ERR_RANGE;
    for (int l = 0; l < 4; ++l)
        if (H[l] < 1 || W[l] < 1) return MRCNN_ERR_RANGE;
    PixelSpace ps;
    const long long bins = (long long)B * N * ph * pw;
    if (pixel_space(H, W, B, &ps) != MRCNN_OK || bins >= (1LL << 29)) return MRCNN_ERR_RANGE;
    *bytes = roialign_bwd_ws_bytes(ps.base[4], bins, C, (long long)B * N);
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
    PixelSpace ps;
    rc = pixel_space(H, W, B, &ps);
    if (rc != MRCNN_OK) return rc;
    const int groups = (ph + kRoiThreads / 32 - 1) / (kRoiThreads / 32);
    const int rows_per_group = (ph + groups - 1) / groups;  // 7x7 -> 1 x 7 rows, 14x14 -> 2 x 7, 28x28 -> 4 x 7
    const size_t smem = (size_t)(kRoiThreads / 32) * C * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(roialign_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    if (ws == nullptr) {  // atomic mode: zero-fill, then scatter every sample with vector reductions
        for (int l = 0; l < 4; ++l) {
            cudaError_t e = cudaMemsetAsync(grad_fmaps[l], 0, (size_t)B * H[l] * W[l] * C * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
        }
        roialign_bwd_kernel<<<B * N * groups, kRoiThreads, smem, st>>>((const float4*)grad_out, (const float4*)boxes,
                                                                      roi_map, tbl, C, N, ph, pw, groups, rows_per_group,
                                                                      ps, nullptr, nullptr);
        return last_error();
    }
    // deterministic mode
    const long long bins_ll = (long long)B * N * ph * pw;
    if (bins_ll >= (1LL << 29)) return MRCNN_ERR_RANGE;
    const int NP = ps.base[4], bins = (int)bins_ll;
    if (!aligned16(ws)) return MRCNN_ERR_ALIGN;
    if (ws_bytes < roialign_bwd_ws_bytes(NP, bins_ll, C, (long long)B * N)) return MRCNN_ERR_WORKSPACE;
    const size_t zeroed = align_up((2 * (size_t)NP + 64) * sizeof(int), 256);
    uint32_t* count = (uint32_t*)ws;
    int* cursor = (int*)ws + NP;
    int* misc = (int*)ws + 2 * (size_t)NP;  // [0] key-list bump pointer, [1] "ordinary samples on a fallback pixel"
    int* start = (int*)((char*)ws + zeroed);
    // a medium pixel holds > kLightMax of the <= 4*bins keys, and there are at most NP pixels
    int* medium = (int*)((char*)start + align_up((size_t)NP * sizeof(int), 256));
    int4* entries = (int4*)((char*)medium + align_up(((size_t)NP / 32 + 1 + 4 * (size_t)bins / 32) * sizeof(int), 256));
    float4* partial = (float4*)((char*)entries + align_up(4 * (size_t)bins * sizeof(int4), 256));
    cudaError_t e = cudaMemsetAsync(ws, 0, zeroed, st);
    if (e != cudaSuccess) return (int)e;
    roialign_bwd_const_kernel<<<B * N, 64, 0, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map, tbl, C, ph, pw,
                                                    partial);
    const int bin_grid = (bins + 255) / 256;
    roialign_bwd_count_kernel<<<bin_grid, 256, 0, st>>>((const float4*)boxes, roi_map, tbl, ps, N, ph, pw, bins, count,
                                                        misc);
    roialign_bwd_alloc_kernel<<<(NP + 1023) / 1024, 256, 0, st>>>(count, NP, start, misc, medium);
    roialign_bwd_fill_kernel<<<bin_grid, 256, 0, st>>>((const float4*)boxes, roi_map, tbl, ps, N, ph, pw, bins, count,
                                                      start, cursor, entries);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int gather_grid = kMediumCtas + min((NP + 31) / 32, 4 * sms);  // 4 resident CTAs per SM (__launch_bounds__)
#define MRCNN_GATHER(V) roialign_bwd_gather_kernel<V><<<gather_grid, 256, 0, st>>>((const float4*)grad_out, partial, bins, \
        tbl, ps, C, count, start, entries, misc, medium)
    if (C == 128) MRCNN_GATHER(1);
    else if (C == 256) MRCNN_GATHER(2);
    else if (C == 512) MRCNN_GATHER(4);
    else MRCNN_GATHER(0);
#undef MRCNN_GATHER
    roialign_bwd_kernel<<<B * N * groups, kRoiThreads, smem, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map,
                                                                  tbl, C, N, ph, pw, groups, rows_per_group, ps, count,
                                                                  misc + 1);
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
    // CTAs per SM of the (persistent) fetch kernel (MRCNN_FETCH_CTAS_PER_SM for measurements: 4 leaves warp slots to a
    // second batch's kernels and gave +2.5 % end to end at one GPU, but 1400-1630 instead of 2100 images/s at two)
    const int grid = min((words + 7) / 8, max(1, min(8, tuning_knob("MRCNN_FETCH_CTAS_PER_SM", 8))) * sms);
#define MRCNN_FETCH(V) roialign_fetch_kernel<V><<<grid, 256, 0, st>>>(host, dev, ps, C, words, need, resident, fetched)
    if (C == 128) MRCNN_FETCH(1);
    else if (C == 256) MRCNN_FETCH(2);
    else if (C == 512) MRCNN_FETCH(4);
    else MRCNN_FETCH(0);
#undef MRCNN_FETCH
    return last_error();
}
