// roialign.cu -- PyramidROIAlign.call (mrcnn_layers.py:583-664) forward and its feature-map gradient.
//
// forward = roialign_prep_kernel: FPN level per ROI (L:596-607, utils.py:825-827) and, per level, the first
//           flattened index at which it appears (tf.unique order over the whole batch, L:613-615, quirk Q2);
//         + roialign_fwd_kernel (<= 32 output bins per CTA, one warp per bin): each warp turns the four
//           first-appearance indices into its ROI's map index, then TF CropAndResize bilinear sampling
//           (crop_and_resize_op.cc; called at L:641) with 128-bit channel-vectorised NHWC loads, written straight
//           into [B,N,ph,pw,C] in input ROI order -- the reference's concat / top_k re-sort / gather passes
//           (L:644-659) have no counterpart here because nothing is ever out of order.
// backward = memset of the four gradient maps + roialign_bwd_kernel: TF CropAndResizeGradImage scatter with
//           red.global.add.v4.f32 (one 16-byte reduction per lane and corner).
#include <limits.h>

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

constexpr int kPrepThreads = 256;
constexpr int kRoiThreads = 256;   // 8 warps per CTA
constexpr int kBinsPerCta = 32;    // upper bound; bins are split evenly over ceil(bins / 32) CTAs per ROI

__device__ __forceinline__ int roi_level_of(float4 b, float denom) {
    const float h = __fsub_rn(b.z, b.x);
    const float w = __fsub_rn(b.w, b.y);
    const float x = __fdiv_rn(__fsqrt_rn(__fmul_rn(h, w)), denom);                 // L:605
    const float lv = __fdiv_rn(det_logf(x), 0.693147182464599609375f);              // utils.log2_graph
    const int r = cast_i32_x86(rintf(lv));                                           // tf.round + int32 cast
    int level = (r < 0) ? max(4 + r, 2) : min(4 + r, 5);                             // L:607, overflow-free
    return min(max(level, 2), 5);
}

// level per ROI + first flattened index at which each level appears (tf.unique order, L:613)
__global__ void __launch_bounds__(kPrepThreads)
roialign_prep_kernel(const float4* __restrict__ boxes, const float* __restrict__ image_meta, int BN,
                     float denominator, int32_t* __restrict__ level_ws, int* __restrict__ first,
                     int32_t* __restrict__ roi_level) {
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
        if ((threadIdx.x & 31) == 0 && m != INT_MAX) atomicMin(&first[l], m);
    }
}

struct RoiParams {
    const float* base;  // feature map of this ROI's image
    float y1, x1, y2, x2, hs, ws;
    int H, W;
};

// per-warp (redundant, divergence-free) ROI setup: level -> map index through the first-appearance table
template <typename Table>
__device__ __forceinline__ int roi_setup(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                                         const int* __restrict__ first, int map_mode, const Table& tbl, int f, int ph,
                                         int pw, RoiParams& p, size_t& image_offset, int N, int C) {
    const int level = level_ws[f];
    int m = level - 2;  // map_mode 1
    if (map_mode == 0) {  // rank of this level in first-appearance order (L:613-619)
        const int4 fa = *reinterpret_cast<const int4*>(first);
        const int mine = (level == 2) ? fa.x : (level == 3) ? fa.y : (level == 4) ? fa.z : fa.w;
        m = (fa.x < mine) + (fa.y < mine) + (fa.z < mine) + (fa.w < mine);
    }
    p.H = (m == 0) ? tbl.H[0] : (m == 1) ? tbl.H[1] : (m == 2) ? tbl.H[2] : tbl.H[3];
    p.W = (m == 0) ? tbl.W[0] : (m == 1) ? tbl.W[1] : (m == 2) ? tbl.W[2] : tbl.W[3];
    image_offset = (size_t)(f / N) * p.H * p.W * C;
    const float4 b = __ldg(boxes + f);
    p.y1 = b.x; p.x1 = b.y; p.y2 = b.z; p.x2 = b.w;
    p.hs = crop_scale(b.x, b.z, p.H, ph);
    p.ws = crop_scale(b.y, b.w, p.W, pw);
    return m;
}

template <int VPL>  // float4 vectors per lane: C == VPL * 128; VPL == 0 -> generic C
__global__ void __launch_bounds__(kRoiThreads)
roialign_fwd_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                    const int* __restrict__ first, int map_mode, MapTable tbl, int C, int N, int ph, int pw,
                    int chunks, int per_chunk, float* __restrict__ out, int32_t* __restrict__ roi_map) {
    const int f = blockIdx.x / chunks, chunk = blockIdx.x - f * chunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RoiParams p;
    size_t off;
    const int m = roi_setup(boxes, level_ws, first, map_mode, tbl, f, ph, pw, p, off, N, C);
    p.base = ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) + off;
    if (chunk == 0 && threadIdx.x == 0) roi_map[f] = m;
    const int bins = ph * pw;
    const int bin_end = min(bins, (chunk + 1) * per_chunk);
    const int c4 = C >> 2;
    float4* orow = reinterpret_cast<float4*>(out) + (size_t)f * bins * c4;
    for (int bin = chunk * per_chunk + warp; bin < bin_end; bin += kRoiThreads / 32) {
        const int y = bin / pw, x = bin - y * pw;
        const Tap ty = make_tap(p.y1, p.y2, p.H, ph, y, p.hs);
        const Tap tx = make_tap(p.x1, p.x2, p.W, pw, x, p.ws);
        float4* o = orow + (size_t)bin * c4;
        if (!(ty.valid && tx.valid)) {  // extrapolation_value = 0
            for (int v = lane; v < c4; v += 32) __stcs(o + v, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        const float4* tl = reinterpret_cast<const float4*>(p.base + ((size_t)ty.lo * p.W + tx.lo) * C);
        const float4* tr = reinterpret_cast<const float4*>(p.base + ((size_t)ty.lo * p.W + tx.hi) * C);
        const float4* bl = reinterpret_cast<const float4*>(p.base + ((size_t)ty.hi * p.W + tx.lo) * C);
        const float4* br = reinterpret_cast<const float4*>(p.base + ((size_t)ty.hi * p.W + tx.hi) * C);
        const float lx = tx.lerp, ly = ty.lerp;
        auto lerp4 = [&](const float4& a, const float4& b, const float4& c, const float4& d) {
            float4 r;
            float top, bot;
            top = __fadd_rn(a.x, __fmul_rn(__fsub_rn(b.x, a.x), lx)); bot = __fadd_rn(c.x, __fmul_rn(__fsub_rn(d.x, c.x), lx));
            r.x = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            top = __fadd_rn(a.y, __fmul_rn(__fsub_rn(b.y, a.y), lx)); bot = __fadd_rn(c.y, __fmul_rn(__fsub_rn(d.y, c.y), lx));
            r.y = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            top = __fadd_rn(a.z, __fmul_rn(__fsub_rn(b.z, a.z), lx)); bot = __fadd_rn(c.z, __fmul_rn(__fsub_rn(d.z, c.z), lx));
            r.z = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            top = __fadd_rn(a.w, __fmul_rn(__fsub_rn(b.w, a.w), lx)); bot = __fadd_rn(c.w, __fmul_rn(__fsub_rn(d.w, c.w), lx));
            r.w = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
            return r;
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

__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// TF CropAndResizeGradImage: dtop = (1-ly) g; tl += (1-lx) dtop; tr += lx dtop; dbot = ly g; bl += (1-lx) dbot;
// br += lx dbot -- skipped exactly where the forward pass extrapolated.
__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ boxes,
                    const int32_t* __restrict__ roi_map, GradTable tbl, int C, int N, int ph, int pw, int chunks,
                    int per_chunk) {
    const int f = blockIdx.x / chunks, chunk = blockIdx.x - f * chunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = roi_map[f];
    const int H = (m == 0) ? tbl.H[0] : (m == 1) ? tbl.H[1] : (m == 2) ? tbl.H[2] : tbl.H[3];
    const int W = (m == 0) ? tbl.W[0] : (m == 1) ? tbl.W[1] : (m == 2) ? tbl.W[2] : tbl.W[3];
    float* gbase = ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
                   (size_t)(f / N) * H * W * C;
    const float4 b = __ldg(boxes + f);
    const float hs = crop_scale(b.x, b.z, H, ph), ws = crop_scale(b.y, b.w, W, pw);
    const int bins = ph * pw;
    const int bin_end = min(bins, (chunk + 1) * per_chunk);
    const int c4 = C >> 2;
    const float4* grow = grad_out + (size_t)f * bins * c4;
    for (int bin = chunk * per_chunk + warp; bin < bin_end; bin += kRoiThreads / 32) {
        const int y = bin / pw, x = bin - y * pw;
        const Tap ty = make_tap(b.x, b.z, H, ph, y, hs);
        const Tap tx = make_tap(b.y, b.w, W, pw, x, ws);
        if (!(ty.valid && tx.valid)) continue;
        float* tl = gbase + ((size_t)ty.lo * W + tx.lo) * C;
        float* tr = gbase + ((size_t)ty.lo * W + tx.hi) * C;
        float* bl = gbase + ((size_t)ty.hi * W + tx.lo) * C;
        float* br = gbase + ((size_t)ty.hi * W + tx.hi) * C;
        const float lx = tx.lerp, ly = ty.lerp;
        const float wy0 = __fsub_rn(1.0f, ly), wx0 = __fsub_rn(1.0f, lx);
        for (int i = lane; i < c4; i += 32) {
            const float4 g = __ldcs(grow + (size_t)bin * c4 + i);
            const float4 dt = make_float4(__fmul_rn(wy0, g.x), __fmul_rn(wy0, g.y), __fmul_rn(wy0, g.z), __fmul_rn(wy0, g.w));
            const float4 db = make_float4(__fmul_rn(ly, g.x), __fmul_rn(ly, g.y), __fmul_rn(ly, g.z), __fmul_rn(ly, g.w));
            red_add_v4(tl + 4 * i, __fmul_rn(wx0, dt.x), __fmul_rn(wx0, dt.y), __fmul_rn(wx0, dt.z), __fmul_rn(wx0, dt.w));
            red_add_v4(tr + 4 * i, __fmul_rn(lx, dt.x), __fmul_rn(lx, dt.y), __fmul_rn(lx, dt.z), __fmul_rn(lx, dt.w));
            red_add_v4(bl + 4 * i, __fmul_rn(wx0, db.x), __fmul_rn(wx0, db.y), __fmul_rn(wx0, db.z), __fmul_rn(wx0, db.w));
            red_add_v4(br + 4 * i, __fmul_rn(lx, db.x), __fmul_rn(lx, db.y), __fmul_rn(lx, db.z), __fmul_rn(lx, db.w));
        }
    }
}

static size_t roialign_ws_bytes(int B, int N) { return align_up((size_t)B * N * sizeof(int32_t), 256) + 256; }

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_roialign_workspace_bytes(int B, int N, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1) return MRCNN_ERR_RANGE;
    *bytes = roialign_ws_bytes(B, N);  // per-ROI level + the four first-appearance indices
    return MRCNN_OK;
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
        (long long)B * N > INT_MAX / 64 || !(denominator > 0.0f))
        return MRCNN_ERR_RANGE;
    if (ws_bytes < roialign_ws_bytes(B, N)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(boxes) || !aligned16(out) || !aligned16(ws)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    MapTable tbl;
    for (int l = 0; l < 4; ++l) { tbl.ptr[l] = fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l]; }
    const int BN = B * N;
    int* first = (int*)ws;                                   // 4 ints, 16-byte aligned
    int32_t* level_ws = (int32_t*)((char*)ws + 256);
    cudaError_t e = cudaMemsetAsync(first, 0x7f, 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    roialign_prep_kernel<<<(BN + kPrepThreads - 1) / kPrepThreads, kPrepThreads, 0, st>>>(
        (const float4*)boxes, image_meta, BN, denominator, level_ws, first, roi_level);
    const int bins = ph * pw;
    const int chunks = (bins + kBinsPerCta - 1) / kBinsPerCta;
    const int per_chunk = (bins + chunks - 1) / chunks;
    const int grid = BN * chunks;
#define MRCNN_FWD(V) roialign_fwd_kernel<V><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, level_ws, first, \
        map_mode, tbl, C, N, ph, pw, chunks, per_chunk, out, roi_map)
    if (C == 128) MRCNN_FWD(1);
    else if (C == 256) MRCNN_FWD(2);
    else if (C == 512) MRCNN_FWD(4);
    else MRCNN_FWD(0);
#undef MRCNN_FWD
    return last_error();
}

MRCNN_EXPORT int mrcnn_roialign_backward(const float* grad_out, const float* boxes, const int32_t* roi_map,
                                         float* const* grad_fmaps, const int* H, const int* W, int C, int B, int N,
                                         int ph, int pw, void* stream) {
    if (!grad_out || !boxes || !roi_map) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)grad_fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || (long long)B * N > INT_MAX / 64) return MRCNN_ERR_RANGE;
    if (!aligned16(boxes) || !aligned16(grad_out)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    GradTable tbl;
    for (int l = 0; l < 4; ++l) {
        tbl.ptr[l] = grad_fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l];
        cudaError_t e = cudaMemsetAsync(grad_fmaps[l], 0, (size_t)B * H[l] * W[l] * C * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int bins = ph * pw;
    const int chunks = (bins + kBinsPerCta - 1) / kBinsPerCta;
    const int per_chunk = (bins + chunks - 1) / chunks;
    roialign_bwd_kernel<<<B * N * chunks, kRoiThreads, 0, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map,
                                                                tbl, C, N, ph, pw, chunks, per_chunk);
    return last_error();
}
