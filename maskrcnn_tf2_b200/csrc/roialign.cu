// roialign.cu -- PyramidROIAlign.call (mrcnn_layers.py:583-664) forward and its feature-map gradient.
//
// forward = roialign_prep_kernel: FPN level per ROI (L:596-607, utils.py:825-827) and, per level, the first
//           flattened index at which it appears (tf.unique order over the whole batch, L:613-615, quirk Q2);
//         + roialign_fwd_kernel (one warp per output row of a ROI): each warp turns the four first-appearance
//           indices into its ROI's map index and walks the bins of its row with TF CropAndResize bilinear sampling
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
__device__ __forceinline__ int roi_map_index(int level, const int* __restrict__ first, int map_mode) {
    if (map_mode != 0) return level - 2;
    const int4 fa = *reinterpret_cast<const int4*>(first);
    const int mine = (level == 2) ? fa.x : (level == 3) ? fa.y : (level == 4) ? fa.z : fa.w;
    return (fa.x < mine) + (fa.y < mine) + (fa.z < mine) + (fa.w < mine);
}

// One warp per output ROW (roi f, row y): lanes span the channels with 128-bit accesses, the warp walks the pw
// bins of its row.  VPL = float4 vectors per lane (C == VPL * 128); VPL == 0 -> any C that is a multiple of 4.
template <int VPL>
__global__ void __launch_bounds__(kRoiThreads, 5)
roialign_fwd_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ level_ws,
                    const int* __restrict__ first, int map_mode, MapTable tbl, int C, int N, int ph, int pw,
                    int total_rows, float* __restrict__ out, int32_t* __restrict__ roi_map) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (kRoiThreads / 32) + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    const int f = row / ph, y = row - f * ph;
    const int m = roi_map_index(level_ws[f], first, map_mode);
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    if (y == 0 && lane == 0) roi_map[f] = m;
    const int c4 = C >> 2;
    const float4* img = reinterpret_cast<const float4*>(
        ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
        (size_t)(f / N) * g.H * g.W * C);
    float4* o = reinterpret_cast<float4*>(out) + (size_t)row * pw * c4;
    const AxisTap ty = axis_tap(g.y0, g.hs, y, g.H);
    const int top = ty.lo * g.W, bot = ty.hi * g.W;  // pixel index of the two sampled rows
    const float ly = ty.lerp;
    for (int x = 0; x < pw; ++x, o += c4) {
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
__device__ __forceinline__ void scatter_corners(float* tl, float* tr, float* bl, float* br, int i, const float4& v,
                                                float wy0, float ly, float wx0, float lx) {
    const float4 dt = make_float4(__fmul_rn(wy0, v.x), __fmul_rn(wy0, v.y), __fmul_rn(wy0, v.z), __fmul_rn(wy0, v.w));
    const float4 db = make_float4(__fmul_rn(ly, v.x), __fmul_rn(ly, v.y), __fmul_rn(ly, v.z), __fmul_rn(ly, v.w));
    if (wy0 != 0.0f && wx0 != 0.0f)
        red_add_v4(tl + 4 * i, __fmul_rn(wx0, dt.x), __fmul_rn(wx0, dt.y), __fmul_rn(wx0, dt.z), __fmul_rn(wx0, dt.w));
    if (wy0 != 0.0f && lx != 0.0f)
        red_add_v4(tr + 4 * i, __fmul_rn(lx, dt.x), __fmul_rn(lx, dt.y), __fmul_rn(lx, dt.z), __fmul_rn(lx, dt.w));
    if (ly != 0.0f && wx0 != 0.0f)
        red_add_v4(bl + 4 * i, __fmul_rn(wx0, db.x), __fmul_rn(wx0, db.y), __fmul_rn(wx0, db.z), __fmul_rn(wx0, db.w));
    if (ly != 0.0f && lx != 0.0f)
        red_add_v4(br + 4 * i, __fmul_rn(lx, db.x), __fmul_rn(lx, db.y), __fmul_rn(lx, db.z), __fmul_rn(lx, db.w));
}

__global__ void __launch_bounds__(kRoiThreads)
roialign_bwd_kernel(const float4* __restrict__ grad_out, const float4* __restrict__ boxes,
                    const int32_t* __restrict__ roi_map, GradTable tbl, int C, int N, int ph, int pw, int groups,
                    int rows_per_group) {
    extern __shared__ __align__(16) float4 s_acc[];  // [8 warps][C/4], constant-tap ROIs only
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x / groups, grp = blockIdx.x - f * groups;
    const int y = grp * rows_per_group + warp;
    const bool has_row = (warp < rows_per_group) && (y < ph);
    const int m = roi_map[f];
    const RoiGeom g = roi_geom(__ldg(boxes + f), m, tbl.H, tbl.W, ph, pw);
    const int c4 = C >> 2;
    float* gimg = ((m == 0) ? tbl.ptr[0] : (m == 1) ? tbl.ptr[1] : (m == 2) ? tbl.ptr[2] : tbl.ptr[3]) +
                  (size_t)(f / N) * g.H * g.W * C;
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
        float* tl = gimg + (size_t)(top + tx.lo) * C;
        float* tr = gimg + (size_t)(top + tx.hi) * C;
        float* bl = gimg + (size_t)(bot + tx.lo) * C;
        float* br = gimg + (size_t)(bot + tx.hi) * C;
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = threadIdx.x; i < c4; i += kRoiThreads) {
            float4 acc = s_acc[i];
            for (int w = 1; w < kRoiThreads / 32; ++w) {
                const float4 v = s_acc[w * c4 + i];
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            scatter_corners(tl, tr, bl, br, i, acc, wy0, ly, wx0, lx);
        }
        return;
    }
    if (!has_row || !ty.valid) return;
    if (g.ws == 0.0f) {  // every bin of the row hits the same pixels: sum the row first
        const AxisTap tx = axis_tap(g.x0, g.ws, 0, g.W);
        if (!tx.valid) return;
        float* tl = gimg + (size_t)(top + tx.lo) * C;
        float* tr = gimg + (size_t)(top + tx.hi) * C;
        float* bl = gimg + (size_t)(bot + tx.lo) * C;
        float* br = gimg + (size_t)(bot + tx.hi) * C;
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int x = 0; x < pw; ++x) {
                const float4 v = __ldcs(gr + (size_t)x * c4 + i);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            scatter_corners(tl, tr, bl, br, i, acc, wy0, ly, wx0, lx);
        }
        return;
    }
    for (int x = 0; x < pw; ++x, gr += c4) {
        const AxisTap tx = axis_tap(g.x0, g.ws, x, g.W);
        if (!tx.valid) continue;
        float* tl = gimg + (size_t)(top + tx.lo) * C;
        float* tr = gimg + (size_t)(top + tx.hi) * C;
        float* bl = gimg + (size_t)(bot + tx.lo) * C;
        float* br = gimg + (size_t)(bot + tx.hi) * C;
        const float lx = tx.lerp, wx0 = __fsub_rn(1.0f, tx.lerp);
        for (int i = lane; i < c4; i += 32) scatter_corners(tl, tr, bl, br, i, __ldcs(gr + i), wy0, ly, wx0, lx);
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
        (long long)B * N * ph > INT_MAX / 2 || !(denominator > 0.0f))
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
    const int total_rows = BN * ph;
    const int grid = (total_rows + kRoiThreads / 32 - 1) / (kRoiThreads / 32);
#define MRCNN_FWD(V) roialign_fwd_kernel<V><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, level_ws, first, \
        map_mode, tbl, C, N, ph, pw, total_rows, out, roi_map)
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
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || (long long)B * N * ph > INT_MAX / 2 || C > 8192) return MRCNN_ERR_RANGE;
    if (!aligned16(boxes) || !aligned16(grad_out)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    GradTable tbl;
    for (int l = 0; l < 4; ++l) {
        tbl.ptr[l] = grad_fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l];
        cudaError_t e = cudaMemsetAsync(grad_fmaps[l], 0, (size_t)B * H[l] * W[l] * C * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    const int groups = (ph + kRoiThreads / 32 - 1) / (kRoiThreads / 32);
    const int rows_per_group = (ph + groups - 1) / groups;  // 7x7 -> 1 x 7 rows, 14x14 -> 2 x 7, 28x28 -> 4 x 7
    const size_t smem = (size_t)(kRoiThreads / 32) * C * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(roialign_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    roialign_bwd_kernel<<<B * N * groups, kRoiThreads, smem, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map,
                                                                  tbl, C, N, ph, pw, groups, rows_per_group);
    return last_error();
}
