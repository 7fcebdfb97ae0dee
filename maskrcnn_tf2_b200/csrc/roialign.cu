// roialign.cu -- PyramidROIAlign.call (mrcnn_layers.py:583-664) forward and its feature-map gradient.
//
// forward = roialign_prep_kernel (one CTA): FPN level per ROI (L:596-607, utils.py:825-827), first-appearance
//           level -> map table over the flattened batch (L:613-615, quirk Q2), roi_map[B,N];
//         + roialign_fwd_kernel (one CTA per ROI, one warp per output bin): TF CropAndResize bilinear sampling
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

constexpr int kPrepThreads = 1024;
constexpr int kRoiThreads = 256;  // 8 warps per ROI

__device__ __forceinline__ int roi_level_of(float4 b, float denom) {
    const float h = __fsub_rn(b.z, b.x);
    const float w = __fsub_rn(b.w, b.y);
    const float x = __fdiv_rn(__fsqrt_rn(__fmul_rn(h, w)), denom);                 // L:605
    const float lv = __fdiv_rn(det_logf(x), 0.693147182464599609375f);              // utils.log2_graph
    const int r = cast_i32_x86(rintf(lv));                                           // tf.round + int32 cast
    int level = (r < 0) ? max(4 + r, 2) : min(4 + r, 5);                             // L:607, overflow-free
    return min(max(level, 2), 5);
}

__global__ void __launch_bounds__(kPrepThreads)
roialign_prep_kernel(const float4* __restrict__ boxes, const float* __restrict__ image_meta, int BN,
                     float denominator, int map_mode, int32_t* __restrict__ roi_map,
                     int32_t* __restrict__ roi_level) {
    __shared__ int first[4];
    __shared__ int table[4];
    const int tid = threadIdx.x;
    if (tid < 4) first[tid] = INT_MAX;
    __syncthreads();
    const float image_area = __fmul_rn(image_meta[4], image_meta[5]);               // L:600,604 (image 0)
    const float denom = __fdiv_rn(denominator, __fsqrt_rn(image_area));
    int mine[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
    for (int f = tid; f < BN; f += kPrepThreads) {
        const int level = roi_level_of(__ldg(boxes + f), denom);
        roi_map[f] = level;
#pragma unroll
        for (int l = 0; l < 4; ++l)
            if (level == l + 2 && mine[l] == INT_MAX) mine[l] = f;
    }
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const int m = __reduce_min_sync(0xffffffffu, mine[l]);
        if ((tid & 31) == 0 && m != INT_MAX) atomicMin(&first[l], m);
    }
    __syncthreads();
    if (tid < 4) {
        int rank = tid;  // map_mode 1: level - 2
        if (map_mode == 0) {  // rank of this level in first-appearance order (tf.unique, L:613)
            rank = 0;
            for (int l = 0; l < 4; ++l) rank += (first[l] < first[tid]) ? 1 : 0;
        }
        table[tid] = rank;
    }
    __syncthreads();
    for (int f = tid; f < BN; f += kPrepThreads) {
        const int level = roi_map[f];
        roi_map[f] = table[level - 2];
        if (roi_level) roi_level[f] = level;
    }
}

struct RoiParams {
    const float* base;  // feature map of this ROI's image
    float y1, x1, y2, x2, hs, ws;
    int H, W;
};

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

template <int VPL>  // float4 vectors per lane: C == VPL * 128; VPL == 0 -> generic C
__global__ void __launch_bounds__(kRoiThreads)
roialign_fwd_kernel(const float4* __restrict__ boxes, const int32_t* __restrict__ roi_map, MapTable tbl, int C, int N,
                    int ph, int pw, float* __restrict__ out) {
    __shared__ RoiParams sp;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) {
        const float4 b = __ldg(boxes + f);
        const int m = roi_map[f];
        RoiParams p;
        p.H = tbl.H[m];
        p.W = tbl.W[m];
        p.base = tbl.ptr[m] + (size_t)(f / N) * p.H * p.W * C;
        p.y1 = b.x; p.x1 = b.y; p.y2 = b.z; p.x2 = b.w;
        p.hs = crop_scale(b.x, b.z, p.H, ph);
        p.ws = crop_scale(b.y, b.w, p.W, pw);
        sp = p;
    }
    __syncthreads();
    const RoiParams p = sp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bins = ph * pw;
    const int c4 = C >> 2;
    float4* orow = reinterpret_cast<float4*>(out) + (size_t)f * bins * c4;
    for (int bin = warp; bin < bins; bin += kRoiThreads / 32) {
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
                a[v] = ldg4(tl + i); b[v] = ldg4(tr + i); c[v] = ldg4(bl + i); d[v] = ldg4(br + i);
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v) __stcs(o + lane + 32 * v, lerp4(a[v], b[v], c[v], d[v]));
        } else {
            for (int i = lane; i < c4; i += 32) __stcs(o + i, lerp4(ldg4(tl + i), ldg4(tr + i), ldg4(bl + i), ldg4(br + i)));
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
                    const int32_t* __restrict__ roi_map, GradTable tbl, int C, int N, int ph, int pw) {
    __shared__ RoiParams sp;
    __shared__ float* s_gbase;
    const int f = blockIdx.x;
    if (threadIdx.x == 0) {
        const float4 b = __ldg(boxes + f);
        const int m = roi_map[f];
        RoiParams p;
        p.H = tbl.H[m];
        p.W = tbl.W[m];
        p.base = nullptr;
        s_gbase = tbl.ptr[m] + (size_t)(f / N) * p.H * p.W * C;
        p.y1 = b.x; p.x1 = b.y; p.y2 = b.z; p.x2 = b.w;
        p.hs = crop_scale(b.x, b.z, p.H, ph);
        p.ws = crop_scale(b.y, b.w, p.W, pw);
        sp = p;
    }
    __syncthreads();
    const RoiParams p = sp;
    float* gbase = s_gbase;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bins = ph * pw;
    const int c4 = C >> 2;
    const float4* grow = grad_out + (size_t)f * bins * c4;
    for (int bin = warp; bin < bins; bin += kRoiThreads / 32) {
        const int y = bin / pw, x = bin - y * pw;
        const Tap ty = make_tap(p.y1, p.y2, p.H, ph, y, p.hs);
        const Tap tx = make_tap(p.x1, p.x2, p.W, pw, x, p.ws);
        if (!(ty.valid && tx.valid)) continue;
        float* tl = gbase + ((size_t)ty.lo * p.W + tx.lo) * C;
        float* tr = gbase + ((size_t)ty.lo * p.W + tx.hi) * C;
        float* bl = gbase + ((size_t)ty.hi * p.W + tx.lo) * C;
        float* br = gbase + ((size_t)ty.hi * p.W + tx.hi) * C;
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

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_roialign_workspace_bytes(int B, int N, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1) return MRCNN_ERR_RANGE;
    *bytes = 256;  // nothing needed beyond the caller-owned roi_map; kept non-zero so callers can always allocate
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
    (void)ws; (void)ws_bytes;
    if (!boxes || !image_meta || !out || !roi_map) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || meta_len < 6 || (map_mode != 0 && map_mode != 1) ||
        (long long)B * N > INT_MAX / 2 || !(denominator > 0.0f))
        return MRCNN_ERR_RANGE;
    if (!aligned16(boxes) || !aligned16(out)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    MapTable tbl;
    for (int l = 0; l < 4; ++l) { tbl.ptr[l] = fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l]; }
    roialign_prep_kernel<<<1, kPrepThreads, 0, st>>>((const float4*)boxes, image_meta, B * N, denominator, map_mode,
                                                     roi_map, roi_level);
    const int grid = B * N;
    if (C == 128) roialign_fwd_kernel<1><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, roi_map, tbl, C, N, ph, pw, out);
    else if (C == 256) roialign_fwd_kernel<2><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, roi_map, tbl, C, N, ph, pw, out);
    else if (C == 512) roialign_fwd_kernel<4><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, roi_map, tbl, C, N, ph, pw, out);
    else roialign_fwd_kernel<0><<<grid, kRoiThreads, 0, st>>>((const float4*)boxes, roi_map, tbl, C, N, ph, pw, out);
    return last_error();
}

MRCNN_EXPORT int mrcnn_roialign_backward(const float* grad_out, const float* boxes, const int32_t* roi_map,
                                         float* const* grad_fmaps, const int* H, const int* W, int C, int B, int N,
                                         int ph, int pw, void* stream) {
    if (!grad_out || !boxes || !roi_map) return MRCNN_ERR_NULL;
    int rc = check_maps((const void* const*)grad_fmaps, H, W, C);
    if (rc != MRCNN_OK) return rc;
    if (B < 1 || N < 1 || ph < 1 || pw < 1 || (long long)B * N > INT_MAX / 2) return MRCNN_ERR_RANGE;
    if (!aligned16(boxes) || !aligned16(grad_out)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    GradTable tbl;
    for (int l = 0; l < 4; ++l) {
        tbl.ptr[l] = grad_fmaps[l]; tbl.H[l] = H[l]; tbl.W[l] = W[l];
        cudaError_t e = cudaMemsetAsync(grad_fmaps[l], 0, (size_t)B * H[l] * W[l] * C * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    roialign_bwd_kernel<<<B * N, kRoiThreads, 0, st>>>((const float4*)grad_out, (const float4*)boxes, roi_map, tbl, C, N,
                                                       ph, pw);
    return last_error();
}
