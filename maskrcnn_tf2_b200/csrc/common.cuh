// common.cuh -- device helpers shared by the ROI-stage kernels (sm_100a).
//
// Numerics contract (DESIGN.md "Numerics"): every fp32 operation that the reference performs as a separate
// TensorFlow op is issued here as an individually rounded IEEE operation (__fadd_rn/__fmul_rn/__fdiv_rn,
// which nvcc never contracts into FMA), in the reference's order (utils.py:830-869, TF crop_and_resize_op.cc,
// non_max_suppression_op.cc).  exp/log follow one fixed Cephes-style fmaf sequence so that decoded boxes are
// bit-identical on any IEEE machine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mrcnn_roi_b200.h"

#define MRCNN_EXPORT extern "C" __attribute__((visibility("default")))

namespace mrcnn {

constexpr int kMaxSort = MRCNN_MAX_SORT;

static inline int last_error() { return (int)cudaGetLastError(); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  Every kernel of the hot path is launched with the programmatic-stream-
// serialization attribute and (a) calls pdl_launch_dependents() first -- the next kernel of the stream may start
// being scheduled as soon as every CTA of this grid is resident -- and (b) calls pdl_wait() before its first access
// to global memory: that blocks until the preceding grid has completed and its writes are visible.  Because every
// kernel waits on ALL paths, completion is transitive along the stream (kernel N done => N-1 done), so a kernel
// may also safely overwrite buffers an earlier kernel read.  What overlaps is launch latency, CTA scheduling and
// the shared-memory / barrier set-up of kernel N+1 with the tail of kernel N.  MRCNN_PDL=0 disables the attribute
// (the device-side calls are then no-ops); a launch behind a memset or a foreign kernel degrades to a plain launch;
// launches captured into a CUDA graph are plain as well (pdl_attr).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__attribute__((visibility("hidden"))) bool pdl_enabled();   // api.cu: getenv("MRCNN_PDL") != "0", read once
// integer tuning knob for measurements (getenv on every call: no cached state); `fallback` when unset
__attribute__((visibility("hidden"))) int tuning_knob(const char* name, int fallback);

// fills `attr[0]` with the PDL attribute when enabled; returns the number of attributes written.  Not while `stream` is
// being captured into a CUDA graph: programmatic edges inside a replayed graph measured slower than plain edges
// (config 4, B = 32: 35.6 k against 39.9 k images/s; config 2: no difference), whereas on a live stream the attribute
// hides the launch gaps (config 2 eager: 23.4 k against 22.4 k) -- profiles/r2_pdl_graph.md.
static inline int pdl_attr(cudaLaunchAttribute* attr, cudaStream_t stream) {
    if (!pdl_enabled()) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (cap != cudaStreamCaptureStatusNone) return 0;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    return 1;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    cfg.numAttrs = (unsigned)pdl_attr(attr, stream);
    cfg.attrs = attr;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------------
// deterministic exp / log (<= 1 ulp); same operation sequence on every IEEE-754 machine
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pow2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }

__device__ __forceinline__ float det_expf(float x) {
    if (x != x) return x;
    if (x > 88.72283935546875f) return __uint_as_float(0x7f800000u);
    if (x < -103.972076416015625f) return 0.0f;
    const float m = floorf(__fmaf_rn(x, 1.44269504088896341f, 0.5f));
    float r = __fmaf_rn(m, -0.693359375f, x);
    r = __fmaf_rn(m, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    const float r2 = __fmul_rn(r, r);
    float y = __fmaf_rn(p, r2, r);
    y = __fadd_rn(y, 1.0f);
    const int mi = (int)m;
    const int m1 = mi >> 1;
    const int m2 = mi - m1;
    return __fmul_rn(__fmul_rn(y, pow2i(m1)), pow2i(m2));
}

__device__ __forceinline__ float det_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return __uint_as_float(0x7fc00000u);
    if (x == 0.0f) return __uint_as_float(0xff800000u);
    if (x == __uint_as_float(0x7f800000u)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = __fmul_rn(x, 8388608.0f); e = -23; }
    uint32_t u = __float_as_uint(x);
    e += (int)((u >> 23) & 0xffu) - 126;
    u = (u & 0x007fffffu) | 0x3f000000u;
    float m = __uint_as_float(u);
    if (m < 0.707106781186547524f) { e -= 1; m = __fadd_rn(__fadd_rn(m, m), -1.0f); }
    else { m = __fadd_rn(m, -1.0f); }
    const float z = __fmul_rn(m, m);
    float p = 7.0376836292e-2f;
    p = __fmaf_rn(p, m, -1.1514610310e-1f);
    p = __fmaf_rn(p, m, 1.1676998740e-1f);
    p = __fmaf_rn(p, m, -1.2420140846e-1f);
    p = __fmaf_rn(p, m, 1.4249322787e-1f);
    p = __fmaf_rn(p, m, -1.6668057665e-1f);
    p = __fmaf_rn(p, m, 2.0000714765e-1f);
    p = __fmaf_rn(p, m, -2.4999993993e-1f);
    p = __fmaf_rn(p, m, 3.3333331174e-1f);
    float y = __fmul_rn(__fmul_rn(m, z), p);
    const float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = __fadd_rn(m, y);
    r = __fmaf_rn(fe, 0.693359375f, r);
    return r;
}

// Eigen static_cast<int32>(float) on x86 (cvttss2si): NaN / out of range -> INT32_MIN
__device__ __forceinline__ int32_t cast_i32_x86(float v) {
    if (!(v >= -2147483648.0f && v < 2147483648.0f)) return (int32_t)0x80000000;
    return (int32_t)v;  // truncation
}

// ---------------------------------------------------------------------------------------------------
// utils.py:830-851 apply_box_deltas_graph (deltas already multiplied by std_dev) + utils.py:854-869 clip
// boxes are (y1, x1, y2, x2) in a float4 (x, y, z, w)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 apply_box_deltas(float4 b, float4 d) {
    float height = __fsub_rn(b.z, b.x);
    float width = __fsub_rn(b.w, b.y);
    float cy = __fadd_rn(b.x, __fmul_rn(0.5f, height));
    float cx = __fadd_rn(b.y, __fmul_rn(0.5f, width));
    cy = __fadd_rn(cy, __fmul_rn(d.x, height));
    cx = __fadd_rn(cx, __fmul_rn(d.y, width));
    height = __fmul_rn(height, det_expf(d.z));
    width = __fmul_rn(width, det_expf(d.w));
    float4 o;
    o.x = __fsub_rn(cy, __fmul_rn(0.5f, height));
    o.y = __fsub_rn(cx, __fmul_rn(0.5f, width));
    o.z = __fadd_rn(o.x, height);
    o.w = __fadd_rn(o.y, width);
    return o;
}

__device__ __forceinline__ float4 scale_deltas(float4 d, float4 sd) {
    return make_float4(__fmul_rn(d.x, sd.x), __fmul_rn(d.y, sd.y), __fmul_rn(d.z, sd.z), __fmul_rn(d.w, sd.w));
}

__device__ __forceinline__ float4 clip_box(float4 b, float4 w) {  // w = (wy1, wx1, wy2, wx2)
    b.x = fmaxf(fminf(b.x, w.z), w.x);
    b.y = fmaxf(fminf(b.y, w.w), w.y);
    b.z = fmaxf(fminf(b.z, w.z), w.x);
    b.w = fmaxf(fminf(b.w, w.w), w.y);
    return b;
}

// ---------------------------------------------------------------------------------------------------
// order-preserving key for fp32 scores: a > b  <=>  key(a) > key(b); -0 == +0; NaN lowest
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t score_key(float x) {
    if (x != x) return 0u;
    const uint32_t u = __float_as_uint(__fadd_rn(x, 0.0f));  // -0 -> +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr uint32_t kKeyNegInf = 0x007fffffu;  // score_key(-inf)

__device__ __forceinline__ uint64_t make_composite(uint32_t key, uint32_t idx) {
    return ((uint64_t)key << 32) | (uint64_t)(0xffffffffu - idx);  // sort descending: key desc, idx asc
}
__device__ __forceinline__ uint32_t composite_idx(uint64_t c) { return 0xffffffffu - (uint32_t)c; }
__device__ __forceinline__ uint32_t composite_key(uint64_t c) { return (uint32_t)(c >> 32); }

// in-place bitonic sort of n (power of two) elements in shared memory, descending; all threads of the block
template <typename T>
__device__ __forceinline__ void block_bitonic_sort_desc(T* s, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool desc = ((i & k) == 0);
                const T a = s[i], b = s[p];
                if ((a < b) == desc) { s[i] = b; s[p] = a; }
            }
            __syncthreads();
        }
    }
}

// Register-blocked bitonic sort, descending, of 1024*E 64-bit keys held E per thread (thread t owns ranks
// t*E .. t*E+E-1) by a 1024-thread block.  Compare distances below E stay in registers, distances inside a warp
// use shuffles, and only the log2(32) remaining distances per merge level go through shared memory (double
// buffered: one __syncthreads per such stage; per-thread rows padded by 16 bytes against bank conflicts).
// xch: 2 * 1024 * (E + 2) uint64_t of shared memory.
template <int E, int J>
__device__ __forceinline__ void sort_local_stage(uint64_t (&v)[E], int k, int t, int log_e) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if ((e & J) == 0 && (e | J) < E) {
            const bool desc = (((t << log_e) | e) & k) == 0;
            const uint64_t a = v[e], b = v[e | J];
            if ((a < b) == desc) { v[e] = b; v[e | J] = a; }
        }
    }
}

template <int E>
__device__ __forceinline__ void block_sort_desc_blocked(uint64_t (&v)[E], uint64_t* xch) {
    constexpr int LOG_E = (E == 1) ? 0 : (E == 2) ? 1 : (E == 4) ? 2 : 3;
    constexpr int N = 1024 * E;
    constexpr int ROW = E + 2;
    const int t = threadIdx.x;
    int buf = 0;
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32 * E) {  // partner element lives in another warp
                const int dm = j >> LOG_E;
                uint64_t* mine = xch + (size_t)buf * (1024 * ROW) + (size_t)t * ROW;
#pragma unroll
                for (int e = 0; e < E; ++e) mine[e] = v[e];
                __syncthreads();
                const uint64_t* theirs = xch + (size_t)buf * (1024 * ROW) + (size_t)(t ^ dm) * ROW;
                const bool lower = (t & dm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t p = theirs[e];
                    const bool desc = (((t << LOG_E) | e) & k) == 0;
                    const bool take_max = (lower == desc);
                    v[e] = take_max ? (v[e] > p ? v[e] : p) : (v[e] < p ? v[e] : p);
                }
                buf ^= 1;
            } else if (j >= E) {  // partner lane in the same warp
                const int lm = j >> LOG_E;
                const bool lower = (t & lm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const uint64_t p = __shfl_xor_sync(0xffffffffu, v[e], lm);
                    const bool desc = (((t << LOG_E) | e) & k) == 0;
                    const bool take_max = (lower == desc);
                    v[e] = take_max ? (v[e] > p ? v[e] : p) : (v[e] < p ? v[e] : p);
                }
            } else {  // both elements in this thread's registers (compile-time register indices)
                if (j == 4) sort_local_stage<E, 4>(v, k, t, LOG_E);
                else if (j == 2) sort_local_stage<E, 2>(v, k, t, LOG_E);
                else sort_local_stage<E, 1>(v, k, t, LOG_E);
            }
        }
    }
}
// Sort `sort_n` (power of two, 32..8192) composites held in shared memory `s`, descending, with a 1024-thread
// block.  Arrays of >= 1024 keys go through the register-blocked network (xch = 2*1024*(E+2) u64 of scratch shared
// memory that may NOT alias `s`); smaller ones use the in-place shared-memory network.
__device__ __forceinline__ void block_sort_desc_any(uint64_t* s, int sort_n, uint64_t* xch) {
    const int t = threadIdx.x;
    if (sort_n < 1024) { block_bitonic_sort_desc(s, sort_n); return; }
#define MRCNN_SORT_CASE(E)                                   \
    {                                                        \
        uint64_t v[E];                                       \
        _Pragma("unroll") for (int e = 0; e < E; ++e) v[e] = s[t * E + e]; \
        block_sort_desc_blocked<E>(v, xch);                  \
        _Pragma("unroll") for (int e = 0; e < E; ++e) s[t * E + e] = v[e]; \
        __syncthreads();                                     \
    }
    if (sort_n == 1024) MRCNN_SORT_CASE(1)
    else if (sort_n == 2048) MRCNN_SORT_CASE(2)
    else if (sort_n == 4096) MRCNN_SORT_CASE(4)
    else MRCNN_SORT_CASE(8)
#undef MRCNN_SORT_CASE
}
constexpr size_t block_sort_xch_bytes(int E) { return 2 * 1024 * (size_t)(E + 2) * sizeof(uint64_t); }

// ---------------------------------------------------------------------------------------------------
// TF NonMaxSuppressionV3 IoU test "inter / (a_i + a_j - inter) > thr" (non_max_suppression_op.cc IOU()).
// Boxes are min/max-normalised corners with precomputed areas; callers guarantee both areas > 0.
// The division is only executed when the multiplicative screen lands within 2^-21 relative of the
// threshold; outside that band the screen provably agrees with the rounded quotient.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool iou_gt(const float4& bi, float ai, const float4& bj, float aj, float thr) {
    const float iy0 = fmaxf(bi.x, bj.x), ix0 = fmaxf(bi.y, bj.y);
    const float iy1 = fminf(bi.z, bj.z), ix1 = fminf(bi.w, bj.w);
    const float dh = fmaxf(__fsub_rn(iy1, iy0), 0.0f);
    const float dw = fmaxf(__fsub_rn(ix1, ix0), 0.0f);
    const float inter = __fmul_rn(dh, dw);
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    const float t = __fmul_rn(thr, uni);
    const float d = __fsub_rn(inter, t);
    bool res = d > 0.0f;
    if (fabsf(d) <= __fmul_rn(t, 4.76837158203125e-07f)) res = __fdiv_rn(inter, uni) > thr;
    return res;
}

// Branch-free variant for hot loops: accumulates the multiplicative screen into `hit` and flags `unsure` when a
// pair lands inside the 2^-21 band; callers re-run the (rare) unsure lanes through iou_gt.
__device__ __forceinline__ void iou_screen(const float4& bi, float ai, const float4& bj, float aj, float thr, bool& hit,
                                           bool& unsure) {
    const float iy0 = fmaxf(bi.x, bj.x), ix0 = fmaxf(bi.y, bj.y);
    const float iy1 = fminf(bi.z, bj.z), ix1 = fminf(bi.w, bj.w);
    const float dh = fmaxf(__fsub_rn(iy1, iy0), 0.0f);
    const float dw = fmaxf(__fsub_rn(ix1, ix0), 0.0f);
    const float inter = __fmul_rn(dh, dw);
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    const float t = __fmul_rn(thr, uni);
    const float d = __fsub_rn(inter, t);
    hit |= d > 0.0f;
    unsure |= fabsf(d) <= __fmul_rn(t, 4.76837158203125e-07f);
}

// Cheaper screen for the NMS inner loops.  With tak = thr * area_k and tac = thr * area_c precomputed per box,
// d = inter * (1 + thr) - (tak + tac) is the same real quantity as inter - thr * union (one FFMA).  Its roundings
// (<= 2^-22 relative to tak + tac) plus those of the reference quotient (<= 2^-22) stay inside the band
// m = (tak + tac) * 2^-20: d > m is a certain hit, d < -m a certain miss, anything else is re-run through the exact
// iou_gt.  For an OR over many kept boxes only the running maximum of d is needed: max d > m_max is a certain hit,
// max d < -m_max a certain miss (m_max from the largest tak).  One clamp is enough: a negative width makes inter <= 0.
constexpr float kScreenBand = 9.5367431640625e-07f;  // 2^-20
__device__ __forceinline__ float iou_screen_d(const float4& bk, float tak, const float4& bc, float tac, float c1) {
    const float dh = fmaxf(__fsub_rn(fminf(bk.z, bc.z), fmaxf(bk.x, bc.x)), 0.0f);
    const float dw = __fsub_rn(fminf(bk.w, bc.w), fmaxf(bk.y, bc.y));
    return __fmaf_rn(__fmul_rn(dh, dw), c1, -__fadd_rn(tak, tac));
}
// Two candidates per lane, packed for the f32x2 pipe (FADD2 / FMUL2 / FFMA2 on sm_100): the screen of one kept box
// against both costs 8 FMNMX + 2 clamps + 5 packed operations instead of 20 scalar ones, with the SAME roundings as
// iou_screen_d (sub, mul, add, fma individually rounded; (-a) + (-b) == -(a + b) exactly).
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub_f2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add_f2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul_f2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
struct CandPair {   // candidates c0 (low halves) and c1 (high halves) of one lane
    float y1a, x1a, y2a, x2a, y1b, x1b, y2b, x2b;
    unsigned long long ntac;   // (-thr * area_c0, -thr * area_c1)
};
// d = inter * (1 + thr) - (tak + tac) for both candidates; ntak = -thr * area_k, c1 = (1 + thr, 1 + thr)
__device__ __forceinline__ void iou_screen_d2(const float4& bk, float ntak, const CandPair& c, unsigned long long c1,
                                              float& d0, float& d1) {
    const unsigned long long top = pack_f2(fminf(bk.z, c.y2a), fminf(bk.z, c.y2b));
    const unsigned long long bot = pack_f2(fmaxf(bk.x, c.y1a), fmaxf(bk.x, c.y1b));
    float h0, h1;
    unpack_f2(sub_f2(top, bot), h0, h1);
    const unsigned long long dh = pack_f2(fmaxf(h0, 0.0f), fmaxf(h1, 0.0f));
    const unsigned long long rgt = pack_f2(fminf(bk.w, c.x2a), fminf(bk.w, c.x2b));
    const unsigned long long lft = pack_f2(fmaxf(bk.y, c.x1a), fmaxf(bk.y, c.x1b));
    const unsigned long long inter = mul_f2(dh, sub_f2(rgt, lft));
    unpack_f2(fma_f2(inter, c1, add_f2(pack_f2(ntak, ntak), c.ntac)), d0, d1);
}
// The same for boxes known to lie inside the unit square (ProposalLayer: clip_boxes_graph U:854-869 clips to [0,0,1,1]):
// the overlap height top - bot is then <= 1, so the saturating subtract (FADD.SAT, fma pipe) IS max(top - bot, 0) -- same
// value, two FMNMX fewer on the alu pipe, which is what bounds the NMS sweep.
__device__ __forceinline__ float sub_sat(float a, float b) {
    float r;
    asm("sub.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <bool UNIT>
__device__ __forceinline__ void iou_screen_d2t(const float4& bk, float ntak, const CandPair& c, unsigned long long c1,
                                               float& d0, float& d1) {
    if (UNIT) {
        const unsigned long long dh = pack_f2(sub_sat(fminf(bk.z, c.y2a), fmaxf(bk.x, c.y1a)),
                                              sub_sat(fminf(bk.z, c.y2b), fmaxf(bk.x, c.y1b)));
        const unsigned long long rgt = pack_f2(fminf(bk.w, c.x2a), fminf(bk.w, c.x2b));
        const unsigned long long lft = pack_f2(fmaxf(bk.y, c.x1a), fmaxf(bk.y, c.x1b));
        const unsigned long long inter = mul_f2(dh, sub_f2(rgt, lft));
        unpack_f2(fma_f2(inter, c1, add_f2(pack_f2(ntak, ntak), c.ntac)), d0, d1);
    } else {
        iou_screen_d2(bk, ntak, c, c1, d0, d1);
    }
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {  // FMNMX3 (sm_100)
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {  // FMNMX3 (sm_100)
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// TF IOU(): corners normalised with min/max; returns area in `area`
__device__ __forceinline__ float4 normalise_box(float4 b, float& area) {
    float4 n;
    n.x = fminf(b.x, b.z);
    n.y = fminf(b.y, b.w);
    n.z = fmaxf(b.x, b.z);
    n.w = fmaxf(b.y, b.w);
    area = __fmul_rn(__fsub_rn(n.z, n.x), __fsub_rn(n.w, n.y));
    return n;
}

// TF crop_and_resize sampling tap along one axis (crop_and_resize_op.cc)
struct Tap {
    int lo, hi;
    float lerp;
    bool valid;
};
__device__ __forceinline__ Tap make_tap(float c1, float c2, int size, int crop, int t, float scale) {
    Tap r;
    float in;
    if (crop > 1) in = __fadd_rn(__fmul_rn(c1, (float)(size - 1)), __fmul_rn((float)t, scale));
    else in = (float)(0.5 * (double)__fadd_rn(c1, c2) * (double)(size - 1));
    r.valid = (in >= 0.0f && in <= (float)(size - 1));
    r.lo = r.valid ? (int)floorf(in) : 0;
    r.hi = r.valid ? (int)ceilf(in) : 0;
    r.lerp = __fsub_rn(in, (float)r.lo);
    return r;
}
__device__ __forceinline__ float crop_scale(float c1, float c2, int size, int crop) {
    return (crop > 1) ? __fdiv_rn(__fmul_rn(__fsub_rn(c2, c1), (float)(size - 1)), (float)(crop - 1)) : 0.0f;
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32); returns the
// exclusive prefix, writes the block total to *total.  `warp_sums` = 32 ints of shared memory.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    __syncthreads();  // protect warp_sums reuse
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < nwarps) ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += n;
        }
        warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    return warp_sums[warp] + incl - v;
}

// Per-device facts and occupancy queries.  The launchers keep no state that changes their results: the only
// process-wide data are these caches of pure device queries, keyed by device ordinal (and kernel / cluster size /
// shared-memory size) and filled under a mutex (api.cu), so concurrent first calls from several threads and several
// GPUs per process are safe.
struct DeviceProps {
    int sms;          // cudaDevAttrMultiProcessorCount
    int smem_optin;   // cudaDevAttrMaxSharedMemoryPerBlockOptin
};
__attribute__((visibility("hidden"))) DeviceProps device_props();   // of the current device
// cudaOccupancyMaxActiveClusters for `kernel` at cluster size cs (threads per CTA, dynamic shared memory `smem`);
// also raises the kernel's dynamic shared memory limit to `smem`.  Negative: the query failed.
__attribute__((visibility("hidden"))) int max_active_clusters(const void* kernel, int threads, int cs, size_t smem);
// static shared memory of a kernel (cudaFuncGetAttributes), cached
__attribute__((visibility("hidden"))) size_t static_smem_bytes(const void* kernel);

// Largest cluster size in {8,4,2,1} (bounded by max_cs) for which the driver can keep one cluster per image resident
// at the same time (max active clusters >= B): a batch whose clusters do not all co-schedule runs in waves and loses
// more than the wider clusters gain.  smem_for(cs) gives the dynamic shared memory of a CTA at that size.
template <typename SmemFn>
static int pick_cluster_size(const void* kernel, int threads, int B, int max_cs, SmemFn smem_for) {
    for (int cs = 16; cs > 1; cs >>= 1) {   // 16: non-portable cluster size (max_active_clusters opts the kernel in)
        if (cs > max_cs) continue;
        int active = max_active_clusters(kernel, threads, cs, smem_for(cs));
        if (active < 0) active = device_props().sms / cs;  // the query failed: assume one CTA per SM
        if (active >= B) return cs;
    }
    return 1;
}

// The same over every size a kernel that is generic in its cluster size can take (sizes need not be powers of two;
// above 8 is the non-portable range, which max_active_clusters opts the kernel into).
template <typename SmemFn>
static int pick_cluster_size_any(const void* kernel, int threads, int B, int max_cs, SmemFn smem_for) {
    static const int kSizes[] = {16, 14, 12, 10, 8, 6, 4, 3, 2};
    for (int cs : kSizes) {
        if (cs > max_cs) continue;
        int active = max_active_clusters(kernel, threads, cs, smem_for(cs));
        if (active < 0) active = device_props().sms / cs;
        if (active >= B) return cs;
    }
    return 1;
}

// ---- internal launchers shared between translation units (hidden visibility) -----------------------
struct NmsEpilogue {
    int mode;                    // 0 = indices, 1 = proposal boxes, 2 = detection rows
    // mode 0
    const int32_t* orig_idx;     // [B,M] sorted position -> original row (NULL = identity)
    int32_t* keep;               // [B,max_out]
    int32_t* count;              // [B]
    // mode 1
    float4* proposals;           // [B,max_out]
    // mode 2
    const float4* refined;       // [B,N] boxes by original row
    const float* scores;         // [B,N]
    const int32_t* class_ids;    // [B,N]
    float* detections;           // [B,max_out,6]
    float4* det_boxes;           // [B,max_out] optional: detections[..., :4] (DetectedBoxesExtraction, L:535-550)
    int N;
};

// boxes_sorted [B,M] in candidate order; valid [B] or NULL (= M); rows_ws: nms_rows_ws_bytes(B, M) of scratch or NULL
__attribute__((visibility("hidden"))) size_t nms_rows_ws_bytes(int B, int M);
__attribute__((visibility("hidden"))) int launch_nms_sorted(const float4* boxes_sorted, const int32_t* valid, int B,
                                                            int M, int max_out, float thr, const NmsEpilogue& epi,
                                                            void* rows_ws, cudaStream_t stream,
                                                            bool unit_boxes = false);

// candidates in input order: the kernel orders them itself by `keys` ([B,M] order-preserving score keys, 0 = not a
// candidate) or else by `scores` ([B,M]); only where nms_fused_applies(M, max_out) (single-CTA problems)
__attribute__((visibility("hidden"))) bool nms_fused_applies(int M, int max_out);
__attribute__((visibility("hidden"))) int launch_nms_unsorted(const float4* boxes, const float* scores,
                                                              const uint32_t* keys, const int32_t* valid, int B, int M,
                                                              int max_out, float thr, const NmsEpilogue& epi,
                                                              cudaStream_t stream);

constexpr int kMaxRpnLevels = 8;
struct TopkDecode {              // optional fused epilogue of the top-k final kernel (ProposalLayer)
    const float4* anchors;       // [B,A]
    const float4* deltas;        // [B,A] raw; NULL when the deltas come per pyramid level:
    const float4* level_deltas[kMaxRpnLevels];  // level l: [B, level_start[l+1] - level_start[l]] (rpn_bbox_pred outputs)
    int level_start[kMaxRpnLevels + 1];         // first anchor of each level in the concatenated order (L:1074-1091)
    int levels;
    float4 std_dev;
    float4* boxes_sorted;        // [B,K]
    float4* pre_nms_boxes;       // [B,K] optional copy
};
// raw deltas of anchor a of image b: from the concatenated [B,A,4] tensor, or straight from the level that owns it
__device__ __forceinline__ float4 topk_load_delta(const TopkDecode& dec, int b, int A, int a) {
    if (dec.deltas) return __ldg(dec.deltas + (size_t)b * A + a);
    int l = 0;
    while (l + 1 < dec.levels && a >= dec.level_start[l + 1]) ++l;
    const int n = dec.level_start[l + 1] - dec.level_start[l];
    return __ldg(dec.level_deltas[l] + (size_t)b * n + (a - dec.level_start[l]));
}
__attribute__((visibility("hidden"))) size_t topk_ws_bytes(int B);
__attribute__((visibility("hidden"))) int launch_topk(const float* scores, int stride, int offset, int B, int A, int K,
                                                      int32_t* idx, float* vals, const TopkDecode* dec, void* ws,
                                                      cudaStream_t stream);

}  // namespace mrcnn
