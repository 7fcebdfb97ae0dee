// rpn_targets.cu -- utils.build_rpn_targets (utils.py:154-262) with compute_overlaps / compute_iou (utils.py:114-151)
// for a whole padded batch on the device.  The reference runs it per image in numpy on the data-loader thread
// (preprocess.py:342-348): an [A, G] float64 IoU matrix (261 888 x G at 1024^2), argmax / max / `overlaps == max`
// scans, np.random.choice subsampling and a Python loop over the positive anchors.
//
// Here (all float64, every operation individually rounded in numpy's order, so rpn_match is bit-exact):
//   rpn_colmax_kernel   thread per anchor: IoU against every non-crowd instance, column maxima (np.max(overlaps, 0),
//                       U:208) through shared-memory then global 64-bit atomicMax on order-preserving keys;
//   rpn_match_kernel    thread per anchor: IoU again (cheaper than an [A,G] float64 matrix in HBM), first-max argmax
//                       (U:203-204), crowd test (U:184-188), the three matching rules (U:205-211); writes the
//                       unsampled match, the argmax and two score arrays for the subsampling;
//   top-k x2            (topk.cu) the R/2 positives and the R negatives with the largest injected key -- the stand-in
//                       for np.random.choice (U:215-228); non-members carry distinct scores -(a+1) so the radix
//                       select never sees a tie flood;
//   rpn_finish_kernel   thread per anchor: resets a class to neutral where it is oversubscribed;
//   rpn_bbox_kernel     CTA per image: re-marks the chosen anchors, sorts the kept positives by anchor index and
//                       writes their float64 refinements (U:232-260), zero padded to R rows.
// No [A,G] matrix, no host round trip; HBM traffic is ~5 passes over [B,A] words.
#include "common.cuh"

namespace mrcnn {
namespace {

constexpr int kRtThreads = 256;

__device__ __forceinline__ unsigned long long f64_key(double v) {  // a > b <=> key(a) > key(b); -0 == +0
    const unsigned long long u = (unsigned long long)__double_as_longlong(__dadd_rn(v, 0.0));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct GtBox {       // one ground-truth row staged in shared memory
    double y1, x1, y2, x2, area;
    int role;        // 0 = padding / ignored, 1 = instance, 2 = crowd
    int pad;
};

// U:125-133 compute_iou, numpy float64: maximum/minimum, max(.,0) * max(.,0), (area_g + area_a) - inter, divide
__device__ __forceinline__ double np_iou(const double a0, const double a1, const double a2, const double a3,
                                         const double area_a, const GtBox& g) {
    const double y1 = fmax(g.y1, a0), y2 = fmin(g.y2, a2), x1 = fmax(g.x1, a1), x2 = fmin(g.x2, a3);
    const double dx = fmax(__dsub_rn(x2, x1), 0.0), dy = fmax(__dsub_rn(y2, y1), 0.0);
    const double inter = __dmul_rn(dx, dy);
    const double uni = __dsub_rn(__dadd_rn(g.area, area_a), inter);
    // most (anchor, instance) pairs are disjoint: 0 / uni is +0 for every positive union, without the fp64 division
    return (inter == 0.0 && uni > 0.0) ? 0.0 : __ddiv_rn(inter, uni);
}

// stage the image's GT rows; U:175-182: crowds (class < 0) are split off from the instances (class > 0); class 0
// rows are batch padding (the loader passes real instances only, preprocess.py:342-348)
__device__ __forceinline__ void stage_gt(const int32_t* __restrict__ cls, const int32_t* __restrict__ box, int G,
                                         GtBox* sg) {
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const int c = cls[g];
        const int4 b = *reinterpret_cast<const int4*>(box + 4 * g);
        GtBox r;
        r.y1 = (double)b.x; r.x1 = (double)b.y; r.y2 = (double)b.z; r.x2 = (double)b.w;
        r.area = (double)((b.z - b.x) * (b.w - b.y));  // U:143 int32 product
        r.role = (c < 0) ? 2 : (c > 0 ? 1 : 0);
        r.pad = 0;
        sg[g] = r;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRtThreads)
rpn_colmax_kernel(const double2* __restrict__ anchors, const int32_t* __restrict__ gt_class_ids,
                  const int32_t* __restrict__ gt_boxes, int A, int G, unsigned long long* __restrict__ colmax) {
    extern __shared__ __align__(16) unsigned char rt_smem[];
    GtBox* sg = reinterpret_cast<GtBox*>(rt_smem);
    unsigned long long* scm = reinterpret_cast<unsigned long long*>(sg + G);
    const int b = blockIdx.y, a = blockIdx.x * kRtThreads + threadIdx.x;
    for (int g = threadIdx.x; g < G; g += kRtThreads) scm[g] = 0ull;
    stage_gt(gt_class_ids + (size_t)b * G, gt_boxes + (size_t)b * G * 4, G, sg);
    if (a < A) {
        const double2 lo = __ldg(anchors + 2 * (size_t)a), hi = __ldg(anchors + 2 * (size_t)a + 1);
        const double area_a = __dmul_rn(__dsub_rn(hi.x, lo.x), __dsub_rn(hi.y, lo.y));  // U:142
        for (int g = 0; g < G; ++g) {
            if (sg[g].role != 1) continue;
            const unsigned long long k = f64_key(np_iou(lo.x, lo.y, hi.x, hi.y, area_a, sg[g]));
            if (k > scm[g]) atomicMax(&scm[g], k);
        }
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += kRtThreads)
        if (scm[g]) atomicMax(&colmax[(size_t)b * G + g], scm[g]);
}

__global__ void __launch_bounds__(kRtThreads)
rpn_match_kernel(const double2* __restrict__ anchors, const int32_t* __restrict__ gt_class_ids,
                 const int32_t* __restrict__ gt_boxes, const float* __restrict__ rand_keys, int A, int G,
                 const unsigned long long* __restrict__ colmax, int32_t* __restrict__ match,
                 int32_t* __restrict__ argmax, float* __restrict__ pos_score, float* __restrict__ neg_score,
                 int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char rt_smem[];
    GtBox* sg = reinterpret_cast<GtBox*>(rt_smem);
    unsigned long long* scm = reinterpret_cast<unsigned long long*>(sg + G);
    __shared__ int s_cnt[2];
    const int b = blockIdx.y, a = blockIdx.x * kRtThreads + threadIdx.x;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    for (int g = threadIdx.x; g < G; g += kRtThreads) scm[g] = colmax[(size_t)b * G + g];
    stage_gt(gt_class_ids + (size_t)b * G, gt_boxes + (size_t)b * G * 4, G, sg);
    int m = 0;
    if (a < A) {
        const double2 lo = __ldg(anchors + 2 * (size_t)a), hi = __ldg(anchors + 2 * (size_t)a + 1);
        const double area_a = __dmul_rn(__dsub_rn(hi.x, lo.x), __dsub_rn(hi.y, lo.y));
        double best = 0.0, crowd_max = -1.0;
        int besti = -1;
        bool ties_colmax = false, has_crowd = false;
        for (int g = 0; g < G; ++g) {
            const int role = sg[g].role;
            if (role == 0) continue;
            const double v = np_iou(lo.x, lo.y, hi.x, hi.y, area_a, sg[g]);
            if (role == 1) {
                if (besti < 0 || v > best) { best = v; besti = g; }      // np.argmax: first maximum (U:203)
                ties_colmax |= (f64_key(v) == scm[g]);                   // overlaps == np.max(overlaps, axis=0) (U:208)
            } else {
                has_crowd = true;
                crowd_max = fmax(crowd_max, v);                          // np.amax(crowd_overlaps, axis=1) (U:186)
            }
        }
        const bool no_crowd = !has_crowd || crowd_max < 0.001;           // U:187-191
        if (best < 0.3 && no_crowd) m = -1;                              // U:205
        if (ties_colmax) m = 1;                                          // U:209
        if (besti >= 0 && best >= 0.7) m = 1;                            // U:211
        const size_t o = (size_t)b * A + a;
        match[o] = m;
        argmax[o] = besti < 0 ? 0 : besti;
        const float key = __ldg(rand_keys + o), filler = -(float)(a + 1);  // distinct, below every key >= 0
        pos_score[o] = (m == 1) ? key : filler;
        neg_score[o] = (m == -1) ? key : filler;
    }
    const unsigned pb = __ballot_sync(0xffffffffu, m == 1), nb = __ballot_sync(0xffffffffu, m == -1);
    if ((threadIdx.x & 31) == 0) {
        if (pb) atomicAdd(&s_cnt[0], __popc(pb));
        if (nb) atomicAdd(&s_cnt[1], __popc(nb));
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counts[2 * b + threadIdx.x], s_cnt[threadIdx.x]);
}

// U:215-228: an oversubscribed class is reset to neutral here and its chosen members re-marked by rpn_bbox_kernel
__global__ void __launch_bounds__(kRtThreads)
rpn_finish_kernel(int32_t* __restrict__ match, const int32_t* __restrict__ counts, int A, int R) {
    const int b = blockIdx.y, a = blockIdx.x * kRtThreads + threadIdx.x;
    if (a >= A) return;
    const int npos = counts[2 * b], nneg = counts[2 * b + 1];
    const int keep_pos = min(npos, R / 2);
    const bool drop_pos = npos > R / 2, drop_neg = nneg > R - keep_pos;
    if (!drop_pos && !drop_neg) return;
    const size_t o = (size_t)b * A + a;
    const int m = match[o];
    if ((m == 1 && drop_pos) || (m == -1 && drop_neg)) match[o] = 0;
}

__global__ void __launch_bounds__(kRtThreads)
rpn_bbox_kernel(const double2* __restrict__ anchors, const int32_t* __restrict__ gt_boxes,
                const int32_t* __restrict__ argmax, const int32_t* __restrict__ counts,
                const int32_t* __restrict__ pos_idx, const int32_t* __restrict__ neg_idx, int A, int G, int R, int Kp,
                int Kn, double sd0, double sd1, double sd2, double sd3, double eps, int32_t* __restrict__ match,
                double* __restrict__ rpn_bbox, float* __restrict__ rpn_bbox_f32, int32_t* __restrict__ counts_out) {
    extern __shared__ __align__(16) unsigned char rt_smem[];
    int32_t* s = reinterpret_cast<int32_t*>(rt_smem);  // [sort_n] descending sort of (INT_MAX - anchor index)
    const int b = blockIdx.x, tid = threadIdx.x;
    const int npos = counts[2 * b], nneg = counts[2 * b + 1];
    const int keep_pos = min(min(npos, R / 2), Kp);
    const int keep_neg = min(min(nneg, R - keep_pos), Kn);
    if (npos > R / 2)
        for (int j = tid; j < keep_pos; j += kRtThreads) match[(size_t)b * A + pos_idx[(size_t)b * Kp + j]] = 1;
    if (nneg > R - keep_pos)
        for (int j = tid; j < keep_neg; j += kRtThreads) match[(size_t)b * A + neg_idx[(size_t)b * Kn + j]] = -1;
    if (tid == 0 && counts_out) { counts_out[2 * b] = keep_pos; counts_out[2 * b + 1] = keep_neg; }
    const int sort_n = max(32, 1 << (32 - __clz(max(Kp, 1) - 1)));
    for (int j = tid; j < sort_n; j += kRtThreads)
        s[j] = (j < keep_pos) ? (0x7fffffff - pos_idx[(size_t)b * Kp + j]) : -1;
    __syncthreads();
    block_bitonic_sort_desc(s, sort_n);                // ids = np.where(rpn_match == 1)[0]: ascending (U:232)
    for (int r = tid; r < R; r += kRtThreads) {
        double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;
        if (r < keep_pos) {
            const int a = 0x7fffffff - s[r];
            const int g = argmax[(size_t)b * A + a];                     // closest instance, maybe IoU < 0.7 (U:238)
            const int4 gb = *reinterpret_cast<const int4*>(gt_boxes + ((size_t)b * G + g) * 4);
            const double2 lo = __ldg(anchors + 2 * (size_t)a), hi = __ldg(anchors + 2 * (size_t)a + 1);
            const double gt_h = (double)(gb.z - gb.x), gt_w = (double)(gb.w - gb.y);            // U:242-243 (int32)
            const double gcy = __dadd_rn((double)gb.x, __dmul_rn(0.5, gt_h));
            const double gcx = __dadd_rn((double)gb.y, __dmul_rn(0.5, gt_w));
            const double a_h = __dsub_rn(hi.x, lo.x), a_w = __dsub_rn(hi.y, lo.y);
            const double acy = __dadd_rn(lo.x, __dmul_rn(0.5, a_h)), acx = __dadd_rn(lo.y, __dmul_rn(0.5, a_w));
            o0 = __ddiv_rn(__ddiv_rn(__dsub_rn(gcy, acy), a_h), sd0);                             // U:253-259
            o1 = __ddiv_rn(__ddiv_rn(__dsub_rn(gcx, acx), a_w), sd1);
            o2 = __ddiv_rn(log(__ddiv_rn(gt_h, __dadd_rn(a_h, eps))), sd2);
            o3 = __ddiv_rn(log(__ddiv_rn(gt_w, __dadd_rn(a_w, eps))), sd3);
        }
        double2* o = reinterpret_cast<double2*>(rpn_bbox + ((size_t)b * R + r) * 4);
        o[0] = make_double2(o0, o1);
        o[1] = make_double2(o2, o3);
        if (rpn_bbox_f32)
            reinterpret_cast<float4*>(rpn_bbox_f32)[(size_t)b * R + r] = make_float4((float)o0, (float)o1, (float)o2, (float)o3);
    }
}

struct RtWs {
    unsigned long long* colmax;  // [B,G]
    int32_t* counts;             // [B,2] unsampled (positives, negatives); zeroed together with colmax
    int32_t* argmax;             // [B,A]
    float* pos_score;            // [B,A]
    float* neg_score;            // [B,A]
    int32_t* pos_idx;            // [B,Kp]
    int32_t* neg_idx;            // [B,Kn]
    void* topk;
};
size_t rt_zero_bytes(int B, int G) { return align_up((size_t)B * G * 8 + (size_t)B * 2 * 4, 256); }
size_t rt_ws_bytes(int B, int A, int G, int Kp, int Kn) {
    return rt_zero_bytes(B, G) + 3 * align_up((size_t)B * A * 4, 256) + align_up((size_t)B * Kp * 4, 256) +
           align_up((size_t)B * Kn * 4, 256) + align_up(topk_ws_bytes(B), 256);
}
RtWs rt_carve(void* ws, int B, int A, int G, int Kp, int Kn) {
    RtWs w;
    char* p = (char*)ws;
    w.colmax = (unsigned long long*)p;
    w.counts = (int32_t*)(p + (size_t)B * G * 8);
    p += rt_zero_bytes(B, G);
    w.argmax = (int32_t*)p;   p += align_up((size_t)B * A * 4, 256);
    w.pos_score = (float*)p;  p += align_up((size_t)B * A * 4, 256);
    w.neg_score = (float*)p;  p += align_up((size_t)B * A * 4, 256);
    w.pos_idx = (int32_t*)p;  p += align_up((size_t)B * Kp * 4, 256);
    w.neg_idx = (int32_t*)p;  p += align_up((size_t)B * Kn * 4, 256);
    w.topk = p;
    return w;
}
bool rt_shape_ok(int B, int A, int G, int R) {
    return B >= 1 && A >= 1 && A <= (1 << 24) && G >= 1 && G <= MRCNN_MAX_GT && R >= 2 && R <= kMaxSort;
}

}  // namespace
}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_rpn_targets_workspace_bytes(int B, int A, int G, int R, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (!rt_shape_ok(B, A, G, R)) return MRCNN_ERR_RANGE;
    *bytes = rt_ws_bytes(B, A, G, (R / 2 < A) ? R / 2 : A, R < A ? R : A);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_rpn_targets_forward(const double* anchors, const int32_t* gt_class_ids, const int32_t* gt_boxes,
                                           const float* rand_keys, int B, int A, int G, int R, const double* std_dev,
                                           double eps, int32_t* rpn_match, double* rpn_bbox, float* rpn_bbox_f32,
                                           int32_t* counts, void* ws, size_t ws_bytes, void* stream) {
    if (!anchors || !gt_class_ids || !gt_boxes || !rand_keys || !std_dev || !rpn_match || !rpn_bbox || !ws)
        return MRCNN_ERR_NULL;
    if (!rt_shape_ok(B, A, G, R)) return MRCNN_ERR_RANGE;
    const int Kp = (R / 2 < A) ? R / 2 : A, Kn = R < A ? R : A;
    if (ws_bytes < rt_ws_bytes(B, A, G, Kp, Kn)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(anchors) || !aligned16(gt_boxes) || !aligned16(rpn_bbox) || !aligned16(ws) ||
        (rpn_bbox_f32 && !aligned16(rpn_bbox_f32)))
        return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const RtWs w = rt_carve(ws, B, A, G, Kp, Kn);
    cudaError_t e = cudaMemsetAsync(w.colmax, 0, rt_zero_bytes(B, G), st);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)((A + kRtThreads - 1) / kRtThreads), (unsigned)B);
    const size_t smem = (size_t)G * (sizeof(GtBox) + sizeof(unsigned long long));
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(rpn_colmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(rpn_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    rpn_colmax_kernel<<<grid, kRtThreads, smem, st>>>((const double2*)anchors, gt_class_ids, gt_boxes, A, G, w.colmax);
    rpn_match_kernel<<<grid, kRtThreads, smem, st>>>((const double2*)anchors, gt_class_ids, gt_boxes, rand_keys, A, G,
                                                    w.colmax, rpn_match, w.argmax, w.pos_score, w.neg_score, w.counts);
    int rc = launch_topk(w.pos_score, 1, 0, B, A, Kp, w.pos_idx, nullptr, nullptr, w.topk, st);
    if (rc) return rc;
    rc = launch_topk(w.neg_score, 1, 0, B, A, Kn, w.neg_idx, nullptr, nullptr, w.topk, st);
    if (rc) return rc;
    rpn_finish_kernel<<<grid, kRtThreads, 0, st>>>(rpn_match, w.counts, A, R);
    const int sort_n = next_pow2(Kp < 32 ? 32 : Kp);
    rpn_bbox_kernel<<<B, kRtThreads, (size_t)sort_n * sizeof(int32_t), st>>>(
        (const double2*)anchors, gt_boxes, w.argmax, w.counts, w.pos_idx, w.neg_idx, A, G, R, Kp, Kn, std_dev[0],
        std_dev[1], std_dev[2], std_dev[3], eps, rpn_match, rpn_bbox, rpn_bbox_f32, counts);
    return last_error();
}
