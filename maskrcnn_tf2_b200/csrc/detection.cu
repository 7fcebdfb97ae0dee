// detection.cu -- DetectionLayer.call + refine_detections (mrcnn_layers.py:369-524).
//   detection_refine_kernel : warp per ROI -- argmax over classes (first maximum, L:385), class-specific delta
//                             gather (L:388-393), std-dev scale + decode + clip to the image window (L:396-398,
//                             window normalised with image 0's shape, L:513-515), keep flag (L:402-414)
//   nms.cu nms_lazy_kernel  : ONE class-agnostic NMS (quirk Q3, L:440-468): the kernel first orders the kept ROIs by
//                             (score desc, ROI index asc), the candidate order of the NMS at L:455 -- only the ~15 % of
//                             the ROIs that are candidates are sorted, in shared memory --, and ends with the detection
//                             epilogue that packs [y1,x1,y2,x2,class,score] and zero-pads (L:494-500).
//   detection_sort_kernel   : (N > 2048 only: the ordering as a kernel of its own in front of a cluster NMS)
//                             The two O(n^2)
//                             broadcast intersections (L:411-414, 475-478) and the final top_k (L:486-490) are
//                             identities on this ordering and have no kernel.
#include "common.cuh"

namespace mrcnn {

__global__ void __launch_bounds__(256)
detection_refine_kernel(const float4* __restrict__ rois, const float* __restrict__ probs,
                        const float4* __restrict__ deltas, const float* __restrict__ image_meta, int meta_len, int N,
                        int NC, float4 std_dev, float min_conf, int use_min_conf, float4* __restrict__ refined,
                        float* __restrict__ scores, int32_t* __restrict__ class_ids, uint32_t* __restrict__ keep_key) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= N) return;
    const float* p = probs + ((size_t)b * N + i) * NC;
    float best = __ldg(p);  // class 0 (every lane)
    int arg = 0;
    for (int c = lane; c < NC; c += 32) {
        const float v = __ldg(p + c);
        if (v > best) { best = v; arg = c; }  // strict: first maximum within the lane's ascending classes
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if (lane == 0) {
        // window: (meta.window - [0,0,1,1]) / ([h,w,h,w] - 1) with h,w of image 0 (L:34-39, 513-515)
        const float sh = __fsub_rn(image_meta[4], 1.0f), sw = __fsub_rn(image_meta[5], 1.0f);
        const float* wm = image_meta + (size_t)b * meta_len + 7;
        float4 win;
        win.x = __fdiv_rn(__fsub_rn(wm[0], 0.0f), sh);
        win.y = __fdiv_rn(__fsub_rn(wm[1], 0.0f), sw);
        win.z = __fdiv_rn(__fsub_rn(wm[2], 1.0f), sh);
        win.w = __fdiv_rn(__fsub_rn(wm[3], 1.0f), sw);
        const size_t r = (size_t)b * N + i;
        const float4 d = scale_deltas(__ldg(deltas + r * NC + arg), std_dev);
        const float4 box = clip_box(apply_box_deltas(__ldg(rois + r), d), win);
        refined[r] = box;
        scores[r] = best;
        class_ids[r] = arg;
        const bool keep = (arg > 0) && (!use_min_conf || best >= min_conf);
        const uint32_t key = score_key(best);
        keep_key[r] = (keep && key > kKeyNegInf) ? key : 0u;  // 0 = not an NMS candidate
    }
}

__global__ void __launch_bounds__(1024)
detection_sort_kernel(const float4* __restrict__ refined, const uint32_t* __restrict__ keep_key, int N,
                      float4* __restrict__ boxes_sorted, int32_t* __restrict__ orig_idx, int32_t* __restrict__ ncand) {
    extern __shared__ __align__(16) uint64_t s[];
    __shared__ int s_n;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int sort_n = max(32, 1 << (32 - __clz(max(N, 1) - 1)));
    pdl_launch_dependents();
    pdl_wait();
    if (tid == 0) s_n = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < sort_n; i += blockDim.x) {
        uint64_t comp = 0ull;
        if (i < N) {
            const uint32_t key = keep_key[(size_t)b * N + i];
            if (key) { comp = make_composite(key, (uint32_t)i); ++local; }
        }
        s[i] = comp;
    }
    if (local) atomicAdd(&s_n, local);
    __syncthreads();
    block_sort_desc_any(s, sort_n, s + sort_n);
    const int nc = s_n;
    for (int r = tid; r < nc; r += blockDim.x) {
        const int i = (int)composite_idx(s[r]);
        orig_idx[(size_t)b * N + r] = i;
        boxes_sorted[(size_t)b * N + r] = refined[(size_t)b * N + i];
    }
    if (tid == 0) ncand[b] = nc;
}

struct DetWs {
    float4* refined;
    float* scores;
    int32_t* class_ids;
    uint32_t* keep_key;
    float4* boxes_sorted;
    int32_t* orig_idx;
    int32_t* ncand;
};
static size_t det_ws_bytes(int B, int N) {
    const size_t bn = (size_t)B * N;
    return 2 * align_up(bn * sizeof(float4), 256) + 4 * align_up(bn * 4, 256) + align_up((size_t)B * 4, 256);
}

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_detection_workspace_bytes(int B, int N, int NC, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1 || N > kMaxSort || NC < 1) return MRCNN_ERR_RANGE;
    *bytes = det_ws_bytes(B, N);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_detection_forward(const float* rois, const float* probs, const float* deltas,
                                         const float* image_meta, int meta_len, int B, int N, int NC,
                                         const float* std_dev, float min_conf, int use_min_conf, int max_inst,
                                         float nms_thr, int per_class, float* detections, int32_t* det_count,
                                         float* det_boxes, void* ws, size_t ws_bytes, void* stream) {
    if (!rois || !probs || !deltas || !image_meta || !std_dev || !detections || !ws) return MRCNN_ERR_NULL;
    if (B < 1 || N < 1 || N > kMaxSort || NC < 1 || meta_len < 11 || max_inst < 1 || per_class != 0 ||
        !(nms_thr >= 0.0f && nms_thr <= 1.0f))
        return MRCNN_ERR_RANGE;
    if (ws_bytes < det_ws_bytes(B, N)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(rois) || !aligned16(deltas) || !aligned16(ws) || !aligned16(det_boxes)) return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bn = (size_t)B * N;
    DetWs w;
    char* p = (char*)ws;
    w.refined = (float4*)p;      p += align_up(bn * sizeof(float4), 256);
    w.boxes_sorted = (float4*)p; p += align_up(bn * sizeof(float4), 256);
    w.scores = (float*)p;        p += align_up(bn * 4, 256);
    w.class_ids = (int32_t*)p;   p += align_up(bn * 4, 256);
    w.keep_key = (uint32_t*)p;   p += align_up(bn * 4, 256);
    w.orig_idx = (int32_t*)p;    p += align_up(bn * 4, 256);
    w.ncand = (int32_t*)p;

    const float4 sd = make_float4(std_dev[0], std_dev[1], std_dev[2], std_dev[3]);
    cudaError_t e = launch_pdl(detection_refine_kernel, dim3((N + 7) / 8, B), dim3(256), 0, st, (const float4*)rois, probs,
                               (const float4*)deltas, image_meta, meta_len, N, NC, sd, min_conf, use_min_conf, w.refined,
                               w.scores, w.class_ids, w.keep_key);
    if (e != cudaSuccess) return (int)e;
    NmsEpilogue epi{};
    epi.mode = 2;
    epi.refined = w.refined;
    epi.scores = w.scores;
    epi.class_ids = w.class_ids;
    epi.detections = detections;
    epi.count = det_count;
    epi.det_boxes = (float4*)det_boxes;
    epi.N = N;
    if (nms_fused_applies(N, max_inst))   // N <= 2048: ordering + NMS + packing in one single-CTA-per-image kernel
        return launch_nms_unsorted(w.refined, nullptr, w.keep_key, nullptr, B, N, max_inst, nms_thr, epi, st);
    const int sort_n = next_pow2(N < 32 ? 32 : N);
    const size_t smem = (size_t)sort_n * sizeof(uint64_t) + (sort_n >= 1024 ? block_sort_xch_bytes(sort_n / 1024) : 0);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(detection_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    e = launch_pdl(detection_sort_kernel, dim3(B), dim3(1024), smem, st, (const float4*)w.refined,
                   (const uint32_t*)w.keep_key, N, w.boxes_sorted, w.orig_idx, w.ncand);
    if (e != cudaSuccess) return (int)e;
    epi.orig_idx = w.orig_idx;
    return launch_nms_sorted(w.boxes_sorted, w.ncand, B, N, max_inst, nms_thr, epi, nullptr, st);
}
