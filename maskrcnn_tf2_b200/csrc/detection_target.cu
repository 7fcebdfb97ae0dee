// detection_target.cu -- DetectionTargetLayer.call + detection_targets_graph (mrcnn_layers.py:313-325, 844-1007).
//   dt_select_kernel : CTA per image -- trim zero rows (L:871-874), split crowds (L:879-884), IoU row max / first
//                      argmax over the GT boxes without materialising the [P*G,4] tiles of overlaps_graph
//                      (L:887,982-1007), crowd filter (L:890-892), positive / negative sets (L:895-900),
//                      subsampling in (rand_key asc, row asc) order standing in for tf.random.shuffle (L:904-910),
//                      ROI / class / box-refinement targets (L:912-925, utils.py:775-798), zero padding (L:958-965)
//   dt_mask_kernel   : CTA per target ROI -- 28x28 crop_and_resize of the assigned GT mask read in place from
//                      [MH,MW,G] (L:929-954); the reference's [#pos,MH,MW,1] gather (277 MB/image) never exists.
#include <float.h>

#include "common.cuh"

namespace mrcnn {

struct MaskJob {
    float4 box;   // crop box in mask coordinates
    int channel;  // GT channel, -1 = zero mask
    int pad[3];
};

__device__ __forceinline__ float plain_iou(const float4& a, const float4& b) {  // overlaps_graph L:982-1007
    const float y1 = fmaxf(a.x, b.x), x1 = fmaxf(a.y, b.y);
    const float y2 = fminf(a.z, b.z), x2 = fminf(a.w, b.w);
    const float inter = __fmul_rn(fmaxf(__fsub_rn(x2, x1), 0.0f), fmaxf(__fsub_rn(y2, y1), 0.0f));
    const float a1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float a2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
}

__device__ __forceinline__ bool nonzero_box(const float4& b) {
    return (fabsf(b.x) + fabsf(b.y) + fabsf(b.z) + fabsf(b.w)) != 0.0f;
}

__global__ void __launch_bounds__(1024)
dt_select_kernel(const float4* __restrict__ proposals, const int32_t* __restrict__ gt_class_ids,
                 const float4* __restrict__ gt_boxes, const uint32_t* __restrict__ rand_keys, int P, int G, int T,
                 int positive_cap, float inv_ratio, float4 std_dev, int use_mini_masks, int sort_n,
                 float4* __restrict__ rois, int32_t* __restrict__ class_ids, float4* __restrict__ deltas,
                 MaskJob* __restrict__ jobs, int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) uint64_t s[];         // sort_n composites
    uint16_t* s_arg = reinterpret_cast<uint16_t*>(s + sort_n);        // [P] assigned GT (compacted index)
    float4* s_gt = reinterpret_cast<float4*>(s_arg + ((P + 7) & ~7));  // [G]
    float4* s_crowd = s_gt + G;                                       // [G]
    int32_t* s_cls = reinterpret_cast<int32_t*>(s_crowd + G);         // [G]
    int32_t* s_chan = s_cls + G;                                      // [G]
    __shared__ int warp_sums[32];
    __shared__ int s_ng, s_ncrowd, s_npos, s_nneg, s_tot;
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) { s_npos = 0; s_nneg = 0; }

    // ---- GT: ordered compaction of non-zero, non-crowd boxes (order matters: argmax = first maximum) ----
    {
        int base_gt = 0, base_cr = 0;
        for (int g0 = 0; g0 < G; g0 += blockDim.x) {
            const int g = g0 + tid;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            int cls = 0;
            if (g < G) { q = __ldg(gt_boxes + (size_t)b * G + g); cls = gt_class_ids[(size_t)b * G + g]; }
            const bool nz = (g < G) && nonzero_box(q);
            const bool is_gt = nz && cls > 0, is_cr = nz && cls < 0;
            const int r_gt = block_exclusive_scan(is_gt ? 1 : 0, warp_sums, &s_tot);
            const int tot_gt = s_tot;
            const int r_cr = block_exclusive_scan(is_cr ? 1 : 0, warp_sums, &s_tot);
            const int tot_cr = s_tot;
            if (is_gt) { s_gt[base_gt + r_gt] = q; s_cls[base_gt + r_gt] = cls; s_chan[base_gt + r_gt] = g; }
            if (is_cr) s_crowd[base_cr + r_cr] = q;
            base_gt += tot_gt;
            base_cr += tot_cr;
            __syncthreads();
        }
        if (tid == 0) { s_ng = base_gt; s_ncrowd = base_cr; }
    }
    __syncthreads();
    const int ng = s_ng, ncrowd = s_ncrowd;

    // ---- proposals: IoU row max / argmax, positive / negative flags, shuffle composites ----------------
    int lpos = 0, lneg = 0;
    for (int i = tid; i < sort_n; i += blockDim.x) {
        uint64_t comp = 0ull;
        if (i < P) {
            const float4 pb = __ldg(proposals + (size_t)b * P + i);
            if (nonzero_box(pb)) {
                float best = -FLT_MAX;
                int arg = 0;
                for (int g = 0; g < ng; ++g) {
                    const float v = plain_iou(pb, s_gt[g]);
                    if (v > best) { best = v; arg = g; }
                }
                float cmax = -FLT_MAX;
                for (int c = 0; c < ncrowd; ++c) cmax = fmaxf(cmax, plain_iou(pb, s_crowd[c]));
                s_arg[i] = (uint16_t)arg;
                const bool pos = best >= 0.5f;
                const bool neg = (best < 0.5f) && (cmax < 0.001f);
                if (pos || neg) {
                    const uint32_t key = rand_keys[(size_t)b * P + i];
                    comp = (pos ? (1ull << 63) : 0ull) | ((uint64_t)(~key) << 31) | (uint64_t)(0x7fffffffu - (uint32_t)i);
                    lpos += pos ? 1 : 0;
                    lneg += pos ? 0 : 1;
                }
            }
        }
        s[i] = comp;
    }
    if (lpos) atomicAdd(&s_npos, lpos);
    if (lneg) atomicAdd(&s_nneg, lneg);
    __syncthreads();
    block_bitonic_sort_desc(s, sort_n);  // positives first; inside each class (key asc, row asc)
    const int npos_all = s_npos, nneg_all = s_nneg;
    int npos = min(npos_all, positive_cap);                                           // L:904-906
    int nneg = cast_i32_x86(__fmul_rn(inv_ratio, (float)npos)) - npos;                // L:908-909
    nneg = min(max(nneg, 0), nneg_all);                                               // L:910
    npos = min(npos, T);
    nneg = min(nneg, T - npos);

    for (int t = tid; t < T; t += blockDim.x) {
        const size_t o = (size_t)b * T + t;
        float4 roi = make_float4(0.f, 0.f, 0.f, 0.f), dl = roi;
        int cls = 0;
        MaskJob job;
        job.box = roi;
        job.channel = -1;
        job.pad[0] = job.pad[1] = job.pad[2] = 0;
        if (t < npos) {
            const int i = (int)(0x7fffffffu - (uint32_t)(s[t] & 0x7fffffffull));
            const int g = s_arg[i];
            roi = __ldg(proposals + (size_t)b * P + i);
            const float4 gt = s_gt[g];
            cls = s_cls[g];
            // utils.box_refinement_graph U:775-798, then /= bbox_std_dev (L:925)
            const float height = __fsub_rn(roi.z, roi.x), width = __fsub_rn(roi.w, roi.y);
            const float cy = __fadd_rn(roi.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(roi.y, __fmul_rn(0.5f, width));
            const float gh = __fsub_rn(gt.z, gt.x), gw = __fsub_rn(gt.w, gt.y);
            const float gcy = __fadd_rn(gt.x, __fmul_rn(0.5f, gh)), gcx = __fadd_rn(gt.y, __fmul_rn(0.5f, gw));
            dl.x = __fdiv_rn(__fdiv_rn(__fsub_rn(gcy, cy), height), std_dev.x);
            dl.y = __fdiv_rn(__fdiv_rn(__fsub_rn(gcx, cx), width), std_dev.y);
            dl.z = __fdiv_rn(det_logf(__fdiv_rn(gh, __fadd_rn(height, 1e-3f))), std_dev.z);
            dl.w = __fdiv_rn(det_logf(__fdiv_rn(gw, __fadd_rn(width, 1e-3f))), std_dev.w);
            job.box = roi;
            if (use_mini_masks) {  // L:935-946: ROI in the GT box's normalised mini-mask frame
                job.box.x = __fdiv_rn(__fsub_rn(roi.x, gt.x), gh);
                job.box.y = __fdiv_rn(__fsub_rn(roi.y, gt.y), gw);
                job.box.z = __fdiv_rn(__fsub_rn(roi.z, gt.x), gh);
                job.box.w = __fdiv_rn(__fsub_rn(roi.w, gt.y), gw);
            }
            job.channel = s_chan[g];
        } else if (t < npos + nneg) {
            const int i = (int)(0x7fffffffu - (uint32_t)(s[npos_all + (t - npos)] & 0x7fffffffull));
            roi = __ldg(proposals + (size_t)b * P + i);
        }
        rois[o] = roi;
        class_ids[o] = cls;
        deltas[o] = dl;
        jobs[o] = job;
    }
    if (tid == 0 && counts) { counts[2 * b] = npos; counts[2 * b + 1] = nneg; }
}

__global__ void __launch_bounds__(256)
dt_mask_kernel(const uint8_t* __restrict__ gt_masks, const MaskJob* __restrict__ jobs, int T, int G, int MH, int MW,
               int mask_h, int mask_w, float* __restrict__ masks) {
    const int t = blockIdx.x, b = blockIdx.y;
    const MaskJob job = jobs[(size_t)b * T + t];
    float* out = masks + ((size_t)b * T + t) * mask_h * mask_w;
    const int npix = mask_h * mask_w;
    if (job.channel < 0) {
        for (int i = threadIdx.x; i < npix; i += blockDim.x) out[i] = 0.0f;
        return;
    }
    const uint8_t* m = gt_masks + (size_t)b * MH * MW * G + job.channel;
    const float hs = crop_scale(job.box.x, job.box.z, MH, mask_h);
    const float ws = crop_scale(job.box.y, job.box.w, MW, mask_w);
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int y = i / mask_w, x = i - y * mask_w;
        const Tap ty = make_tap(job.box.x, job.box.z, MH, mask_h, y, hs);
        const Tap tx = make_tap(job.box.y, job.box.w, MW, mask_w, x, ws);
        float v = 0.0f;
        if (ty.valid && tx.valid) {
            const float tl = m[((size_t)ty.lo * MW + tx.lo) * G] ? 1.0f : 0.0f;
            const float tr = m[((size_t)ty.lo * MW + tx.hi) * G] ? 1.0f : 0.0f;
            const float bl = m[((size_t)ty.hi * MW + tx.lo) * G] ? 1.0f : 0.0f;
            const float br = m[((size_t)ty.hi * MW + tx.hi) * G] ? 1.0f : 0.0f;
            const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), tx.lerp));
            const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), tx.lerp));
            v = rintf(__fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ty.lerp)));  // tf.round (L:954)
        }
        out[i] = v;
    }
}

static size_t dt_smem_bytes(int P, int G, int sort_n) {
    return (size_t)sort_n * 8 + (size_t)((P + 7) & ~7) * 2 + (size_t)G * (16 + 16 + 4 + 4);
}

}  // namespace mrcnn

using namespace mrcnn;

MRCNN_EXPORT int mrcnn_detection_target_workspace_bytes(int B, int P, int G, int T, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    if (B < 1 || P < 1 || P > kMaxSort || G < 1 || G > MRCNN_MAX_GT || T < 1) return MRCNN_ERR_RANGE;
    *bytes = align_up((size_t)B * T * sizeof(MaskJob), 256);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_detection_target_forward(const float* proposals, const int32_t* gt_class_ids,
                                                const float* gt_boxes, const uint8_t* gt_masks,
                                                const uint32_t* rand_keys, int B, int P, int G, int MH, int MW, int T,
                                                double roi_positive_ratio, const float* bbox_std_dev, int mask_h,
                                                int mask_w, int use_mini_masks, float* rois, int32_t* class_ids,
                                                float* deltas, float* masks, int32_t* counts, void* ws,
                                                size_t ws_bytes, void* stream) {
    if (!proposals || !gt_class_ids || !gt_boxes || !gt_masks || !rand_keys || !bbox_std_dev || !rois || !class_ids ||
        !deltas || !masks || !ws)
        return MRCNN_ERR_NULL;
    if (B < 1 || P < 1 || P > kMaxSort || G < 1 || G > MRCNN_MAX_GT || T < 1 || MH < 1 || MW < 1 || mask_h < 1 ||
        mask_w < 1 || !(roi_positive_ratio > 0.0 && roi_positive_ratio <= 1.0))
        return MRCNN_ERR_RANGE;
    if (ws_bytes < align_up((size_t)B * T * sizeof(MaskJob), 256)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(proposals) || !aligned16(gt_boxes) || !aligned16(rois) || !aligned16(deltas) || !aligned16(ws))
        return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int positive_cap = (int)((double)T * roi_positive_ratio);   // L:904, Python float arithmetic
    const float inv_ratio = (float)(1.0 / roi_positive_ratio);        // L:908, then an fp32 multiply (L:909)
    const int sort_n = next_pow2(P < 32 ? 32 : P);
    const size_t smem = dt_smem_bytes(P, G, sort_n);
    cudaError_t e = cudaFuncSetAttribute(dt_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const float4 sd = make_float4(bbox_std_dev[0], bbox_std_dev[1], bbox_std_dev[2], bbox_std_dev[3]);
    MaskJob* jobs = (MaskJob*)ws;
    dt_select_kernel<<<B, 1024, smem, st>>>((const float4*)proposals, gt_class_ids, (const float4*)gt_boxes, rand_keys,
                                            P, G, T, positive_cap, inv_ratio, sd, use_mini_masks, sort_n, (float4*)rois,
                                            class_ids, (float4*)deltas, jobs, counts);
    dt_mask_kernel<<<dim3(T, B), 256, 0, st>>>(gt_masks, jobs, T, G, MH, MW, mask_h, mask_w, masks);
    return last_error();
}
