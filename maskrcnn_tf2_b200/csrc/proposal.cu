// proposal.cu -- ProposalLayer.call (mrcnn_layers.py:233-269) as one asynchronous launch sequence:
//   top-k over the foreground column of rpn_probs (L:235,246)  -> topk.cu (hist, hist, compact, final)
//   gather + std-dev scale + decode + clip (L:238,247-261)      -> fused into the top-k final kernel
//   NMS 0.7 + gather + zero pad (L:224-231,268)                 -> nms.cu (nms_lazy_kernel, proposal epilogue)
// No per-image Python loop (utils.batch_slice, utils.py:738-772): the whole batch is one grid per kernel.
#include "common.cuh"

using namespace mrcnn;

namespace {
struct ProposalWs {
    void* topk;
    float4* boxes_sorted;  // [B,K]
};
size_t proposal_ws_bytes(int B, int K) {
    return align_up(topk_ws_bytes(B), 256) + align_up((size_t)B * K * sizeof(float4), 256) + nms_rows_ws_bytes(B, K);
}
}  // namespace

namespace {
// one thread per output proposal row: recompute the forward decode of its source anchor, then TF's gradients
// (tf.minimum passes to x where x <= y, tf.maximum where x >= y; mul / exp in autodiff order)
__global__ void __launch_bounds__(256)
proposal_bwd_kernel(const float4* __restrict__ grad_proposals, const float4* __restrict__ rpn_bbox,
                    const float4* __restrict__ anchors, const int32_t* __restrict__ topk_idx,
                    const int32_t* __restrict__ keep_idx, int A, int K, int P, float4 sd,
                    float4* __restrict__ grad_rpn_bbox) {
    const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int k = keep_idx[(size_t)b * P + p];
    if (k < 0) return;  // zero-padded row (L:229-230)
    const int a = topk_idx[(size_t)b * K + k];
    const float4 an = __ldg(anchors + (size_t)b * A + a);
    const float4 d = scale_deltas(__ldg(rpn_bbox + (size_t)b * A + a), sd);
    const float4 g = __ldg(grad_proposals + (size_t)b * P + p);
    const float height = __fsub_rn(an.z, an.x), width = __fsub_rn(an.w, an.y);
    float cy = __fadd_rn(an.x, __fmul_rn(0.5f, height)), cx = __fadd_rn(an.y, __fmul_rn(0.5f, width));
    cy = __fadd_rn(cy, __fmul_rn(d.x, height));
    cx = __fadd_rn(cx, __fmul_rn(d.y, width));
    const float eh = det_expf(d.z), ew = det_expf(d.w);
    const float h2 = __fmul_rn(height, eh), w2 = __fmul_rn(width, ew);
    const float y1 = __fsub_rn(cy, __fmul_rn(0.5f, h2)), x1 = __fsub_rn(cx, __fmul_rn(0.5f, w2));
    const float y2 = __fadd_rn(y1, h2), x2 = __fadd_rn(x1, w2);
    auto pass = [](float v, float gv) { return (v <= 1.0f && fminf(v, 1.0f) >= 0.0f) ? gv : 0.0f; };  // clip to [0,1]
    const float g0 = pass(y1, g.x), g1 = pass(x1, g.y), g2 = pass(y2, g.z), g3 = pass(x2, g.w);
    const float gy1 = __fadd_rn(g0, g2), gx1 = __fadd_rn(g1, g3);
    const float gh2 = __fadd_rn(g2, __fmul_rn(-0.5f, gy1)), gw2 = __fadd_rn(g3, __fmul_rn(-0.5f, gx1));
    float4 o;
    o.x = __fmul_rn(__fmul_rn(gy1, height), sd.x);
    o.y = __fmul_rn(__fmul_rn(gx1, width), sd.y);
    o.z = __fmul_rn(__fmul_rn(__fmul_rn(gh2, height), eh), sd.z);
    o.w = __fmul_rn(__fmul_rn(__fmul_rn(gw2, width), ew), sd.w);
    grad_rpn_bbox[(size_t)b * A + a] = o;  // every anchor is selected at most once
}
}  // namespace

MRCNN_EXPORT int mrcnn_proposal_backward(const float* grad_proposals, const float* rpn_bbox, const float* anchors,
                                         const int32_t* topk_idx, const int32_t* keep_idx, int B, int A, int K, int P,
                                         const float* std_dev, float* grad_rpn_bbox, void* stream) {
    if (!grad_proposals || !rpn_bbox || !anchors || !topk_idx || !keep_idx || !std_dev || !grad_rpn_bbox)
        return MRCNN_ERR_NULL;
    if (B < 1 || A < 1 || K < 1 || K > A || K > kMaxSort || P < 1) return MRCNN_ERR_RANGE;
    if (!aligned16(grad_proposals) || !aligned16(rpn_bbox) || !aligned16(anchors) || !aligned16(grad_rpn_bbox))
        return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(grad_rpn_bbox, 0, (size_t)B * A * 4 * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    const float4 sd = make_float4(std_dev[0], std_dev[1], std_dev[2], std_dev[3]);
    proposal_bwd_kernel<<<dim3((P + 255) / 256, B), 256, 0, st>>>((const float4*)grad_proposals, (const float4*)rpn_bbox,
                                                               (const float4*)anchors, topk_idx, keep_idx, A, K, P, sd,
                                                               (float4*)grad_rpn_bbox);
    return last_error();
}

MRCNN_EXPORT int mrcnn_proposal_workspace_bytes(int B, int A, int pre_nms_limit, int P, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    const int K = pre_nms_limit < A ? pre_nms_limit : A;
    if (B < 1 || A < 1 || K < 1 || K > kMaxSort || P < 1) return MRCNN_ERR_RANGE;
    *bytes = proposal_ws_bytes(B, K);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_proposal_forward(const float* rpn_probs, const float* rpn_bbox, const float* anchors, int B,
                                        int A, int pre_nms_limit, int P, const float* std_dev, float nms_thr,
                                        float* proposals, int32_t* topk_idx, int32_t* keep_idx, int32_t* keep_count,
                                        float* pre_nms_boxes, void* ws, size_t ws_bytes, void* stream) {
    if (!rpn_probs || !rpn_bbox || !anchors || !std_dev || !proposals || !ws) return MRCNN_ERR_NULL;
    const int K = pre_nms_limit < A ? pre_nms_limit : A;  // L:245
    if (B < 1 || A < 1 || K < 1 || K > kMaxSort || P < 1 || !(nms_thr >= 0.0f && nms_thr <= 1.0f))
        return MRCNN_ERR_RANGE;
    if (ws_bytes < proposal_ws_bytes(B, K)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(rpn_bbox) || !aligned16(anchors) || !aligned16(proposals) || !aligned16(ws) ||
        (pre_nms_boxes && !aligned16(pre_nms_boxes)))
        return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    ProposalWs w;
    char* p = (char*)ws;
    w.topk = p;                  p += align_up(topk_ws_bytes(B), 256);
    w.boxes_sorted = (float4*)p; p += align_up((size_t)B * K * sizeof(float4), 256);
    void* rows_ws = p;

    TopkDecode dec{};
    dec.anchors = (const float4*)anchors;
    dec.deltas = (const float4*)rpn_bbox;
    dec.std_dev = make_float4(std_dev[0], std_dev[1], std_dev[2], std_dev[3]);
    dec.boxes_sorted = w.boxes_sorted;
    dec.pre_nms_boxes = (float4*)pre_nms_boxes;
    int rc = launch_topk(rpn_probs, 2, 1, B, A, K, topk_idx, nullptr, &dec, w.topk, st);
    if (rc != 0) return rc;

    NmsEpilogue epi{};
    epi.mode = 1;
    epi.proposals = (float4*)proposals;
    epi.keep = keep_idx;
    epi.count = keep_count;
    return launch_nms_sorted(w.boxes_sorted, nullptr, B, K, P, nms_thr, epi, rows_ws, st, /*unit_boxes=*/true);
}


// ---- ProposalLayer fed by the RPN head's per-level outputs ---------------------------
// The reference concatenates the five levels' rpn_class_logits / rpn_class / rpn_bbox (model.py:465-478) after a Keras
// softmax over [B,A_l,2] per level (mrcnn_layers.py:1081) and hands the concatenated tensors to ProposalLayer.  Here one
// kernel reads the level logits where they lie and writes the dense foreground-probability row [B,A] the top-k needs
// (TF SoftmaxEigenImpl: exp(l - max) * (1 / sum), exp = the shared deterministic sequence), and the decode epilogue of
// the top-k fetches the winners' deltas straight from the level tensors: no concatenated copy, no [B,A,2] softmax pass.
namespace {
struct RpnLevels {
    const float2* logits[kMaxRpnLevels];
    int start[kMaxRpnLevels + 1];
    int levels;
};

__global__ void __launch_bounds__(256)
rpn_fg_score_kernel(RpnLevels lv, int A, float* __restrict__ fg /*[B,A]*/, float2* __restrict__ probs /*[B,A] or null*/) {
    const int a = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    if (a >= A) return;
    int l = 0;
    while (l + 1 < lv.levels && a >= lv.start[l + 1]) ++l;
    const int n = lv.start[l + 1] - lv.start[l];
    const float2 x = __ldg(lv.logits[l] + (size_t)b * n + (a - lv.start[l]));
    const float m = fmaxf(x.x, x.y);
    const float e0 = det_expf(__fsub_rn(x.x, m)), e1 = det_expf(__fsub_rn(x.y, m));
    const float inv = __fdiv_rn(1.0f, __fadd_rn(e0, e1));
    const float p1 = __fmul_rn(e1, inv);
    fg[(size_t)b * A + a] = p1;
    if (probs) probs[(size_t)b * A + a] = make_float2(__fmul_rn(e0, inv), p1);
}
}  // namespace

MRCNN_EXPORT int mrcnn_proposal_levels_workspace_bytes(int B, int A, int pre_nms_limit, int P, size_t* bytes) {
    if (!bytes) return MRCNN_ERR_NULL;
    const int K = pre_nms_limit < A ? pre_nms_limit : A;
    if (B < 1 || A < 1 || K < 1 || K > kMaxSort || P < 1) return MRCNN_ERR_RANGE;
    *bytes = proposal_ws_bytes(B, K) + align_up((size_t)B * A * sizeof(float), 256);
    return MRCNN_OK;
}

MRCNN_EXPORT int mrcnn_proposal_forward_levels(const float* const* rpn_class_logits, const float* const* rpn_bbox,
                                               const int* level_anchors, int levels, const float* anchors, int B,
                                               int pre_nms_limit, int P, const float* std_dev, float nms_thr,
                                               float* proposals, float* rpn_probs, int32_t* topk_idx, int32_t* keep_idx,
                                               int32_t* keep_count, void* ws, size_t ws_bytes, void* stream) {
    if (!rpn_class_logits || !rpn_bbox || !level_anchors || !anchors || !std_dev || !proposals || !ws)
        return MRCNN_ERR_NULL;
    if (levels < 1 || levels > kMaxRpnLevels) return MRCNN_ERR_RANGE;
    RpnLevels lv{};
    TopkDecode dec{};
    long long acc = 0;
    for (int l = 0; l < levels; ++l) {
        if (!rpn_class_logits[l] || !rpn_bbox[l]) return MRCNN_ERR_NULL;
        if (level_anchors[l] < 1) return MRCNN_ERR_RANGE;
        if ((reinterpret_cast<uintptr_t>(rpn_class_logits[l]) & 7u) || !aligned16(rpn_bbox[l])) return MRCNN_ERR_ALIGN;
        lv.logits[l] = (const float2*)rpn_class_logits[l];
        dec.level_deltas[l] = (const float4*)rpn_bbox[l];
        lv.start[l] = dec.level_start[l] = (int)acc;
        acc += level_anchors[l];
        if (acc > (1 << 24)) return MRCNN_ERR_RANGE;
    }
    const int A = (int)acc;
    lv.start[levels] = dec.level_start[levels] = A;
    lv.levels = dec.levels = levels;
    const int K = pre_nms_limit < A ? pre_nms_limit : A;  // L:245
    if (B < 1 || K < 1 || K > kMaxSort || P < 1 || !(nms_thr >= 0.0f && nms_thr <= 1.0f)) return MRCNN_ERR_RANGE;
    if (ws_bytes < proposal_ws_bytes(B, K) + align_up((size_t)B * A * sizeof(float), 256)) return MRCNN_ERR_WORKSPACE;
    if (!aligned16(anchors) || !aligned16(proposals) || !aligned16(ws) || (reinterpret_cast<uintptr_t>(rpn_probs) & 7u))
        return MRCNN_ERR_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    char* p = (char*)ws;
    void* topk_ws = p;            p += align_up(topk_ws_bytes(B), 256);
    float4* boxes_sorted = (float4*)p; p += align_up((size_t)B * K * sizeof(float4), 256);
    void* rows_ws = p;                 p += nms_rows_ws_bytes(B, K);
    float* fg = (float*)p;
    rpn_fg_score_kernel<<<dim3((A + 255) / 256, B), 256, 0, st>>>(lv, A, fg, (float2*)rpn_probs);
    dec.anchors = (const float4*)anchors;
    dec.deltas = nullptr;
    dec.std_dev = make_float4(std_dev[0], std_dev[1], std_dev[2], std_dev[3]);
    dec.boxes_sorted = boxes_sorted;
    dec.pre_nms_boxes = nullptr;
    int rc = launch_topk(fg, 1, 0, B, A, K, topk_idx, nullptr, &dec, topk_ws, st);
    if (rc != 0) return rc;
    NmsEpilogue epi{};
    epi.mode = 1;
    epi.proposals = (float4*)proposals;
    epi.keep = keep_idx;
    epi.count = keep_count;
    return launch_nms_sorted(boxes_sorted, nullptr, B, K, P, nms_thr, epi, rows_ws, st, /*unit_boxes=*/true);
}
