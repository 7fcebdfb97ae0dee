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
    return align_up(topk_ws_bytes(B), 256) + align_up((size_t)B * K * sizeof(float4), 256);
}
}  // namespace

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
    w.boxes_sorted = (float4*)p;

    TopkDecode dec;
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
    return launch_nms_sorted(w.boxes_sorted, nullptr, B, K, P, nms_thr, epi, st);
}
