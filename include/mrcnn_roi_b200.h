/*
 * mrcnn_roi_b200.h -- extern "C" launcher ABI of libmrcnn_roi_b200.so: the B200 (sm_100a) replacement for
 * the ROI-stage hot path of miguelalejo/maskrcnn_tf2.
 *
 * The reference has no FFI for this path: its four Keras layers (src/layers/mrcnn_layers.py) call stock
 * TensorFlow ops.  The boundary a maintainer binds is therefore the layer `call` bodies; each entry point
 * below names the reference lines it replaces.  INTEGRATION.md shows the TF custom-op shim and the ctypes
 * binding that sit on top of this header.
 *
 * Conventions (all entry points):
 *   - every tensor pointer is DEVICE memory unless the comment says "host"; tensors are dense, row-major,
 *     fp32 / int32 / uint8 / uint32 as typed; float4-read tensors (boxes, anchors, deltas, feature maps,
 *     outputs) must be 16-byte aligned (MRCNN_ERR_ALIGN otherwise);
 *   - the caller owns every buffer including `ws`; launchers never allocate, never keep a pointer after
 *     returning, never synchronise the host and never touch any stream but `stream` (a cudaStream_t passed
 *     as void*; NULL = the legacy default stream).  They are re-entrant and hold no mutable globals, so
 *     one process may drive several GPUs / streams concurrently (cudaSetDevice is the caller's job);
 *   - outputs are fully written (zero / -1 padding included); `ws` needs no initialisation;
 *   - return value: 0 = success; negative = argument error (below); positive = a cudaError_t observed by
 *     cudaGetLastError() after the launches.  No exceptions, no abort, no printf.
 *   - workspace size queries are pure host functions of the shape arguments.
 */
#ifndef MRCNN_ROI_B200_H_
#define MRCNN_ROI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRCNN_OK 0
#define MRCNN_ERR_NULL (-1)      /* a required pointer is NULL */
#define MRCNN_ERR_RANGE (-2)     /* a size/threshold is outside the supported range */
#define MRCNN_ERR_WORKSPACE (-3) /* ws_bytes smaller than the matching *_workspace_bytes() */
#define MRCNN_ERR_ALIGN (-4)     /* a vector-accessed pointer is not 16-byte aligned */

/* hard limits of this build */
#define MRCNN_MAX_SORT 8192      /* K (pre-NMS top-k), NMS candidates M and DetectionTarget proposals P */
#define MRCNN_MAX_GT 1024        /* DetectionTarget GT instances per image */

const char* mrcnn_roi_b200_version(void);
const char* mrcnn_status_string(int rc);

/* ---- tf.nn.top_k(scores, K, sorted=True).indices  (mrcnn_layers.py:246) ------------------------------
 * scores: element (b, a) is at scores[(b*A + a)*stride + offset]; stride/offset let the caller point at
 * the foreground column of rpn_probs [B,A,2] (stride 2, offset 1) without a slice copy (L:235).
 * idx [B,K] int32: descending value, ties -> lower index (TF TopKV2).  vals [B,K] optional (NULL ok).
 * Requires 1 <= K <= min(A, MRCNN_MAX_SORT).  NaN scores are ordered below every number. */
int mrcnn_topk_workspace_bytes(int B, int A, int K, size_t* bytes);
int mrcnn_topk_forward(const float* scores, int stride, int offset, int B, int A, int K, int32_t* idx,
                       float* vals, void* ws, size_t ws_bytes, void* stream);

/* ---- tf.image.non_max_suppression(boxes, scores, max_out, thr)  (mrcnn_layers.py:225,455) -------------
 * Batched: boxes [B,M,4], scores [B,M]; valid [B] (optional) limits image b to its first valid[b] rows.
 * keep [B,max_out] int32 indices into the M rows in selection order, -1 padded; count [B].
 * TF NonMaxSuppressionV3 CPU semantics: candidates score > -inf, order (score desc, index asc), suppressed
 * iff IoU > thr with IoU = inter/(a_i+a_j-inter) as a true fp32 division, zero-area boxes have IoU 0.
 * Requires 1 <= M <= MRCNN_MAX_SORT, 0 <= thr <= 1. */
int mrcnn_nms_workspace_bytes(int B, int M, size_t* bytes);
int mrcnn_nms_forward(const float* boxes, const float* scores, const int32_t* valid, int B, int M, int max_out,
                      float thr, int32_t* keep, int32_t* count, void* ws, size_t ws_bytes, void* stream);

/* ---- ProposalLayer.call  (mrcnn_layers.py:233-269, nms 224-231; utils.py:830-869) ----------------------
 * rpn_probs [B,A,2], rpn_bbox [B,A,4] (raw, multiplied by std_dev inside, L:238), anchors [B,A,4]
 * normalised.  K = min(pre_nms_limit, A) (L:245).  proposals [B,P,4]: decoded, clipped to [0,1], NMS'ed,
 * zero-padded rows after the kept ones (L:229-230).
 * Optional debug outputs (NULL ok): topk_idx [B,K]; keep_idx [B,P] (positions in the top-k order, -1
 * padded); keep_count [B]; pre_nms_boxes [B,K,4].  std_dev: host pointer to 4 floats. */
int mrcnn_proposal_workspace_bytes(int B, int A, int pre_nms_limit, int P, size_t* bytes);
int mrcnn_proposal_forward(const float* rpn_probs, const float* rpn_bbox, const float* anchors, int B, int A,
                           int pre_nms_limit, int P, const float* std_dev, float nms_thr, float* proposals,
                           int32_t* topk_idx, int32_t* keep_idx, int32_t* keep_count, float* pre_nms_boxes,
                           void* ws, size_t ws_bytes, void* stream);

/* ---- ProposalLayer fed by the RPN head's per-level outputs (rpn_graph mrcnn_layers.py:1052-1093 called per pyramid
 * level and concatenated at model.py:465-478) ---------------------------------------------------------------------
 * rpn_class_logits / rpn_bbox: host arrays of `levels` device pointers; level l holds [B,A_l,2] raw logits (the
 * reshaped rpn_class_raw output) and [B,A_l,4] raw deltas (rpn_bbox_pred), level_anchors[l] = A_l; the concatenated
 * anchor order is level-major, which is the order of `anchors` [B,A,4], A = sum A_l.  The foreground probability is
 * Keras' softmax over the two logits (TF SoftmaxEigenImpl: exp(l - max) * (1 / sum)); no concatenated tensor and no
 * [B,A,2] softmax pass is materialised unless rpn_probs [B,A,2] (optional: the model's `rpn_class` output) is given.
 * Everything after the scores is mrcnn_proposal_forward: same proposals, same optional index outputs. */
int mrcnn_proposal_levels_workspace_bytes(int B, int A, int pre_nms_limit, int P, size_t* bytes);
int mrcnn_proposal_forward_levels(const float* const* rpn_class_logits, const float* const* rpn_bbox,
                                  const int* level_anchors, int levels, const float* anchors, int B, int pre_nms_limit,
                                  int P, const float* std_dev, float nms_thr, float* proposals, float* rpn_probs,
                                  int32_t* topk_idx, int32_t* keep_idx, int32_t* keep_count, void* ws, size_t ws_bytes,
                                  void* stream);

/* ---- gradient of ProposalLayer.call w.r.t. rpn_bbox (TF autodiff through L:227 gather, utils.py:854-869 clip,
 * utils.py:830-851 decode, L:238 scale; the reference has no stop_gradient on the proposals: model.py:155-157,168).
 * topk_idx [B,K] and keep_idx [B,P] are the tensors mrcnn_proposal_forward returned for the same inputs.
 * grad_rpn_bbox [B,A,4] is zero-filled by the launcher; anchors and rpn_probs receive no gradient. */
int mrcnn_proposal_backward(const float* grad_proposals, const float* rpn_bbox, const float* anchors,
                            const int32_t* topk_idx, const int32_t* keep_idx, int B, int A, int K, int P,
                            const float* std_dev, float* grad_rpn_bbox, void* stream);

/* ---- PyramidROIAlign.call  (mrcnn_layers.py:583-664; utils.py:825-827) ---------------------------------
 * boxes [B,N,4] normalised; image_meta [B,meta_len] (only row 0, columns 4..5 = image h,w are read,
 * L:600); fmaps: host array of 4 device pointers P2..P5, each [B,H[l],W[l],C] NHWC fp32; C % 4 == 0.
 * out [B,N,ph,pw,C] written directly in input ROI order.  roi_map [B,N] int32 (required; saved for the
 * backward pass): index 0..3 of the feature map each ROI was sampled from.
 * map_mode 0 = reference behaviour: map index = first-appearance rank of the ROI's FPN level over the
 * flattened [B*N] batch (L:613-619,641); 1 = canonical level-2.  denominator: 244.0 in the reference
 * (L:574).  roi_level [B,N] optional. */
int mrcnn_roialign_workspace_bytes(int B, int N, size_t* bytes);
int mrcnn_roialign_forward(const float* boxes, const float* image_meta, int meta_len, const float* const* fmaps,
                           const int* H, const int* W, int C, int B, int N, int ph, int pw, float denominator,
                           int map_mode, float* out, int32_t* roi_map, int32_t* roi_level, void* ws,
                           size_t ws_bytes, void* stream);

/* ---- PyramidROIAlign with the feature maps in PINNED HOST memory: demand-driven staging ---------------------------
 * For callers whose P2..P5 start in host memory (this repository's end-to-end benchmark; a TF input pipeline that
 * keeps maps on the host).  Instead of copying all four maps (89 MB per image at 1024^2) before
 * mrcnn_roialign_forward, mrcnn_roialign_fetch_hostmaps marks the map pixels the given ROIs sample (the forward
 * kernel's own taps) and copies exactly those, once each, from host_fmaps (host array of 4 pointers to page-locked,
 * device-accessible host maps: cudaHostAlloc / cudaHostRegister under UVA) into dev_fmaps (host array of 4 device
 * staging maps of the same shapes).  mrcnn_roialign_forward on dev_fmaps then returns bit-identical results.
 * resident: device bitmap of mrcnn_roialign_resident_words() 32-bit words (one bit per map pixel of the batch): a set
 * bit = the staging pixel already holds the host pixel; pass reset != 0 when the host maps hold new contents (first
 * call of a step), 0 to fetch only what is still missing (the mask branch's call on the detections).
 * fetched: optional device counter, incremented by the number of pixels copied (each C * 4 bytes). */
int mrcnn_roialign_resident_words(int B, const int* H, const int* W, size_t* words);
int mrcnn_roialign_fetch_workspace_bytes(int B, int N, const int* H, const int* W, size_t* bytes);
int mrcnn_roialign_fetch_hostmaps(const float* boxes, const float* image_meta, int meta_len,
                                  const float* const* host_fmaps, float* const* dev_fmaps, const int* H, const int* W,
                                  int C, int B, int N, int ph, int pw, float denominator, int map_mode,
                                  uint32_t* resident, int reset, unsigned long long* fetched, void* ws,
                                  size_t ws_bytes, void* stream);

/* ---- gradient of PyramidROIAlign w.r.t. the four feature maps (TF CropAndResizeGradImage through the
 * reference's gather/concat, M:142,168; boxes get no gradient, L:628-629).  grad_fmaps: host array of 4
 * device pointers; every element is written by the launcher (no pre-clearing needed).
 * ws != NULL (size from mrcnn_roialign_backward_workspace_bytes): deterministic mode -- every gradient-map pixel
 * is written once with the sum of its samples in TF CropAndResizeGradImage's own sequential order
 * ((box, y, x, corner) ascending), bit-identical to the CPU kernel and reproducible run to run.  The exception:
 * pixels that collect more than 1024 samples or lie under a zero-size ROI (zero-padded ROIs pile thousands of
 * samples on pixel (0,0)) are accumulated with fp32 vector atomics.
 * ws == NULL: atomic mode -- zero-fill, then one fp32 vector reduction per sample corner (order not fixed). */
int mrcnn_roialign_backward_workspace_bytes(int B, int N, int ph, int pw, const int* H, const int* W, int C,
                                            size_t* bytes);
int mrcnn_roialign_backward(const float* grad_out, const float* boxes, const int32_t* roi_map,
                            float* const* grad_fmaps, const int* H, const int* W, int C, int B, int N, int ph,
                            int pw, void* ws, size_t ws_bytes, void* stream);

/* ---- DetectionLayer.call + refine_detections  (mrcnn_layers.py:369-524) -------------------------------
 * rois [B,N,4], probs [B,N,NC], deltas [B,N,NC,4], image_meta [B,meta_len] (window = columns 7..10 of each
 * row, normalised with image 0's h,w, L:513-515).  use_min_conf mirrors the Python truthiness test at
 * L:404.  per_class 0 = reference behaviour (one class-agnostic NMS, L:440-468); 1 is rejected
 * (MRCNN_ERR_RANGE) in this build.  detections [B,max_inst,6] = (y1,x1,y2,x2,class,score), zero padded;
 * det_count [B] optional.  det_boxes [B,max_inst,4] optional: detections[..., :4], the tensor
 * DetectedBoxesExtraction (L:535-550) feeds to the mask branch's PyramidROIAlign, written by the same kernel so the
 * slice copy disappears.  Requires N <= MRCNN_MAX_SORT. */
int mrcnn_detection_workspace_bytes(int B, int N, int NC, size_t* bytes);
int mrcnn_detection_forward(const float* rois, const float* probs, const float* deltas, const float* image_meta,
                            int meta_len, int B, int N, int NC, const float* std_dev, float min_conf,
                            int use_min_conf, int max_inst, float nms_thr, int per_class, float* detections,
                            int32_t* det_count, float* det_boxes, void* ws, size_t ws_bytes, void* stream);

/* ---- DetectionTargetLayer.call + detection_targets_graph  (mrcnn_layers.py:313-325, 844-1007) ----------
 * proposals [B,P,4]; gt_class_ids [B,G] int32 (crowds negative); gt_boxes [B,G,4] normalised; gt_masks
 * [B,MH,MW,G] uint8 (tf.bool layout); rand_keys [B,P] uint32: the injected stand-in for the reference's
 * unseeded tf.random.shuffle (L:905,910) -- positives / negatives are taken in (key asc, row asc) order.
 * Outputs: rois [B,T,4], class_ids [B,T] int32, deltas [B,T,4], masks [B,T,mask_h,mask_w] fp32 in {0,1},
 * all zero padded; counts [B,2] (positives, negatives) optional.  roi_positive_ratio is a double because the
 * reference evaluates int(T * ratio) and 1.0 / ratio in Python floats (L:904,908).  bbox_std_dev: host
 * pointer to 4 floats.  Requires P <= MRCNN_MAX_SORT, G <= MRCNN_MAX_GT. */
int mrcnn_detection_target_workspace_bytes(int B, int P, int G, int T, size_t* bytes);
int mrcnn_detection_target_forward(const float* proposals, const int32_t* gt_class_ids, const float* gt_boxes,
                                   const uint8_t* gt_masks, const uint32_t* rand_keys, int B, int P, int G,
                                   int MH, int MW, int T, double roi_positive_ratio, const float* bbox_std_dev,
                                   int mask_h, int mask_w, int use_mini_masks, float* rois, int32_t* class_ids,
                                   float* deltas, float* masks, int32_t* counts, void* ws, size_t ws_bytes,
                                   void* stream);

/* ---- utils.generate_pyramid_anchors (utils.py:54-111) + AnchorsLayer.get_anchors / NormBoxesLayer.call
 * (mrcnn_layers.py:34-39,116-132): the anchor constant, built on the device -----------------------------------------
 * scales / strides / feat_h / feat_w: host arrays of `levels` entries (rpn_anchor_scales, backbone_strides,
 * utils.compute_backbone_shapes); ratios: host array of `nratios` doubles.  Order: level-major, row-major (y, x),
 * ratio innermost.  anchors_px [A,4] float64 (optional): generate_pyramid_anchors' output bit for bit (what the data
 * loader hands to build_rpn_targets); anchors_norm [B,A,4] fp32 (optional): the AnchorsLayer output (fp32 cast, then
 * (a - [0,0,1,1]) / ([h,w,h,w] - 1) in fp32, broadcast to the batch) that ProposalLayer consumes.
 * mrcnn_anchors_count gives A.  Requires levels <= 8, nratios <= 8, A <= 2^24. */
int mrcnn_anchors_count(const int* feat_h, const int* feat_w, int levels, int nratios, int anchor_stride, int* count);
int mrcnn_anchors_forward(const double* scales, const double* ratios, const int* feat_h, const int* feat_w,
                          const int* strides, int levels, int nratios, int anchor_stride, int img_h, int img_w, int B,
                          double* anchors_px, float* anchors_norm, void* stream);

/* ---- utils.build_rpn_targets  (utils.py:154-262; compute_overlaps / compute_iou utils.py:114-151) -------------
 * The data loader's per-image numpy routine (preprocess.py:342-348), for a padded batch in one launch sequence.
 * anchors [A,4] float64 PIXEL boxes, as utils.generate_pyramid_anchors returns them and the loader keeps them
 * (preprocess.py:82,297), shared by the batch; gt_class_ids [B,G] int32 (0 = padding row, negative = COCO crowd);
 * gt_boxes [B,G,4] int32 pixel boxes (utils.extract_bboxes); rand_keys [B,A] fp32, finite and >= 0 (e.g. uniform
 * [0,1)): the injected stand-in for np.random.choice (utils.py:219,227) -- of an oversubscribed class the anchors
 * with the largest key are kept, ties -> lower anchor index.
 * Arithmetic is float64 in numpy's operation order, so rpn_match [B,A] int32 (1 / -1 / 0) and the first two
 * columns of rpn_bbox are bit-identical to the reference's; the log columns differ by the last ulp of log().
 * rpn_bbox [B,R,4] float64 (R = rpn_train_anchors_per_image): refinements of the kept positives in ascending anchor
 * order, divided by std_dev (host pointer to 4 doubles), zero padded.  rpn_bbox_f32 [B,R,4] optional: the same
 * rounded to fp32 (what the model's float32 input_rpn_bbox receives, model.py:420).  counts [B,2] optional: kept
 * positives / negatives.  An image without instances (the loader skips those, preprocess.py:336-338) yields
 * negatives only.  Requires 2 <= R <= MRCNN_MAX_SORT, G <= MRCNN_MAX_GT, A <= 2^24, y2 >= y1 and x2 >= x1. */
int mrcnn_rpn_targets_workspace_bytes(int B, int A, int G, int R, size_t* bytes);
int mrcnn_rpn_targets_forward(const double* anchors, const int32_t* gt_class_ids, const int32_t* gt_boxes,
                              const float* rand_keys, int B, int A, int G, int R, const double* std_dev, double eps,
                              int32_t* rpn_match, double* rpn_bbox, float* rpn_bbox_f32, int32_t* counts, void* ws,
                              size_t ws_bytes, void* stream);

/* ---- helpers exported for the parity tests (device arrays of n elements) ------------------------------ */
int mrcnn_test_expf(const float* x, float* y, int n, void* stream);
int mrcnn_test_logf(const float* x, float* y, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MRCNN_ROI_B200_H_ */
