"""Tensor-level entry points: torch CUDA tensors in, torch CUDA tensors out, all work done by the C-ABI launchers.

Every function is asynchronous on torch's current stream and never synchronises the host.  Workspaces are cached
per (op, shape, device) so repeated calls (and CUDA-graph capture) do not allocate.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import c_float, c_int, c_size_t, c_void_p, check, ptr

_ws_cache = {}
_ws_namespace = [0]


class workspace_namespace:
    """Launchers are stateless and the caller owns every workspace; this module caches them per (op, shape, device).
    Calls that may run CONCURRENTLY (several CUDA streams in one process) must not share a workspace: wrap each
    stream's calls in `with workspace_namespace(i):` to give them their own set."""

    def __init__(self, key):
        self.key = key

    def __enter__(self):
        self.prev = _ws_namespace[0]
        _ws_namespace[0] = self.key
        return self

    def __exit__(self, *exc):
        _ws_namespace[0] = self.prev
        return False


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name, ndim=None, pinned_ok=False):
    """pinned_ok: the launcher reads this tensor sparsely by index (a few rows out of many), so a page-locked host
    tensor -- device-accessible under unified addressing -- may be passed as is: only the rows used cross the bus."""
    if not isinstance(t, torch.Tensor) or not (t.is_cuda or (pinned_ok and t.is_pinned())):
        raise TypeError(f"{name}: expected a CUDA tensor (the ROI-stage kernels have no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(key, nbytes, device):
    key = key + (device.index, _ws_namespace[0])
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _query(fn, *args):
    out = c_size_t(0)
    check(fn(*args, ctypes.byref(out)), fn.__name__)
    return out.value


def det_expf(x):
    x = _req(x, torch.float32, "x")
    y = torch.empty_like(x)
    check(_lib.lib().mrcnn_test_expf(ptr(x), ptr(y), x.numel(), _stream()), "mrcnn_test_expf")
    return y


def det_logf(x):
    x = _req(x, torch.float32, "x")
    y = torch.empty_like(x)
    check(_lib.lib().mrcnn_test_logf(ptr(x), ptr(y), x.numel(), _stream()), "mrcnn_test_logf")
    return y


def topk(scores, k, column=None, return_values=False):
    """tf.nn.top_k(scores, k, sorted=True).indices per row.  scores [B,A], or [B,A,S] with `column` selecting
    one of the S interleaved columns (e.g. the foreground probability of rpn_probs)."""
    L = _lib.lib()
    if column is None:
        scores = _req(scores, torch.float32, "scores", 2)
        stride, offset = 1, 0
    else:
        scores = _req(scores, torch.float32, "scores", 3)
        stride, offset = scores.shape[2], int(column)
    B, A = scores.shape[0], scores.shape[1]
    k = int(k)
    nbytes = _query(L.mrcnn_topk_workspace_bytes, B, A, k)
    ws = _workspace(("topk", B, A, k), nbytes, scores.device)
    idx = torch.empty((B, k), dtype=torch.int32, device=scores.device)
    vals = torch.empty((B, k), dtype=torch.float32, device=scores.device) if return_values else None
    check(L.mrcnn_topk_forward(ptr(scores), stride, offset, B, A, k, ptr(idx), ptr(vals), ptr(ws), ws.numel(),
                               _stream()), "mrcnn_topk_forward")
    return (idx, vals) if return_values else idx


def nms(boxes, scores, max_output_size, iou_threshold, valid=None):
    """Batched tf.image.non_max_suppression.  boxes [B,M,4], scores [B,M] -> keep [B,max_out] (-1 padded), count [B]."""
    L = _lib.lib()
    boxes = _req(boxes, torch.float32, "boxes", 3)
    scores = _req(scores, torch.float32, "scores", 2)
    B, M = scores.shape
    if valid is not None:
        valid = _req(valid, torch.int32, "valid", 1)
    nbytes = _query(L.mrcnn_nms_workspace_bytes, B, M)
    ws = _workspace(("nms", B, M), nbytes, boxes.device)
    keep = torch.empty((B, int(max_output_size)), dtype=torch.int32, device=boxes.device)
    count = torch.empty((B,), dtype=torch.int32, device=boxes.device)
    check(L.mrcnn_nms_forward(ptr(boxes), ptr(scores), ptr(valid), B, M, int(max_output_size),
                              c_float(float(iou_threshold)), ptr(keep), ptr(count), ptr(ws), ws.numel(), _stream()),
          "mrcnn_nms_forward")
    return keep, count


def proposal_forward(rpn_probs, rpn_bbox, anchors, pre_nms_limit, proposal_count, std_dev, nms_threshold,
                     debug=False):
    """ProposalLayer.call.  Returns proposals [B,P,4]; with debug=True a dict with the intermediate indices too."""
    L = _lib.lib()
    rpn_probs = _req(rpn_probs, torch.float32, "rpn_probs", 3)
    rpn_bbox = _req(rpn_bbox, torch.float32, "rpn_bbox", 3, pinned_ok=True)   # only the K winners' rows are read
    anchors = _req(anchors, torch.float32, "anchors", 3)
    B, A, two = rpn_probs.shape
    if two != 2 or tuple(rpn_bbox.shape) != (B, A, 4) or tuple(anchors.shape) != (B, A, 4):
        raise ValueError("expected rpn_probs [B,A,2], rpn_bbox [B,A,4], anchors [B,A,4]")
    P = int(proposal_count)
    K = min(int(pre_nms_limit), A)
    dev = rpn_probs.device
    nbytes = _query(L.mrcnn_proposal_workspace_bytes, B, A, int(pre_nms_limit), P)
    ws = _workspace(("proposal", B, A, K, P), nbytes, dev)
    proposals = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    tk = ki = kc = pb = None
    if debug:
        tk = torch.empty((B, K), dtype=torch.int32, device=dev)
        ki = torch.empty((B, P), dtype=torch.int32, device=dev)
        kc = torch.empty((B,), dtype=torch.int32, device=dev)
        pb = torch.empty((B, K, 4), dtype=torch.float32, device=dev)
    check(L.mrcnn_proposal_forward(ptr(rpn_probs), ptr(rpn_bbox), ptr(anchors), B, A, int(pre_nms_limit), P,
                                   _lib.float4(std_dev), c_float(float(nms_threshold)), ptr(proposals), ptr(tk),
                                   ptr(ki), ptr(kc), ptr(pb), ptr(ws), ws.numel(), _stream()),
          "mrcnn_proposal_forward")
    if debug:
        return dict(proposals=proposals, topk_idx=tk, keep_idx=ki, keep_count=kc, pre_nms_boxes=pb)
    return proposals


def proposal_backward(grad_proposals, rpn_bbox, anchors, topk_idx, keep_idx, std_dev):
    """Gradient of ProposalLayer.call w.r.t. rpn_bbox [B,A,4] (the reference does not stop it)."""
    L = _lib.lib()
    grad_proposals = _req(grad_proposals, torch.float32, "grad_proposals", 3)
    rpn_bbox = _req(rpn_bbox, torch.float32, "rpn_bbox", 3)
    anchors = _req(anchors, torch.float32, "anchors", 3)
    topk_idx = _req(topk_idx, torch.int32, "topk_idx", 2)
    keep_idx = _req(keep_idx, torch.int32, "keep_idx", 2)
    B, A, _ = rpn_bbox.shape
    K, P = topk_idx.shape[1], keep_idx.shape[1]
    grad = torch.empty_like(rpn_bbox)
    check(L.mrcnn_proposal_backward(ptr(grad_proposals), ptr(rpn_bbox), ptr(anchors), ptr(topk_idx), ptr(keep_idx), B, A,
                                    K, P, _lib.float4(std_dev), ptr(grad), _stream()), "mrcnn_proposal_backward")
    return grad


def _map_args(maps):
    if len(maps) != 4:
        raise ValueError("PyramidROIAlign needs exactly four feature maps (P2..P5)")
    maps = [_req(m, torch.float32, f"feature_maps[{i}]", 4) for i, m in enumerate(maps)]
    ptrs = (c_void_p * 4)(*[m.data_ptr() for m in maps])
    Hs = (c_int * 4)(*[m.shape[1] for m in maps])
    Ws = (c_int * 4)(*[m.shape[2] for m in maps])
    return maps, ptrs, Hs, Ws


def roialign_forward(boxes, image_meta, feature_maps, pool_shape, denominator=244.0, map_mode=0,
                     return_level=False):
    """PyramidROIAlign.call.  Returns (pooled [B,N,ph,pw,C], roi_map [B,N]) (+ roi_level when asked)."""
    L = _lib.lib()
    boxes = _req(boxes, torch.float32, "boxes", 3)
    image_meta = _req(image_meta, torch.float32, "image_meta", 2)
    maps, ptrs, Hs, Ws = _map_args(feature_maps)
    B, N, _ = boxes.shape
    C = maps[0].shape[3]
    for m in maps:
        if m.shape[0] != B or m.shape[3] != C:
            raise ValueError("feature maps must be [B,H,W,C] with the batch size of `boxes` and equal C")
    ph, pw = int(pool_shape[0]), int(pool_shape[1])
    dev = boxes.device
    out = torch.empty((B, N, ph, pw, C), dtype=torch.float32, device=dev)
    roi_map = torch.empty((B, N), dtype=torch.int32, device=dev)
    roi_level = torch.empty((B, N), dtype=torch.int32, device=dev) if return_level else None
    nbytes = _query(L.mrcnn_roialign_workspace_bytes, B, N)
    ws = _workspace(("roialign", B, N, ph, pw), nbytes, dev)
    check(L.mrcnn_roialign_forward(ptr(boxes), ptr(image_meta), image_meta.shape[1], ptrs, Hs, Ws, C, B, N, ph, pw,
                                   c_float(float(denominator)), int(map_mode), ptr(out), ptr(roi_map), ptr(roi_level),
                                   ptr(ws), ws.numel(), _stream()), "mrcnn_roialign_forward")
    if return_level:
        return out, roi_map, roi_level
    return out, roi_map


def deterministic_ops():
    """TensorFlow's switch for reproducible kernels (TF_DETERMINISTIC_OPS=1), honoured by the ROIAlign gradient."""
    return os.environ.get("TF_DETERMINISTIC_OPS", "0").strip().lower() in ("1", "true")


def roialign_backward(grad_out, boxes, roi_map, fmap_shapes, deterministic=None):
    """Feature-map gradients of PyramidROIAlign: list of four [B,H,W,C] tensors.  deterministic=True sums every
    pixel's samples in TF CropAndResizeGradImage's sequential order (bit-identical to the CPU kernel, reproducible);
    False scatters with fp32 vector atomics (order not fixed; faster at 14x14).  None -> TF_DETERMINISTIC_OPS."""
    if deterministic is None:
        deterministic = deterministic_ops()
    L = _lib.lib()
    grad_out = _req(grad_out, torch.float32, "grad_out", 5)
    boxes = _req(boxes, torch.float32, "boxes", 3)
    roi_map = _req(roi_map, torch.int32, "roi_map", 2)
    B, N, ph, pw, C = grad_out.shape
    grads = [torch.empty(tuple(s), dtype=torch.float32, device=grad_out.device) for s in fmap_shapes]
    ptrs = (c_void_p * 4)(*[g.data_ptr() for g in grads])
    Hs = (c_int * 4)(*[s[1] for s in fmap_shapes])
    Ws = (c_int * 4)(*[s[2] for s in fmap_shapes])
    ws, nbytes = None, 0
    if deterministic:
        nbytes = _query(L.mrcnn_roialign_backward_workspace_bytes, B, N, ph, pw, Hs, Ws, C)
        ws = _workspace(("roialign_bwd", B, N, ph, pw, C, tuple(tuple(s) for s in fmap_shapes)), nbytes, grad_out.device)
    check(L.mrcnn_roialign_backward(ptr(grad_out), ptr(boxes), ptr(roi_map), ptrs, Hs, Ws, C, B, N, ph, pw, ptr(ws),
                                    nbytes, _stream()), "mrcnn_roialign_backward")
    return grads


def detection_forward(rois, probs, deltas, image_meta, bbox_std_dev, min_confidence, max_instances, nms_threshold,
                      return_count=False, return_boxes=False):
    """DetectionLayer.call -> detections [B,max_instances,6] (, count [B]) (, boxes [B,max_instances,4] =
    detections[..., :4], the DetectedBoxesExtraction output, written by the same kernel)."""
    L = _lib.lib()
    rois = _req(rois, torch.float32, "rois", 3)
    probs = _req(probs, torch.float32, "mrcnn_class", 3)
    deltas = _req(deltas, torch.float32, "mrcnn_bbox", 4, pinned_ok=True)   # only the argmax class's row per ROI
    image_meta = _req(image_meta, torch.float32, "image_meta", 2)
    B, N, NC = probs.shape
    if tuple(rois.shape) != (B, N, 4) or tuple(deltas.shape) != (B, N, NC, 4):
        raise ValueError("expected rois [B,N,4], mrcnn_class [B,N,NC], mrcnn_bbox [B,N,NC,4]")
    dev = rois.device
    nbytes = _query(L.mrcnn_detection_workspace_bytes, B, N, NC)
    ws = _workspace(("detection", B, N, NC), nbytes, dev)
    det = torch.empty((B, int(max_instances), 6), dtype=torch.float32, device=dev)
    cnt = torch.empty((B,), dtype=torch.int32, device=dev) if return_count else None
    boxes = torch.empty((B, int(max_instances), 4), dtype=torch.float32, device=dev) if return_boxes else None
    use_conf = 1 if min_confidence else 0
    check(L.mrcnn_detection_forward(ptr(rois), ptr(probs), ptr(deltas), ptr(image_meta), image_meta.shape[1], B, N,
                                    NC, _lib.float4(bbox_std_dev), c_float(float(min_confidence or 0.0)), use_conf,
                                    int(max_instances), c_float(float(nms_threshold)), 0, ptr(det), ptr(cnt), ptr(boxes),
                                    ptr(ws), ws.numel(), _stream()), "mrcnn_detection_forward")
    out = (det,) + ((cnt,) if return_count else ()) + ((boxes,) if return_boxes else ())
    return out if len(out) > 1 else det


def detection_target_forward(proposals, gt_class_ids, gt_boxes, gt_masks, rand_keys, train_rois_per_image,
                             roi_positive_ratio, bbox_std_dev, mask_shape, use_mini_masks=False, return_counts=False):
    """DetectionTargetLayer.call -> rois [B,T,4], class_ids [B,T] int32, deltas [B,T,4], masks [B,T,mh,mw]."""
    L = _lib.lib()
    proposals = _req(proposals, torch.float32, "proposals", 3)
    gt_class_ids = _req(gt_class_ids, torch.int32, "gt_class_ids", 2)
    gt_boxes = _req(gt_boxes, torch.float32, "gt_boxes", 3)
    gt_masks = _req(gt_masks, torch.uint8, "gt_masks", 4)
    rand_keys = _req(rand_keys, torch.int32, "rand_keys", 2)  # raw 32 bits, compared as unsigned
    B, P, _ = proposals.shape
    G = gt_class_ids.shape[1]
    _, MH, MW, G2 = gt_masks.shape
    if G2 != G or tuple(gt_boxes.shape) != (B, G, 4) or tuple(rand_keys.shape) != (B, P):
        raise ValueError("expected gt_boxes [B,G,4], gt_masks [B,MH,MW,G], rand_keys [B,P]")
    T = int(train_rois_per_image)
    mh, mw = int(mask_shape[0]), int(mask_shape[1])
    dev = proposals.device
    nbytes = _query(L.mrcnn_detection_target_workspace_bytes, B, P, G, T)
    ws = _workspace(("dtarget", B, P, G, T), nbytes, dev)
    rois = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    cls = torch.empty((B, T), dtype=torch.int32, device=dev)
    dl = torch.empty((B, T, 4), dtype=torch.float32, device=dev)
    masks = torch.empty((B, T, mh, mw), dtype=torch.float32, device=dev)
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev) if return_counts else None
    check(L.mrcnn_detection_target_forward(ptr(proposals), ptr(gt_class_ids), ptr(gt_boxes), ptr(gt_masks),
                                           ptr(rand_keys), B, P, G, MH, MW, T, ctypes.c_double(roi_positive_ratio),
                                           _lib.float4(bbox_std_dev), mh, mw, 1 if use_mini_masks else 0, ptr(rois),
                                           ptr(cls), ptr(dl), ptr(masks), ptr(counts), ptr(ws), ws.numel(), _stream()),
          "mrcnn_detection_target_forward")
    if return_counts:
        return rois, cls, dl, masks, counts
    return rois, cls, dl, masks


def rpn_targets_forward(anchors_px, gt_class_ids, gt_boxes, rand_keys, rpn_train_anchors_per_image, rpn_bbox_std,
                        eps=1e-3, return_f32=False, return_counts=False):
    """utils.build_rpn_targets (utils.py:154-262) for a padded batch -> rpn_match [B,A] int32, rpn_bbox [B,R,4] float64
    (+ the same in fp32, + counts [B,2] of kept positives / negatives).

    anchors_px [A,4] float64 pixel anchors (shared by the batch); gt_class_ids [B,G] int32 (0 = padding row, negative =
    crowd); gt_boxes [B,G,4] int32 pixel boxes; rand_keys [B,A] fp32 >= 0 in place of np.random.choice."""
    L = _lib.lib()
    anchors_px = _req(anchors_px, torch.float64, "anchors_px", 2)
    gt_class_ids = _req(gt_class_ids, torch.int32, "gt_class_ids", 2)
    gt_boxes = _req(gt_boxes, torch.int32, "gt_boxes", 3)
    rand_keys = _req(rand_keys, torch.float32, "rand_keys", 2)
    A, (B, G), R = anchors_px.shape[0], gt_class_ids.shape, int(rpn_train_anchors_per_image)
    if anchors_px.shape[1] != 4 or tuple(gt_boxes.shape) != (B, G, 4) or tuple(rand_keys.shape) != (B, A):
        raise ValueError("expected anchors_px [A,4], gt_boxes [B,G,4], rand_keys [B,A]")
    dev = anchors_px.device
    nbytes = _query(L.mrcnn_rpn_targets_workspace_bytes, B, A, G, R)
    ws = _workspace(("rpn_targets", B, A, G, R), nbytes, dev)
    match = torch.empty((B, A), dtype=torch.int32, device=dev)
    bbox = torch.empty((B, R, 4), dtype=torch.float64, device=dev)
    bbox32 = torch.empty((B, R, 4), dtype=torch.float32, device=dev) if return_f32 else None
    counts = torch.empty((B, 2), dtype=torch.int32, device=dev) if return_counts else None
    sd = (ctypes.c_double * 4)(*[float(v) for v in rpn_bbox_std])
    check(L.mrcnn_rpn_targets_forward(ptr(anchors_px), ptr(gt_class_ids), ptr(gt_boxes), ptr(rand_keys), B, A, G, R, sd,
                                      ctypes.c_double(eps), ptr(match), ptr(bbox), ptr(bbox32), ptr(counts), ptr(ws),
                                      ws.numel(), _stream()), "mrcnn_rpn_targets_forward")
    out = [match, bbox]
    if return_f32:
        out.append(bbox32)
    if return_counts:
        out.append(counts)
    return tuple(out)


class HostMapStage:
    """Device-side staging of feature maps that live in PINNED HOST memory (demand-driven H2D, csrc/roialign.cu).

    `stage.fetch(boxes, image_meta, pool_shape, reset=...)` copies exactly the map pixels those ROIs sample, once
    each, straight out of the pinned host tensors; `stage.maps` are then valid inputs for roialign_forward /
    PyramidROIAlign with the same boxes.  reset=True declares new host contents (first call of a step); later calls of
    the step (the mask branch) fetch only pixels that are not resident yet.  `fetched_pixels()` synchronises and
    returns the running count of pixels copied (each C * 4 bytes)."""

    def __init__(self, host_maps, device):
        if len(host_maps) != 4:
            raise ValueError("PyramidROIAlign needs exactly four feature maps (P2..P5)")
        for i, m in enumerate(host_maps):
            if not isinstance(m, torch.Tensor) or m.is_cuda or not m.is_pinned() or m.dtype != torch.float32 or \
                    m.dim() != 4 or not m.is_contiguous():
                raise TypeError(f"host_maps[{i}]: expected a contiguous pinned (page-locked) fp32 CPU tensor [B,H,W,C]")
        self.host_maps = list(host_maps)
        self.device = device
        self.maps = [torch.empty(m.shape, dtype=torch.float32, device=device) for m in host_maps]
        self.B, self.C = host_maps[0].shape[0], host_maps[0].shape[3]
        self._Hs = (c_int * 4)(*[m.shape[1] for m in host_maps])
        self._Ws = (c_int * 4)(*[m.shape[2] for m in host_maps])
        self._host_ptrs = (c_void_p * 4)(*[m.data_ptr() for m in host_maps])   # device-accessible under UVA
        self._dev_ptrs = (c_void_p * 4)(*[m.data_ptr() for m in self.maps])
        words = _query(_lib.lib().mrcnn_roialign_resident_words, self.B, self._Hs, self._Ws)
        self.resident = torch.zeros(int(words), dtype=torch.int32, device=device)
        self.counter = torch.zeros(1, dtype=torch.int64, device=device)

    def fetch(self, boxes, image_meta, pool_shape, reset, denominator=244.0, map_mode=0):
        L = _lib.lib()
        boxes = _req(boxes, torch.float32, "boxes", 3)
        image_meta = _req(image_meta, torch.float32, "image_meta", 2)
        B, N, _ = boxes.shape
        if B != self.B:
            raise ValueError("boxes and host maps disagree on the batch size")
        nbytes = _query(L.mrcnn_roialign_fetch_workspace_bytes, B, N, self._Hs, self._Ws)
        ws = _workspace(("roialign_fetch", B, N, id(self)), nbytes, self.device)
        check(L.mrcnn_roialign_fetch_hostmaps(ptr(boxes), ptr(image_meta), image_meta.shape[1], self._host_ptrs,
                                              self._dev_ptrs, self._Hs, self._Ws, self.C, B, N, int(pool_shape[0]),
                                              int(pool_shape[1]), c_float(float(denominator)), int(map_mode),
                                              ptr(self.resident), 1 if reset else 0, ptr(self.counter), ptr(ws),
                                              ws.numel(), _stream()), "mrcnn_roialign_fetch_hostmaps")
        return self.maps

    def fetched_pixels(self):
        return int(self.counter.item())


def anchors_forward(scales, ratios, feature_shapes, strides, anchor_stride, image_hw, batch, device,
                    return_px=False):
    """utils.generate_pyramid_anchors + AnchorsLayer.get_anchors on the device -> anchors [batch,A,4] fp32 normalised
    (and, with return_px, the float64 pixel anchors [A,4] the data loader / build_rpn_targets use)."""
    L = _lib.lib()
    n = len(scales)
    fh = (c_int * n)(*[int(s[0]) for s in feature_shapes])
    fw = (c_int * n)(*[int(s[1]) for s in feature_shapes])
    st = (c_int * n)(*[int(s) for s in strides])
    sc = (ctypes.c_double * n)(*[float(s) for s in scales])
    ra = (ctypes.c_double * len(ratios))(*[float(r) for r in ratios])
    count = c_int(0)
    check(L.mrcnn_anchors_count(fh, fw, n, len(ratios), int(anchor_stride), ctypes.byref(count)), "mrcnn_anchors_count")
    A = count.value
    norm = torch.empty((int(batch), A, 4), dtype=torch.float32, device=device)
    px = torch.empty((A, 4), dtype=torch.float64, device=device) if return_px else None
    check(L.mrcnn_anchors_forward(sc, ra, fh, fw, st, n, len(ratios), int(anchor_stride), int(image_hw[0]),
                                  int(image_hw[1]), int(batch), ptr(px), ptr(norm), _stream()), "mrcnn_anchors_forward")
    return (norm, px) if return_px else norm


def proposal_forward_levels(rpn_class_logits, rpn_bbox, anchors, pre_nms_limit, proposal_count, std_dev, nms_threshold,
                            return_probs=False, debug=False):
    """ProposalLayer fed by the RPN head's per-level outputs: rpn_class_logits = list of [B,A_l,2] raw logits,
    rpn_bbox = list of [B,A_l,4] raw deltas (level-major = the order of `anchors` [B,A,4]).  The softmax and the three
    Concatenate layers of model.py:465-478 are fused away.  Returns proposals [B,P,4] (, rpn_probs [B,A,2])."""
    L = _lib.lib()
    n = len(rpn_class_logits)
    if n != len(rpn_bbox) or n < 1:
        raise ValueError("rpn_class_logits and rpn_bbox must be lists with one tensor per pyramid level")
    lg = [_req(t, torch.float32, f"rpn_class_logits[{i}]", 3) for i, t in enumerate(rpn_class_logits)]
    bb = [_req(t, torch.float32, f"rpn_bbox[{i}]", 3) for i, t in enumerate(rpn_bbox)]
    anchors = _req(anchors, torch.float32, "anchors", 3)
    B, A = anchors.shape[0], anchors.shape[1]
    counts = [t.shape[1] for t in lg]
    for a_, b_ in zip(lg, bb):
        if a_.shape[0] != B or a_.shape[2] != 2 or tuple(b_.shape) != (B, a_.shape[1], 4):
            raise ValueError("expected per-level rpn_class_logits [B,A_l,2] and rpn_bbox [B,A_l,4]")
    if sum(counts) != A:
        raise ValueError("the levels' anchor counts must add up to anchors.shape[1]")
    P, K = int(proposal_count), min(int(pre_nms_limit), A)
    dev = anchors.device
    nbytes = _query(L.mrcnn_proposal_levels_workspace_bytes, B, A, int(pre_nms_limit), P)
    ws = _workspace(("proposal_levels", B, A, K, P), nbytes, dev)
    proposals = torch.empty((B, P, 4), dtype=torch.float32, device=dev)
    probs = torch.empty((B, A, 2), dtype=torch.float32, device=dev) if return_probs else None
    tk = ki = kc = None
    if debug:
        tk = torch.empty((B, K), dtype=torch.int32, device=dev)
        ki = torch.empty((B, P), dtype=torch.int32, device=dev)
        kc = torch.empty((B,), dtype=torch.int32, device=dev)
    lp = (c_void_p * n)(*[t.data_ptr() for t in lg])
    bp = (c_void_p * n)(*[t.data_ptr() for t in bb])
    cn = (c_int * n)(*counts)
    check(L.mrcnn_proposal_forward_levels(lp, bp, cn, n, ptr(anchors), B, int(pre_nms_limit), P, _lib.float4(std_dev),
                                          c_float(float(nms_threshold)), ptr(proposals), ptr(probs), ptr(tk), ptr(ki),
                                          ptr(kc), ptr(ws), ws.numel(), _stream()), "mrcnn_proposal_forward_levels")
    if debug:
        return dict(proposals=proposals, rpn_probs=probs, topk_idx=tk, keep_idx=ki, keep_count=kc)
    return (proposals, probs) if return_probs else proposals
