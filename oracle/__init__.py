"""numpy-facing wrapper around the C oracle (oracle/mrcnn_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of mrcnn_oracle.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.  PARITY: the reference holds no golden
vectors for this path.  The layer-level restatement is pinned by outputs of the reference's own layer code executed on a
numpy stand-in for tf.* (tests/test_reference_layers.py, test_reference_differential.py, test_reference_gradients.py);
the bodies of the TensorFlow kernels (top_k, non_max_suppression, crop_and_resize) remain PARITY UNPINNED by TensorFlow
itself and are pinned by known answers, properties and independent implementations (tests/test_oracle_*.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborc.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile liborc.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "mrcnn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liborc.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_expf.restype = ctypes.c_float
        _lib.orc_expf.argtypes = [ctypes.c_float]
        _lib.orc_logf.restype = ctypes.c_float
        _lib.orc_logf.argtypes = [ctypes.c_float]
        _lib.orc_tf_iou.restype = ctypes.c_float
        _lib.orc_nms.restype = ctypes.c_int
        _lib.orc_roi_level.restype = ctypes.c_int32
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(int(n)))


def max_threads():
    return int(lib().orc_max_threads())


def expf(x):
    x = _f32(x)
    out = np.empty_like(x)
    L = lib()
    flat_in, flat_out = x.reshape(-1), out.reshape(-1)
    for i in range(flat_in.size):
        flat_out[i] = L.orc_expf(ctypes.c_float(float(flat_in[i])))
    return out


def logf(x):
    x = _f32(x)
    out = np.empty_like(x)
    L = lib()
    flat_in, flat_out = x.reshape(-1), out.reshape(-1)
    for i in range(flat_in.size):
        flat_out[i] = L.orc_logf(ctypes.c_float(float(flat_in[i])))
    return out


def topk(scores, k):
    """TF TopKV2 indices for one row (descending, ties -> lower index)."""
    s = _f32(scores)
    k = min(int(k), s.size)
    idx = np.empty(k, dtype=np.int32)
    lib().orc_topk(_p(s, _f32p), ctypes.c_int(s.size), ctypes.c_int(k), _p(idx, _i32p))
    return idx


def apply_box_deltas(boxes, deltas):
    b, d = _f32(boxes), _f32(deltas)
    out = np.empty_like(b)
    lib().orc_apply_box_deltas(_p(b, _f32p), _p(d, _f32p), ctypes.c_int(b.shape[0]), _p(out, _f32p))
    return out


def clip_boxes(boxes, window):
    b = _f32(boxes).copy()
    w = _f32(window)
    lib().orc_clip_boxes(_p(b, _f32p), ctypes.c_int(b.shape[0]), _p(w, _f32p))
    return b


def tf_iou(boxes, i, j):
    b = _f32(boxes)
    return float(lib().orc_tf_iou(_p(b, _f32p), ctypes.c_int(i), ctypes.c_int(j)))


def nms(boxes, scores, max_out, iou_thr):
    """tf.image.non_max_suppression(boxes, scores, max_out, iou_thr) -> int32 indices."""
    b, s = _f32(boxes).reshape(-1, 4), _f32(scores).reshape(-1)
    keep = np.empty(max(int(max_out), 1), dtype=np.int32)
    n = lib().orc_nms(_p(b, _f32p), _p(s, _f32p), ctypes.c_int(s.size), ctypes.c_int(int(max_out)),
                      ctypes.c_float(iou_thr), _p(keep, _i32p))
    return keep[:n].copy()


def crop_and_resize(image, boxes, box_ind, crop_size):
    img, b = _f32(image), _f32(boxes).reshape(-1, 4)
    bi = np.ascontiguousarray(box_ind, dtype=np.int32)
    B, H, W, C = img.shape
    ph, pw = crop_size
    out = np.zeros((b.shape[0], ph, pw, C), dtype=np.float32)
    lib().orc_crop_and_resize(_p(img, _f32p), B, H, W, C, _p(b, _f32p), _p(bi, _i32p), b.shape[0], ph, pw,
                              _p(out, _f32p))
    return out


def crop_and_resize_grad_image(grads, boxes, box_ind, image_shape):
    g, b = _f32(grads), _f32(boxes).reshape(-1, 4)
    bi = np.ascontiguousarray(box_ind, dtype=np.int32)
    B, H, W, C = image_shape
    nb, ph, pw, _ = g.shape
    out = np.zeros((B, H, W, C), dtype=np.float32)
    lib().orc_crop_and_resize_grad_image(_p(g, _f32p), B, H, W, C, _p(b, _f32p), _p(bi, _i32p), nb, ph, pw,
                                         _p(out, _f32p))
    return out


def proposal_layer(rpn_probs, rpn_bbox, anchors, pre_nms_limit, proposal_count, std_dev, nms_thr):
    """ProposalLayer.call -> dict(proposals, topk_idx, keep_idx, keep_count, pre_nms_boxes)."""
    pr, bb, an = _f32(rpn_probs), _f32(rpn_bbox), _f32(anchors)
    B, A, _ = pr.shape
    K = min(int(pre_nms_limit), A)
    P = int(proposal_count)
    sd = _f32(std_dev)
    out = np.empty((B, P, 4), dtype=np.float32)
    tk = np.empty((B, K), dtype=np.int32)
    ki = np.empty((B, P), dtype=np.int32)
    kc = np.empty((B,), dtype=np.int32)
    pb = np.empty((B, K, 4), dtype=np.float32)
    lib().orc_proposal_layer(_p(pr, _f32p), _p(bb, _f32p), _p(an, _f32p), B, A, int(pre_nms_limit), P,
                             _p(sd, _f32p), ctypes.c_float(nms_thr), _p(out, _f32p), _p(tk, _i32p),
                             _p(ki, _i32p), _p(kc, _i32p), _p(pb, _f32p))
    return dict(proposals=out, topk_idx=tk, keep_idx=ki, keep_count=kc, pre_nms_boxes=pb)


def proposal_layer_grad(grad_proposals, rpn_bbox, anchors, topk_idx, keep_idx, std_dev):
    """Gradient of ProposalLayer w.r.t. rpn_bbox [B,A,4] (reference behaviour Q7)."""
    g, bb, an = _f32(grad_proposals), _f32(rpn_bbox), _f32(anchors)
    tk = np.ascontiguousarray(topk_idx, dtype=np.int32)
    ki = np.ascontiguousarray(keep_idx, dtype=np.int32)
    B, A, _ = bb.shape
    sd = _f32(std_dev)
    out = np.empty_like(bb)
    lib().orc_proposal_layer_grad(_p(g, _f32p), _p(bb, _f32p), _p(an, _f32p), _p(tk, _i32p), _p(ki, _i32p), B, A,
                                  tk.shape[1], ki.shape[1], _p(sd, _f32p), _p(out, _f32p))
    return out


def roi_level(box, img_h, img_w, denominator=244.0):
    b = _f32(box)
    return int(lib().orc_roi_level(_p(b, _f32p), ctypes.c_float(img_h), ctypes.c_float(img_w),
                                   ctypes.c_float(denominator)))


def _fmap_args(fmaps):
    maps = [_f32(m) for m in fmaps]
    assert len(maps) == 4
    ptrs = (_f32p * 4)(*[_p(m, _f32p) for m in maps])
    Hs = (ctypes.c_int * 4)(*[m.shape[1] for m in maps])
    Ws = (ctypes.c_int * 4)(*[m.shape[2] for m in maps])
    return maps, ptrs, Hs, Ws


def pyramid_roi_align(boxes, img_h, img_w, fmaps, pool_shape, denominator=244.0, map_mode=0):
    """PyramidROIAlign.call -> dict(out [B,N,ph,pw,C], level [B,N], roi_map [B,N])."""
    bx = _f32(boxes)
    B, N, _ = bx.shape
    maps, ptrs, Hs, Ws = _fmap_args(fmaps)
    C = maps[0].shape[3]
    ph, pw = pool_shape
    out = np.zeros((B, N, ph, pw, C), dtype=np.float32)
    lvl = np.empty((B, N), dtype=np.int32)
    rm = np.empty((B, N), dtype=np.int32)
    lib().orc_pyramid_roi_align(_p(bx, _f32p), B, N, ctypes.c_float(img_h), ctypes.c_float(img_w), ptrs, Hs, Ws,
                                C, ph, pw, ctypes.c_float(denominator), int(map_mode), _p(out, _f32p),
                                _p(lvl, _i32p), _p(rm, _i32p))
    return dict(out=out, level=lvl, roi_map=rm)


def pyramid_roi_align_grad(grad_out, boxes, img_h, img_w, fmap_shapes, denominator=244.0, map_mode=0):
    """Gradient of PyramidROIAlign w.r.t. its four feature maps -> list of 4 arrays."""
    g, bx = _f32(grad_out), _f32(boxes)
    B, N, ph, pw, C = g.shape
    grads = [np.zeros(s, dtype=np.float32) for s in fmap_shapes]
    ptrs = (_f32p * 4)(*[_p(m, _f32p) for m in grads])
    Hs = (ctypes.c_int * 4)(*[s[1] for s in fmap_shapes])
    Ws = (ctypes.c_int * 4)(*[s[2] for s in fmap_shapes])
    lib().orc_pyramid_roi_align_grad(_p(g, _f32p), _p(bx, _f32p), B, N, ctypes.c_float(img_h),
                                     ctypes.c_float(img_w), Hs, Ws, C, ph, pw, ctypes.c_float(denominator),
                                     int(map_mode), ptrs)
    return grads


def detection_layer(rois, probs, deltas, image_meta, std_dev, min_conf, max_inst, nms_thr):
    """DetectionLayer.call -> dict(detections [B,max_inst,6], count [B])."""
    r, p, d, m = _f32(rois), _f32(probs), _f32(deltas), _f32(image_meta)
    B, N, NC = p.shape
    sd = _f32(std_dev)
    det = np.empty((B, int(max_inst), 6), dtype=np.float32)
    cnt = np.empty((B,), dtype=np.int32)
    use = 1 if min_conf else 0
    lib().orc_detection_layer(_p(r, _f32p), _p(p, _f32p), _p(d, _f32p), _p(m, _f32p), B, N, NC, m.shape[1],
                              _p(sd, _f32p), ctypes.c_float(min_conf or 0.0), use, int(max_inst),
                              ctypes.c_float(nms_thr), _p(det, _f32p), _p(cnt, _i32p))
    return dict(detections=det, count=cnt)


def detection_target_layer(proposals, gt_class_ids, gt_boxes, gt_masks, rand_keys, train_rois_per_image,
                           roi_positive_ratio, bbox_std_dev, mask_shape, use_mini_masks=False):
    """DetectionTargetLayer.call with an injected shuffle -> dict(rois, class_ids, deltas, masks, counts)."""
    pr, gb = _f32(proposals), _f32(gt_boxes)
    gc = np.ascontiguousarray(gt_class_ids, dtype=np.int32)
    gm = np.ascontiguousarray(gt_masks, dtype=np.uint8)
    rk = np.ascontiguousarray(rand_keys, dtype=np.uint32)
    B, P, _ = pr.shape
    G = gc.shape[1]
    _, MH, MW, _ = gm.shape
    T = int(train_rois_per_image)
    mh, mw = mask_shape
    sd = _f32(bbox_std_dev)
    rois = np.empty((B, T, 4), dtype=np.float32)
    cls = np.empty((B, T), dtype=np.int32)
    dl = np.empty((B, T, 4), dtype=np.float32)
    mk = np.empty((B, T, mh, mw), dtype=np.float32)
    cnt = np.empty((B, 2), dtype=np.int32)
    lib().orc_detection_target_layer(_p(pr, _f32p), _p(gc, _i32p), _p(gb, _f32p), _p(gm, _u8p), _p(rk, _u32p),
                                     B, P, G, MH, MW, T, ctypes.c_double(roi_positive_ratio), _p(sd, _f32p),
                                     mh, mw, 1 if use_mini_masks else 0, _p(rois, _f32p), _p(cls, _i32p),
                                     _p(dl, _f32p), _p(mk, _f32p), _p(cnt, _i32p))
    return dict(rois=rois, class_ids=cls, deltas=dl, masks=mk, counts=cnt)


def build_rpn_targets(anchors_px, gt_class_ids, gt_boxes, rand_keys, rpn_train_anchors_per_image, rpn_bbox_std,
                      eps=1e-3):
    """utils.build_rpn_targets (utils.py:154-262) over a padded batch (class id 0 = padding row) with injected keys
    in place of np.random.choice -> dict(rpn_match [B,A] int32, rpn_bbox [B,R,4] float64, counts [B,2])."""
    an = np.ascontiguousarray(anchors_px, dtype=np.float64)
    gc = np.ascontiguousarray(gt_class_ids, dtype=np.int32)
    gb = np.ascontiguousarray(gt_boxes, dtype=np.int32)
    rk = _f32(rand_keys)
    B, G = gc.shape
    A, R = an.shape[0], int(rpn_train_anchors_per_image)
    sd = np.ascontiguousarray(rpn_bbox_std, dtype=np.float64)
    match = np.empty((B, A), dtype=np.int32)
    bbox = np.empty((B, R, 4), dtype=np.float64)
    cnt = np.empty((B, 2), dtype=np.int32)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib().orc_build_rpn_targets(_p(an, f64p), _p(gc, _i32p), _p(gb, _i32p), _p(rk, _f32p), B, A, G, R, _p(sd, f64p),
                                ctypes.c_double(eps), _p(match, _i32p), _p(bbox, f64p), _p(cnt, _i32p))
    return dict(rpn_match=match, rpn_bbox=bbox, counts=cnt)


def rpn_softmax(logits):
    """Keras softmax over the last axis of [..., 2] RPN logits (TF SoftmaxEigenImpl order) -> probs, same shape."""
    x = _f32(logits)
    out = np.empty_like(x)
    lib().orc_softmax2(_p(x, _f32p), ctypes.c_int(x.size // 2), _p(out, _f32p))
    return out
