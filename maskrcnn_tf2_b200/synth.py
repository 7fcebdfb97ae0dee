"""Deterministic COCO-shape synthetic inputs for the ROI stage, as numpy arrays.

Everything the hot path consumes but other reference code produces is re-derived here from the cited lines:
anchors (utils.py:54-111,725-735 + AnchorsLayer/NormBoxesLayer, mrcnn_layers.py:34-39,116-132), image_meta
(utils.py:494-516) and the tensor layouts of the RPN / head outputs.  Seeds: default_rng(1000*config + image).
"""
import math

import numpy as np


def backbone_shapes(img_size, strides):
    """utils.compute_backbone_shapes (utils.py:725-735)."""
    return [(int(math.ceil(img_size / s)), int(math.ceil(img_size / s))) for s in strides]


def pyramid_anchors_px(img_size, scales=(32, 64, 128, 256, 512), ratios=(0.5, 1, 2), strides=(4, 8, 16, 32, 64),
                       anchor_stride=1):
    """float64 pixel anchors [A,4] as utils.generate_pyramid_anchors returns them (utils.py:54-111): level-major,
    then row-major (y, x), ratio innermost.  These are what the data loader hands to build_rpn_targets
    (preprocess.py:82,297,343)."""
    out = []
    ratios = np.asarray(ratios, dtype=np.float64)
    for scale, (fh, fw), stride in zip(scales, backbone_shapes(img_size, strides), strides):
        hs = scale / np.sqrt(ratios)
        ws = scale * np.sqrt(ratios)
        cy = np.arange(0, fh, anchor_stride, dtype=np.float64) * stride
        cx = np.arange(0, fw, anchor_stride, dtype=np.float64) * stride
        # [fh, fw, R]
        CY = np.broadcast_to(cy[:, None, None], (cy.size, cx.size, ratios.size))
        CX = np.broadcast_to(cx[None, :, None], (cy.size, cx.size, ratios.size))
        HH = np.broadcast_to(hs[None, None, :], CY.shape)
        WW = np.broadcast_to(ws[None, None, :], CY.shape)
        boxes = np.stack([CY - 0.5 * HH, CX - 0.5 * WW, CY + 0.5 * HH, CX + 0.5 * WW], axis=-1).reshape(-1, 4)
        out.append(boxes)
    return np.concatenate(out, axis=0)


def pyramid_anchors(img_size, scales=(32, 64, 128, 256, 512), ratios=(0.5, 1, 2), strides=(4, 8, 16, 32, 64),
                    anchor_stride=1):
    """Normalised fp32 anchors [A,4]: float64 pixel boxes (pyramid_anchors_px) -> fp32 ->
    (a - [0,0,1,1]) / ([h,w,h,w] - 1) in fp32 (mrcnn_layers.py:34-39,116-132)."""
    a = pyramid_anchors_px(img_size, scales, ratios, strides, anchor_stride).astype(np.float32)
    scale = np.array([img_size, img_size, img_size, img_size], dtype=np.float32) - np.float32(1.0)
    shift = np.array([0, 0, 1, 1], dtype=np.float32)
    return ((a - shift) / scale).astype(np.float32)


def image_meta(batch, img_size, num_classes):
    """utils.compose_image_meta layout: [id, orig h,w,c, h,w,c, window y1,x1,y2,x2, scale, active classes]."""
    m = np.zeros((batch, 12 + num_classes), dtype=np.float32)
    for b in range(batch):
        m[b, 0] = b
        m[b, 1:4] = (img_size, img_size, 3)
        m[b, 4:7] = (img_size, img_size, 3)
        m[b, 7:11] = (0, 0, img_size, img_size)
        m[b, 11] = 1.0
        m[b, 12:] = 1.0
    return m


def _softmax(x):
    x = x - x.max(axis=-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=-1, keepdims=True)


def gt_instances(rng, img_size, n_real=20, max_gt=100, num_classes=81, crowd=False):
    """Pixel GT boxes: side U(32,512), centres uniform, class U{1..NC-1}; optional crowd box (negative class)."""
    boxes = np.zeros((max_gt, 4), dtype=np.float32)
    cls = np.zeros((max_gt,), dtype=np.int32)
    for g in range(n_real):
        h, w = rng.uniform(32, min(512, img_size / 2), size=2)
        cy = rng.uniform(h / 2, img_size - h / 2)
        cx = rng.uniform(w / 2, img_size - w / 2)
        boxes[g] = np.round([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2])
        cls[g] = rng.integers(1, num_classes)
    if crowd and n_real > 0:
        cls[n_real - 1] = -cls[n_real - 1]
    return boxes, cls


def norm_boxes(boxes_px, img_size):
    scale = np.float32(img_size - 1)
    shift = np.array([0, 0, 1, 1], dtype=np.float32)
    return ((boxes_px.astype(np.float32) - shift) / scale).astype(np.float32)


def ellipse_masks(boxes_px, img_size, mask_size=None):
    """uint8 [H,W,G] filled ellipses inscribed in each GT box (zero channels for zero boxes).
    mask_size=(mh,mw): mini-masks in the box frame instead (use_mini_masks, config.py:38-39)."""
    G = boxes_px.shape[0]
    if mask_size is None:
        m = np.zeros((img_size, img_size, G), dtype=np.uint8)
        yy, xx = np.mgrid[0:img_size, 0:img_size]
        for g in range(G):
            y1, x1, y2, x2 = boxes_px[g]
            if y2 <= y1 or x2 <= x1:
                continue
            cy, cx, ry, rx = (y1 + y2) / 2, (x1 + x2) / 2, (y2 - y1) / 2, (x2 - x1) / 2
            ys, xs = slice(int(y1), int(y2) + 1), slice(int(x1), int(x2) + 1)
            m[ys, xs, g] = (((yy[ys, xs] - cy) / ry) ** 2 + ((xx[ys, xs] - cx) / rx) ** 2 <= 1.0)
        return m
    mh, mw = mask_size
    m = np.zeros((mh, mw, G), dtype=np.uint8)
    yy, xx = np.mgrid[0:mh, 0:mw]
    for g in range(G):
        if boxes_px[g, 2] <= boxes_px[g, 0]:
            continue
        m[:, :, g] = (((yy - (mh - 1) / 2) / (mh / 2)) ** 2 + ((xx - (mw - 1) / 2) / (mw / 2)) ** 2 <= 1.0)
    return m


def _iou_matrix(a, b):
    y1 = np.maximum(a[:, None, 0], b[None, :, 0]); x1 = np.maximum(a[:, None, 1], b[None, :, 1])
    y2 = np.minimum(a[:, None, 2], b[None, :, 2]); x2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.maximum(y2 - y1, 0) * np.maximum(x2 - x1, 0)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-12)


# (objects, delta noise, logit noise): 'clustered' = busy COCO image, NMS keeps 1000 after examining ~3500 of the
# 6000 candidates (~70 % suppressed); 'sparse' = few objects, ~200 survivors of all 6000 (zero padding exercised)
_REGIMES = {"clustered": (40, 0.15, 1.0), "sparse": (20, 0.05, 0.5)}


def rpn_outputs(rng, anchors, regime, img_size, std_dev=(0.1, 0.1, 0.2, 0.2)):
    """One image's rpn_probs [A,2] and raw rpn_bbox [A,4].
    regime 'iid': softmax(2 N(0,1)), N(0,0.5) deltas -- NMS nearly trivial (best case).
    regime 'clustered' / 'sparse': what a trained RPN emits -- objectness follows the best IoU with a set of
    objects and the deltas regress towards them, so proposals pile up and NMS suppresses most of them."""
    A = anchors.shape[0]
    if regime == 'iid':
        probs = _softmax(2.0 * rng.standard_normal((A, 2), dtype=np.float32)).astype(np.float32)
        bbox = (0.5 * rng.standard_normal((A, 4), dtype=np.float32)).astype(np.float32)
        return probs, bbox
    n_gt, delta_noise, logit_noise = _REGIMES[regime]
    gt_px, _ = gt_instances(rng, img_size, n_real=n_gt, max_gt=n_gt)
    gt = norm_boxes(gt_px, img_size).astype(np.float64)
    an = anchors.astype(np.float64)
    iou = _iou_matrix(an, gt)
    best = iou.argmax(axis=1)
    miou = iou[np.arange(A), best]
    logit = 12.0 * miou - 4.0 + logit_noise * rng.standard_normal(A)
    p = 1.0 / (1.0 + np.exp(-logit))
    probs = np.stack([1.0 - p, p], axis=1).astype(np.float32)
    g = gt[best]
    ah, aw = an[:, 2] - an[:, 0], an[:, 3] - an[:, 1]
    gh, gw = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
    t = np.stack([((g[:, 0] + 0.5 * gh) - (an[:, 0] + 0.5 * ah)) / ah,
                  ((g[:, 1] + 0.5 * gw) - (an[:, 1] + 0.5 * aw)) / aw,
                  np.log(gh / ah), np.log(gw / aw)], axis=1)
    t = np.clip(t, -4.0, 4.0)
    d = 0.8 * t + delta_noise * rng.standard_normal((A, 4))
    bbox = (d / np.asarray(std_dev, dtype=np.float64)).astype(np.float32)
    return probs, bbox


def feature_maps(rng, batch, img_size, channels=256, strides=(4, 8, 16, 32)):
    """P2..P5, NHWC fp32 N(0,1)."""
    return [rng.standard_normal((batch, s[0], s[1], channels), dtype=np.float32)
            for s in backbone_shapes(img_size, strides)]


def head_outputs(rng, batch, n_rois, num_classes):
    """mrcnn_class = softmax(3 N(0,1)) [B,N,NC]; mrcnn_bbox = N(0,1) [B,N,NC,4]."""
    probs = _softmax(3.0 * rng.standard_normal((batch, n_rois, num_classes), dtype=np.float32)).astype(np.float32)
    bbox = rng.standard_normal((batch, n_rois, num_classes, 4), dtype=np.float32)
    return probs, bbox


def inference_batch(config_id, batch, img_size=1024, num_classes=81, regime='clustered', n_rois=1000,
                    channels=256, first_image=0):
    """All inputs of the inference ROI stage for `batch` images (numpy, host)."""
    anchors1 = pyramid_anchors(img_size)
    A = anchors1.shape[0]
    probs = np.empty((batch, A, 2), dtype=np.float32)
    bbox = np.empty((batch, A, 4), dtype=np.float32)
    fmaps = [np.empty((batch, s[0], s[1], channels), dtype=np.float32)
             for s in backbone_shapes(img_size, (4, 8, 16, 32))]
    mc = np.empty((batch, n_rois, num_classes), dtype=np.float32)
    mb = np.empty((batch, n_rois, num_classes, 4), dtype=np.float32)
    for b in range(batch):
        rng = np.random.default_rng(1000 * config_id + first_image + b)
        probs[b], bbox[b] = rpn_outputs(rng, anchors1, regime, img_size)
        for lvl, fm in enumerate(feature_maps(rng, 1, img_size, channels)):
            fmaps[lvl][b] = fm[0]
        c, d = head_outputs(rng, 1, n_rois, num_classes)
        mc[b], mb[b] = c[0], d[0]
    anchors = np.ascontiguousarray(np.broadcast_to(anchors1, (batch,) + anchors1.shape))
    return dict(rpn_probs=probs, rpn_bbox=bbox, anchors=anchors, feature_maps=fmaps, mrcnn_class=mc, mrcnn_bbox=mb,
                image_meta=image_meta(batch, img_size, num_classes))


def training_targets_batch(config_id, batch, img_size=1024, num_classes=81, max_gt=100, n_real=20, mini_mask=None,
                           first_image=0):
    """GT tensors for DetectionTargetLayer: class ids [B,G] int32, normalised boxes [B,G,4], masks uint8."""
    cls = np.zeros((batch, max_gt), dtype=np.int32)
    boxes = np.zeros((batch, max_gt, 4), dtype=np.float32)
    masks = []
    for b in range(batch):
        rng = np.random.default_rng(1000 * config_id + 500 + first_image + b)
        bp, c = gt_instances(rng, img_size, n_real, max_gt, num_classes, crowd=(rng.uniform() < 0.1))
        cls[b] = c
        nb = norm_boxes(bp, img_size)
        boxes[b] = nb
        masks.append(ellipse_masks(bp, img_size, mini_mask))
    return dict(gt_class_ids=cls, gt_boxes=boxes, gt_masks=np.stack(masks, axis=0))
