"""Generates tests/golden/reference_layers_golden.npz by EXECUTING THE REFERENCE'S OWN LAYER CODE in this container
(`python tests/golden/make_reference_layers_golden.py`; needs /root/reference, so it cannot run on the GPU box -- the
vectors it writes are committed).

`src/layers/mrcnn_layers.py` is loaded from where it lies and its `ProposalLayer.call`, `PyramidROIAlign.call` and
`DetectionLayer.call` / `refine_detections` (L:233-269, 583-664, 369-524) -- with `utils.batch_slice`,
`apply_box_deltas_graph`, `clip_boxes_graph`, `parse_image_meta_graph`, `log2_graph` and `NormBoxesLayer` underneath --
run UNMODIFIED on top of `numpy_tf` below: a numpy stand-in for the ~55 `tf.*` functions those lines call
(TensorFlow itself cannot be installed here).  What this pins, and what it does not:

  * pinned by the reference's own Python: everything the layers DO with the ops -- operation order of the fp32
    arithmetic, the std-dev scaling, per-image slicing, the level formula with its 244 denominator (Q1), the
    first-appearance map table and the `batch*100000+box` re-sort (Q2), the single class-agnostic NMS over the kept
    set (Q3), the broadcast intersections, padding, the window normalised with image 0's shape, the final reshape;
  * NOT pinned (restated here from the TF kernels' documented behaviour): the bodies of
    `tf.nn.top_k`, `tf.image.non_max_suppression`, `tf.image.crop_and_resize` -- written below in plain numpy,
    independently of oracle/; and `tf.exp` / `tf.math.log`, which call the oracle's correctly-rounded-to-<1ulp
    routines so that the stored boxes are bit-comparable (numpy's own expf/logf differ from them in the last bit on
    some inputs, as TF's Eigen versions differ from each other).

Nothing is copied from the reference: only its outputs are stored.
"""
import contextlib
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, ROOT)

f32 = np.float32
SHUFFLE = {"calls": 0, "per_image": []}      # [(uint32 keys [P], original row of every non-zero proposal)] per image


# ------------------------------------------------------------------------------------------------ numpy_tf
class T(np.ndarray):
    """ndarray with the two Tensor methods the reference calls."""

    def set_shape(self, shape):
        pass

    def numpy(self):
        return np.asarray(self)


def _t(x, dtype=None):
    return np.asarray(x, dtype=dtype).view(T)


def _np_dtype(d):
    if isinstance(d, str):
        return np.dtype(d)
    return d


def make_numpy_tf():
    import oracle
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = np.float32, np.float64, np.int32, np.int64, np.bool_
    tf.newaxis = None

    def _elem(fn):
        def run(x, *a, **k):
            x = np.asarray(x)
            if x.dtype.kind != "f":
                x = x.astype(f32)                       # python floats become float32 tensors, as in TF
            return _t(fn(x.astype(f32)).reshape(x.shape))
        return run

    tf.exp = _elem(oracle.expf)                         # see the module docstring
    math = types.SimpleNamespace(
        log=_elem(oracle.logf), maximum=lambda a, b, **k: _t(np.maximum(a, b)),
        minimum=lambda a, b, **k: _t(np.minimum(a, b)), multiply=lambda a, b, **k: _t(np.multiply(a, b)),
        divide=lambda a, b, **k: _t(np.divide(a, b)),
        # Eigen's MaxReducer starts from lowest(): that is what a reduction over an empty axis returns
        reduce_max=lambda x, axis=None, **k: _t(np.max(x, axis=axis, initial=np.finfo(f32).min)
                                                if np.asarray(x).size == 0 else np.max(x, axis=axis)),
        reduce_sum=lambda x, axis=None, **k: _t(np.sum(x, axis=axis, dtype=np.asarray(x).dtype)))
    tf.math = math
    tf.maximum, tf.minimum, tf.reduce_sum = math.maximum, math.minimum, math.reduce_sum
    def sqrt(x, **k):                                                # IEEE correctly rounded everywhere
        with np.errstate(invalid="ignore"):                          # h*w < 0 (flipped ROI) -> NaN -> level 2, as in TF
            return _t(np.sqrt(np.asarray(x, f32)))
    tf.sqrt = sqrt
    tf.abs = lambda x, **k: _t(np.abs(x))
    tf.round = lambda x, **k: _t(np.round(x))                         # half to even, like tf.round
    tf.equal = lambda a, b, **k: _t(np.equal(a, b))
    tf.greater = lambda a, b, **k: _t(np.greater(a, b))
    tf.logical_and = lambda a, b, **k: _t(np.logical_and(a, b))
    tf.identity = lambda x, **k: _t(x)
    tf.stop_gradient = lambda x, **k: _t(x)
    tf.constant = lambda v, dtype=None, **k: _t(v, _np_dtype(dtype) if dtype is not None else
                                                 (f32 if np.asarray(v).dtype.kind == "f" else None))

    def cast(x, dtype, **k):
        with np.errstate(invalid="ignore"):             # -inf / NaN -> INT_MIN, what the x86 TF kernel yields too
            return _t(np.asarray(x).astype(_np_dtype(dtype)))
    tf.cast = cast
    tf.shape = lambda x, **k: _t(np.asarray(np.shape(x), np.int32))
    tf.reshape = lambda x, shape, **k: _t(np.reshape(x, tuple(int(s) for s in np.asarray(shape).reshape(-1))))
    tf.squeeze = lambda x, axis=None, **k: _t(np.squeeze(x, axis=axis))
    tf.expand_dims = lambda x, axis, **k: _t(np.expand_dims(x, axis))
    tf.transpose = lambda x, perm=None, **k: _t(np.transpose(x, perm))
    tf.tile = lambda x, m, **k: _t(np.tile(x, tuple(int(v) for v in m)))
    tf.stack = lambda xs, axis=0, **k: _t(np.stack([np.asarray(x) for x in xs], axis=axis))
    tf.concat = lambda xs, axis=0, **k: _t(np.concatenate([np.asarray(x) for x in xs], axis=axis))
    tf.range = lambda *a, **k: _t(np.arange(*[int(v) for v in a], dtype=np.int32))
    tf.argmax = lambda x, axis=None, output_type=np.int64, **k: _t(np.argmax(x, axis=axis).astype(output_type))

    def split(x, n, axis=0, **k):
        x = np.asarray(x)
        if isinstance(n, (int, np.integer)):
            return [_t(p) for p in np.split(x, int(n), axis=axis)]
        return [_t(p) for p in np.split(x, np.cumsum([int(v) for v in n])[:-1], axis=axis)]
    tf.split = split

    def pad(x, paddings, mode="CONSTANT", constant_values=0, **k):
        assert mode == "CONSTANT"
        p = [(int(a), int(b)) for a, b in np.asarray(paddings).reshape(-1, 2)]
        return _t(np.pad(np.asarray(x), p, mode="constant", constant_values=constant_values))
    tf.pad = pad

    def gather(params, indices, axis=0, **k):
        return _t(np.take(np.asarray(params), np.asarray(indices).astype(np.int64), axis=axis))
    tf.gather = gather

    def gather_nd(params, indices, **k):
        idx = np.asarray(indices).astype(np.int64)
        return _t(np.asarray(params)[tuple(idx[..., d] for d in range(idx.shape[-1]))])
    tf.gather_nd = gather_nd

    def where(cond, x=None, y=None, **k):
        if x is None:
            return _t(np.argwhere(np.asarray(cond)).astype(np.int64))     # row-major coordinates, int64
        return _t(np.where(cond, x, y))
    tf.where = where

    def boolean_mask(x, mask, axis=None, **k):
        return _t(np.asarray(x)[np.asarray(mask).astype(bool)])
    tf.boolean_mask = boolean_mask

    def unique(x, **k):                                                   # values in first-occurrence order
        x = np.asarray(x)
        vals, first, inv = np.unique(x, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        rank = np.empty_like(order)
        rank[order] = np.arange(order.size)
        return _t(vals[order]), _t(rank[inv].astype(np.int32))
    tf.unique = unique

    class TopK(tuple):
        values = property(lambda s: s[0])
        indices = property(lambda s: s[1])

    def top_k(x, k=1, sorted=True, **kw):
        """TopKV2: descending, equal values -> lower index first; int32 indices."""
        x = np.asarray(x)
        k = int(k)
        idx = np.argsort(-x.astype(np.float64) if x.dtype.kind == "f" else -x.astype(np.int64), axis=-1,
                         kind="stable")[..., :k].astype(np.int32)
        return TopK((_t(np.take_along_axis(x, idx.astype(np.int64), axis=-1)), _t(idx)))
    tf.nn = types.SimpleNamespace(top_k=top_k)

    def iou_many(boxes, i, js):
        """NonMaxSuppressionV3's IOU() of box i against boxes js: corners min/max-normalised, non-positive areas -> 0;
        every operation an individually rounded fp32 numpy operation, in the kernel's order."""
        a, b = boxes[i], boxes[js]
        ymin_i, xmin_i, ymax_i, xmax_i = min(a[0], a[2]), min(a[1], a[3]), max(a[0], a[2]), max(a[1], a[3])
        ymin_j, xmin_j = np.minimum(b[:, 0], b[:, 2]), np.minimum(b[:, 1], b[:, 3])
        ymax_j, xmax_j = np.maximum(b[:, 0], b[:, 2]), np.maximum(b[:, 1], b[:, 3])
        area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
        area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
        ih = np.maximum(np.minimum(ymax_i, ymax_j) - np.maximum(ymin_i, ymin_j), f32(0))
        iw = np.maximum(np.minimum(xmax_i, xmax_j) - np.maximum(xmin_i, xmin_j), f32(0))
        inter = ih * iw
        with np.errstate(divide="ignore", invalid="ignore"):
            v = inter / (area_i + area_j - inter)
        return np.where((area_i <= 0) | (area_j <= 0), f32(0), v)

    def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5, score_threshold=float("-inf"), **k):
        """Greedy hard NMS: candidates by (score desc, index asc); kept iff IoU with every kept box is not > thr."""
        boxes, scores = np.asarray(boxes, f32), np.asarray(scores, f32)
        thr = f32(iou_threshold)
        order = [int(i) for i in np.argsort(-scores.astype(np.float64), kind="stable") if scores[i] > score_threshold]
        keep = []
        for i in order:
            if len(keep) >= int(max_output_size):
                break
            if not keep or not np.any(iou_many(boxes, i, np.asarray(keep)) > thr):
                keep.append(i)
        return _t(np.asarray(keep, np.int32))

    def crop_and_resize(image, boxes, box_indices, crop_size, method="bilinear", extrapolation_value=0.0, **k):
        """CropAndResize, bilinear: one sample per output bin at y1*(H-1) + y*(y2-y1)*(H-1)/(ph-1)."""
        assert method == "bilinear"
        image, boxes = np.asarray(image, f32), np.asarray(boxes, f32)
        ph, pw = int(crop_size[0]), int(crop_size[1])
        _, H, W, C = image.shape
        out = np.zeros((boxes.shape[0], ph, pw, C), f32)

        def taps(c1, c2, size, crop):
            if crop > 1:
                scale = (c2 - c1) * f32(size - 1) / f32(crop - 1)
                pos = c1 * f32(size - 1) + np.arange(crop, dtype=f32) * scale
            else:
                pos = np.asarray([f32(0.5) * (c1 + c2) * f32(size - 1)], f32)
            ok = (pos >= 0) & (pos <= f32(size - 1))
            safe = np.where(ok, pos, f32(0))
            lo, hi = np.floor(safe).astype(np.int64), np.ceil(safe).astype(np.int64)
            return ok, lo, hi, (safe - lo.astype(f32)).astype(f32)

        for n in range(boxes.shape[0]):
            img = image[int(box_indices[n])]
            y1, x1, y2, x2 = boxes[n]
            oky, ylo, yhi, ly = taps(y1, y2, H, ph)
            okx, xlo, xhi, lx = taps(x1, x2, W, pw)
            tl, tr = img[ylo[:, None], xlo[None, :]], img[ylo[:, None], xhi[None, :]]
            bl, br = img[yhi[:, None], xlo[None, :]], img[yhi[:, None], xhi[None, :]]
            lxb, lyb = lx[None, :, None], ly[:, None, None]
            top = tl + (tr - tl) * lxb
            bot = bl + (br - bl) * lxb
            val = top + (bot - top) * lyb
            out[n] = np.where((oky[:, None] & okx[None, :])[..., None], val, f32(extrapolation_value))
        return _t(out)

    tf.image = types.SimpleNamespace(non_max_suppression=non_max_suppression, crop_and_resize=crop_and_resize)
    tf.Assert = lambda *a, **k: None
    tf.Variable = lambda v, **k: _t(v)
    tf.control_dependencies = lambda deps: contextlib.nullcontext()

    def shuffle(x, **k):
        """tf.random.shuffle is unseeded in the reference (any permutation is reference behaviour).  The
        stand-in draws THE permutation the B200 layer draws from the same per-proposal random keys: ascending
        (key of the proposal's original row, index) -- see SHUFFLE below."""
        keys, rows = SHUFFLE["per_image"][SHUFFLE["calls"] // 2]      # two calls per image: positives, negatives
        SHUFFLE["calls"] += 1
        x = np.asarray(x)
        order = sorted(range(x.size), key=lambda i: (int(keys[rows[int(x[i])]]), int(x[i])))
        return _t(x[order])
    tf.random = types.SimpleNamespace(shuffle=shuffle)
    tf.function = lambda fn=None, **k: fn if fn is not None else (lambda f: f)
    tf.map_fn = lambda fn, elems, **k: _t(np.stack([np.asarray(fn(e)) for e in elems]))
    tf.cond = lambda pred, true_fn, false_fn, **k: true_fn() if bool(pred) else false_fn()

    class Layer:
        def __init__(self, name=None, **kwargs):
            self.name, self.built = name, False

        def build(self, input_shape):
            self.built = True

        def __call__(self, inputs, **kwargs):
            # Keras autocasts floating inputs to the layer's dtype (float32): AnchorsLayer hands float64 numpy anchors
            # to NormBoxesLayer this way (mrcnn_layers.py:116-132)
            def autocast(v):
                if isinstance(v, (list, tuple)):
                    return type(v)(autocast(e) for e in v)
                if isinstance(v, np.ndarray) and v.dtype == np.float64:
                    return _t(v.astype(f32))
                return v
            return self.call(autocast(inputs), **kwargs)

        def get_config(self):
            return {"name": self.name}

    class _Any(types.ModuleType):
        def __getattr__(self, name):                    # Conv2D, BatchNormalization ...: only ever subclassed / built
            if name.startswith("__"):
                raise AttributeError(name)
            return type(name, (Layer,), {})

    keras = types.ModuleType("tensorflow.keras")
    keras.layers = _Any("tensorflow.keras.layers")
    keras.layers.Layer = Layer
    keras.utils = types.SimpleNamespace(register_keras_serializable=lambda *a, **k: (lambda cls: cls))
    keras.backend = _Any("tensorflow.keras.backend")
    tf.keras = keras
    return tf


def load_reference_layers(tf=None):
    """Loads /root/reference/src/{common/utils.py, layers/mrcnn_layers.py} on top of numpy_tf (or of another stand-in
    module with the same surface: tests/test_reference_gradients.py passes a torch-backed one for autograd)."""
    tf = tf or make_numpy_tf()
    stubs = {"tensorflow": tf, "tensorflow.keras": tf.keras, "tensorflow.keras.layers": tf.keras.layers,
             "tensorflow.keras.backend": getattr(tf.keras, "backend", types.ModuleType("tensorflow.keras.backend"))}
    for name in ("skimage", "skimage.transform", "efficientnet", "efficientnet.keras", "layers", "layers.backbones",
                 "layers.backbones.models_factory", "common"):
        stubs[name] = types.ModuleType(name)
    stubs["layers.backbones.models_factory"].Classifiers = object
    stubs["efficientnet"].keras = stubs["efficientnet.keras"]
    try:
        import distutils.version  # noqa: F401
    except Exception:
        d, dv = types.ModuleType("distutils"), types.ModuleType("distutils.version")
        dv.LooseVersion = str
        stubs["distutils"], stubs["distutils.version"] = d, dv
    saved = {n: sys.modules.get(n) for n in stubs}
    sys.modules.update(stubs)
    try:
        def load(name, path):
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod
        utils = load("common.utils", os.path.join(REF, "common", "utils.py"))
        stubs["common"].utils = utils
        layers = load("reference_mrcnn_layers", os.path.join(REF, "layers", "mrcnn_layers.py"))
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
        sys.modules.pop("common.utils", None)
        sys.modules.pop("reference_mrcnn_layers", None)
    return layers


def build():
    from maskrcnn_tf2_b200 import synth
    L = load_reference_layers()
    S, B, NC, C = 128, 3, 5, 8
    K, P, D = 600, 100, 20
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": K, "images_per_gpu": B,
           "rpn_bbox_std_dev": np.array([0.1, 0.1, 0.2, 0.2], dtype="float32"),      # config.py:90-91
           "bbox_std_dev": np.array([0.1, 0.1, 0.2, 0.2], dtype="float32")}
    anchors1 = synth.pyramid_anchors(S)
    A = anchors1.shape[0]
    rng = np.random.default_rng(20261018)
    probs, bbox = zip(*[synth.rpn_outputs(np.random.default_rng(900 + b), anchors1, "clustered", S) for b in range(B)])
    probs, bbox = np.stack(probs).astype(f32), np.stack(bbox).astype(f32)
    probs[1, :, 1] = np.round(probs[1, :, 1] * 64) / 64                   # image 1: heavy score ties
    probs[1, :, 0] = 1 - probs[1, :, 1]
    bbox[2] *= 0.05                                                        # image 2: near-duplicate boxes, < P survivors
    anchors = np.ascontiguousarray(np.broadcast_to(anchors1, (B, A, 4))).astype(f32)
    meta = synth.image_meta(B, S, NC).astype(f32)
    meta[1, 7:11] = [10, 6, 120, 100]                                     # a padded image: window inside the canvas
    fmaps = [rng.standard_normal((B, S // s, S // s, C)).astype(f32) for s in (4, 8, 16, 32)]
    mrcnn_class, mrcnn_bbox = synth.head_outputs(rng, B, P, NC)
    mrcnn_class[2, :, 1:] *= 0.01                                          # image 2: (almost) background only
    mrcnn_class[2, :, 0] = 1 - mrcnn_class[2, :, 1:].sum(-1)

    t = lambda a: np.asarray(a).view(T)
    rois = np.asarray(L.ProposalLayer(proposal_count=P, config=cfg)([t(probs), t(bbox), t(anchors)]))
    assert rois.shape == (B, P, 4) and rois.dtype == f32
    # more slots than survivors: the zero padding of L:229-230
    rois_p400 = np.asarray(L.ProposalLayer(proposal_count=400, config=cfg)([t(probs), t(bbox), t(anchors)]))
    pooled = np.asarray(L.PyramidROIAlign([7, 7], name="roi_align_classifier")([t(rois), t(meta)] + [t(f) for f in fmaps]))
    det = np.asarray(L.DetectionLayer(proposals=P, detection_min_confidence=0.7, detection_max_instances=D,
                                      detection_nms_threshold=0.3, bbox_std_dev=cfg["bbox_std_dev"], images_per_gpu=B,
                                      batch_size=B)([t(rois), t(mrcnn_class), t(mrcnn_bbox), t(meta)]))
    assert det.shape == (B, D, 6) and det.dtype == f32
    det_boxes = np.ascontiguousarray(det[..., :4])
    mask_pooled = np.asarray(L.PyramidROIAlign([14, 14], name="roi_align_mask")([t(det_boxes), t(meta)] + [t(f) for f in fmaps]))
    # quirk Q2 case: the batch's first ROI is a large one, so its (coarser) level appears first and takes feature map 0
    # (P2), and level 2 -- met second -- samples P3
    q2_boxes = np.array([[[0.05, 0.05, 0.95, 0.95], [0.4, 0.4, 0.45, 0.46], [0.1, 0.2, 0.5, 0.7], [0, 0, 0, 0]],
                         [[0.3, 0.3, 0.33, 0.34], [0.0, 0.0, 1.0, 1.0], [0.2, 0.1, 0.35, 0.3], [0.6, 0.6, 1.2, 1.3]],
                         [[0, 0, 0, 0], [0, 0, 0, 0], [0.5, 0.5, 0.75, 0.8], [0.1, 0.1, 0.2, 0.2]]], f32)
    q2_pooled = np.asarray(L.PyramidROIAlign([3, 5], name="roi_align_q2")([t(q2_boxes), t(meta)] + [t(f) for f in fmaps]))
    # DetectionLayer without a confidence filter (detection_min_confidence = 0 skips L:404-414)
    det0 = np.asarray(L.DetectionLayer(proposals=P, detection_min_confidence=0, detection_max_instances=D,
                                       detection_nms_threshold=0.3, bbox_std_dev=cfg["bbox_std_dev"], images_per_gpu=B,
                                       batch_size=B)([t(rois), t(mrcnn_class), t(mrcnn_bbox), t(meta)]))
    kept = [int((rois[b].any(-1)).sum()) for b in range(B)]
    dets = [int((det[b, :, 4] > 0).sum()) for b in range(B)]
    print(f"P=400: kept {[int((rois_p400[b].any(-1)).sum()) for b in range(B)]}")
    print(f"proposals kept per image {kept} of {P}; detections per image {dets} of {D}; "
          f"no-confidence detections {[int((det0[b, :, 4] > 0).sum()) for b in range(B)]}")
    out = dict(img_size=S, pre_nms_limit=K, proposal_count=P, max_instances=D, num_classes=NC,
               rpn_probs=probs, rpn_bbox=bbox, anchors=anchors, image_meta=meta, mrcnn_class=mrcnn_class.astype(f32),
               mrcnn_bbox=mrcnn_bbox.astype(f32), rois=rois, pooled=pooled, detections=det, mask_pooled=mask_pooled,
               q2_boxes=q2_boxes, q2_pooled=q2_pooled, detections_noconf=det0, rois_p400=rois_p400)
    for i, f in enumerate(fmaps):
        out[f"fmap{i}"] = f
    path = os.path.join(HERE, "reference_layers_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def build_targets():
    """DetectionTargetLayer.call -> detection_targets_graph (+ trim_zeros_graph, overlaps_graph,
    utils.box_refinement_graph; mrcnn_layers.py:313-325, 844-1007, utils.py:775-798), full-size and mini masks."""
    L = load_reference_layers()
    rng = np.random.default_rng(20261019)
    B, P, G, T_, MH = 3, 300, 12, 32, 40
    out = dict(train_rois_per_image=T_, roi_positive_ratio=0.33, mask_shape=np.array([14, 14]))

    def boxes(n, lo, hi):
        c = rng.uniform(0.2, 0.8, (n, 2))
        s = rng.uniform(lo, hi, (n, 2))
        return np.clip(np.concatenate([c - s / 2, c + s / 2], 1), 0, 1).astype(f32)

    gtb = np.zeros((B, G, 4), f32)
    gtc = np.zeros((B, G), np.int32)
    props = np.zeros((B, P, 4), f32)
    for b in range(B):
        n_real = 6 - b
        gtb[b, :n_real] = boxes(n_real, 0.15, 0.5)
        gtc[b, :n_real] = rng.integers(1, 81, n_real)
        jit = np.repeat(gtb[b, :n_real], 30, 0) + rng.normal(0, 0.02, (30 * n_real, 4)).astype(f32)
        n_prop = P - 40 - 10 * b                                   # the rest stays zero padding (trimmed, L:871)
        props[b, :jit.shape[0]] = np.clip(jit, 0, 1)
        props[b, jit.shape[0]:n_prop] = boxes(n_prop - jit.shape[0], 0.05, 0.4)
        perm = rng.permutation(n_prop)
        props[b, :n_prop] = props[b, :n_prop][perm]
    gtc[0, 2] *= -1                                                # a crowd box (negative class id, L:879-884)
    gtb[1, 3], gtb[1, 5] = gtb[1, 5].copy(), gtb[1, 3].copy()      # a zero GT row in the middle of the list (image 1)
    gtc[1, 3], gtc[1, 5] = gtc[1, 5], gtc[1, 3]
    props[2, 5] = 0                                                # a zero proposal in the middle of the list
    keys = rng.integers(0, 2 ** 32, (B, P), dtype=np.uint64).astype(np.uint32)
    keys[0, ::3] = keys[0, 1::3]                                   # tied keys: the index decides
    t = lambda a: np.asarray(a).view(T)
    for mini in (False, True):
        mh = 16 if mini else MH
        yy, xx = np.mgrid[0:mh, 0:mh]
        masks = np.zeros((B, mh, mh, G), bool)
        for b in range(B):
            for g in range(G):
                cy, cx, ry, rx = rng.uniform(0.3, 0.7) * mh, rng.uniform(0.3, 0.7) * mh, rng.uniform(0.15, 0.45) * mh, \
                    rng.uniform(0.15, 0.45) * mh
                masks[b, :, :, g] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        cfg = {"train_rois_per_image": T_, "roi_positive_ratio": 0.33, "use_mini_masks": mini, "mask_shape": (14, 14),
               "bbox_std_dev": np.array([0.1, 0.1, 0.2, 0.2], dtype="float32"), "images_per_gpu": B}
        SHUFFLE["calls"] = 0
        SHUFFLE["per_image"] = [(keys[b], np.flatnonzero(np.abs(props[b]).sum(1) != 0)) for b in range(B)]
        rois, cls, deltas, tm = L.DetectionTargetLayer(cfg)([t(props), t(gtc), t(gtb), t(masks)])
        assert SHUFFLE["calls"] == 2 * B
        tag = "mini" if mini else "full"
        out.update({f"gt_masks_{tag}": masks, f"rois_{tag}": np.asarray(rois, f32), f"class_ids_{tag}": np.asarray(cls),
                    f"deltas_{tag}": np.asarray(deltas, f32), f"masks_{tag}": np.asarray(tm, f32)})
        print(tag, "positives per image", [(np.asarray(cls)[b] != 0).sum() for b in range(B)], "rois per image",
              [int(np.asarray(rois)[b].any(-1).sum()) for b in range(B)], "mask mean", float(np.asarray(tm).mean()))
    out.update(proposals=props, gt_class_ids=gtc, gt_boxes=gtb, rand_keys=keys)
    path = os.path.join(HERE, "reference_target_layer_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def full_size_inputs(batch=2):
    """Config 2 of BASELINE.json at full size (1024^2, A = 261 888, 81 classes, C = 256), `batch` images: regenerated
    from seeds wherever needed (the GPU test rebuilds the same arrays), never stored."""
    from maskrcnn_tf2_b200 import synth
    return synth.inference_batch(2, batch, img_size=1024, num_classes=81, regime="clustered", n_rois=1000, channels=256)


def run_full_size(L, x):
    """The inference ROI stage of model.py:556-573 through the reference's own layers at COCO shape."""
    B = x["rpn_probs"].shape[0]
    sd = np.array([0.1, 0.1, 0.2, 0.2], dtype="float32")
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": 6000, "images_per_gpu": B, "rpn_bbox_std_dev": sd,
           "bbox_std_dev": sd}
    t = lambda a: np.asarray(a).view(T)
    fm = [t(f) for f in x["feature_maps"]]
    meta = t(x["image_meta"].astype(f32))
    rois = np.asarray(L.ProposalLayer(proposal_count=1000, config=cfg)([t(x["rpn_probs"]), t(x["rpn_bbox"]),
                                                                       t(x["anchors"])]))
    pooled = np.asarray(L.PyramidROIAlign([7, 7], name="roi_align_classifier")([t(rois), meta] + fm))
    det = np.asarray(L.DetectionLayer(proposals=1000, detection_min_confidence=0.7, detection_max_instances=100,
                                      detection_nms_threshold=0.3, bbox_std_dev=sd, images_per_gpu=B, batch_size=B)(
        [t(rois), t(x["mrcnn_class"]), t(x["mrcnn_bbox"]), meta]))
    mask_pooled = np.asarray(L.PyramidROIAlign([14, 14], name="roi_align_mask")(
        [t(np.ascontiguousarray(det[..., :4])), meta] + fm))
    return dict(rois=rois, pooled=pooled, detections=det, mask_pooled=mask_pooled)


def build_full_size():
    """SHA-256 of the reference layers' outputs at BASELINE.json's full sizes (the arrays themselves are 120 MB)."""
    import json
    import time
    L = load_reference_layers()
    x = full_size_inputs()
    t0 = time.time()
    out = run_full_size(L, x)
    kept = [int(out["rois"][b].any(-1).sum()) for b in range(out["rois"].shape[0])]
    dets = [int((out["detections"][b, :, 4] > 0).sum()) for b in range(out["rois"].shape[0])]
    rec = {"inputs": "synth.inference_batch(2, 2, img_size=1024, num_classes=81, regime='clustered', n_rois=1000, "
                     "channels=256)",
           "input_sha256": {k: digest(x[k]) for k in ("rpn_probs", "rpn_bbox", "anchors", "mrcnn_class", "mrcnn_bbox",
                                                      "image_meta")},
           "fmap_sha256": [digest(f) for f in x["feature_maps"]],
           "proposals_kept": kept, "detections": dets,
           "shape": {k: list(v.shape) for k, v in out.items()},
           "sha256": {k: digest(v.astype(f32)) for k, v in out.items()}}
    path = os.path.join(HERE, "reference_layers_full_size_sha256.json")
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)
    print(f"full size: reference layers took {time.time() - t0:.1f} s; kept {kept}, detections {dets}; wrote {path}")


def training_inputs():
    """The training-step case of tests/test_gpu_configs.py (config 3 of BASELINE.json at 512^2, two images): 2000
    proposals per image, 100 GT slots with 20 real instances and full-size masks, 200 target ROIs."""
    from maskrcnn_tf2_b200 import synth
    x = synth.inference_batch(3, 2, img_size=512, regime="clustered")
    g = synth.training_targets_batch(3, 2, img_size=512)
    keys = np.random.default_rng(7).integers(0, 2 ** 32, (2, 2000), dtype=np.uint64).astype(np.uint32)
    return x, g, keys


def run_training(L, x, g, keys):
    B = 2
    sd = np.array([0.1, 0.1, 0.2, 0.2], dtype="float32")
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": 6000, "images_per_gpu": B, "rpn_bbox_std_dev": sd,
           "bbox_std_dev": sd, "train_rois_per_image": 200, "roi_positive_ratio": 0.33, "use_mini_masks": False,
           "mask_shape": (28, 28)}
    t = lambda a: np.asarray(a).view(T)
    props = np.asarray(L.ProposalLayer(proposal_count=2000, config=cfg)([t(x["rpn_probs"]), t(x["rpn_bbox"]),
                                                                        t(x["anchors"])]))
    SHUFFLE["calls"] = 0
    SHUFFLE["per_image"] = [(keys[b], np.flatnonzero(np.abs(props[b]).sum(1) != 0)) for b in range(B)]
    rois, cls, deltas, masks = L.DetectionTargetLayer(cfg)([t(props), t(g["gt_class_ids"]), t(g["gt_boxes"]),
                                                            t(g["gt_masks"].astype(bool))])
    return dict(proposals=props, rois=np.asarray(rois, f32), class_ids=np.asarray(cls).astype(np.int32),
                deltas=np.asarray(deltas, f32), masks=np.asarray(masks, f32))


def build_full_size_training():
    import json
    L = load_reference_layers()
    x, g, keys = training_inputs()
    out = run_training(L, x, g, keys)
    path = os.path.join(HERE, "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))
    rec["training"] = {
        "inputs": "synth.inference_batch(3, 2, img_size=512, regime='clustered'); synth.training_targets_batch(3, 2, "
                  "img_size=512); keys = default_rng(7).integers(0, 2**32, (2, 2000), uint64).astype(uint32)",
        "input_sha256": {"rpn_probs": digest(x["rpn_probs"]), "rpn_bbox": digest(x["rpn_bbox"]),
                         "gt_class_ids": digest(g["gt_class_ids"]), "gt_boxes": digest(g["gt_boxes"]),
                         "gt_masks": digest(g["gt_masks"]), "keys": digest(keys)},
        "positives": [int((out["class_ids"][b] != 0).sum()) for b in range(2)],
        "rois_per_image": [int(out["rois"][b].any(-1).sum()) for b in range(2)],
        "sha256": {k: digest(v) for k, v in out.items()}}
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)
    print("training case:", rec["training"]["positives"], rec["training"]["rois_per_image"], "-> added to", path)


def training_full_inputs(mini):
    """Config 3 of BASELINE.json at FULL size: 1024^2, 8 images, 2000 proposals per image into DetectionTargetLayer
    (T = 200, 100 GT slots with 20 real instances; full-size 1024^2 masks, or 32x32 mini-masks)."""
    from maskrcnn_tf2_b200 import synth
    B = 8
    x = synth.inference_batch(3, B, img_size=1024, regime="clustered")
    g = synth.training_targets_batch(3, B, img_size=1024, mini_mask=(32, 32) if mini else None)
    keys = np.random.default_rng(9).integers(0, 2 ** 32, (B, 2000), dtype=np.uint64).astype(np.uint32)
    return x, g, keys


def run_training_full(L, x, g, keys, mini, props=None):
    """ProposalLayer(2000) -> DetectionTargetLayer -> PyramidROIAlign 7x7 and 14x14 on the target ROIs (model.py:502-504,
    mrcnn_layers.py:1145,1215), all through the reference's own layer code."""
    B = x["rpn_probs"].shape[0]
    sd = np.array([0.1, 0.1, 0.2, 0.2], dtype="float32")
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": 6000, "images_per_gpu": B, "rpn_bbox_std_dev": sd,
           "bbox_std_dev": sd, "train_rois_per_image": 200, "roi_positive_ratio": 0.33, "use_mini_masks": bool(mini),
           "mask_shape": (28, 28)}
    t = lambda a: np.asarray(a).view(T)
    if props is None:
        props = np.asarray(L.ProposalLayer(proposal_count=2000, config=cfg)([t(x["rpn_probs"]), t(x["rpn_bbox"]),
                                                                            t(x["anchors"])]))
    SHUFFLE["calls"] = 0
    SHUFFLE["per_image"] = [(keys[b], np.flatnonzero(np.abs(props[b]).sum(1) != 0)) for b in range(B)]
    rois, cls, deltas, masks = L.DetectionTargetLayer(cfg)([t(props), t(g["gt_class_ids"]), t(g["gt_boxes"]),
                                                            t(g["gt_masks"].astype(bool))])
    out = dict(proposals=props, rois=np.asarray(rois, f32), class_ids=np.asarray(cls).astype(np.int32),
               deltas=np.asarray(deltas, f32), masks=np.asarray(masks, f32))
    if not mini:
        fm = [t(f) for f in x["feature_maps"]]
        meta = t(x["image_meta"].astype(f32))
        out["pooled7"] = np.asarray(L.PyramidROIAlign([7, 7], name="roi_align_classifier")([t(out["rois"]), meta] + fm), f32)
        out["pooled14"] = np.asarray(L.PyramidROIAlign([14, 14], name="roi_align_mask")([t(out["rois"]), meta] + fm), f32)
    return out


def build_full_size_training_config3():
    """SHA-256 of the reference layers' outputs for config 3 at full size (and its mini-mask variant)."""
    import json
    import time
    L = load_reference_layers()
    path = os.path.join(HERE, "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))
    props = None
    for mini in (False, True):
        t0 = time.time()
        x, g, keys = training_full_inputs(mini)
        out = run_training_full(L, x, g, keys, mini, props)
        props = out["proposals"]
        B = props.shape[0]
        rec["training_full_mini" if mini else "training_full"] = {
            "inputs": "synth.inference_batch(3, 8, img_size=1024, regime='clustered'); synth.training_targets_batch(3, 8, "
                      "img_size=1024, mini_mask=%s); keys = default_rng(9).integers(0, 2**32, (8, 2000), uint64)"
                      ".astype(uint32)" % ("(32, 32)" if mini else "None"),
            "input_sha256": {"rpn_probs": digest(x["rpn_probs"]), "rpn_bbox": digest(x["rpn_bbox"]),
                             "gt_class_ids": digest(g["gt_class_ids"]), "gt_boxes": digest(g["gt_boxes"]),
                             "gt_masks": digest(g["gt_masks"]), "keys": digest(keys)},
            "positives": [int((out["class_ids"][b] != 0).sum()) for b in range(B)],
            "rois_per_image": [int(out["rois"][b].any(-1).sum()) for b in range(B)],
            "proposals_kept": [int(props[b].any(-1).sum()) for b in range(B)],
            "sha256": {k: digest(v) for k, v in out.items()}}
        print("config 3 full size, mini =", mini, rec["training_full_mini" if mini else "training_full"]["positives"],
              f"{time.time() - t0:.1f} s")
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)


def build_full_size_config5():
    """Config 5 of BASELINE.json: the batch-64 COCO-shape inference stage in the reference's own first-appearance mode
    (Q2) -- digests for the whole batch on one replica AND for each 32 / 16 / 8-image shard run as its own batch (what
    a 2 / 4 / 8-GPU split computes; the Q2 table is local to a replica).  Feature maps of 64 images are
    5.7 GB, so ROIAlign outputs are hashed image by image from per-shard runs."""
    import json
    import time
    from maskrcnn_tf2_b200 import synth
    L = load_reference_layers()
    path = os.path.join(HERE, "reference_layers_full_size_sha256.json")
    rec = json.load(open(path))
    out = {"inputs": "synth.inference_batch(5, n, img_size=1024, num_classes=81, regime='clustered', first_image=lo) "
                     "for every shard [lo, lo+n) of 64 images, n in (64, 32, 16, 8)", "shards": {}}
    t0 = time.time()
    for n in (64, 32, 16, 8):
        for lo in range(0, 64, n):
            if n < 64 and lo >= 2 * n:          # two shards per split size pin the property; all of them take minutes
                continue
            x = synth.inference_batch(5, n, img_size=1024, num_classes=81, regime="clustered", first_image=lo)
            o = run_full_size(L, x)
            out["shards"][f"{lo}+{n}"] = {k: digest(v.astype(f32)) for k, v in o.items()}
            print("config 5 shard", lo, n, f"{time.time() - t0:.0f} s", flush=True)
            del x, o
    rec["config5"] = out
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)



if __name__ == "__main__":
    build()
    build_targets()
    build_full_size()
    build_full_size_training()
    build_full_size_training_config3()
    build_full_size_config5()
