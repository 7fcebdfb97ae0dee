"""The reference's ROI-stage Keras layer API (src/layers/mrcnn_layers.py) on top of the B200 launchers.

Same class names, constructor arguments, `call(inputs)` argument order, output shapes, layer names and error
behaviour as the reference, so code written against `mrcnnl.ProposalLayer(...)([...])` reads the same here.  The
host language of the reference is Python/TF2 and TensorFlow is not installable in this image, so the tensors are
torch CUDA tensors (device memory + streams only); the TensorFlow custom-op shim that binds the very same
launchers is in tf_shim/ (see INTEGRATION.md).  There is no CPU path: CPU tensors raise.
"""
import numpy as np
import torch

from . import functional as F


class _Layer:
    """Minimal stand-in for tf.keras.layers.Layer: name, build(), __call__ -> call(), get_config()."""

    def __init__(self, name=None, **kwargs):
        self.name = name
        self.built = False
        self._kwargs = kwargs

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, **kwargs):
        if not self.built:
            self.build(None)
        return self.call(inputs, **kwargs)

    def get_config(self):
        return {"name": self.name}


class AnchorsLayer(_Layer):
    """mrcnn_layers.py:104-145: the anchor pyramid of the configured image size, normalised and broadcast to the
    batch, as a constant of the model -- built by one kernel on `device` (csrc/anchors.cu) instead of numpy.
    `call` returns it whatever the input (the reference ignores `inputs` too, L:138-139); `anchors_px` holds the
    float64 pixel anchors the data loader passes to build_rpn_targets (preprocess.py:82,297)."""

    def __init__(self, config, training=False, name="anchors", device=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.config = config
        self.training = training
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        S = config['img_size']
        shapes = [(-(-S // s), -(-S // s)) for s in config['backbone_strides']]      # utils.compute_backbone_shapes
        self.anchors, self.anchors_px = F.anchors_forward(
            config['rpn_anchor_scales'], config['rpn_anchor_ratios'], shapes, config['backbone_strides'],
            config['rpn_anchor_stride'], config['image_shape'][:2], config['batch_size'], dev, return_px=True)

    def call(self, inputs=None, **kwargs):
        return self.anchors


class _ProposalFn(torch.autograd.Function):
    """rpn_bbox receives the reference's gradient (no stop_gradient on proposals, model.py:155-157,168 through
    mrcnn_layers.py:227); rpn_probs and anchors receive none (top_k / NMS indices are not differentiable)."""

    @staticmethod
    def forward(ctx, rpn_probs, rpn_bbox, anchors, pre_nms_limit, proposal_count, std_dev, nms_threshold):
        need_grad = ctx.needs_input_grad[1]
        out = F.proposal_forward(rpn_probs, rpn_bbox, anchors, pre_nms_limit, proposal_count, std_dev, nms_threshold,
                                 debug=need_grad)
        if not need_grad:
            return out
        ctx.save_for_backward(rpn_bbox, anchors, out["topk_idx"], out["keep_idx"])
        ctx.std_dev = std_dev
        return out["proposals"]

    @staticmethod
    def backward(ctx, grad_proposals):
        rpn_bbox, anchors, topk_idx, keep_idx = ctx.saved_tensors
        g = F.proposal_backward(grad_proposals.contiguous(), rpn_bbox, anchors, topk_idx, keep_idx, ctx.std_dev)
        return None, g, None, None, None, None, None


class ProposalLayer(_Layer):
    """mrcnn_layers.py:202-280.  inputs = [rpn_probs [B,A,2], rpn_bbox [B,A,4], anchors [B,A,4]] ->
    proposals [B, proposal_count, 4] in normalised coordinates, zero padded."""

    def __init__(self, proposal_count, config, name='roi', **kwargs):
        super().__init__(name=name, **kwargs)
        self.config = config
        self.proposal_count = proposal_count
        self.nms_threshold = self.config['rpn_nms_threshold']

    def call(self, inputs, **kwargs):
        rpn_probs, rpn_bbox, anchors = inputs[0], inputs[1], inputs[2]
        return _ProposalFn.apply(rpn_probs, rpn_bbox, anchors, self.config['pre_nms_limit'], self.proposal_count,
                                 np.asarray(self.config['rpn_bbox_std_dev'], dtype=np.float32), self.nms_threshold)

    def call_levels(self, rpn_class_logits, rpn_bbox, anchors, return_probs=False):
        """The same layer fed by the RPN head's per-level outputs (lists of [B,A_l,2] logits and [B,A_l,4] deltas, one
        entry per pyramid level): the per-level softmax and the Concatenate layers of model.py:465-478 are fused into
        the launch (inference path: no gradient)."""
        return F.proposal_forward_levels(rpn_class_logits, rpn_bbox, anchors, self.config['pre_nms_limit'],
                                         self.proposal_count,
                                         np.asarray(self.config['rpn_bbox_std_dev'], dtype=np.float32),
                                         self.nms_threshold, return_probs=return_probs)

    def compute_output_shape(self, input_shape):
        return None, self.proposal_count, 4


class _PyramidROIAlignFn(torch.autograd.Function):
    """Feature maps receive CropAndResizeGradImage gradients; boxes and image_meta receive none (L:628-629)."""

    @staticmethod
    def forward(ctx, boxes, image_meta, pool_shape, denominator, map_mode, *feature_maps):
        out, roi_map = F.roialign_forward(boxes, image_meta, feature_maps, pool_shape, denominator, map_mode)
        ctx.save_for_backward(boxes, roi_map)
        ctx.shapes = [tuple(m.shape) for m in feature_maps]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        boxes, roi_map = ctx.saved_tensors
        grads = F.roialign_backward(grad_out.contiguous(), boxes, roi_map, ctx.shapes)
        return (None, None, None, None, None) + tuple(grads)


class PyramidROIAlign(_Layer):
    """mrcnn_layers.py:553-671.  inputs = [boxes [B,N,4], image_meta [B,meta], P2, P3, P4, P5] ->
    [B, N, pool_h, pool_w, C].  `denominator` keeps the reference's 244.0 default (L:574); `map_mode=0` keeps
    its first-appearance level->map assignment (L:613-619), `map_mode=1` selects the canonical level-2."""

    def __init__(self, pool_shape, denominator=244.0, name='roi_align', map_mode=0, **kwargs):
        super().__init__(name=name, **kwargs)
        self.pool_shape = tuple(pool_shape)
        self.denominator = denominator
        self.map_mode = map_mode

    def call(self, inputs, host_stage=None, new_maps=True, **kwargs):
        """host_stage (functional.HostMapStage): the feature maps live in pinned host memory; only the pixels these
        boxes sample are staged on the device before pooling (inference only; inputs[2:] are then ignored).
        new_maps=False: same host contents as the previous call with this stage (fetch only what is missing)."""
        boxes, image_meta = inputs[0], inputs[1]
        feature_maps = list(inputs[2:])
        if host_stage is not None:
            feature_maps = host_stage.fetch(boxes, image_meta, self.pool_shape, reset=new_maps,
                                            denominator=self.denominator, map_mode=self.map_mode)
        return _PyramidROIAlignFn.apply(boxes, image_meta, self.pool_shape, self.denominator, self.map_mode,
                                        *feature_maps)

    def compute_output_shape(self, input_shape):
        return input_shape[0][:2] + self.pool_shape + (input_shape[2][-1],)


class DetectionLayer(_Layer):
    """mrcnn_layers.py:343-531.  inputs = [rois [B,N,4], mrcnn_class [B,N,NC], mrcnn_bbox [B,N,NC,4],
    image_meta [B,meta]] -> [batch_size, detection_max_instances, 6] = (y1,x1,y2,x2,class_id,score)."""

    def __init__(self, proposals, detection_min_confidence, detection_max_instances, detection_nms_threshold,
                 bbox_std_dev, images_per_gpu, batch_size, name='mrcnn_detection', **kwargs):
        super().__init__(name=name, **kwargs)
        self.detection_min_confidence = detection_min_confidence
        self.detection_max_instances = detection_max_instances
        self.detection_nms_threshold = detection_nms_threshold
        self.bbox_std_dev = bbox_std_dev
        self.batch_size = batch_size
        self.proposals = proposals
        self.images_per_gpu = images_per_gpu

    def call(self, inputs, **kwargs):
        rois, mrcnn_class, mrcnn_bbox, image_meta = inputs[0], inputs[1], inputs[2], inputs[3]
        if rois.shape[1] != self.proposals:
            # the reference builds tf.range(self.proposals) indices (L:388-391) and fails on a mismatch
            raise ValueError(f"DetectionLayer built for {self.proposals} proposals, got {rois.shape[1]}")
        det, boxes = F.detection_forward(rois, mrcnn_class, mrcnn_bbox, image_meta,
                                         np.asarray(self.bbox_std_dev, dtype=np.float32), self.detection_min_confidence,
                                         self.detection_max_instances, self.detection_nms_threshold, return_boxes=True)
        det = det.reshape(self.batch_size, self.detection_max_instances, 6)  # L:524
        # detections[..., :4] came out of the same kernel: DetectedBoxesExtraction hands it on without a slice copy
        det._mrcnn_detected_boxes = boxes.reshape(self.batch_size, self.detection_max_instances, 4)
        return det

    def compute_output_shape(self, input_shape):
        return None, self.detection_max_instances, 6


class DetectedBoxesExtraction(_Layer):
    """mrcnn_layers.py:535-550: detections[..., :4], the boxes the mask branch's PyramidROIAlign crops
    (model.py:566-573).  When the input is the tensor DetectionLayer just produced, the contiguous [B,D,4] copy its
    kernel wrote alongside is returned; any other tensor is sliced."""

    def __init__(self, config=None, name='detected_boxes_extraction', **kwargs):
        super().__init__(name=name, **kwargs)
        self.config = config

    def call(self, inputs, **kwargs):
        boxes = getattr(inputs, "_mrcnn_detected_boxes", None)
        return boxes if boxes is not None else inputs[..., :4].contiguous()

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (4,)


class DetectionTargetLayer(_Layer):
    """mrcnn_layers.py:283-340.  inputs = [proposals [B,P,4], gt_class_ids [B,G] int32, gt_boxes [B,G,4],
    gt_masks [B,H,W,G] bool/uint8] -> [rois [B,T,4], target_class_ids [B,T], target_deltas [B,T,4],
    target_mask [B,T,mh,mw]].

    The reference subsamples with an unseeded tf.random.shuffle (L:905,910).  Here the permutation comes from one
    uint32 key per proposal row: drawn from `generator` (a torch.Generator on the device; default: torch's global
    CUDA generator), or injected through call(..., rand_keys=...) for reproducible tests."""

    def __init__(self, config, name='proposal_targets', generator=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.config = config
        self.generator = generator

    def call(self, inputs, rand_keys=None, **kwargs):
        proposals, gt_class_ids, gt_boxes, gt_masks = inputs[0], inputs[1], inputs[2], inputs[3]
        if proposals.shape[1] <= 0:
            raise ValueError("roi_assertion: DetectionTargetLayer needs at least one proposal (L:866-868)")
        if gt_masks.dtype == torch.bool:
            gt_masks = gt_masks.view(torch.uint8)
        if gt_class_ids.dtype != torch.int32:
            gt_class_ids = gt_class_ids.to(torch.int32)  # Keras casts to the Input dtype (model.py:424)
        if rand_keys is None:
            B, P = proposals.shape[0], proposals.shape[1]
            rand_keys = torch.randint(-2 ** 31, 2 ** 31, (B, P), dtype=torch.int64, device=proposals.device,
                                      generator=self.generator).to(torch.int32)
        cfg = self.config
        rois, cls, deltas, masks = F.detection_target_forward(
            proposals, gt_class_ids, gt_boxes, gt_masks, rand_keys, cfg['train_rois_per_image'],
            cfg['roi_positive_ratio'], np.asarray(cfg['bbox_std_dev'], dtype=np.float32), cfg['mask_shape'],
            bool(cfg['use_mini_masks']))
        return [rois, cls, deltas, masks]

    def compute_output_shape(self, input_shape):
        T = self.config['train_rois_per_image']
        return [(None, T, 4), (None, T), (None, T, 4),
                (None, T, self.config['mask_shape'][0], self.config['mask_shape'][1])]

    def compute_mask(self, inputs, mask=None):
        return [None, None, None, None]
