"""Drop-in Keras layers for src/layers/mrcnn_layers.py of miguelalejo/maskrcnn_tf2, backed by libmrcnn_roi_ops.so.

NOT IMPORTABLE IN THIS REPOSITORY'S CI (TensorFlow is not installable in the build image).  Where a CUDA-12
TensorFlow exists: build tf_shim/mrcnn_roi_ops.cc (see its header), put this file on the reference's import path
and, in src/layers/mrcnn_layers.py, replace the four class bodies by

    from mrcnn_layers_b200 import ProposalLayer, PyramidROIAlign, DetectionLayer, DetectionTargetLayer

Class names, constructor signatures, layer names, `call` argument order and output shapes are the reference's
(mrcnn_layers.py:202-280, 283-340, 343-531, 553-671), so mask_rcnn_functional() (model.py:398), fpn_classifier_graph
/ fpn_mask_graph (mrcnn_layers.py:1145,1215) and the ONNX/TRT surgery that keys on layer names
(inference_optimize.py:455-465) keep working unchanged.  The same classes, over torch tensors, are what this
repository tests: maskrcnn_tf2_b200/layers.py.
"""
import os

import numpy as np
import tensorflow as tf
from tensorflow.keras import layers as tfl

_ops = tf.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmrcnn_roi_ops.so"))


@tf.RegisterGradient("MrcnnPyramidRoiAlign")
def _pyramid_roi_align_grad(op, grad_pooled, _grad_roi_map):
    # boxes and image_meta receive no gradient (tf.stop_gradient at mrcnn_layers.py:628-629)
    d2, d3, d4, d5 = _ops.mrcnn_pyramid_roi_align_grad(grad_pooled, op.inputs[0], op.outputs[1], *op.inputs[2:6])
    return [None, None, d2, d3, d4, d5]


@tf.RegisterGradient("MrcnnProposal")
def _proposal_grad(op, grad_proposals, _grad_topk, _grad_keep):
    # the reference back-propagates mrcnn_bbox_loss through the proposals into rpn_bbox (no stop_gradient at
    # mrcnn_layers.py:227); rpn_probs (top_k / NMS indices) and the anchors receive none
    g = _ops.mrcnn_proposal_grad(grad_proposals, op.inputs[1], op.inputs[2], op.outputs[1], op.outputs[2],
                                 std_dev=op.get_attr("std_dev"))
    return [None, g, None]


tf.no_gradient("MrcnnProposalGrad")
tf.no_gradient("MrcnnDetection")
tf.no_gradient("MrcnnDetectionTarget")


@tf.keras.utils.register_keras_serializable()
class ProposalLayer(tfl.Layer):
    def __init__(self, proposal_count, config, name='roi', **kwargs):
        super(ProposalLayer, self).__init__(name=name, **kwargs)
        self.config = config
        self.proposal_count = proposal_count
        self.nms_threshold = self.config['rpn_nms_threshold']

    def call(self, inputs, **kwargs):
        proposals, _, _ = _ops.mrcnn_proposal(inputs[0], inputs[1], inputs[2], proposal_count=self.proposal_count,
                                              pre_nms_limit=self.config['pre_nms_limit'],
                                              nms_threshold=self.nms_threshold,
                                              std_dev=[float(v) for v in self.config['rpn_bbox_std_dev']])
        return proposals

    def call_levels(self, rpn_class_logits, rpn_bbox, anchors):
        """Inference path fed by the per-level RPN outputs (lists, one [B,A_l,2] / [B,A_l,4] tensor per pyramid
        level, before model.py:465-478 concatenates them): returns (proposals, rpn_class [B,A,2])."""
        return _ops.mrcnn_proposal_levels(list(rpn_class_logits), list(rpn_bbox), anchors,
                                          proposal_count=self.proposal_count,
                                          pre_nms_limit=self.config['pre_nms_limit'], nms_threshold=self.nms_threshold,
                                          std_dev=[float(v) for v in self.config['rpn_bbox_std_dev']])

    def build(self, input_shape):
        self.built = True
        super(ProposalLayer, self).build(input_shape)

    def compute_output_shape(self, input_shape):
        return None, self.proposal_count, 4

    def get_config(self):
        return super(ProposalLayer, self).get_config()


@tf.keras.utils.register_keras_serializable()
class PyramidROIAlign(tfl.Layer):
    def __init__(self, pool_shape, denominator=244.0, name='roi_align', **kwargs):
        super(PyramidROIAlign, self).__init__(name=name, **kwargs)
        self.pool_shape = tuple(pool_shape)
        self.denominator = denominator

    def build(self, input_shape):
        self.built = True
        super(PyramidROIAlign, self).build(input_shape)

    def call(self, inputs, **kwargs):
        pooled, _ = _ops.mrcnn_pyramid_roi_align(inputs[0], inputs[1], *inputs[2:6], pool_height=self.pool_shape[0],
                                                 pool_width=self.pool_shape[1], denominator=self.denominator,
                                                 map_mode=0)
        return pooled

    def compute_output_shape(self, input_shape):
        return input_shape[0][:2] + self.pool_shape + (input_shape[2][-1],)

    def get_config(self):
        return super(PyramidROIAlign, self).get_config()


@tf.keras.utils.register_keras_serializable()
class DetectionLayer(tfl.Layer):
    def __init__(self, proposals, detection_min_confidence, detection_max_instances, detection_nms_threshold,
                 bbox_std_dev, images_per_gpu, batch_size, name='mrcnn_detection', **kwargs):
        super(DetectionLayer, self).__init__(name=name, **kwargs)
        self.detection_min_confidence = detection_min_confidence
        self.detection_max_instances = detection_max_instances
        self.detection_nms_threshold = detection_nms_threshold
        self.bbox_std_dev = bbox_std_dev
        self.batch_size = batch_size
        self.proposals = proposals
        self.images_per_gpu = images_per_gpu

    def build(self, input_shape):
        self.built = True
        super(DetectionLayer, self).build(input_shape)

    def call(self, inputs, **kwargs):
        det, boxes = _ops.mrcnn_detection(inputs[0], inputs[1], inputs[2], inputs[3],
                                          min_confidence=float(self.detection_min_confidence or 0.0),
                                          use_min_confidence=bool(self.detection_min_confidence),
                                          max_instances=self.detection_max_instances,
                                          nms_threshold=self.detection_nms_threshold,
                                          std_dev=[float(v) for v in np.asarray(self.bbox_std_dev)])
        det = tf.reshape(det, [self.batch_size, self.detection_max_instances, 6])
        # detections[..., :4] written by the same kernel: DetectedBoxesExtraction below hands it on without a slice
        det._mrcnn_detected_boxes = tf.reshape(boxes, [self.batch_size, self.detection_max_instances, 4])
        return det

    def compute_output_shape(self, input_shape):
        return None, self.detection_max_instances, 6

    def get_config(self):
        return super(DetectionLayer, self).get_config()


class DetectedBoxesExtraction(tfl.Layer):
    """mrcnn_layers.py:535-550 (detections[..., :4] for the mask branch, model.py:566-573)."""

    def __init__(self, config=None, name='detected_boxes_extraction', **kwargs):
        super(DetectedBoxesExtraction, self).__init__(name=name, **kwargs)
        self.config = config

    def build(self, input_shape):
        self.built = True
        super(DetectedBoxesExtraction, self).build(input_shape)

    def call(self, inputs, **kwargs):
        boxes = getattr(inputs, "_mrcnn_detected_boxes", None)
        return boxes if boxes is not None else inputs[..., :4]

    def get_config(self):
        return super(DetectedBoxesExtraction, self).get_config()


@tf.keras.utils.register_keras_serializable()
class DetectionTargetLayer(tfl.Layer):
    def __init__(self, config, name='proposal_targets', **kwargs):
        super(DetectionTargetLayer, self).__init__(name=name, **kwargs)
        self.config = config

    def call(self, inputs, **kwargs):
        proposals, gt_class_ids, gt_boxes, gt_masks = inputs[0], inputs[1], inputs[2], inputs[3]
        shape = tf.shape(proposals)[:2]
        keys = tf.random.uniform(shape, minval=tf.int32.min, maxval=tf.int32.max, dtype=tf.int32)  # tf.random.shuffle
        cfg = self.config
        return list(_ops.mrcnn_detection_target(
            proposals, tf.cast(gt_class_ids, tf.int32), gt_boxes, tf.cast(gt_masks, tf.bool), keys,
            train_rois_per_image=cfg['train_rois_per_image'], roi_positive_ratio=cfg['roi_positive_ratio'],
            mask_height=cfg['mask_shape'][0], mask_width=cfg['mask_shape'][1],
            use_mini_masks=bool(cfg['use_mini_masks']), std_dev=[float(v) for v in cfg['bbox_std_dev']]))

    def compute_output_shape(self, input_shape):
        T = self.config['train_rois_per_image']
        return [(None, T, 4), (None, T), (None, T, 4),
                (None, T, self.config['mask_shape'][0], self.config['mask_shape'][1])]

    def compute_mask(self, inputs, mask=None):
        return [None, None, None, None]

    def get_config(self):
        return super(DetectionTargetLayer, self).get_config()


def build_rpn_targets(anchors, gt_class_ids, gt_boxes, rpn_train_anchors_per_image, rpn_bbox_std, eps=1e-3):
    """Device twin of utils.build_rpn_targets (utils.py:154-262) for a padded batch: anchors [A,4] float64 pixel
    boxes, gt_class_ids [B,G] (0 = padding row, negative = crowd), gt_boxes [B,G,4] pixel boxes ->
    (rpn_match [B,A,1] int32, rpn_bbox [B,R,4] float32), the shapes and dtypes model.py:419-420 declares for
    input_rpn_match / input_rpn_bbox.  np.random.choice (utils.py:219,227) becomes one uniform key per anchor."""
    a = tf.shape(anchors)[0]
    keys = tf.random.uniform(tf.stack([tf.shape(gt_class_ids)[0], a]), dtype=tf.float32)
    match, _, bbox32 = _ops.mrcnn_rpn_targets(
        tf.cast(anchors, tf.float64), tf.cast(gt_class_ids, tf.int32), tf.cast(gt_boxes, tf.int32), keys,
        rpn_train_anchors_per_image=int(rpn_train_anchors_per_image),
        rpn_bbox_std_dev=[float(v) for v in rpn_bbox_std], eps=float(eps))
    return match, bbox32
