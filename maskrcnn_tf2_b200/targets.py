"""Device twin of the data loader's RPN target builder, under the reference's own name and signature.

`utils.build_rpn_targets(anchors, gt_class_ids, gt_boxes, rpn_train_anchors_per_image, rpn_bbox_std, eps=1e-3)`
(utils.py:154-262) is what `SegmentationDataGenerator` calls once per image (preprocess.py:342-348) before stacking
`batch_rpn_match [B,A,1]` / `batch_rpn_bbox [B,R,4]` (preprocess.py:369-372,411-412).  Here the same call takes CUDA
tensors -- one image ([G] class ids, [G,4] boxes) exactly like the reference, or a whole padded batch ([B,G], [B,G,4];
class id 0 marks padding rows) -- and runs on the device (csrc/rpn_targets.cu): there is no CPU path.
"""
import torch

from . import functional as F


def build_rpn_targets(anchors, gt_class_ids, gt_boxes, rpn_train_anchors_per_image, rpn_bbox_std, eps=1e-3,
                      rand_keys=None, generator=None):
    """Returns (rpn_match, rpn_bbox) as the reference does: rpn_match [A] int32 (1 positive / -1 negative / 0 neutral),
    rpn_bbox [R,4] float64 -- with a leading batch dimension when the inputs have one.

    np.random.choice (utils.py:219,227) becomes one fp32 key in [0,1) per anchor, drawn from `generator` (a
    torch.Generator on the device; default: torch's global CUDA generator) or injected as `rand_keys` [A] / [B,A]:
    an oversubscribed class keeps its anchors with the largest keys."""
    batched = gt_class_ids.dim() == 2
    cls = gt_class_ids if batched else gt_class_ids[None]
    box = gt_boxes if batched else gt_boxes[None]
    if cls.dtype != torch.int32:
        cls = cls.to(torch.int32)
    if box.dtype != torch.int32:
        box = box.to(torch.int32)                       # utils.extract_bboxes returns int32 pixel boxes
    if anchors.dtype != torch.float64:
        anchors = anchors.to(torch.float64)             # generate_pyramid_anchors returns float64
    B, A = cls.shape[0], anchors.shape[0]
    if rand_keys is None:
        rand_keys = torch.rand((B, A), dtype=torch.float32, device=anchors.device, generator=generator)
    elif rand_keys.dim() == 1:
        rand_keys = rand_keys[None]
    match, bbox = F.rpn_targets_forward(anchors, cls, box, rand_keys, rpn_train_anchors_per_image, rpn_bbox_std, eps)
    return (match, bbox) if batched else (match[0], bbox[0])
