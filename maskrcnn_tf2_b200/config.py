"""The configuration keys the ROI-stage layers read, with the reference's defaults (src/common/config.py:9-179).

Only the hot path's keys are mirrored; the rest of the reference's CONFIG dict (backbone, optimiser, callbacks ...)
belongs to code that stays in TensorFlow.  A reference CONFIG dict can be passed to the layers unchanged: extra
keys are ignored.
"""
import numpy as np

CONFIG = {
    'image_shape': (1024, 1024, 3),          # C:20 (512 in the reference's balloon default)
    'img_size': 1024,                        # C:21
    'num_classes': 81,                       # C:23-24 (COCO_CONFIG, samples/coco/coco.py:30-115)
    'meta_shape': 1 + 3 + 3 + 4 + 1 + 81,    # C:22
    'use_mini_masks': False,                 # C:38
    'mini_mask_shape': (32, 32),             # C:39
    'mask_shape': (28, 28),                  # C:43
    'batch_size': 8,                         # C:47
    'images_per_gpu': 8,                     # C:48
    'backbone_strides': [4, 8, 16, 32, 64],  # C:70
    'top_down_pyramid_size': 256,            # C:72
    'rpn_anchor_scales': (32, 64, 128, 256, 512),  # C:75
    'rpn_anchor_ratios': [0.5, 1, 2],        # C:79
    'rpn_anchor_stride': 1,                  # C:84
    'max_gt_instances': 100,                 # C:87
    'rpn_bbox_std_dev': np.array([0.1, 0.1, 0.2, 0.2], dtype='float32'),  # C:90
    'bbox_std_dev': np.array([0.1, 0.1, 0.2, 0.2], dtype='float32'),      # C:91
    'rpn_nms_threshold': 0.7,                # C:95
    'detection_min_confidence': 0.7,         # C:108
    'detection_nms_threshold': 0.3,          # C:110
    'detection_max_instances': 100,          # C:112
    'pre_nms_limit': 6000,                   # C:115
    'post_nms_rois_training': 2000,          # C:118
    'post_nms_rois_inference': 1000,         # C:119
    'train_rois_per_image': 200,             # C:126
    'roi_positive_ratio': 0.33,              # C:129
    'pool_size': 7,                          # C:132
    'mask_pool_size': 14,                    # C:133
}


def make_config(**overrides):
    cfg = dict(CONFIG)
    cfg.update(overrides)
    if 'num_classes' in overrides and 'meta_shape' not in overrides:
        cfg['meta_shape'] = 1 + 3 + 3 + 4 + 1 + cfg['num_classes']
    if 'img_size' in overrides and 'image_shape' not in overrides:
        cfg['image_shape'] = (cfg['img_size'], cfg['img_size'], 3)
    if 'batch_size' in overrides and 'images_per_gpu' not in overrides:
        cfg['images_per_gpu'] = cfg['batch_size']
    return cfg
