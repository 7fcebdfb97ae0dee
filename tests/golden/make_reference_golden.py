"""Generates tests/golden/reference_numpy_golden.npz by RUNNING THE REFERENCE'S OWN CODE in this container
(`python tests/golden/make_reference_golden.py`; needs /root/reference, so it cannot run on the GPU box -- the
vectors it writes are committed).

The reference's hot path proper is Python over TensorFlow kernels and TensorFlow is not installable here, but
`src/common/utils.py` also holds the plain-numpy producers and twins of that path, and those do run once the
module's unused heavy imports (`tensorflow`, `skimage`) are stubbed:

  * `compute_backbone_shapes` + `generate_pyramid_anchors` (utils.py:54-111,725-735) -- the pixel anchors that
    AnchorsLayer (mrcnn_layers.py:116-132) feeds ProposalLayer with;
  * `norm_boxes` (utils.py:691-705) -- numpy twin of NormBoxesLayer (mrcnn_layers.py:34-39);
  * `compose_image_meta` (utils.py:494-516) -- the image_meta layout PyramidROIAlign / DetectionLayer parse;
  * `compute_overlaps` (utils.py:114-151) -- numpy twin of overlaps_graph (mrcnn_layers.py:982-1007);
  * `box_refinement` (utils.py:468-491) -- numpy twin of box_refinement_graph (utils.py:775-798).

Nothing is copied from the reference: the module is loaded from where it lies and only its outputs are stored.
"""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_UTILS = "/root/reference/src/common/utils.py"


def load_reference_utils():
    for name in ("tensorflow", "skimage", "skimage.transform"):      # imported at module top, unused by the numpy code
        sys.modules.setdefault(name, types.ModuleType(name))
    try:
        import distutils.version  # noqa: F401
    except Exception:                                                # python >= 3.12 without setuptools' shim
        d, dv = types.ModuleType("distutils"), types.ModuleType("distutils.version")
        dv.LooseVersion = str
        sys.modules["distutils"], sys.modules["distutils.version"] = d, dv
    spec = importlib.util.spec_from_file_location("reference_utils", REF_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_config(img_size):
    # src/common/config.py:70-133 values that the functions below read
    return {"img_size": img_size, "backbone_strides": [4, 8, 16, 32, 64], "rpn_anchor_scales": (32, 64, 128, 256, 512),
            "rpn_anchor_ratios": [0.5, 1, 2], "rpn_anchor_stride": 1}


def build():
    U = load_reference_utils()
    g = {}
    for S in (128, 256, 1024):
        cfg = reference_config(S)
        shapes = U.compute_backbone_shapes(cfg)
        px = U.generate_pyramid_anchors(cfg["rpn_anchor_scales"], cfg["rpn_anchor_ratios"], shapes,
                                        cfg["backbone_strides"], cfg["rpn_anchor_stride"])
        g[f"backbone_shapes_{S}"] = np.asarray(shapes, np.int64)
        if S == 128:
            g["anchors_px_128"] = px                                  # float64 [4092,4], as the reference returns it
            g["anchors_norm_numpy_128"] = U.norm_boxes(px, (S, S))
        # all sizes: digest of the float64 bytes (the 1024 set is 8 MB, too large to commit)
        g[f"anchors_px_sha256_{S}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(px).tobytes()).digest(), np.uint8)
        g[f"anchors_count_{S}"] = np.int64(px.shape[0])
    # image_meta
    g["image_meta"] = U.compose_image_meta(7, (480, 640, 3), (128, 0, 896, 1024), 1.6, np.ones(81, np.int32),
                                           {"img_size": 1024}).astype(np.float64)
    # overlaps + refinement on jittered ground truth (pixel boxes)
    rng = np.random.default_rng(20261019)
    G = 12
    cy, cx = np.meshgrid(np.arange(4) * 250 + 130, np.arange(3) * 330 + 170, indexing="ij")   # well separated
    hw = rng.uniform(60, 200, (G, 2))
    gt = np.stack([cy.ravel() - hw[:, 0] / 2, cx.ravel() - hw[:, 1] / 2, cy.ravel() + hw[:, 0] / 2,
                   cx.ravel() + hw[:, 1] / 2], 1)
    gt = np.round(gt).astype(np.float32)
    reps = 10
    props = np.repeat(gt, reps, 0) + rng.normal(0, 6.0, (G * reps, 4)).astype(np.float32)
    far = rng.uniform(0, 1024, (40, 2))
    far = np.concatenate([far, far + rng.uniform(20, 120, (40, 2))], 1).astype(np.float32)
    props = np.concatenate([props, far]).astype(np.float32)
    gtn, prn = U.norm_boxes(gt, (1024, 1024)), U.norm_boxes(props, (1024, 1024))
    g["gt_boxes_px"], g["proposals_px"] = gt, props
    g["gt_boxes_norm"], g["proposals_norm"] = gtn, prn
    ov = U.compute_overlaps(prn.astype(np.float64), gtn.astype(np.float64))
    g["overlaps"] = ov                                                # float64 [P,G]
    arg = ov.argmax(1)
    g["refinement"] = U.box_refinement(prn, gtn[arg])                 # fp32 [P,4] against each row's best GT
    return g


if __name__ == "__main__":
    out = os.path.join(HERE, "reference_numpy_golden.npz")
    np.savez_compressed(out, **build())
    print("wrote", out, os.path.getsize(out), "bytes")
