"""Pins against outputs of the REFERENCE'S OWN numpy code (tests/golden/reference_numpy_golden.npz, produced by
tests/golden/make_reference_golden.py importing /root/reference/src/common/utils.py in the build container).
These are the only parts of the path's neighbourhood the reference can execute without TensorFlow: the anchor
producer, the box normalisation, the image_meta layout, and the numpy twins of overlaps_graph /
box_refinement_graph.  The TF kernels themselves (top_k, NMS, crop_and_resize) stay pinned only by the known-answer
and differential tests."""
import hashlib
import os

import numpy as np
import pytest

from maskrcnn_tf2_b200 import synth

SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


@pytest.fixture(scope="module")
def R():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_numpy_golden.npz"))


def _norm_fp32(px, S):
    # NormBoxesLayer.call (mrcnn_layers.py:34-39) on the Keras-autocast fp32 anchors, in fp32
    a = px.astype(np.float32)
    return ((a - np.array([0, 0, 1, 1], np.float32)) / (np.float32(S) - np.float32(1.0))).astype(np.float32)


def _pixel_anchors(S):
    """Undo nothing: rebuild the float64 pixel anchors the way synth does, for the digest comparison."""
    out = []
    ratios = np.asarray((0.5, 1, 2), np.float64)
    for scale, (fh, fw), stride in zip((32, 64, 128, 256, 512), synth.backbone_shapes(S, (4, 8, 16, 32, 64)),
                                       (4, 8, 16, 32, 64)):
        hs, ws = scale / np.sqrt(ratios), scale * np.sqrt(ratios)
        cy, cx = np.arange(fh, dtype=np.float64) * stride, np.arange(fw, dtype=np.float64) * stride
        shp = (fh, fw, 3)
        CY, CX = np.broadcast_to(cy[:, None, None], shp), np.broadcast_to(cx[None, :, None], shp)
        HH, WW = np.broadcast_to(hs[None, None, :], shp), np.broadcast_to(ws[None, None, :], shp)
        out.append(np.stack([CY - 0.5 * HH, CX - 0.5 * WW, CY + 0.5 * HH, CX + 0.5 * WW], -1).reshape(-1, 4))
    return np.concatenate(out, 0)


def test_anchor_producer_bit_exact_against_the_reference(R):
    # the synthetic-input generator must emit exactly the anchors AnchorsLayer would hand to ProposalLayer
    assert np.array_equal(synth.pyramid_anchors(128), _norm_fp32(R["anchors_px_128"], 128))
    assert np.array_equal(_pixel_anchors(128), R["anchors_px_128"])
    for S in (128, 256, 1024):
        assert synth.backbone_shapes(S, (4, 8, 16, 32, 64)) == [tuple(r) for r in R[f"backbone_shapes_{S}"].tolist()]
        px = _pixel_anchors(S)
        assert px.shape[0] == int(R[f"anchors_count_{S}"])
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(px).tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, R[f"anchors_px_sha256_{S}"])
        assert np.array_equal(synth.pyramid_anchors(S), _norm_fp32(px, S))
    # the numpy norm_boxes divides in float64 and rounds once: within one fp32 ulp of the layer's fp32 arithmetic
    assert np.abs(synth.pyramid_anchors(128) - R["anchors_norm_numpy_128"]).max() <= 2.0 ** -23
    assert np.abs(synth.norm_boxes(R["gt_boxes_px"], 1024) - R["gt_boxes_norm"]).max() <= 2.0 ** -23


def test_image_meta_layout_matches_compose_image_meta(R):
    ref = R["image_meta"]
    assert ref.shape[0] == 12 + 81
    m = synth.image_meta(1, 1024, 81)[0]
    # same slots: [id | original h,w,c | h,w,c | window | scale | active classes]
    assert np.array_equal(m[4:7], ref[4:7]) and m.shape == ref.shape and np.array_equal(m[12:], ref[12:])
    assert list(ref[1:4]) == [480, 640, 3] and list(ref[7:11]) == [128, 0, 896, 1024] and ref[11] == 1.6


def test_oracle_detection_targets_against_reference_overlaps_and_refinement(orc, R):
    """DetectionTargetLayer in the oracle vs the reference's numpy compute_overlaps / box_refinement: positive set
    (max IoU >= 0.5), first-max GT assignment, and the regression targets (fp32 numpy vs the oracle's fp32 + fixed
    log: within the north-star tolerance)."""
    props, gt, ov = R["proposals_norm"], R["gt_boxes_norm"], R["overlaps"]
    P, G = ov.shape
    assert np.abs(ov.max(1) - 0.5).min() > 1e-3                    # the fixture keeps IoUs away from the threshold
    pos_ref = np.nonzero(ov.max(1) >= 0.5)[0]
    arg_ref = ov.argmax(1)
    cls = (np.arange(G) + 1).astype(np.int32)
    T, ratio = 400, 0.5                                            # int(T*ratio)=200 >= #positives: nothing dropped
    keys = np.arange(P, dtype=np.uint32)                           # injected shuffle = identity order
    masks = np.zeros((1, 8, 8, G), np.uint8)
    t = orc.detection_target_layer(props[None], cls[None], gt[None], masks, keys[None], T, ratio, SD, (28, 28))
    npos, nneg = t["counts"][0]
    assert npos == len(pos_ref) and nneg == min(P - npos, npos)    # int(fp32(1/0.5) * npos) - npos negatives
    assert np.array_equal(t["rois"][0, :npos], props[pos_ref])
    assert np.array_equal(t["class_ids"][0, :npos], cls[arg_ref[pos_ref]])
    want = R["refinement"][pos_ref] / SD
    assert np.allclose(t["deltas"][0, :npos], want, rtol=1e-5, atol=1e-6)
    neg_ref = np.nonzero(ov.max(1) < 0.5)[0]
    assert np.array_equal(t["rois"][0, npos:npos + nneg], props[neg_ref[:nneg]])


@pytest.mark.gpu
def test_cuda_detection_targets_against_reference_overlaps_and_refinement(R, dev):
    import torch
    from maskrcnn_tf2_b200 import functional as F
    props, gt, ov = R["proposals_norm"], R["gt_boxes_norm"], R["overlaps"]
    P, G = ov.shape
    pos_ref, arg_ref = np.nonzero(ov.max(1) >= 0.5)[0], ov.argmax(1)
    cls = (np.arange(G) + 1).astype(np.int32)
    T_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = F.detection_target_forward(T_(props[None]), T_(cls[None]), T_(gt[None]), T_(np.zeros((1, 8, 8, G), np.uint8)),
                                     T_(np.arange(P, dtype=np.int32)[None]), 400, 0.5, SD, (28, 28))
    rois, class_ids, deltas = [o.cpu().numpy() for o in out[:3]]
    n = len(pos_ref)
    assert np.array_equal(rois[0, :n], props[pos_ref]) and np.array_equal(class_ids[0, :n], cls[arg_ref[pos_ref]])
    assert np.allclose(deltas[0, :n], R["refinement"][pos_ref] / SD, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_device_anchor_producer_bit_exact_against_the_reference(R, dev):
    """csrc/anchors.cu: float64 pixel anchors identical to generate_pyramid_anchors' output (full vector at 128,
    SHA-256 at 128 / 256 / 1024), AnchorsLayer output identical to the fp32 NormBoxesLayer arithmetic."""
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import AnchorsLayer
    for S in (128, 256, 1024):
        cfg = make_config(img_size=S, batch_size=3)
        layer = AnchorsLayer(cfg, training=False, device=dev)
        px = layer.anchors_px.cpu().numpy()
        assert px.dtype == np.float64 and px.shape[0] == int(R[f"anchors_count_{S}"])
        digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(px).tobytes()).digest(), np.uint8)
        assert np.array_equal(digest, R[f"anchors_px_sha256_{S}"])
        if S == 128:
            assert np.array_equal(px, R["anchors_px_128"])
        norm = layer(None).cpu().numpy()
        assert norm.shape == (3, px.shape[0], 4)
        assert np.array_equal(norm[0], _norm_fp32(px, S)) and np.array_equal(norm[2], norm[0])
        assert np.array_equal(norm[1], synth.pyramid_anchors(S))
