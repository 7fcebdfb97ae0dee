"""The launchers never synchronise the host, never allocate and keep no state, so a sequence of layer calls
can be captured into a CUDA graph and replayed on new input contents: the replay must reproduce the eager calls bit for
bit, including the programmatic-dependent-launch edges between the kernels (functional.py / graphs.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_captured_inference_stage_replays_bit_exact_on_new_inputs(orc, dev):
    from maskrcnn_tf2_b200 import make_config, synth
    from maskrcnn_tf2_b200.graphs import CapturedStage
    from maskrcnn_tf2_b200.layers import DetectedBoxesExtraction, DetectionLayer, ProposalLayer, PyramidROIAlign
    B, S, NC = 2, 512, 81
    cfg = make_config(img_size=S, num_classes=NC, batch_size=B)
    batches = [synth.inference_batch(6, B, img_size=S, num_classes=NC, regime="clustered", first_image=f) for f in (0, 2)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: t(v) for k, v in batches[0].items() if k != "feature_maps"}
    maps = [t(f) for f in batches[0]["feature_maps"]]
    proposal = ProposalLayer(1000, cfg)
    align7, align14 = PyramidROIAlign([7, 7], name="roi_align_classifier"), PyramidROIAlign([14, 14], name="roi_align_mask")
    detect = DetectionLayer(1000, 0.7, 100, 0.3, cfg["bbox_std_dev"], B, B)
    boxes_of = DetectedBoxesExtraction(cfg)

    def stage():
        rois = proposal([d["rpn_probs"], d["rpn_bbox"], d["anchors"]])
        pooled = align7([rois, d["image_meta"]] + maps)
        det = detect([rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"]])
        return rois, pooled, det, align14([boxes_of(det), d["image_meta"]] + maps)

    captured = CapturedStage(stage, device=dev)
    for x in (batches[1], batches[0], batches[1]):
        for k, v in x.items():                      # new contents in the static input buffers
            if k == "feature_maps":
                for m, f in zip(maps, v):
                    m.copy_(t(f))
            else:
                d[k].copy_(t(v))
        got = [o.clone() for o in captured.replay()]
        want = stage()                              # eager calls on the same buffers
        for g, w in zip(got, want):
            assert torch.equal(g, w)
        ref = orc.proposal_layer(x["rpn_probs"], x["rpn_bbox"], x["anchors"], 6000, 1000, cfg["rpn_bbox_std_dev"], 0.7)
        assert np.array_equal(got[0].cpu().numpy(), ref["proposals"])


def test_captured_training_backward_replays_bit_exact(dev):
    from maskrcnn_tf2_b200 import functional as F
    from maskrcnn_tf2_b200 import synth
    from maskrcnn_tf2_b200.graphs import CapturedStage
    rng = np.random.default_rng(5)
    B, Nr = 2, 64
    boxes = torch.from_numpy(rng.uniform(0, 1, (B, Nr, 4)).astype(np.float32)).to(dev)
    boxes[..., 2:] = boxes[..., :2] + 0.05 + 0.4 * boxes[..., 2:]
    boxes.clamp_(0, 1)
    meta = torch.from_numpy(synth.image_meta(B, 1024, 81)).to(dev)
    maps = [torch.randn((B, s, s, 256), device=dev) for s in (64, 32, 16, 8)]
    shapes = [tuple(m.shape) for m in maps]
    grad = torch.randn((B, Nr, 7, 7, 256), device=dev)

    def step():
        out, rmap = F.roialign_forward(boxes, meta, maps, (7, 7))
        return [out] + F.roialign_backward(grad, boxes, rmap, shapes, deterministic=True)

    captured = CapturedStage(step, device=dev)
    for _ in range(2):
        grad.normal_()
        got = [o.clone() for o in captured.replay()]
        for g, w in zip(got, step()):
            assert torch.equal(g, w)
