"""The TensorFlow custom-op shim (maskrcnn_tf2_b200/tf_shim/mrcnn_roi_ops.cc) cannot be built against TensorFlow in
this image, so it is compiled against tests/tf_stub/ -- a functional stand-in for the TF op API -- instead:

* CPU (`-m "not gpu"`): the file compiles with -Wall -Wextra -Werror (every launcher call is type-checked against
  include/mrcnn_roi_b200.h), registers the eight ops with GPU kernels only, its shape functions return the reference
  layers' output shapes, the kernels reject bad attributes, and the Python side of the shim
  (tf_shim/mrcnn_layers_b200.py, parsed, not imported) passes exactly the inputs / attributes the ops declare.
* GPU (`-m gpu`): every OpKernel::Compute runs on the B200 through a fake OpKernelContext and returns, bit for bit,
  what the ctypes path (maskrcnn_tf2_b200.functional) returns for the same inputs.
"""
import ast
import os
import re

import numpy as np
import pytest
import torch

import tf_stub
from conftest import random_boxes

SD = [0.1, 0.1, 0.2, 0.2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY_SHIM = os.path.join(ROOT, "maskrcnn_tf2_b200", "tf_shim", "mrcnn_layers_b200.py")
OPS = ["MrcnnProposal", "MrcnnProposalGrad", "MrcnnPyramidRoiAlign", "MrcnnPyramidRoiAlignGrad", "MrcnnDetection",
       "MrcnnDetectionTarget", "MrcnnRpnTargets", "MrcnnProposalLevels"]


@pytest.fixture(scope="module")
def stub_ready():
    """The GPU cases need the stub-built shim; where it cannot be built they skip (the CPU case above all asserts that it
    builds, with -Werror, in the container that has the toolchain)."""
    try:
        tf_stub.lib()
    except Exception as e:      # noqa: BLE001
        pytest.skip(f"stub-built shim unavailable: {e}")
    return True


# --------------------------------------------------------------------------------------------------- CPU
def test_shim_compiles_against_the_stub_and_registers_gpu_only_kernels():
    tf_stub.build(force=True)        # g++ -Wall -Wextra -Werror
    sig = tf_stub.signatures()
    assert sorted(sig) == sorted(OPS)
    for op, s in sig.items():
        assert s["devices"] == ["GPU"], f"{op}: no CPU kernel may be registered (north star: no CPU fallback)"
    assert [o for o, _ in sig["MrcnnProposal"]["outputs"]] == ["proposals", "topk_idx", "keep_idx"]
    assert sig["MrcnnProposal"]["attrs"]["pre_nms_limit"] == ("int", "6000")          # config.py defaults
    assert sig["MrcnnProposal"]["attrs"]["nms_threshold"] == ("float", "0.7")
    assert sig["MrcnnPyramidRoiAlign"]["attrs"]["denominator"] == ("float", "244.0")  # quirk Q1
    assert sig["MrcnnDetection"]["attrs"]["nms_threshold"] == ("float", "0.3")
    assert sig["MrcnnDetectionTarget"]["attrs"]["roi_positive_ratio"] == ("float", "0.33")
    assert sig["MrcnnProposalLevels"]["inputs"][0] == ("rpn_class_logits", "float", "N")


def test_shape_functions_return_the_reference_layers_output_shapes():
    B, A = 8, 261888
    assert tf_stub.infer_shapes("MrcnnProposal", [(B, A, 2), (B, A, 4), (B, A, 4)], proposal_count=1000) == \
        [(B, 1000, 4), (B, -1), (B, 1000)]                       # compute_output_shape, mrcnn_layers.py:275-276
    maps = [(B, 256 >> l, 256 >> l, 256) for l in range(4)]
    assert tf_stub.infer_shapes("MrcnnPyramidRoiAlign", [(B, 1000, 4), (B, 93)] + maps, pool_height=7,
                                pool_width=7) == [(B, 1000, 7, 7, 256), (B, 1000)]          # L:666-667
    assert tf_stub.infer_shapes("MrcnnPyramidRoiAlignGrad", [(B, 200, 14, 14, 256), (B, 200, 4), (B, 200)] + maps) == maps
    assert tf_stub.infer_shapes("MrcnnDetection", [(B, 1000, 4), (B, 1000, 81), (B, 1000, 81, 4), (B, 93)]) == \
        [(B, 100, 6), (B, 100, 4)]                               # L:526-527
    assert tf_stub.infer_shapes("MrcnnDetectionTarget",
                                [(B, 2000, 4), (B, 100), (B, 100, 4), (B, 1024, 1024, 100), (B, 2000)]) == \
        [(B, 200, 4), (B, 200), (B, 200, 4), (B, 200, 28, 28)]   # L:327-333
    assert tf_stub.infer_shapes("MrcnnRpnTargets", [(A, 4), (B, 100), (B, 100, 4), (B, A)]) == \
        [(B, A, 1), (B, 256, 4), (B, 256, 4)]
    lv = [(B, n, 2) for n in (196608, 49152, 12288, 3072, 768)]
    lb = [(B, n, 4) for n in (196608, 49152, 12288, 3072, 768)]
    assert tf_stub.infer_shapes("MrcnnProposalLevels", lv + lb + [(B, A, 4)], N=5, proposal_count=1000) == \
        [(B, 1000, 4), (B, A, 2)]
    assert tf_stub.infer_shapes("MrcnnProposalGrad", [(B, 1000, 4), (B, A, 4), (B, A, 4), (B, 6000), (B, 1000)]) == \
        [(B, A, 4)]
    with pytest.raises(ValueError):
        tf_stub.infer_shapes("MrcnnProposal", [(B, A, 2), (B, A, 4), (B, A, 4)])      # required attr missing


def test_kernel_construction_validates_attributes(stub_ready):
    for op, bad in [("MrcnnProposal", dict(proposal_count=10, std_dev=[0.1, 0.2])),
                    ("MrcnnProposalGrad", dict(std_dev=[0.1])),
                    ("MrcnnDetection", dict(std_dev=[0.1, 0.1, 0.2])),
                    ("MrcnnDetectionTarget", dict(std_dev=[1.0] * 5)),
                    ("MrcnnRpnTargets", dict(rpn_bbox_std_dev=[0.1, 0.1])),
                    ("MrcnnProposalLevels", dict(N=9, proposal_count=10)),
                    ("MrcnnProposal", dict()),                                       # proposal_count is required
                    ("MrcnnPyramidRoiAlign", dict(pool_height=7))]:                  # pool_width is required
        with pytest.raises(ValueError):
            tf_stub.StubOp(op, **bad)
    for op, good in [("MrcnnProposal", dict(proposal_count=1000)), ("MrcnnPyramidRoiAlign", dict(pool_height=7, pool_width=7)),
                     ("MrcnnDetection", dict()), ("MrcnnDetectionTarget", dict()), ("MrcnnRpnTargets", dict()),
                     ("MrcnnPyramidRoiAlignGrad", dict()), ("MrcnnProposalGrad", dict()),
                     ("MrcnnProposalLevels", dict(N=5, proposal_count=1000))]:
        assert tf_stub.StubOp(op, **good).h


def _camel(snake):
    return "".join(p.capitalize() for p in snake.split("_"))


def test_python_side_of_the_shim_passes_what_the_ops_declare():
    """mrcnn_layers_b200.py is parsed (TensorFlow is not importable): every `_ops.<op>(...)` call must pass one
    positional argument per declared input, only declared attributes, every required attribute, and unpack as many
    results as the op has outputs; every @tf.RegisterGradient / tf.no_gradient names a registered op."""
    sig = tf_stub.signatures()
    tree = ast.parse(open(PY_SHIM).read())
    seen = set()
    for node in ast.walk(tree):
        calls = []
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Call):
            calls.append((node.value, node.targets[0]))
        elif isinstance(node, ast.Call):
            calls.append((node, None))
        for call, target in calls:
            f = call.func
            if not (isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) and f.value.id == "_ops"):
                continue
            op = _camel(f.attr)
            assert op in sig, f"_ops.{f.attr}: no op {op} registered"
            seen.add(op)
            n_pos = 0
            for a in call.args:
                if isinstance(a, ast.Starred):      # *inputs[2:6]
                    sl = a.value.slice
                    n_pos += sl.upper.value - sl.lower.value
                else:
                    n_pos += 1
            assert n_pos == len(sig[op]["inputs"]), (op, n_pos)
            kws = {k.arg for k in call.keywords}
            assert kws <= set(sig[op]["attrs"]), (op, kws - set(sig[op]["attrs"]))
            required = {a for a, (_, d) in sig[op]["attrs"].items() if d is None} - {"N"}   # N: inferred from the list
            assert required <= kws, (op, required - kws)
            if isinstance(target, ast.Tuple):
                assert len(target.elts) == len(sig[op]["outputs"]), op
    assert seen == set(OPS)
    src = open(PY_SHIM).read()
    for name in re.findall(r'(?:RegisterGradient|no_gradient)\("(\w+)"\)', src):
        assert name in sig, name
    # the reference's class names / layer names survive in the Python shim
    for cls, lname in [("ProposalLayer", "roi"), ("DetectionLayer", "mrcnn_detection"),
                       ("DetectionTargetLayer", "proposal_targets"), ("PyramidROIAlign", "roi_align")]:
        assert re.search(rf"class {cls}\(", src) and f"name='{lname}'" in src, cls


# --------------------------------------------------------------------------------------------------- GPU
def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.fixture(scope="module")
def F():
    from maskrcnn_tf2_b200 import functional
    return functional


@pytest.mark.gpu
def test_proposal_ops_run_through_the_shim(F, dev, stub_ready):
    from maskrcnn_tf2_b200 import synth
    B, S = 2, 256
    anchors = synth.pyramid_anchors(S)
    probs, bbox = zip(*[synth.rpn_outputs(np.random.default_rng(40 + b), anchors, "clustered", S) for b in range(B)])
    tp, tb = T(np.stack(probs), dev), T(np.stack(bbox), dev)
    ta = T(np.broadcast_to(anchors, (B,) + anchors.shape), dev)
    ref = F.proposal_forward(tp, tb, ta, 6000, 1000, SD, 0.7, debug=True)
    out, topk, keep = tf_stub.StubOp("MrcnnProposal", proposal_count=1000, pre_nms_limit=6000, nms_threshold=0.7,
                                     std_dev=SD)(tp, tb, ta)
    assert torch.equal(out, ref["proposals"]) and torch.equal(topk, ref["topk_idx"])
    assert torch.equal(keep, ref["keep_idx"])
    g = torch.randn_like(out)
    (grad,) = tf_stub.StubOp("MrcnnProposalGrad", std_dev=SD)(g, tb, ta, topk, keep)
    assert torch.equal(grad, F.proposal_backward(g, tb, ta, topk, keep, SD))
    # per-level entry: logits whose softmax is `probs`, split at the level boundaries
    counts = [3 * h * w for h, w in synth.backbone_shapes(S, [4, 8, 16, 32, 64])]
    starts = np.concatenate([[0], np.cumsum(counts)])
    logits = torch.log(tp.clamp_min(1e-30))
    lv = [logits[:, s:e].contiguous() for s, e in zip(starts[:-1], starts[1:])]
    lb = [tb[:, s:e].contiguous() for s, e in zip(starts[:-1], starts[1:])]
    want, want_probs = F.proposal_forward_levels(lv, lb, ta, 6000, 1000, SD, 0.7, return_probs=True)
    got, got_probs = tf_stub.StubOp("MrcnnProposalLevels", N=5, proposal_count=1000)(*lv, *lb, ta)
    assert torch.equal(got, want) and torch.equal(got_probs, want_probs)
    with pytest.raises(RuntimeError, match="rpn_bbox / anchors"):
        tf_stub.StubOp("MrcnnProposal", proposal_count=10)(tp, tb[:, :-1], ta)
    with pytest.raises(RuntimeError, match="add up"):
        tf_stub.StubOp("MrcnnProposalLevels", N=5, proposal_count=10)(*lv, *lb, ta[:, :-3])


@pytest.mark.gpu
def test_roialign_ops_run_through_the_shim(F, dev, stub_ready):
    from maskrcnn_tf2_b200 import synth
    rng = np.random.default_rng(41)
    B, Nr, C = 2, 120, 64
    boxes = np.stack([random_boxes(rng, Nr, min_size=0.02, max_size=0.8) for _ in range(B)])
    boxes[:, -10:] = 0
    fm = [T(rng.standard_normal((B, s, s, C)).astype(np.float32), dev) for s in (64, 32, 16, 8)]
    tb, meta = T(boxes, dev), T(synth.image_meta(B, 256, 81), dev)
    for pool in ((7, 7), (14, 14)):
        want, want_map = F.roialign_forward(tb, meta, fm, pool)
        got, got_map = tf_stub.StubOp("MrcnnPyramidRoiAlign", pool_height=pool[0], pool_width=pool[1])(tb, meta, *fm)
        assert torch.equal(got, want) and torch.equal(got_map, want_map)
        g = torch.randn_like(want)
        grads = tf_stub.StubOp("MrcnnPyramidRoiAlignGrad")(g, tb, got_map, *fm)
        ref = F.roialign_backward(g, tb, want_map, [tuple(f.shape) for f in fm], deterministic=True)
        for l, (a, b) in enumerate(zip(grads, ref)):
            # the shim always passes a workspace: the sequential-order (reproducible) gradient.  Pixel (0,0) collects
            # every bin of the zero-padded ROIs and is the one place that still accumulates atomically
            a0, b0 = a[:, 0, 0, :].clone(), b[:, 0, 0, :].clone()
            a[:, 0, 0, :] = 0
            b[:, 0, 0, :] = 0
            assert torch.equal(a, b), (pool, l, float((a - b).abs().max()))
            assert torch.allclose(a0, b0, rtol=1e-4, atol=1e-3), (pool, l, float((a0 - b0).abs().max()))
    with pytest.raises(RuntimeError, match="feature maps"):
        tf_stub.StubOp("MrcnnPyramidRoiAlign", pool_height=7, pool_width=7)(tb, meta, fm[0], fm[1], fm[2], fm[3][:1])


@pytest.mark.gpu
def test_detection_op_runs_through_the_shim(F, dev, stub_ready):
    from maskrcnn_tf2_b200 import synth
    rng = np.random.default_rng(42)
    B, Nr, NC = 3, 1000, 81
    rois = T(np.stack([random_boxes(rng, Nr, clusters=6) for _ in range(B)]), dev)
    probs, deltas = synth.head_outputs(rng, B, Nr, NC)
    tp, td, meta = T(probs, dev), T(deltas, dev), T(synth.image_meta(B, 1024, NC), dev)
    for conf in (0.7, 0.0):
        want, want_boxes = F.detection_forward(rois, tp, td, meta, SD, conf, 100, 0.3, return_boxes=True)
        got, got_boxes = tf_stub.StubOp("MrcnnDetection", min_confidence=conf, use_min_confidence=bool(conf),
                                        max_instances=100, nms_threshold=0.3, std_dev=SD)(rois, tp, td, meta)
        assert torch.equal(got, want) and torch.equal(got_boxes, want_boxes)
        assert torch.equal(got[..., :4], got_boxes)
    with pytest.raises(RuntimeError, match="mrcnn_bbox"):
        tf_stub.StubOp("MrcnnDetection")(rois, tp, td[:, :, :5], meta)


@pytest.mark.gpu
def test_training_side_ops_run_through_the_shim(F, dev, stub_ready):
    rng = np.random.default_rng(43)
    B, P, G, MH = 2, 600, 20, 56
    props = np.stack([random_boxes(rng, P, min_size=0.04, max_size=0.5, clusters=8) for _ in range(B)])
    props[:, -60:] = 0
    gtb = np.zeros((B, G, 4), np.float32)
    gtc = np.zeros((B, G), np.int32)
    for b in range(B):
        pick = rng.choice(P - 60, 6, replace=False)
        gtb[b, :6] = props[b, pick] + rng.normal(0, 0.004, (6, 4)).astype(np.float32)
        gtc[b, :6] = rng.integers(1, 81, 6)
    gtc[0, 5] *= -1
    masks = rng.uniform(0, 1, (B, MH, MH, G)) < 0.5
    keys = rng.integers(0, 2 ** 32, (B, P), dtype=np.uint64).astype(np.uint32).view(np.int32)
    tp, tc, tb, tk = T(props, dev), T(gtc, dev), T(gtb, dev), T(keys, dev)
    want = F.detection_target_forward(tp, tc, tb, T(masks.astype(np.uint8), dev), tk, 100, 0.33, SD, (28, 28))
    got = tf_stub.StubOp("MrcnnDetectionTarget", train_rois_per_image=100, roi_positive_ratio=0.33, mask_height=28,
                         mask_width=28, use_mini_masks=False, std_dev=SD)(tp, tc, tb, T(masks, dev), tk)   # tf.bool masks
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    assert int((want[1] > 0).sum()) > 0
    # RPN targets (data-loader side, float64)
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import AnchorsLayer
    anchors_px = AnchorsLayer(make_config(img_size=256, batch_size=B), device=dev).anchors_px
    A = anchors_px.shape[0]
    gcls = T(np.array([[3, 1, 0, -2], [7, 0, 0, 0]], np.int32), dev)
    gbox = T(np.array([[[20, 30, 120, 140], [100, 60, 220, 250], [0, 0, 0, 0], [10, 10, 200, 90]],
                       [[64, 64, 192, 192], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]], np.int32), dev)
    rk = T(rng.random((B, A), dtype=np.float32), dev)
    # the reference's loader hands over the float32 config array (config.py:90): the divisor is (double)0.1f
    w_match, w_bbox, w_bbox32 = F.rpn_targets_forward(anchors_px, gcls, gbox, rk, 64, np.array(SD, np.float32),
                                                      return_f32=True)
    match, bbox, bbox32 = tf_stub.StubOp("MrcnnRpnTargets", rpn_train_anchors_per_image=64, rpn_bbox_std_dev=SD,
                                         eps=1e-3)(anchors_px, gcls, gbox, rk)
    assert tuple(match.shape) == (B, A, 1) and torch.equal(match[..., 0], w_match)
    assert torch.equal(bbox, w_bbox) and torch.equal(bbox32, w_bbox32)


@pytest.mark.gpu
def test_shim_rejects_mismatched_input_shapes_before_any_launch(F, dev, stub_ready):
    """Every OpKernel::Compute checks the shapes its launcher dereferences (OP_REQUIRES -> InvalidArgument), so a
    mismatch -- e.g. a different max_gt_instances for gt_masks and gt_class_ids -- is an error status, never an
    out-of-bounds device read."""
    rng = np.random.default_rng(44)
    B, P, G, MH = 2, 64, 8, 16
    props = T(np.stack([random_boxes(rng, P) for _ in range(B)]), dev)
    cls = T(rng.integers(0, 5, (B, G)).astype(np.int32), dev)
    boxes = T(np.stack([random_boxes(rng, G) for _ in range(B)]), dev)
    masks = T(rng.uniform(0, 1, (B, MH, MH, G)) < 0.5, dev)
    keys = T(rng.integers(0, 2 ** 31, (B, P)).astype(np.int32), dev)
    op = tf_stub.StubOp("MrcnnDetectionTarget", train_rois_per_image=16, roi_positive_ratio=0.33, mask_height=28,
                        mask_width=28, use_mini_masks=False, std_dev=SD)
    op(props, cls, boxes, masks, keys)                                              # the matching call is fine
    bad_masks = T(rng.uniform(0, 1, (B, MH, MH, G + 3)) < 0.5, dev)                  # another max_gt_instances
    for args in ((props, cls, boxes, bad_masks, keys), (props, cls, boxes[:, :5], masks, keys),
                 (props, cls, boxes, masks, keys[:, :10]), (props, cls[:1], boxes, masks, keys)):
        with pytest.raises(RuntimeError, match="status"):
            op(*args)
    fm = [torch.zeros((B, s, s, 8), device=dev) for s in (16, 8, 4, 2)]
    meta = T(np.zeros((B, 17), np.float32) + 64.0, dev)
    ra = tf_stub.StubOp("MrcnnPyramidRoiAlign", pool_height=7, pool_width=7, denominator=244.0, map_mode=0)
    ra(props, meta, *fm)
    with pytest.raises(RuntimeError, match="status"):
        ra(props, meta[:1], *fm)                                                     # image_meta of another batch size
    with pytest.raises(RuntimeError, match="status"):
        ra(props, meta[:, :4], *fm)                                                  # image_meta too short
    pooled, roi_map = ra(props, meta, *fm)
    grad_op = tf_stub.StubOp("MrcnnPyramidRoiAlignGrad")
    grad_op(torch.ones_like(pooled), props, roi_map, *fm)
    with pytest.raises(RuntimeError, match="status"):
        grad_op(torch.ones_like(pooled), props[:, :10], roi_map, *fm)                # boxes of another N
    with pytest.raises(RuntimeError, match="status"):
        grad_op(torch.ones_like(pooled), props, roi_map[:, :10], *fm)                # roi_map of another N


# ------------------------------------------------------------------------ the Python side, executed (fake `tensorflow`)
def test_python_shim_executes_under_the_fake_tensorflow_and_keeps_the_reference_surface():
    """tf_shim/mrcnn_layers_b200.py is run as it is against tests/tf_stub/fake_tf.py: the module body (op library,
    gradient registrations), the constructors, names, shapes and configs of SURVEY.md 8(b)."""
    import inspect
    from tf_stub import fake_tf
    from maskrcnn_tf2_b200 import layers as torch_layers, make_config
    with fake_tf.installed() as tf:
        shim = fake_tf.import_shim()
        assert set(tf.gradients) == {"MrcnnPyramidRoiAlign", "MrcnnProposal"}
        assert tf.no_gradients == {"MrcnnProposalGrad", "MrcnnDetection", "MrcnnDetectionTarget"}
        assert set(tf.serializable) >= {"ProposalLayer", "PyramidROIAlign", "DetectionLayer", "DetectionTargetLayer"}
        cfg = make_config()
        want = {   # constructor parameters of the reference classes (mrcnn_layers.py:218, 574, 353-355, 309)
            "ProposalLayer": ["proposal_count", "config", "name", "kwargs"],
            "PyramidROIAlign": ["pool_shape", "denominator", "name", "kwargs"],
            "DetectionLayer": ["proposals", "detection_min_confidence", "detection_max_instances",
                               "detection_nms_threshold", "bbox_std_dev", "images_per_gpu", "batch_size", "name",
                               "kwargs"],
            "DetectionTargetLayer": ["config", "name", "kwargs"],
        }
        for cls, params in want.items():
            got = list(inspect.signature(getattr(shim, cls).__init__).parameters)[1:]
            assert got == params, (cls, got)
            # the torch mirror takes the same arguments in the same order (plus keyword-only-in-practice extensions)
            got = list(inspect.signature(getattr(torch_layers, cls).__init__).parameters)[1:]
            assert got[:len(params) - 1] == params[:-1] and got[-1] == "kwargs", (cls, got)
        p = shim.ProposalLayer(proposal_count=1000, config=cfg)
        assert p.name == "roi" and p.compute_output_shape(None) == (None, 1000, 4) and p.nms_threshold == 0.7
        r = shim.PyramidROIAlign([7, 7], name="roi_align_classifier")
        assert r.name == "roi_align_classifier" and r.denominator == 244.0
        assert r.compute_output_shape([(None, 1000, 4), (None, 93), (None, 256, 256, 256)]) == (None, 1000, 7, 7, 256)
        assert shim.PyramidROIAlign([14, 14]).name == "roi_align"
        d = shim.DetectionLayer(1000, 0.7, 100, 0.3, cfg["bbox_std_dev"], 8, 8)
        assert d.name == "mrcnn_detection" and d.compute_output_shape(None) == (None, 100, 6)
        t = shim.DetectionTargetLayer(cfg)
        assert t.name == "proposal_targets" and t.compute_mask(None) == [None] * 4
        assert t.compute_output_shape(None) == [(None, 200, 4), (None, 200), (None, 200, 4), (None, 200, 28, 28)]
        for layer in (p, r, d, t):
            assert layer.get_config()["name"] == layer.name
            layer.build(None)
            assert layer.built
    assert "tensorflow" not in __import__("sys").modules or not getattr(__import__("sys").modules["tensorflow"],
                                                                        "__fake__", False)


@pytest.mark.gpu
def test_python_shim_layers_run_the_roi_stage_through_the_cpp_shim(F, dev, stub_ready):
    """Inference wiring of model.py:556-573 with the shim's own Keras classes (ProposalLayer -> PyramidROIAlign ->
    DetectionLayer -> DetectedBoxesExtraction -> PyramidROIAlign), every op going Python shim -> C++ OpKernel ->
    launcher, against the torch mirror of the same classes; then the training layer and both registered gradients."""
    from tf_stub import fake_tf
    from maskrcnn_tf2_b200 import layers as TL, make_config, synth
    B, S, NC = 2, 256, 5
    cfg = make_config(img_size=S, num_classes=NC, batch_size=B)
    x = synth.inference_batch(77, B, img_size=S, num_classes=NC, regime="clustered", n_rois=1000, channels=64)
    t = lambda a: T(a, dev)
    probs, bbox, anchors, meta = t(x["rpn_probs"]), t(x["rpn_bbox"]), t(x["anchors"]), t(x["image_meta"])
    fm = [t(f) for f in x["feature_maps"]]
    cls, dl = t(x["mrcnn_class"]), t(x["mrcnn_bbox"])
    det_args = (1000, cfg["detection_min_confidence"], cfg["detection_max_instances"],
                cfg["detection_nms_threshold"], cfg["bbox_std_dev"], B, B)
    with fake_tf.installed() as tf:
        shim = fake_tf.import_shim()
        rois = shim.ProposalLayer(1000, cfg)([probs, bbox, anchors])
        pooled = shim.PyramidROIAlign([7, 7], name="roi_align_classifier")([rois, meta] + fm)
        det = shim.DetectionLayer(*det_args)([rois, cls, dl, meta])
        boxes = shim.DetectedBoxesExtraction(cfg)(det)
        mask_pooled = shim.PyramidROIAlign([14, 14], name="roi_align_mask")([boxes, meta] + fm)
        assert [c[0] for c in tf.library.calls] == ["MrcnnProposal", "MrcnnPyramidRoiAlign", "MrcnnDetection",
                                                    "MrcnnPyramidRoiAlign"]
        # the same through the torch mirror of the classes
        w_rois = TL.ProposalLayer(1000, cfg)([probs, bbox, anchors])
        w_pooled = TL.PyramidROIAlign([7, 7], name="roi_align_classifier")([w_rois, meta] + fm)
        w_det = TL.DetectionLayer(*det_args)([w_rois, cls, dl, meta])
        w_mask = TL.PyramidROIAlign([14, 14], name="roi_align_mask")([w_det[..., :4].contiguous(), meta] + fm)
        assert torch.equal(rois, w_rois) and torch.equal(pooled, w_pooled) and torch.equal(det, w_det)
        assert torch.equal(boxes, w_det[..., :4]) and torch.equal(mask_pooled, w_mask)
        assert int((w_det[..., 4] > 0).sum()) > 0

        # registered gradients: ROIAlign (feature maps only) and Proposal (rpn_bbox only)
        g = torch.randn_like(pooled)
        roi_map = F.roialign_forward(rois, meta, fm, (7, 7))[1]
        grads = tf.gradients["MrcnnPyramidRoiAlign"](fake_tf.FakeOp([rois, meta] + fm, [pooled, roi_map], {}), g, None)
        assert grads[0] is None and grads[1] is None and len(grads) == 6
        ref = F.roialign_backward(g, rois, roi_map, [tuple(f.shape) for f in fm], deterministic=True)
        for a, b in zip(grads[2:], ref):
            # a wiring check: pixels with > 1024 samples and pixel (0,0) (zero-padded ROIs) accumulate atomically
            assert torch.allclose(a, b, rtol=1e-3, atol=1e-2)
        dbg = F.proposal_forward(probs, bbox, anchors, 6000, 1000, cfg["rpn_bbox_std_dev"], 0.7, debug=True)
        gp = torch.randn_like(rois)
        sd = [float(v) for v in cfg["rpn_bbox_std_dev"]]
        out = tf.gradients["MrcnnProposal"](fake_tf.FakeOp([probs, bbox, anchors],
                                                           [rois, dbg["topk_idx"], dbg["keep_idx"]], {"std_dev": sd}),
                                            gp, None, None)
        assert out[0] is None and out[2] is None
        assert torch.equal(out[1], F.proposal_backward(gp, bbox, anchors, dbg["topk_idx"], dbg["keep_idx"], sd))

        # training layer: the keys the shim draws are captured and replayed through the ctypes path
        rng = np.random.default_rng(78)
        P, G = 500, 20
        props = np.stack([random_boxes(rng, P, min_size=0.04, max_size=0.5, clusters=8) for _ in range(B)])
        gtb = np.zeros((B, G, 4), np.float32)
        gtc = np.zeros((B, G), np.int32)
        for b in range(B):
            pick = rng.choice(P, 6, replace=False)
            gtb[b, :6] = props[b, pick] + rng.normal(0, 0.004, (6, 4)).astype(np.float32)
            gtc[b, :6] = rng.integers(1, NC, 6)
        masks = rng.uniform(0, 1, (B, 56, 56, G)) < 0.5
        drawn = []
        uniform = tf.random.uniform
        tf.random.uniform = lambda *a, **k: drawn.append(uniform(*a, **k)) or drawn[-1]
        tcfg = make_config(img_size=S, num_classes=NC, batch_size=B, train_rois_per_image=64)
        got = shim.DetectionTargetLayer(tcfg)([t(props), t(gtc.astype(np.int64)), t(gtb), t(masks.astype(np.uint8))])
        assert len(got) == 4 and len(drawn) == 1 and drawn[0].dtype == torch.int32
        want = F.detection_target_forward(t(props), t(gtc), t(gtb), t(masks.astype(np.uint8)), drawn[0], 64,
                                          tcfg["roi_positive_ratio"], tcfg["bbox_std_dev"], tcfg["mask_shape"],
                                          use_mini_masks=bool(tcfg["use_mini_masks"]))
        for a, b in zip(got, want):
            assert torch.equal(a, b)
        assert int((want[1] > 0).sum()) > 0
        # loader-side helper of the shim
        anchors_px = TL.AnchorsLayer(cfg, device=dev).anchors_px
        gcls = t(np.array([[3, 1, 0, -2], [2, 0, 0, 0]], np.int64))
        gbox = t(np.array([[[20, 30, 120, 140], [100, 60, 220, 250], [0, 0, 0, 0], [10, 10, 200, 90]],
                           [[64, 64, 192, 192], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]], np.float32))
        drawn.clear()
        match, bbox32 = shim.build_rpn_targets(anchors_px, gcls, gbox, 64, cfg["rpn_bbox_std_dev"])
        w_match, _, w_bbox32 = F.rpn_targets_forward(anchors_px, gcls.to(torch.int32), gbox.to(torch.int32), drawn[0],
                                                     64, cfg["rpn_bbox_std_dev"], return_f32=True)
        assert torch.equal(match[..., 0], w_match) and torch.equal(bbox32, w_bbox32)
