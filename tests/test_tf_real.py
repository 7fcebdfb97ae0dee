"""Self-activating tests under a REAL TensorFlow.

TensorFlow is not installable in this image (no network), so every test here skips today.  The day a
CUDA-12 TensorFlow wheel is importable they run without further changes and pin what the CPU oracle cannot:

  * the shim builds against TensorFlow's own headers (`tf.sysconfig` flags) and loads with `tf.load_op_library`;
  * the reference's OWN layers (loaded unmodified from /root/reference/src, eager, on TF's kernels -- TopKV2,
    NonMaxSuppressionV3, CropAndResize) run beside the shim layers on the committed golden inputs and on seeded
    COCO-shape inputs, compared by the north-star rule: top-k / NMS-derived indices and everything integer exact,
    floats within 1e-5 relative / 1e-6 absolute;
  * gradients dispatched through @tf.RegisterGradient agree with TF's autodiff of the reference layers;
  * `mask_rcnn_functional()` (model.py:398) builds in both modes with the shim layers patched into the reference's
    `mrcnn_layers` module, and `tf2onnx.convert.from_keras(**onnx_export.tf2onnx_kwargs())` converts the inference
    model when tf2onnx is importable too (inference_optimize.py:12-20).

The GPU tests additionally need a CUDA device TensorFlow can see.
"""
import os
import subprocess
import sys
import types

import numpy as np
import pytest

tf = pytest.importorskip("tensorflow", reason="TensorFlow is not installable in this image (SURVEY section 0)")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIM_DIR = os.path.join(ROOT, "maskrcnn_tf2_b200", "tf_shim")
REF_SRC = "/root/reference/src"
RTOL, ATOL = 1e-5, 1e-6
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)

needs_reference = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree is not on this machine")


def _has_tf_gpu():
    try:
        return bool(tf.config.list_physical_devices("GPU"))
    except Exception:
        return False


needs_tf_gpu = pytest.mark.skipif(not _has_tf_gpu(), reason="TensorFlow sees no CUDA device (the ops register GPU kernels only)")


@pytest.fixture(scope="module")
def shim():
    """Builds libmrcnn_roi_ops.so against the installed TensorFlow and imports the drop-in layer module."""
    from maskrcnn_tf2_b200 import _lib
    _lib.build()
    out = os.path.join(SHIM_DIR, "libmrcnn_roi_ops.so")
    src = os.path.join(SHIM_DIR, "mrcnn_roi_ops.cc")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
        cmd = (["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-DGOOGLE_CUDA=1", src, "-o", out]
               + list(tf.sysconfig.get_compile_flags()) + list(tf.sysconfig.get_link_flags())
               + ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(cuda, "include"),
                  "-L" + os.path.join(ROOT, "maskrcnn_tf2_b200"), "-lmrcnn_roi_b200",
                  "-Wl,-rpath," + os.path.join(ROOT, "maskrcnn_tf2_b200")])
        subprocess.check_call(cmd)
    sys.path.insert(0, SHIM_DIR)
    try:
        import mrcnn_layers_b200
    finally:
        sys.path.pop(0)
    return mrcnn_layers_b200


@pytest.fixture(scope="module")
def reference():
    """The reference's own src/layers/mrcnn_layers.py, unmodified; the backbone packages it imports at module top
    (efficientnet, classification_models: unused by the four layers) are stubbed when absent."""
    if not os.path.isdir(REF_SRC):
        pytest.skip("the reference tree is not on this machine")
    for name in ("efficientnet", "efficientnet.keras", "classification_models", "classification_models.keras"):
        try:
            __import__(name)
        except Exception:
            sys.modules.setdefault(name, types.ModuleType(name))
    if getattr(sys.modules.get("classification_models.keras"), "Classifiers", None) is None:
        sys.modules["classification_models.keras"].Classifiers = types.SimpleNamespace(get=lambda *_a, **_k: (None, None))
    sys.path.insert(0, REF_SRC)
    try:
        from layers import mrcnn_layers
    finally:
        sys.path.pop(0)
    return mrcnn_layers


def _golden():
    return np.load(os.path.join(HERE, "golden", "reference_layers_golden.npz"))


def _close(a, b):
    return np.allclose(np.asarray(a), np.asarray(b), rtol=RTOL, atol=ATOL)


def test_shim_builds_against_real_tensorflow_and_registers_every_op(shim):
    from maskrcnn_tf2_b200.tf_shim import onnx_export
    ops = shim._ops
    for op_type in onnx_export.EXPORTED_OPS:
        snake = "".join("_" + ch.lower() if ch.isupper() else ch for ch in op_type).lstrip("_")
        assert hasattr(ops, snake), f"{op_type} is not registered in libmrcnn_roi_ops.so"
    for cls in ("ProposalLayer", "PyramidROIAlign", "DetectionLayer", "DetectionTargetLayer"):
        assert issubclass(getattr(shim, cls), tf.keras.layers.Layer)


@needs_reference
@needs_tf_gpu
def test_reference_layers_and_shim_layers_agree_on_the_golden_inputs(shim, reference):
    g = _golden()
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": 6000, "images_per_gpu": int(g["rpn_probs"].shape[0]),
           "rpn_bbox_std_dev": SD, "bbox_std_dev": SD}
    inputs = [tf.constant(g["rpn_probs"]), tf.constant(g["rpn_bbox"]), tf.constant(g["anchors"])]
    P = int(g["rois"].shape[1])
    ref_rois = reference.ProposalLayer(proposal_count=P, config=cfg)(inputs).numpy()
    got_rois = shim.ProposalLayer(proposal_count=P, config=cfg)(inputs).numpy()
    # NMS keep indices / top-k indices exact <=> the same rows survive in the same order; boxes within tolerance
    assert np.array_equal(np.abs(ref_rois).sum(-1) > 0, np.abs(got_rois).sum(-1) > 0)
    assert _close(got_rois, ref_rois)
    assert _close(ref_rois, g["rois"])          # ... and TensorFlow itself agrees with the committed stand-in vectors
    fm = [tf.constant(g[f"fmap{i}"]) for i in range(4)]
    meta = tf.constant(g["image_meta"])
    for pool, key in (((7, 7), "pooled"), ((14, 14), "mask_pooled")):
        boxes = tf.constant(g["rois"] if key == "pooled" else np.ascontiguousarray(g["detections"][..., :4]))
        ref = reference.PyramidROIAlign(list(pool), name="roi_align_" + key)([boxes, meta] + fm).numpy()
        got = shim.PyramidROIAlign(list(pool), name="roi_align_b200_" + key)([boxes, meta] + fm).numpy()
        assert _close(got, ref) and _close(ref, g[key])
    B = int(g["rois"].shape[0])
    args = dict(proposals=P, detection_min_confidence=0.7, detection_max_instances=int(g["detections"].shape[1]),
                detection_nms_threshold=0.3, bbox_std_dev=SD, images_per_gpu=B, batch_size=B)
    det_in = [tf.constant(g["rois"]), tf.constant(g["mrcnn_class"]), tf.constant(g["mrcnn_bbox"]), meta]
    ref = reference.DetectionLayer(**args)(det_in).numpy()
    got = shim.DetectionLayer(**args)(det_in).numpy()
    assert np.array_equal(got[..., 4], ref[..., 4])             # class ids: exact
    assert _close(got, ref) and _close(ref, g["detections"])


@needs_reference
@needs_tf_gpu
def test_gradients_through_the_registered_gradient_functions(shim, reference):
    rng = np.random.default_rng(3)
    g = _golden()
    fm = [tf.Variable(g[f"fmap{i}"]) for i in range(4)]
    meta, boxes = tf.constant(g["image_meta"]), tf.constant(g["rois"])
    w = tf.constant(rng.standard_normal(g["pooled"].shape).astype(np.float32))
    grads = []
    for layer in (reference.PyramidROIAlign([7, 7], name="ra_ref"), shim.PyramidROIAlign([7, 7], name="ra_b200")):
        with tf.GradientTape() as tape:
            loss = tf.reduce_sum(layer([boxes, meta] + fm) * w)
        grads.append([t.numpy() for t in tape.gradient(loss, fm)])
    for a, b in zip(*grads):
        assert _close(b, a)
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": 6000, "images_per_gpu": int(g["rpn_probs"].shape[0]),
           "rpn_bbox_std_dev": SD, "bbox_std_dev": SD}
    bbox = tf.Variable(g["rpn_bbox"])
    wp = tf.constant(rng.standard_normal(g["rois"].shape).astype(np.float32))
    grads = []
    for cls in (reference.ProposalLayer, shim.ProposalLayer):
        with tf.GradientTape() as tape:
            out = cls(proposal_count=int(g["rois"].shape[1]), config=cfg)([tf.constant(g["rpn_probs"]), bbox,
                                                                          tf.constant(g["anchors"])])
            loss = tf.reduce_sum(out * wp)
        grads.append(tape.gradient(loss, bbox).numpy())     # Q7: the reference does not stop this gradient
    assert np.array_equal(grads[0] != 0, grads[1] != 0) and _close(grads[1], grads[0])


@needs_reference
def test_mask_rcnn_functional_builds_with_the_drop_in_layers(shim, reference, monkeypatch):
    """model.py:398-586 wires mrcnnl.ProposalLayer / DetectionTargetLayer / DetectionLayer and, through
    fpn_classifier_graph / fpn_mask_graph, mrcnnl.PyramidROIAlign: patch the four names, build both models."""
    for name in ("ProposalLayer", "PyramidROIAlign", "DetectionLayer", "DetectionTargetLayer"):
        monkeypatch.setattr(reference, name, getattr(shim, name))
    sys.path.insert(0, REF_SRC)
    try:
        import model as ref_model
        from common.config import CONFIG
    finally:
        sys.path.pop(0)
    for training in (False, True):
        cfg = dict(CONFIG, training=training)
        m = ref_model.mask_rcnn_functional(cfg)
        names = {layer.name for layer in m.layers}
        assert "roi" in names and ("proposal_targets" in names if training else "mrcnn_detection" in names)


@needs_reference
def test_tf2onnx_converts_the_inference_model_with_the_custom_op_table(shim, reference, monkeypatch):
    tf2onnx = pytest.importorskip("tf2onnx")
    from maskrcnn_tf2_b200.tf_shim import onnx_export
    for name in ("ProposalLayer", "PyramidROIAlign", "DetectionLayer"):
        monkeypatch.setattr(reference, name, getattr(shim, name))
    sys.path.insert(0, REF_SRC)
    try:
        import model as ref_model
        from common.config import CONFIG
    finally:
        sys.path.pop(0)
    m = ref_model.mask_rcnn_functional(dict(CONFIG, training=False))
    onnx_model, _ = tf2onnx.convert.from_keras(m, **onnx_export.tf2onnx_kwargs({"opset": 11}))
    types_seen = {n.op_type for n in onnx_model.graph.node if n.domain == onnx_export.DOMAIN}
    assert {"MrcnnProposal", "MrcnnPyramidRoiAlign", "MrcnnDetection"} <= types_seen
    layer_names = " ".join(n.name for n in onnx_model.graph.node)
    for key in ("mrcnn_detection", "roi_align_classifier", "roi_align_mask"):   # inference_optimize.py:455-465 greps these
        assert key in layer_names
