"""Host-side logic that needs no GPU: the layer classes' reference-facing surface, the config mirror, the
synthetic-input generator and the shard arithmetic."""
import numpy as np
import pytest
import torch


def test_layers_keep_reference_signatures_and_names():
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import DetectionLayer, DetectionTargetLayer, ProposalLayer, PyramidROIAlign
    cfg = make_config()
    p = ProposalLayer(proposal_count=1000, config=cfg)
    assert p.name == "roi" and p.proposal_count == 1000 and p.nms_threshold == 0.7
    assert p.compute_output_shape(None) == (None, 1000, 4)
    r = PyramidROIAlign([7, 7], name="roi_align_classifier")
    assert r.pool_shape == (7, 7) and r.denominator == 244.0 and r.name == "roi_align_classifier"
    assert PyramidROIAlign([14, 14]).name == "roi_align"
    d = DetectionLayer(proposals=1000, detection_min_confidence=0.7, detection_max_instances=100,
                       detection_nms_threshold=0.3, bbox_std_dev=cfg['bbox_std_dev'], images_per_gpu=8, batch_size=8)
    assert d.name == "mrcnn_detection" and d.compute_output_shape(None) == (None, 100, 6)
    t = DetectionTargetLayer(cfg)
    assert t.name == "proposal_targets"
    assert t.compute_output_shape(None) == [(None, 200, 4), (None, 200), (None, 200, 4), (None, 200, 28, 28)]
    assert t.compute_mask(None) == [None, None, None, None]
    for layer in (p, r, d, t):
        assert layer.get_config()["name"] == layer.name


def test_layers_refuse_cpu_tensors_no_fallback():
    from maskrcnn_tf2_b200 import make_config
    from maskrcnn_tf2_b200.layers import ProposalLayer, PyramidROIAlign
    cfg = make_config(img_size=256)
    with pytest.raises(TypeError, match="no CPU path"):
        ProposalLayer(100, cfg)([torch.zeros(1, 10, 2), torch.zeros(1, 10, 4), torch.zeros(1, 10, 4)])
    with pytest.raises(TypeError, match="no CPU path"):
        PyramidROIAlign([7, 7])([torch.zeros(1, 4, 4), torch.zeros(1, 93)] + [torch.zeros(1, 8, 8, 4)] * 4)


def test_product_path_never_imports_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "maskrcnn_tf2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liborc" not in text, f
                assert "mrcnn_oracle" not in text, f


def test_config_mirror_defaults():
    from maskrcnn_tf2_b200 import CONFIG, make_config
    assert CONFIG['pre_nms_limit'] == 6000 and CONFIG['post_nms_rois_inference'] == 1000
    assert CONFIG['rpn_nms_threshold'] == 0.7 and CONFIG['detection_nms_threshold'] == 0.3
    assert CONFIG['rpn_bbox_std_dev'].dtype == np.float32
    c = make_config(num_classes=2, img_size=512, batch_size=1)
    assert c['meta_shape'] == 14 and c['image_shape'] == (512, 512, 3) and c['images_per_gpu'] == 1
    assert CONFIG['num_classes'] == 81            # make_config copies


def test_synthetic_anchors_follow_the_reference_layout():
    from maskrcnn_tf2_b200 import synth
    a = synth.pyramid_anchors(1024)
    assert a.shape == (261888, 4) and a.dtype == np.float32
    assert synth.pyramid_anchors(512).shape[0] == 65472 and synth.pyramid_anchors(256).shape[0] == 16368
    # first cell of P2: scale 32, ratios 0.5/1/2 innermost, centred at pixel (0,0), normalised by 1023 after -[0,0,1,1]
    h = 32 / np.sqrt(0.5)
    w = 32 * np.sqrt(0.5)
    ref0 = (np.array([-h / 2, -w / 2, h / 2, w / 2], np.float32) - np.array([0, 0, 1, 1], np.float32)) / np.float32(1023)
    assert np.allclose(a[0], ref0, atol=1e-7)
    assert np.allclose(a[1], (np.array([-16, -16, 16, 16], np.float32) - [0, 0, 1, 1]) / 1023, atol=1e-7)
    # second cell of the first row moves x by the stride (4 px)
    assert np.allclose(a[3] - a[0], [0, 4 / 1023, 0, 4 / 1023], atol=1e-7)
    # P3 block starts after 256*256*3 anchors with scale 64
    p3 = a[256 * 256 * 3 + 1]
    assert np.allclose(p3, (np.array([-32, -32, 32, 32], np.float32) - [0, 0, 1, 1]) / 1023, atol=1e-7)


def test_synthetic_regimes_exercise_nms_differently(orc):
    from maskrcnn_tf2_b200 import synth
    a = synth.pyramid_anchors(512)
    kept = {}
    for regime in ("iid", "clustered", "sparse"):
        p, d = synth.rpn_outputs(np.random.default_rng(5), a, regime, 512)
        assert p.shape == (65472, 2) and np.allclose(p.sum(1), 1.0, atol=1e-6)
        r = orc.proposal_layer(p[None], d[None], a[None], 6000, 1000, [0.1, 0.1, 0.2, 0.2], 0.7)
        kept[regime] = int(r["keep_count"][0])
    assert kept["iid"] == 1000 and kept["sparse"] < 700      # sparse: heavy suppression, zero padding exercised
    assert kept["clustered"] > kept["sparse"]


def test_shard_ranges_cover_the_batch_once():
    from maskrcnn_tf2_b200.sharding import shard_range, shard_sizes
    for total, world in [(64, 1), (64, 2), (64, 8), (10, 4), (3, 8)]:
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert shard_sizes(total, world) == [b - a for a, b in spans]
        assert max(shard_sizes(total, world)) - min(shard_sizes(total, world)) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_host_resident_inputs_must_be_page_locked():
    """The demand-driven paths dereference host pointers on the device: pageable memory is refused up front (before
    any CUDA call), there is no silent copy and no CPU path."""
    import torch
    from maskrcnn_tf2_b200 import functional as F
    maps = [torch.zeros((1, 8 // s, 8 // s, 4)) for s in (1, 2, 4, 8)]
    with pytest.raises(TypeError):
        F.HostMapStage(maps, torch.device("cpu"))
    with pytest.raises(TypeError):
        F._req(torch.zeros(4), torch.float32, "rpn_bbox", pinned_ok=True)


def test_onnx_export_table_matches_the_op_registrations():
    """tf_shim/onnx_export.py (tf2onnx kwargs for maskrcnn_to_onnx, inference_optimize.py:12-20) must name op types,
    result-output counts and attributes exactly as tf_shim/mrcnn_roi_ops.cc registers them."""
    import importlib.util
    import os
    import re
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "maskrcnn_tf2_b200", "tf_shim")
    spec = importlib.util.spec_from_file_location("onnx_export", os.path.join(shim, "onnx_export.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # imports neither tensorflow nor tf2onnx
    src = open(os.path.join(shim, "mrcnn_roi_ops.cc")).read()
    regs = {}
    for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?)\.SetShapeFn', src, re.S):
        body = re.sub(r'//[^\n]*', '', m.group(2))
        regs[m.group(1)] = (re.findall(r'\.Output\("(\w+):', body), re.findall(r'\.Attr\("(\w+):', body))
    for op, (n_out, attrs) in mod.EXPORTED_OPS.items():
        outs, reg_attrs = regs[op]
        assert 1 <= n_out <= len(outs), op
        assert set(attrs) == set(reg_attrs), (op, attrs, reg_attrs)
    kw = mod.tf2onnx_kwargs({"opset": 11, "custom_ops": {"Foo": "bar"}})
    assert kw["opset"] == 11 and kw["custom_ops"]["Foo"] == "bar"
    assert set(mod.EXPORTED_OPS) <= set(kw["custom_ops"]) and set(mod.EXPORTED_OPS) == set(kw["custom_op_handlers"])

    class _Node:  # the handler keeps the node and moves it into the custom domain; gradient-only outputs unconsumed
        type, output, domain = "MrcnnProposal", ["p:0", "p:1", "p:2"], ""

    class _Ctx:
        def __init__(self, used):
            self.used = used

        def find_output_consumers(self, out):
            return [1] if out in self.used else []

    fn, extra = kw["custom_op_handlers"]["MrcnnProposal"]
    assert fn(_Ctx({"p:0"}), _Node(), "n", extra).domain == mod.DOMAIN
    with pytest.raises(ValueError):
        fn(_Ctx({"p:1"}), _Node(), "n", extra)


def test_bench_reference_arm_prints_one_contract_line_and_only_on_rank_zero():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys; under torchrun every rank but 0 exits 0 without work or output."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--batch", "2", "--img-size", "256", "--num-classes", "5"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "roi_stage_images_per_sec" and line["unit"] == "images/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
    other = subprocess.run(cmd + ["--gpus", "2"], env=dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2"),
                           capture_output=True, text=True, timeout=300)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_bench_algorithmic_bytes_match_the_survey_figures():
    """ROIAlign forward = output written once + min(4 x output, all map pixels) + boxes, per image."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert round(bench.algorithmic_bytes(1024, 1000, 7, 7) / 1e6, 2) == 139.32      # 7x7, N=1000
    assert round(bench.algorithmic_bytes(1024, 100, 14, 14) / 1e6, 2) == 100.35     # 14x14, N=100
    assert round(bench.algorithmic_bytes(512, 1000, 7, 7) / 1e6, 2) == 72.47
    assert round(bench.algorithmic_bytes(256, 100, 14, 14) / 1e6, 2) == 25.64
    # the tight count the roofline uses never exceeds the closed form
    import numpy as np
    boxes = np.array([[[0.1, 0.1, 0.2, 0.2], [0.0, 0.0, 1.0, 1.0]]], np.float32)
    roi_map = np.array([[0, 3]], np.int32)
    touched = bench.touched_map_bytes(boxes, roi_map, 1024, 7, 7)
    assert 0 < touched <= 2 * 4 * 7 * 7 * 256 * 4
