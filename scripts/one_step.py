"""Two passes of the inference ROI stage at config 2 (B=8, COCO shape) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskrcnn_tf2_b200 import make_config, synth
from maskrcnn_tf2_b200.layers import DetectedBoxesExtraction, DetectionLayer, ProposalLayer, PyramidROIAlign
B = int(os.environ.get("B", "8"))
regime = os.environ.get("REGIME", "clustered")
cfg = make_config(batch_size=B)
x = synth.inference_batch(2, B, regime=regime)
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(v).to(dev) for k, v in x.items() if k != "feature_maps"}
maps = [torch.from_numpy(f).to(dev) for f in x["feature_maps"]]
prop = ProposalLayer(1000, cfg); a7 = PyramidROIAlign([7, 7]); a14 = PyramidROIAlign([14, 14])
det = DetectionLayer(1000, 0.7, 100, 0.3, cfg["bbox_std_dev"], B, B)
for _ in range(2):
    rois = prop([d["rpn_probs"], d["rpn_bbox"], d["anchors"]])
    p7 = a7([rois, d["image_meta"]] + maps)
    dt = det([rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"]])
    p14 = a14([DetectedBoxesExtraction(cfg)(dt), d["image_meta"]] + maps)
torch.cuda.synchronize()
print("ok", float(p7.sum()), float(p14.sum()))
