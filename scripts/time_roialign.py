import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import functional as F, synth
dev = torch.device("cuda:0")
B = 8
x = synth.inference_batch(2, B, regime="clustered")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
maps = [t(f) for f in x["feature_maps"]]
rois = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 1000, [0.1, 0.1, 0.2, 0.2], 0.7)
det = F.detection_forward(rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"], [0.1, 0.1, 0.2, 0.2], 0.7, 100, 0.3)
boxes = det[..., :4].contiguous()
def timed(fn, n=30):
    for _ in range(5): fn()
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    # back-to-back pairs so that launch latency is hidden behind the previous call
    for a, b in e:
        a.record(); fn(); fn(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / 3 for a, b in e)
    return ts[len(ts) // 2] * 1e3
print("fwd variant", os.environ.get("MRCNN_ROIALIGN_FWD", "0"),
      "7x7: %.1f us" % timed(lambda: F.roialign_forward(rois, d["image_meta"], maps, (7, 7))),
      "14x14: %.1f us" % timed(lambda: F.roialign_forward(boxes, d["image_meta"], maps, (14, 14))))
