"""Phase timeline of the fused-ordering NMS kernel of DetectionLayer, CTA 0 (debug build `make prof`)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get("MRCNN_PROF_LIB", "libmrcnn_roi_b200_prof.so"))
from maskrcnn_tf2_b200 import functional as F, synth
L = _lib.lib()
B = 8
x = synth.inference_batch(2, B)
t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).cuda()
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
rois = F.proposal_forward(t(x["rpn_probs"]), t(x["rpn_bbox"]), t(x["anchors"]), 6000, 1000, SD, 0.7)
mc, mb, meta = t(x["mrcnn_class"]), t(x["mrcnn_bbox"]), t(x["image_meta"])
for _ in range(3):
    det = F.detection_forward(rois, mc, mb, meta, SD, 0.7, 100, 0.3)
torch.cuda.synchronize()
tl = (ctypes.c_longlong * (128 * 8))()
L.mrcnn_debug_nms_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
L.mrcnn_debug_nms_timeline(tl)
tl = np.array(list(tl), dtype=np.int64).reshape(128, 8)
g = tl[100] - tl[100, 0]
print("entry 0 | dependency wait over %d | candidates compacted %d | ranked + staged %d | diag(0) + cluster barrier %d | "
      "sweep over %d | block barrier %d | outputs written %d" % tuple(g[1:8]))
for i in range(8):
    if tl[i, 3] > 0:
        r = tl[i] - tl[100, 0]
        print(f"  tile {i}: start {r[0]} far arrived {r[1]} kept known {r[2]} released {r[3]}")
