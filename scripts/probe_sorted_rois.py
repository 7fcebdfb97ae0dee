"""Probe: does PyramidROIAlign get faster when the ROIs of an image are processed grouped by level / in spatial order?
Run with map_mode=1 (level - 2): under the reference's first-appearance table (Q2, map_mode=0) permuting the INPUT
changes which map a level samples, i.e. the semantics, not just the order.  Result on B200 (config 2): 117.8 us in every
order (Morton order 115.8), so the order is not worth a sort; and 117.8 us against 140.3 us for the same boxes under the Q2
table -- the headline workload pays 22 us for the quirk (levels sampling the wrong, finer maps).  DESIGN section 4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import functional as F, synth
dev = torch.device("cuda:0")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, iters=30):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for x, y in ev:
        flush.zero_(); x.record(); fn(); y.record()
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return t[len(t) // 2] * 1e3
B = 8
x = synth.inference_batch(2, B)
t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)
d = {k: t(v) for k, v in x.items() if k != "feature_maps"}
maps = [t(f) for f in x["feature_maps"]]
rois = F.proposal_forward(d["rpn_probs"], d["rpn_bbox"], d["anchors"], 6000, 1000, SD, 0.7)
r = rois.cpu().numpy()
def level(b):
    h, w = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    return np.clip(np.round(4 + np.log2(np.sqrt(np.maximum(h * w, 1e-12)) / (224.0 / 1024.0))), 2, 5)
out = {}
det = F.detection_forward(rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"], SD, 0.7, 100, 0.3, return_boxes=True)[1]
dn = det.cpu().numpy()
for name in ("score order", "by level"):
    dd = dn.copy()
    for b in range(B):
        if name == "by level":
            dd[b] = dn[b][np.argsort(level(dn[b]) * 1e6 + np.arange(100), kind="stable")]
    td = t(dd)
    print(f"{name:36s} ROIAlign 14x14: {timed(lambda: F.roialign_forward(td, d["image_meta"], maps, (14, 14), map_mode=1)):7.1f} us", flush=True)
# the whole batch grouped by level (image order inside a level)
lv_all = np.stack([level(r[b]) for b in range(B)])
flat = r.reshape(-1, 4)
for name in ("score order", "by level", "by level, descending", "blocks of 1024 grouped by level", "by level, then 8x8 cells (y, x)", "by level, Morton 16x16"):
    rr = r.copy()
    for b in range(B):
        lv = level(r[b]); cy = (r[b, :, 0] + r[b, :, 2]) / 2; cx = (r[b, :, 1] + r[b, :, 3]) / 2
        if name == "score order": key = np.arange(1000)
        elif name == "by level": key = lv * 1e6 + np.arange(1000)
        elif name == "by level, descending": key = -lv * 1e6 + np.arange(1000)
        elif name.startswith("blocks of 1024"): key = np.arange(1000)
        elif name.startswith("by level, then"): key = lv * 1e6 + np.floor(cy * 8) * 1e3 + np.floor(cx * 8) * 10 + np.arange(1000) * 1e-4
        else:
            iy, ix = np.clip((cy * 16).astype(int), 0, 15), np.clip((cx * 16).astype(int), 0, 15)
            m = np.zeros(1000)
            for bit in range(4): m += ((iy >> bit) & 1) * (2 ** (2 * bit + 1)) + ((ix >> bit) & 1) * (2 ** (2 * bit))
            key = lv * 1e6 + m * 1e3 + np.arange(1000) * 1e-3
        rr[b] = r[b][np.argsort(key, kind="stable")]
    if name.startswith("blocks of 1024"):
        ff = flat.copy(); lvf = level(flat)
        for k0 in range(0, len(ff), 1024):
            sl = slice(k0, min(k0 + 1024, len(ff)))
            ff[sl] = flat[sl][np.argsort(lvf[sl] * 1e6 + np.arange(sl.stop - sl.start), kind="stable")]
        rr = ff.reshape(B, 1000, 4)
    tr = t(rr)
    out[name] = timed(lambda: F.roialign_forward(tr, d["image_meta"], maps, (7, 7), map_mode=1))
    print(f"{name:36s} ROIAlign 7x7: {out[name]:7.1f} us", flush=True)
