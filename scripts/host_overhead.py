"""Host-side cost of each layer call (no device sync inside the timed calls) + CUDA-graph replay of the stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import make_config, synth, functional as F
from maskrcnn_tf2_b200.layers import DetectionLayer, ProposalLayer, PyramidROIAlign

dev = torch.device("cuda:0")
B, S, NC = 8, 1024, 81
cfg = make_config(img_size=S, num_classes=NC, batch_size=B)
x = synth.inference_batch(2, B, img_size=S, num_classes=NC, regime="clustered")
d = {k: torch.from_numpy(v).to(dev) for k, v in x.items() if k != "feature_maps"}
maps = [torch.from_numpy(f).to(dev) for f in x["feature_maps"]]
proposal = ProposalLayer(1000, cfg); a7 = PyramidROIAlign([7, 7]); a14 = PyramidROIAlign([14, 14])
det = DetectionLayer(1000, 0.7, 100, 0.3, cfg["bbox_std_dev"], B, B)

def stage(timing=None):
    t0 = time.perf_counter()
    rois = proposal([d["rpn_probs"], d["rpn_bbox"], d["anchors"]]); t1 = time.perf_counter()
    p7 = a7([rois, d["image_meta"]] + maps); t2 = time.perf_counter()
    dt = det([rois, d["mrcnn_class"], d["mrcnn_bbox"], d["image_meta"]]); t3 = time.perf_counter()
    p14 = a14([dt[..., :4].contiguous(), d["image_meta"]] + maps); t4 = time.perf_counter()
    if timing is not None:
        timing.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
    return rois, p7, dt, p14

for _ in range(5): stage()
torch.cuda.synchronize()
tm = []
for _ in range(20):
    stage(tm); torch.cuda.synchronize()
tm = np.array(tm) * 1e6
print("host us per call (median): proposal %.0f  align7 %.0f  detection %.0f  align14 %.0f" % tuple(np.median(tm, 0)))

# individual pieces
t0 = time.perf_counter()
for _ in range(100): torch.empty((B, 1000, 7, 7, 256), device=dev)
print("torch.empty 401MB: %.1f us" % ((time.perf_counter() - t0) * 1e4))
t0 = time.perf_counter()
for _ in range(100): F._map_args(maps)
print("_map_args: %.1f us" % ((time.perf_counter() - t0) * 1e4))

# eager loop, device timing
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): stage()
e1.record(); torch.cuda.synchronize()
print("eager: %.3f ms/step" % (e0.elapsed_time(e1) / 20))

# CUDA graph replay
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): stage()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    outs = stage()
g.replay(); torch.cuda.synchronize()
e0.record()
for _ in range(20): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph: %.3f ms/step" % (e0.elapsed_time(e1) / 20))
ref = stage()
torch.cuda.synchronize()
print("graph outputs equal eager:", all(torch.equal(a, b) for a, b in zip(outs, ref)))
