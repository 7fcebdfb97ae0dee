"""Per-phase cycle counters of nms_lazy_kernel (debug build `make -C maskrcnn_tf2_b200/csrc prof`)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maskrcnn_tf2_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get("MRCNN_PROF_LIB", "libmrcnn_roi_b200_prof.so"))
from maskrcnn_tf2_b200 import functional as F, synth
L = _lib.lib()
L.mrcnn_debug_nms_profile.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
regime = sys.argv[2] if len(sys.argv) > 2 else "clustered"
a = synth.pyramid_anchors(1024)
pr, bb = zip(*[synth.rpn_outputs(np.random.default_rng(2000 + b), a, regime, 1024) for b in range(B)])
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
probs, bbox, anch = t(np.stack(pr)), t(np.stack(bb)), t(np.broadcast_to(a, (B,) + a.shape))
for _ in range(3):
    out = F.proposal_forward(probs, bbox, anch, 6000, 1000, [0.1, 0.1, 0.2, 0.2], 0.7, debug=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
L.mrcnn_debug_nms_profile(buf)
v = list(buf)
tiles = max(v[4], 1)
print(f"B={B} {regime}: tiles {v[4]} kept {v[5]}  cycles/tile: wait far {v[0]/tiles:.0f}  resolve {v[1]/tiles:.0f}  "
      f"total {sum(v[:4])/tiles:.0f}")
w = v[8:13]
print("  far warp 1 of CTA 0, cycles/tile: prologue %.0f  wait release %.0f  second instalment + vote %.0f  send %.0f  "
      "first instalment %.0f  total %.0f; exact-division fallbacks: %d"
      % tuple([x / tiles for x in w] + [sum(w) / tiles, v[14]]))

tl = (ctypes.c_longlong * (128 * 8))()
L.mrcnn_debug_nms_timeline.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
L.mrcnn_debug_nms_timeline(tl)
tl = np.array(list(tl), dtype=np.int64).reshape(128, 8)
t0 = tl[20, 0]
print("timeline of CTA 0 (cycles since resolve(20) started); columns: resolver start | far set arrived | kept known | released "
      "|| far warp 0: loop top | release seen | sent || row warp 0: rows sent")
for t in range(20, 32):
    r = tl[t] - t0
    print(f"  tile {t}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} || {r[7]:7d} {r[4]:7d} {r[5]:7d} || {r[6]:7d}")
