"""maskrcnn_tf2_b200 -- B200-native (sm_100a) ROI stage of the TF2 Mask R-CNN in miguelalejo/maskrcnn_tf2.

Layout: csrc/ (CUDA kernels + extern "C" launchers, built into libmrcnn_roi_b200.so), _lib.py (ctypes binding),
functional.py (tensor-level entry points), layers.py (the reference's Keras layer API: ProposalLayer,
PyramidROIAlign, DetectionLayer, DetectionTargetLayer), synth.py (COCO-shape synthetic inputs), tf_shim/ (the
TensorFlow custom-op shim sources; TensorFlow is not installable in this image).
"""
from .config import CONFIG, make_config  # noqa: F401

__all__ = ["CONFIG", "make_config"]
__version__ = "0.1.0"
