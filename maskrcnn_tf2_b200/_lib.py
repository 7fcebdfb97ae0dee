"""ctypes binding of libmrcnn_roi_b200.so (include/mrcnn_roi_b200.h).

There is no CPU fallback: if the library is missing or a launcher returns non-zero, this module raises.
torch is used only for device memory, streams and workspace allocation.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrcnn_roi_b200.so")
CSRC = os.path.join(_HERE, "csrc")
_lib = None

c_int, c_float, c_double, c_void_p, c_size_t = (ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p,
                                                ctypes.c_size_t)

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/mrcnn_roi_b200.h exactly
PROTOTYPES = {
    "mrcnn_roi_b200_version": [],
    "mrcnn_status_string": [c_int],
    "mrcnn_topk_workspace_bytes": [c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_topk_forward": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                           c_void_p],
    "mrcnn_nms_workspace_bytes": [c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_nms_forward": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                          c_size_t, c_void_p],
    "mrcnn_proposal_workspace_bytes": [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_proposal_forward": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_float),
                               c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p],
    "mrcnn_proposal_levels_workspace_bytes": [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_proposal_forward_levels": [ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_int), c_int,
                                      c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_float), c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "mrcnn_proposal_backward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                ctypes.POINTER(c_float), c_void_p, c_void_p],
    "mrcnn_roialign_workspace_bytes": [c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_roialign_forward": [c_void_p, c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int),
                               ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "mrcnn_roialign_resident_words": [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_size_t)],
    "mrcnn_roialign_fetch_workspace_bytes": [c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                             ctypes.POINTER(c_size_t)],
    "mrcnn_roialign_fetch_hostmaps": [c_void_p, c_void_p, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                      ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "mrcnn_roialign_backward_workspace_bytes": [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int),
                                                ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_roialign_backward": [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int),
                                ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p],
    "mrcnn_detection_workspace_bytes": [c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_detection_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                ctypes.POINTER(c_float), c_float, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p],
    "mrcnn_detection_target_workspace_bytes": [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_detection_target_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, c_double, ctypes.POINTER(c_float), c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "mrcnn_anchors_count": [ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_int, c_int, ctypes.POINTER(c_int)],
    "mrcnn_anchors_forward": [ctypes.POINTER(c_double), ctypes.POINTER(c_double), ctypes.POINTER(c_int),
                              ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p],
    "mrcnn_rpn_targets_workspace_bytes": [c_int, c_int, c_int, c_int, ctypes.POINTER(c_size_t)],
    "mrcnn_rpn_targets_forward": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  ctypes.POINTER(c_double), c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p],
    "mrcnn_test_expf": [c_void_p, c_void_p, c_int, c_void_p],
    "mrcnn_test_logf": [c_void_p, c_void_p, c_int, c_void_p],
}
_RESTYPE = {"mrcnn_roi_b200_version": ctypes.c_char_p, "mrcnn_status_string": ctypes.c_char_p}


class MrcnnError(RuntimeError):
    pass


def build(force=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (csrc/Makefile), in-tree."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "mrcnn_roi_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MrcnnError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback for the ROI-stage kernels)")
        L = ctypes.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPE.get(name, c_int)
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().mrcnn_status_string(rc).decode()
        raise MrcnnError(f"{what} failed with status {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def float4(values):
    arr = (c_float * 4)(*[float(v) for v in values])
    return arr
