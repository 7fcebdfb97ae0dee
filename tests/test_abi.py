"""The C-ABI library: loads, exports exactly the symbols include/mrcnn_roi_b200.h declares, and rejects bad
arguments with the documented negative codes before touching the device (no compute without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrcnn_roi_b200.h")


@pytest.fixture(scope="module")
def lib():
    from maskrcnn_tf2_b200 import _lib
    _lib.build()
    return _lib.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrcnn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from maskrcnn_tf2_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == names       # the ctypes table mirrors the header one-to-one
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r" T (mrcnn_[a-z0-9_]+)", out)))
    assert exported == names                      # nothing else leaks out of the library


def test_library_is_sm100a_only_and_torch_free():
    from maskrcnn_tf2_b200 import _lib
    out = subprocess.check_output(["cuobjdump", "-lelf", _lib.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}
    ldd = subprocess.check_output(["ldd", _lib.LIB_PATH]).decode()
    assert "torch" not in ldd and "tensorflow" not in ldd and "liborc" not in ldd


def test_version_and_status_strings(lib):
    assert b"sm_100a" in lib.mrcnn_roi_b200_version()
    assert lib.mrcnn_status_string(0) == b"ok"
    assert b"NULL" in lib.mrcnn_status_string(-1)
    assert b"workspace" in lib.mrcnn_status_string(-3)


def test_workspace_queries_are_pure_host_functions(lib):
    n = ctypes.c_size_t(0)
    assert lib.mrcnn_proposal_workspace_bytes(8, 261888, 6000, 1000, ctypes.byref(n)) == 0
    # the sorted boxes (the top-k cluster kernel keeps its candidates on chip: no global scratch)
    assert n.value >= 8 * 6000 * 16
    assert lib.mrcnn_topk_workspace_bytes(2, 1000, 100, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.mrcnn_nms_workspace_bytes(2, 1000, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.mrcnn_detection_workspace_bytes(8, 1000, 81, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.mrcnn_detection_target_workspace_bytes(8, 2000, 100, 200, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.mrcnn_roialign_workspace_bytes(8, 1000, ctypes.byref(n)) == 0
    assert lib.mrcnn_proposal_workspace_bytes(8, 261888, 6000, 1000, None) == -1
    assert lib.mrcnn_proposal_workspace_bytes(8, 261888, 9000, 1000, ctypes.byref(n)) == -2    # K > 8192
    assert lib.mrcnn_topk_workspace_bytes(1, 10, 11, ctypes.byref(n)) == -2                      # K > A
    assert lib.mrcnn_nms_workspace_bytes(1, 8193, ctypes.byref(n)) == -2
    assert lib.mrcnn_detection_target_workspace_bytes(1, 2000, 2000, 200, ctypes.byref(n)) == -2  # G > 1024


def test_launchers_reject_bad_arguments_without_launching(lib):
    std = (ctypes.c_float * 4)(0.1, 0.1, 0.2, 0.2)
    fake = ctypes.c_void_p(0x1000)          # aligned, never dereferenced: validation fails first
    odd = ctypes.c_void_p(0x1004)
    assert lib.mrcnn_topk_forward(None, 1, 0, 1, 10, 5, fake, None, fake, 1 << 20, None) == -1
    assert lib.mrcnn_topk_forward(fake, 1, 0, 1, 10, 50, fake, None, fake, 1 << 20, None) == -2
    assert lib.mrcnn_topk_forward(fake, 2, 2, 1, 10, 5, fake, None, fake, 1 << 20, None) == -2   # offset >= stride
    assert lib.mrcnn_topk_forward(fake, 1, 0, 1, 10, 5, fake, None, fake, 16, None) == -3
    assert lib.mrcnn_nms_forward(fake, fake, None, 1, 100, 10, 1.5, fake, fake, fake, 1 << 30, None) == -2
    assert lib.mrcnn_nms_forward(odd, fake, None, 1, 100, 10, 0.5, fake, fake, fake, 1 << 30, None) == -4
    assert lib.mrcnn_proposal_forward(fake, fake, fake, 1, 1000, 6000, 100, std, 0.7, None, None, None, None, None,
                                      fake, 1 << 30, None) == -1
    assert lib.mrcnn_proposal_forward(fake, fake, fake, 1, 1000, 6000, 100, std, 0.7, fake, None, None, None, None,
                                      fake, 64, None) == -3
    assert lib.mrcnn_proposal_forward(fake, odd, fake, 1, 1000, 6000, 100, std, 0.7, fake, None, None, None, None,
                                      fake, 1 << 30, None) == -4
    assert lib.mrcnn_proposal_backward(fake, fake, fake, fake, None, 1, 1000, 600, 100, std, fake, None) == -1
    assert lib.mrcnn_proposal_backward(fake, fake, fake, fake, fake, 1, 100, 600, 100, std, fake, None) == -2   # K > A
    maps = (ctypes.c_void_p * 4)(0x1000, 0x2000, 0x3000, 0x4000)
    hw = (ctypes.c_int * 4)(8, 4, 2, 1)
    assert lib.mrcnn_roialign_forward(fake, fake, 93, maps, hw, hw, 255, 1, 10, 7, 7, 244.0, 0, fake, fake, None,
                                      fake, 1 << 20, None) == -2                                # C % 4 != 0
    assert lib.mrcnn_roialign_forward(fake, fake, 93, maps, hw, hw, 256, 1, 10, 7, 7, 244.0, 2, fake, fake, None,
                                      fake, 1 << 20, None) == -2                                # map_mode
    assert lib.mrcnn_roialign_forward(fake, fake, 93, maps, hw, hw, 256, 1, 10, 7, 7, 244.0, 0, fake, None, None,
                                      fake, 1 << 20, None) == -1                                # roi_map required
    assert lib.mrcnn_roialign_forward(fake, fake, 93, maps, hw, hw, 256, 1, 10, 7, 7, 244.0, 0, fake, fake, None,
                                      fake, 16, None) == -3                                     # workspace
    assert lib.mrcnn_roialign_backward(fake, fake, fake, maps, hw, hw, 256, 1, 10, 0, 7, None, 0, None) == -2
    assert lib.mrcnn_roialign_backward(fake, fake, fake, maps, hw, hw, 256, 1, 10, 7, 7, fake, 64, None) == -3  # workspace
    need = ctypes.c_size_t(0)
    assert lib.mrcnn_roialign_backward_workspace_bytes(8, 200, 7, 7, hw, hw, 256, ctypes.byref(need)) == 0
    assert need.value >= 4 * 8 * 200 * 49 * 4 + 3 * 8 * 85 * 4
    assert lib.mrcnn_roialign_backward_workspace_bytes(8, 200, 7, 7, hw, hw, 256, None) == -1
    assert lib.mrcnn_detection_forward(fake, fake, fake, fake, 93, 1, 1000, 81, std, 0.7, 1, 100, 0.3, 1, fake, None,
                                       None, fake, 1 << 30, None) == -2                         # per_class unsupported
    assert lib.mrcnn_detection_forward(fake, fake, fake, fake, 93, 1, 9000, 81, std, 0.7, 1, 100, 0.3, 0, fake, None,
                                       None, fake, 1 << 30, None) == -2
    assert lib.mrcnn_detection_forward(fake, fake, fake, fake, 93, 1, 1000, 81, std, 0.7, 1, 100, 0.3, 0, fake, None,
                                       odd, fake, 1 << 30, None) == -4                          # det_boxes alignment
    assert lib.mrcnn_detection_target_forward(fake, fake, fake, fake, fake, 1, 2000, 100, 64, 64, 200, 0.0, std, 28,
                                              28, 0, fake, fake, fake, fake, None, fake, 1 << 30, None) == -2
    assert lib.mrcnn_detection_target_forward(fake, fake, fake, fake, None, 1, 2000, 100, 64, 64, 200, 0.33, std, 28,
                                              28, 0, fake, fake, fake, fake, None, fake, 1 << 30, None) == -1
