"""Builds maskrcnn_tf2_b200/tf_shim/mrcnn_roi_ops.cc against the TF-API stand-in of this directory and drives the
resulting library through ctypes (test infrastructure; see tf_stub.h)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(_HERE))
SHIM = os.path.join(ROOT, "maskrcnn_tf2_b200", "tf_shim", "mrcnn_roi_ops.cc")
LIB_DIR = os.path.join(ROOT, "maskrcnn_tf2_b200")
OUT = os.path.join(_HERE, "_build", "libmrcnn_roi_ops_stub.so")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
_lib = None

DT = {"float": 1, "double": 2, "int32": 3, "uint8": 4, "bool": 10}


def build(force=False):
    srcs = [SHIM, os.path.join(_HERE, "tf_stub.h"), os.path.join(_HERE, "tf_stub_harness.cc"),
            os.path.join(ROOT, "include", "mrcnn_roi_b200.h")]
    stale = not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs)
    if force or stale:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([
            "g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-Wall", "-Wextra", "-Werror", "-DGOOGLE_CUDA=1",
            "-I" + _HERE, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CUDA, "include"),
            SHIM, os.path.join(_HERE, "tf_stub_harness.cc"), "-o", OUT,
            "-L" + LIB_DIR, "-lmrcnn_roi_b200", "-Wl,-rpath," + LIB_DIR,
            "-L" + os.path.join(CUDA, "lib64"), "-lcudart"])
    return OUT


def lib():
    global _lib
    if _lib is None:
        build()
        import torch  # noqa: F401  (brings libcudart.so.12 into the process before the stub library asks for it)
        L = ctypes.CDLL(OUT)
        L.tfstub_kernel_create.restype = ctypes.c_void_p
        L.tfstub_kernel_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.tfstub_kernel_compute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                            ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_longlong), ctypes.c_void_p, ctypes.c_char_p,
                                            ctypes.c_int]
        L.tfstub_output_info.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                         ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong), ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_void_p)]
        L.tfstub_output_copy.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_void_p]
        L.tfstub_kernel_destroy.argtypes = [ctypes.c_void_p]
        L.tfstub_infer_shapes.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_longlong), ctypes.c_int,
                                          ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong), ctypes.c_int,
                                          ctypes.c_char_p, ctypes.c_int]
        _lib = L
    return _lib


def _attr_text(attrs):
    def one(v):
        if isinstance(v, bool):
            return "true" if v else "false"
        if isinstance(v, (list, tuple, np.ndarray)):
            return ",".join(repr(float(x)) for x in v)
        return repr(v) if isinstance(v, float) else str(v)
    return ";".join(f"{k}={one(v)}" for k, v in attrs.items()).encode()


def signatures():
    """{op: dict(inputs=[(name, type, number_attr)], outputs=[(name, type)], attrs={name: (type, default|None)},
    devices=[...])} as registered by the shim's REGISTER_OP / REGISTER_KERNEL_BUILDER statements."""
    L = lib()
    buf = ctypes.create_string_buffer(8192)
    out = {}
    for i in range(L.tfstub_num_ops()):
        L.tfstub_op_signature(i, buf, len(buf))
        name, ins, outs, attrs = buf.value.decode().split("|")
        L.tfstub_kernel_devices(name.encode(), buf, len(buf))
        devices = [d for d in buf.value.decode().split(",") if d]
        parse_in = []
        for a in filter(None, ins.split(",")):
            n, t = a.split(":")
            t, _, num = t.partition("*")
            parse_in.append((n, t, num or None))
        parse_attr = {}
        for a in filter(None, attrs.split(";")):
            n, t = a.split(":", 1)
            t, eq, d = t.partition("=")
            parse_attr[n] = (t, d if eq else None)
        out[name] = dict(inputs=parse_in, outputs=[tuple(a.split(":")) for a in filter(None, outs.split(","))],
                         attrs=parse_attr, devices=devices)
    return out


def infer_shapes(op, in_shapes, **attrs):
    """Runs the op's shape function; returns the list of output shapes (-1 = unknown dimension)."""
    L = lib()
    nd = (ctypes.c_int * len(in_shapes))(*[len(s) for s in in_shapes])
    flat = [int(d) for s in in_shapes for d in s]
    dims = (ctypes.c_longlong * max(len(flat), 1))(*flat)
    ond = (ctypes.c_int * 8)()
    od = (ctypes.c_longlong * 64)()
    err = ctypes.create_string_buffer(512)
    n = L.tfstub_infer_shapes(op.encode(), _attr_text(attrs), len(in_shapes), nd, dims, 8, ond, od, 64, err, 512)
    if n < 0:
        raise ValueError(err.value.decode())
    shapes, off = [], 0
    for i in range(n):
        shapes.append(tuple(od[off:off + ond[i]]))
        off += ond[i]
    return shapes


class StubOp:
    """One OpKernel instance of the shim (GPU registration), constructed from attributes as TF would."""

    def __init__(self, op, **attrs):
        self.L = lib()
        self.op = op
        self.sig = signatures()[op]
        err = ctypes.create_string_buffer(512)
        self.h = self.L.tfstub_kernel_create(op.encode(), _attr_text(attrs), err, 512)
        if not self.h:
            raise ValueError(err.value.decode())

    def __call__(self, *tensors):
        """tensors: torch CUDA tensors (the op's flat input list).  Returns the outputs as torch tensors."""
        import torch
        tdt = {torch.float32: 1, torch.float64: 2, torch.int32: 3, torch.uint8: 4, torch.bool: 10}
        back = {1: torch.float32, 2: torch.float64, 3: torch.int32, 4: torch.uint8, 10: torch.bool}
        ts = [t.contiguous() for t in tensors]
        n = len(ts)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        dts = (ctypes.c_int * n)(*[tdt[t.dtype] for t in ts])
        nds = (ctypes.c_int * n)(*[t.dim() for t in ts])
        flat = [int(d) for t in ts for d in t.shape]
        dims = (ctypes.c_longlong * max(len(flat), 1))(*flat)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        err = ctypes.create_string_buffer(1024)
        rc = self.L.tfstub_kernel_compute(self.h, n, ptrs, dts, nds, dims, stream, err, 1024)
        if rc != 0:
            raise RuntimeError(f"{self.op}: status {rc}: {err.value.decode()}")
        outs = []
        for i in range(len(self.sig["outputs"])):
            dt, nd, od, p = ctypes.c_int(), ctypes.c_int(), (ctypes.c_longlong * 8)(), ctypes.c_void_p()
            assert self.L.tfstub_output_info(self.h, i, dt, nd, od, 8, p) == 0, f"output {i} was not allocated"
            o = torch.empty(tuple(od[:nd.value]), dtype=back[dt.value], device=ts[0].device)
            assert self.L.tfstub_output_copy(self.h, i, o.data_ptr(), o.numel() * o.element_size(), stream) == 0
            outs.append(o)
        torch.cuda.current_stream().synchronize()
        return outs

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tfstub_kernel_destroy(self.h)
            self.h = None
