"""A stand-in for the few `tensorflow` Python names maskrcnn_tf2_b200/tf_shim/mrcnn_layers_b200.py uses, over torch CUDA
tensors, so that the Python side of the shim can be EXECUTED (not just parsed) where TensorFlow cannot be installed:
`tf.load_op_library` returns an object whose `mrcnn_<op>(*inputs, **attrs)` methods run the C++ OpKernels of
tf_shim/mrcnn_roi_ops.cc through tests/tf_stub (tf_stub.StubOp).  TEST INFRASTRUCTURE ONLY.

    with fake_tf.installed() as tf:
        shim = fake_tf.import_shim()          # the real mrcnn_layers_b200.py, executed against this module
        rois = shim.ProposalLayer(1000, cfg)([rpn_probs, rpn_bbox, anchors])
"""
import contextlib
import importlib.util
import os
import sys
import types

import torch

from . import ROOT, StubOp, signatures

PY_SHIM = os.path.join(ROOT, "maskrcnn_tf2_b200", "tf_shim", "mrcnn_layers_b200.py")


class _DType:
    def __init__(self, torch_dtype):
        self.torch = torch_dtype
        if not torch_dtype.is_floating_point and torch_dtype != torch.bool:
            self.min, self.max = torch.iinfo(torch_dtype).min, torch.iinfo(torch_dtype).max


class _OpLibrary:
    """What tf.load_op_library returns: one snake_case callable per registered op."""

    def __init__(self):
        self._sig = signatures()
        self.calls = []          # (op, attrs) of every call, for the tests

    def __getattr__(self, snake):
        op = "".join(p.capitalize() for p in snake.split("_"))
        if op not in self._sig:
            raise AttributeError(f"no op {op} in the library")
        sig = self._sig[op]

        def run(*inputs, **attrs):
            flat = []
            if len(inputs) != len(sig["inputs"]):
                raise TypeError(f"{op} takes {len(sig['inputs'])} inputs, got {len(inputs)}")
            for value, (_, _, number_attr) in zip(inputs, sig["inputs"]):
                if number_attr:                      # "N * float": a list input; TF infers N from its length
                    value = list(value)
                    if attrs.setdefault(number_attr, len(value)) != len(value):
                        raise ValueError(f"{op}: list inputs disagree on {number_attr}")
                    flat.extend(value)
                else:
                    flat.append(value)
            unknown = set(attrs) - set(sig["attrs"])
            if unknown:
                raise TypeError(f"{op} has no attr {sorted(unknown)}")
            self.calls.append((op, dict(attrs)))
            outs = StubOp(op, **attrs)(*flat)
            return outs[0] if len(outs) == 1 else tuple(outs)
        return run


class _Layer:
    """tf.keras.layers.Layer as far as the shim's classes rely on it."""

    def __init__(self, name=None, **kwargs):
        if kwargs:
            raise TypeError(f"unexpected Layer kwargs {sorted(kwargs)}")
        self.name = name
        self.built = False

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, **kwargs):
        if not self.built:
            self.build(None)
        return self.call(inputs, **kwargs)

    def get_config(self):
        return {"name": self.name, "trainable": True, "dtype": "float32"}


def make_module():
    tf = types.ModuleType("tensorflow")
    tf.__fake__ = True
    tf.int32, tf.float32, tf.float64, tf.bool = (_DType(torch.int32), _DType(torch.float32), _DType(torch.float64),
                                                 _DType(torch.bool))
    tf.gradients = {}        # op name -> python gradient function (tf.RegisterGradient)
    tf.no_gradients = set()
    tf.serializable = []
    tf.library = None

    def load_op_library(path):
        assert os.path.basename(path) == "libmrcnn_roi_ops.so", path
        tf.library = _OpLibrary()
        return tf.library

    def register_gradient(op):
        def deco(fn):
            tf.gradients[op] = fn
            return fn
        return deco

    def uniform(shape, minval=0, maxval=None, dtype=None, seed=None):
        shape = tuple(int(s) for s in shape)
        dtype = dtype or tf.float32
        dev = torch.device("cuda", torch.cuda.current_device())
        if dtype.torch.is_floating_point:
            hi = 1.0 if maxval is None else maxval
            return torch.rand(shape, dtype=dtype.torch, device=dev) * (hi - minval) + minval
        return torch.randint(int(minval), int(maxval), shape, dtype=torch.int64, device=dev).to(dtype.torch)

    tf.load_op_library = load_op_library
    tf.RegisterGradient = register_gradient
    tf.no_gradient = tf.no_gradients.add
    tf.cast = lambda x, dtype: x.to(dtype.torch)
    tf.shape = lambda x: tuple(x.shape)
    tf.reshape = lambda x, shape: x.reshape([int(s) for s in shape])
    tf.stack = lambda values: tuple(int(v) for v in values)
    tf.stop_gradient = lambda x: x.detach()
    tf.random = types.SimpleNamespace(uniform=uniform)
    keras = types.ModuleType("tensorflow.keras")
    keras.layers = types.ModuleType("tensorflow.keras.layers")
    keras.layers.Layer = _Layer

    def register_keras_serializable(package="Custom", name=None):
        def deco(cls):
            tf.serializable.append(cls.__name__)
            return cls
        return deco

    keras.utils = types.SimpleNamespace(register_keras_serializable=register_keras_serializable)
    tf.keras = keras
    return tf


@contextlib.contextmanager
def installed():
    """Puts the stand-in into sys.modules as `tensorflow` (+ .keras, .keras.layers) for the duration."""
    names = ["tensorflow", "tensorflow.keras", "tensorflow.keras.layers"]
    saved = {n: sys.modules.get(n) for n in names}
    tf = make_module()
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": tf.keras, "tensorflow.keras.layers": tf.keras.layers})
    try:
        yield tf
    finally:
        for n in names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]


def import_shim():
    """Executes the real tf_shim/mrcnn_layers_b200.py against whatever `tensorflow` is in sys.modules."""
    spec = importlib.util.spec_from_file_location("mrcnn_layers_b200_under_fake_tf", PY_SHIM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class FakeOp:
    """The `op` argument TF hands to a registered gradient function."""

    def __init__(self, inputs, outputs, attrs):
        self.inputs, self.outputs, self._attrs = list(inputs), list(outputs), dict(attrs)

    def get_attr(self, name):
        return self._attrs[name]
