"""Gradients of the reference's own layer code, by autograd.

The reference gets its ROI-stage gradients from TensorFlow's autodiff of the layer code: d(proposals)/d(rpn_bbox)
through gather -> std-dev scale -> decode -> clip -> NMS gather (no stop_gradient), and
d(pooled)/d(feature maps) through crop_and_resize with the boxes stopped (mrcnn_layers.py:628-629).  Here the same
Python -- ProposalLayer.call and PyramidROIAlign.call of /root/reference, unmodified -- runs on a TORCH-backed stand-in
for the tf.* functions it uses, torch's autograd differentiates it, and the result is compared with the oracle's
hand-written gradients (oracle.proposal_layer_grad, oracle.pyramid_roi_align_grad), which the CUDA kernels are in turn
compared with on the GPU.  Live only where /root/reference exists (this container).  Tolerance 1e-5 rel / 1e-6 abs of
the accumulated magnitude: autograd's summation order is not TF's.  exp / log call the oracle's routines in the
forward pass so that NMS decisions cannot flip on a last-bit difference."""
import importlib.util
import os
import types

import numpy as np
import pytest
import torch

from conftest import random_boxes

pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/src/layers/mrcnn_layers.py"),
                                reason="needs the reference checkout (not present on the GPU box)")
SD = np.array([0.1, 0.1, 0.2, 0.2], np.float32)


class RT(torch.Tensor):
    """torch.Tensor with TF's value semantics where the reference relies on them: numpy operands, `x op= y`
    rebinding instead of mutating, reversed slices, set_shape."""

    def set_shape(self, shape):
        pass

    def __getitem__(self, key):
        if isinstance(key, slice) and key.step == -1 and key.start is None and key.stop is None:
            return torch.flip(self, [0])
        return super().__getitem__(key)


def _conv(other):
    return torch.as_tensor(other) if isinstance(other, np.ndarray) else other


for _name in ("__mul__", "__rmul__", "__add__", "__radd__", "__sub__", "__rsub__", "__truediv__", "__rtruediv__",
              "__ge__", "__gt__", "__lt__", "__le__"):
    setattr(RT, _name, (lambda base: lambda self, other: base(self, _conv(other)))(getattr(torch.Tensor, _name)))
for _name, _op in (("__imul__", "__mul__"), ("__iadd__", "__add__"), ("__isub__", "__sub__"), ("__itruediv__", "__truediv__")):
    setattr(RT, _name, (lambda op: lambda self, other: getattr(self, op)(other))(_op))     # tensors are immutable in TF


def rt(x, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x if dtype is None else x.to(dtype)
    else:
        t = torch.as_tensor(np.asarray(x), dtype=dtype)
    return t.as_subclass(RT)


class _OracleUnary(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, which):
        import oracle
        fn = oracle.expf if which == "exp" else oracle.logf
        y = torch.from_numpy(fn(x.detach().numpy().astype(np.float32)).reshape(tuple(x.shape)))
        ctx.which = which
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return (g * y if ctx.which == "exp" else g / x), None


def make_torch_tf(numpy_tf):
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool
    tf.newaxis = None
    is_t = lambda v: isinstance(v, torch.Tensor)

    def unary(which):
        return lambda x, **k: rt(_OracleUnary.apply(rt(x, torch.float32) if not is_t(x) else x, which))
    tf.exp = unary("exp")

    def minmax(fn, pyfn):
        def run(a, b, **k):
            if not is_t(a) and not is_t(b):
                return pyfn(int(a), int(b))
            a = a if is_t(a) else torch.as_tensor(a, dtype=b.dtype)
            b = b if is_t(b) else torch.as_tensor(b, dtype=a.dtype)
            return rt(fn(a, b))
        return run
    tf.maximum, tf.minimum = minmax(torch.maximum, max), minmax(torch.minimum, min)
    tf.math = types.SimpleNamespace(log=unary("log"), maximum=tf.maximum, minimum=tf.minimum)
    tf.sqrt = lambda x, **k: rt(torch.sqrt(x))
    tf.round = lambda x, **k: rt(torch.round(x))
    tf.equal = lambda a, b, **k: rt(torch.eq(a, b))
    tf.cast = lambda x, dtype, **k: rt(x if is_t(x) else torch.as_tensor(x)).to(dtype).as_subclass(RT)
    tf.constant = lambda v, dtype=None, **k: rt(v, dtype)
    tf.stop_gradient = lambda x, **k: rt(x.detach())
    tf.identity = lambda x, **k: x
    tf.shape = lambda x, **k: np.asarray(tuple(x.shape), np.int64)
    tf.reshape = lambda x, shape, **k: rt(x.reshape(tuple(int(s) for s in np.asarray(shape).reshape(-1))))
    tf.squeeze = lambda x, axis=None, **k: rt(x.squeeze(axis))
    tf.expand_dims = lambda x, axis, **k: rt(x.unsqueeze(axis))
    tf.stack = lambda xs, axis=0, **k: rt(torch.stack(list(xs), dim=axis))
    tf.range = lambda *a, **k: rt(torch.arange(*[int(v) for v in a], dtype=torch.int32))

    def concat(xs, axis=0, **k):
        if not any(is_t(x) for x in xs):
            return np.concatenate([np.asarray(x) for x in xs], axis=axis)
        return rt(torch.cat([x if is_t(x) else torch.as_tensor(x) for x in xs], dim=axis))
    tf.concat = concat

    def split(x, n, axis=0, **k):
        x = x if is_t(x) else rt(x)
        return [rt(p) for p in torch.split(x, x.shape[axis] // int(n), dim=axis)]
    tf.split = split

    def pad(x, paddings, mode="CONSTANT", constant_values=0, **k):
        p = [(int(a), int(b)) for a, b in np.asarray(paddings.detach() if is_t(paddings) else paddings).reshape(-1, 2)]
        flat = [v for ab in reversed(p) for v in ab]                      # torch pads the last dimension first
        return rt(torch.nn.functional.pad(x, flat, value=constant_values))
    tf.pad = pad
    tf.gather = lambda params, indices, axis=0, **k: rt(torch.index_select(
        params, axis, (indices if is_t(indices) else torch.as_tensor(np.asarray(indices))).to(torch.int64).reshape(-1)))
    tf.gather_nd = lambda params, idx, **k: rt(params[tuple(idx[..., d].to(torch.int64) for d in range(idx.shape[-1]))])
    tf.where = lambda cond, **k: rt(torch.nonzero(cond))

    def unique(x, **k):
        vals, idx = numpy_tf.unique(x.detach().numpy())
        return rt(np.asarray(vals)), rt(np.asarray(idx))
    tf.unique = unique

    class TopK(tuple):
        values = property(lambda s: s[0])
        indices = property(lambda s: s[1])

    def top_k(x, k=1, sorted=True, **kw):
        idx = torch.from_numpy(np.asarray(numpy_tf.nn.top_k(x.detach().numpy(), int(k)).indices).astype(np.int64))
        return TopK((rt(torch.gather(x, -1, idx)), rt(idx.to(torch.int32))))
    tf.nn = types.SimpleNamespace(top_k=top_k)

    def nms(boxes, scores, max_output_size, iou_threshold=0.5, **k):
        keep = numpy_tf.image.non_max_suppression(boxes.detach().numpy(), scores.detach().numpy(), int(max_output_size),
                                                  float(iou_threshold))
        return rt(np.asarray(keep))

    def crop_and_resize(image, boxes, box_indices, crop_size, method="bilinear", **k):
        """Differentiable w.r.t. `image` only (the reference stops the boxes): same taps as CropAndResize."""
        ph, pw = int(crop_size[0]), int(crop_size[1])
        _, H, W, C = image.shape
        b = boxes.detach().numpy().astype(np.float32)
        out = []
        f32 = np.float32

        def taps(c1, c2, size, crop):
            if crop > 1:
                pos = c1 * f32(size - 1) + np.arange(crop, dtype=f32) * ((c2 - c1) * f32(size - 1) / f32(crop - 1))
            else:
                pos = np.asarray([f32(0.5) * (c1 + c2) * f32(size - 1)], f32)
            ok = (pos >= 0) & (pos <= f32(size - 1))
            safe = np.where(ok, pos, f32(0))
            lo, hi = np.floor(safe).astype(np.int64), np.ceil(safe).astype(np.int64)
            return torch.from_numpy(ok), torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(safe - lo.astype(f32))

        for n in range(b.shape[0]):
            img = image[int(box_indices[n])]
            oky, ylo, yhi, ly = taps(b[n, 0], b[n, 2], H, ph)
            okx, xlo, xhi, lx = taps(b[n, 1], b[n, 3], W, pw)
            tl, tr = img[ylo][:, xlo], img[ylo][:, xhi]
            bl, br = img[yhi][:, xlo], img[yhi][:, xhi]
            top = tl + (tr - tl) * lx[None, :, None]
            bot = bl + (br - bl) * lx[None, :, None]
            val = top + (bot - top) * ly[:, None, None]
            out.append(val * (oky[:, None] & okx[None, :])[..., None].to(val.dtype))
        if not out:
            return rt(torch.zeros((0, ph, pw, C)))
        return rt(torch.stack(out))
    tf.image = types.SimpleNamespace(non_max_suppression=nms, crop_and_resize=crop_and_resize)
    tf.function = lambda fn=None, **k: fn if fn is not None else (lambda f: f)
    tf.keras = numpy_tf.keras                                             # the Layer base class and decorators
    return tf


@pytest.fixture(scope="module")
def ref_torch():
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_reference_layers_golden.py")
    spec = importlib.util.spec_from_file_location("make_reference_layers_golden", path)
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    return gen.load_reference_layers(make_torch_tf(gen.make_numpy_tf()))


@pytest.mark.parametrize("seed", range(4))
def test_proposal_gradient_is_the_autodiff_of_the_reference_layer(orc, ref_torch, seed):
    from maskrcnn_tf2_b200 import synth
    L = ref_torch
    rng = np.random.default_rng(9300 + seed)
    S, B, K, P = 64, 2, int(rng.choice([100, 400])), int(rng.choice([30, 200]))
    anchors1 = synth.pyramid_anchors(S)
    A = anchors1.shape[0]
    probs, bbox = zip(*[synth.rpn_outputs(np.random.default_rng(seed * 7 + b), anchors1, "clustered", S) for b in range(B)])
    probs, bbox = np.stack(probs).astype(np.float32), (np.stack(bbox) * (1 + 2 * (seed % 2))).astype(np.float32)
    anchors = np.ascontiguousarray(np.broadcast_to(anchors1, (B, A, 4))).astype(np.float32)
    cfg = {"rpn_nms_threshold": 0.7, "pre_nms_limit": K, "images_per_gpu": B, "rpn_bbox_std_dev": SD}
    tb = rt(bbox).requires_grad_(True)
    rois = L.ProposalLayer(proposal_count=P, config=cfg)([rt(probs), tb, rt(anchors)])
    r = orc.proposal_layer(probs, bbox, anchors, K, P, SD, 0.7)
    assert np.array_equal(rois.detach().numpy(), r["proposals"])          # forward: bit-exact on this stand-in too
    g = rng.standard_normal((B, P, 4)).astype(np.float32)
    (rois * torch.from_numpy(g)).sum().backward()
    want = orc.proposal_layer_grad(g, bbox, anchors, r["topk_idx"], r["keep_idx"], SD)
    got = tb.grad.numpy()
    assert np.array_equal(got != 0, want != 0)                            # the same rows receive gradient (clip pass-through rules)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert np.count_nonzero(want) > 0


@pytest.mark.parametrize("seed", range(4))
def test_roialign_gradient_is_the_autodiff_of_the_reference_layer(orc, ref_torch, seed):
    from maskrcnn_tf2_b200 import synth
    L = ref_torch
    rng = np.random.default_rng(9400 + seed)
    S, B, N, C = 128, 2, int(rng.integers(5, 40)), 3
    pool = (int(rng.choice([1, 3, 7])), int(rng.choice([2, 7])))
    boxes = np.stack([random_boxes(rng, N, min_size=0.03, max_size=0.9) for _ in range(B)])
    boxes[:, -2:] = 0                                                      # zero-padded ROIs: pixel (0,0) of "level 2"
    boxes[:, 1] += rng.uniform(-0.3, 0.3, 4).astype(np.float32)            # partly outside
    meta = synth.image_meta(B, S, 5).astype(np.float32)
    fm = [rng.standard_normal((B, S // s, S // s, C)).astype(np.float32) for s in (4, 8, 16, 32)]
    tfm = [rt(f).requires_grad_(True) for f in fm]
    tbx = rt(boxes).requires_grad_(True)
    pooled = L.PyramidROIAlign(list(pool))([tbx, rt(meta)] + tfm)
    want_fwd = orc.pyramid_roi_align(boxes, float(S), float(S), fm, pool)["out"]
    assert np.array_equal(pooled.detach().numpy(), want_fwd)
    g = rng.standard_normal(want_fwd.shape).astype(np.float32)
    (pooled * torch.from_numpy(g)).sum().backward()
    assert tbx.grad is None or not tbx.grad.any()                          # boxes are stopped (L:628-629)
    want = orc.pyramid_roi_align_grad(g, boxes, float(S), float(S), [f.shape for f in fm])
    mag = orc.pyramid_roi_align_grad(np.abs(g), boxes, float(S), float(S), [f.shape for f in fm])
    for l in range(4):
        got = tfm[l].grad.numpy() if tfm[l].grad is not None else np.zeros_like(fm[l])
        assert np.all(np.abs(got - want[l]) <= 1e-6 + 1e-5 * mag[l]), l
    assert sum(np.count_nonzero(w) for w in want) > 0
