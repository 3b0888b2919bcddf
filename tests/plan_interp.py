"""Test helpers (torch fp32 evaluation of planned ops):

  run_graph_cpu            interprets a plan.Graph op list on the CPU — arena offsets, channel slices,
                           packed KRSC weight blobs, merged convs — so the graph builder / memory planner /
                           weight packer are verified without a GPU.
  teacher_forced_errors    on the GPU: runs every op of a real engine ONE AT A TIME and compares its output
                           with a torch fp32 evaluation of the SAME device inputs (no error accumulation
                           across layers), in units of fp16 ulps.
"""
import numpy as np
import torch
import torch.nn.functional as F

import yolox_b200 as yb
from yolox_b200 import _capi

ACT = {0: "none", 1: "silu", 2: "hard_swish", 3: "relu", 4: "lrelu"}


def _act(x, code):
    from oracle.model_ref import activation
    return activation(x, ACT[code])


def ulp16(t):
    """Spacing of fp16 numbers at |t| (normal range; subnormals share 2^-24)."""
    e = torch.floor(torch.log2(torch.clamp(t.abs(), min=2.0 ** -14)))
    return torch.pow(2.0, e - 10)


def eval_op(o, image, src, res, wblob, bblob, q, want_band=False, up=None):
    """Evaluate one planned op in fp32.  o: _capi.Op; image NCHW fp32 (S2D only); src/res NHWC fp32;
    wblob fp32 view of the fp16 weight blob, bblob fp32.  q: rounding applied where the engine rounds to
    fp16 (identity for an exact evaluation).  Returns NHWC [n,h,w,dst.c]."""
    def s2d(x, unshuffle, padded):
        tl, tr, bl, br = x[..., ::2, ::2], x[..., ::2, 1::2], x[..., 1::2, ::2], x[..., 1::2, 1::2]
        if unshuffle:
            y = torch.stack((tl, tr, bl, br), dim=2).reshape(x.shape[0], 12, x.shape[2] // 2, x.shape[3] // 2)
        else:
            y = torch.cat((tl, bl, tr, br), dim=1)
        y = F.pad(y, (0, 0, 0, 0, 0, 4))
        if padded:                         # padded rows: [0 | pixels | 0 0 0]
            y = F.pad(y, (1, 3))
        return q(y.permute(0, 2, 3, 1))
    if o.kind == _capi.OP_S2D:
        return s2d(image, o.aux & 1, o.aux & 2)
    if o.kind == _capi.OP_CONV and (o.aux & 4):      # image-fed stem: the kernel does the (fp16-rounded) s2d itself
        src = s2d(image, o.aux & 8, False)
    if o.kind == _capi.OP_CONV and (o.aux & 1):
        # row-packed stem conv: pixel x of the GEMM sees buffer columns x, x+1, x+2 (= image pixels x-1, x, x+1)
        W = src.shape[2] - 4
        x48 = torch.cat([src[:, :, 0:W], src[:, :, 1:W + 1], src[:, :, 2:W + 2]], dim=3).permute(0, 3, 1, 2)
        w = wblob[o.w_offset // 2: o.w_offset // 2 + o.cout_pad * 3 * 48].view(o.cout_pad, 3, 1, 48).permute(0, 3, 1, 2)
        b = bblob[o.b_offset // 4: o.b_offset // 4 + o.cout_pad]
        y0 = q(F.conv2d(x48, w, b, stride=1, padding=(1, 0)))
        y = _act(y0, o.act)
        band = None
        if want_band:
            u = ulp16(y0)
            band = torch.maximum((_act(y0 + u, o.act) - y).abs(), (_act(y0 - u, o.act) - y).abs())
            band = band.permute(0, 2, 3, 1)[..., :o.dst.c]
        y = q(y.permute(0, 2, 3, 1)[..., :o.dst.c])
        return (y, band) if want_band else y
    if o.kind == _capi.OP_CONV:
        k = o.ksize
        x = src.permute(0, 3, 1, 2)
        if o.up.c > 0:  # fused torch.cat([nn.Upsample(2, "nearest")(up), src], 1)
            x = torch.cat([F.interpolate(up.permute(0, 3, 1, 2), scale_factor=2, mode="nearest"), x], 1)
        x = F.pad(x, (0, 0, 0, 0, 0, o.cin_pad - x.shape[1]))
        w = wblob[o.w_offset // 2: o.w_offset // 2 + o.cout_pad * k * k * o.cin_pad]
        w = w.view(o.cout_pad, k, k, o.cin_pad).permute(0, 3, 1, 2)
        b = bblob[o.b_offset // 4: o.b_offset // 4 + o.cout_pad]
        y0 = q(F.conv2d(x, w, b, stride=o.stride, padding=(k - 1) // 2))
        y = _act(y0, o.act)
        band = None
        if want_band:  # output change when the fp32 sum rounds to a NEIGHBOURING fp16 value (order of accumulation)
            u = ulp16(y0)
            band = torch.maximum((_act(y0 + u, o.act) - y).abs(), (_act(y0 - u, o.act) - y).abs())
            band = band.permute(0, 2, 3, 1)[..., :o.dst.c]
        y = y.permute(0, 2, 3, 1)[..., :o.dst.c]
        if o.res.c > 0:
            if want_band:  # the activation output is itself rounded to fp16 before the residual add
                band = band + ulp16(y)
            y = q(y) + res
        return (q(y), band) if want_band else q(y)
    if o.kind == _capi.OP_DWCONV:
        k = o.ksize
        x = src.permute(0, 3, 1, 2)
        c = x.shape[1]
        w = wblob[o.w_offset // 2: o.w_offset // 2 + k * k * c].view(k, k, c).permute(2, 0, 1).unsqueeze(1)
        b = bblob[o.b_offset // 4: o.b_offset // 4 + c]
        y0 = q(F.conv2d(x, w, b, stride=o.stride, padding=k // 2, groups=c))
        y = _act(y0, o.act)
        if want_band:  # as for the dense conv: the kernel applies the activation to the fp32 sum and rounds once
            u = ulp16(y0)
            band = torch.maximum((_act(y0 + u, o.act) - y).abs(), (_act(y0 - u, o.act) - y).abs()).permute(0, 2, 3, 1)
            return q(y.permute(0, 2, 3, 1)), band
        return q(y.permute(0, 2, 3, 1))
    if o.kind == _capi.OP_SPP:
        x = src.permute(0, 3, 1, 2)
        return torch.cat([F.max_pool2d(x, ks, 1, ks // 2) for ks in (5, 9, 13)], 1).permute(0, 2, 3, 1)
    if o.kind == _capi.OP_UPSAMPLE:
        return F.interpolate(src.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    raise AssertionError(f"unknown op kind {o.kind}")


class ArenaSim:
    """Flat fp32 'arena' addressed in fp16 element units, with the same views as the C side."""

    def __init__(self, nbytes):
        self.mem = torch.full((nbytes // 2,), float("nan"))

    def _index(self, v):
        off = v.offset // 2
        n = torch.arange(v.n).view(-1, 1, 1, 1) * v.nstride
        h = torch.arange(v.h).view(1, -1, 1, 1) * (v.w * v.pitch)
        w = torch.arange(v.w).view(1, 1, -1, 1) * v.pitch
        c = torch.arange(v.c).view(1, 1, 1, -1)
        return off + n + h + w + c

    def read(self, v):  # -> [n,h,w,c]
        return self.mem[self._index(v)]

    def write(self, v, t):
        self.mem[self._index(v)] = t


def run_graph_cpu(g, image, quantize=False):
    """image: NCHW fp32.  quantize=True rounds every stored activation to fp16 like the engine."""
    q = (lambda t: t.half().float()) if quantize else (lambda t: t)
    arena = ArenaSim(g.arena_bytes)
    wblob, bblob = g.weight_blob.float(), g.bias_blob
    ops = g.c_ops()
    for i, pop in enumerate(g.ops):
        o = ops[i]
        src = arena.read(o.src) if (o.kind != _capi.OP_S2D and not (o.kind == _capi.OP_CONV and (o.aux & 4))) else None
        res = arena.read(o.res) if (o.kind == _capi.OP_CONV and o.res.c > 0) else None
        up = arena.read(o.up) if (o.kind == _capi.OP_CONV and o.up.c > 0) else None
        for t, what in ((src, "src"), (res, "residual"), (up, "upsample source")):
            assert t is None or not torch.isnan(t).any(), f"op {i} {pop.name}: {what} reads uninitialised arena memory"
        arena.write(o.dst, eval_op(o, image, src, res, wblob, bblob, q, up=up))
    outs = g.outputs
    A, B = outs["reg"].h, g.batch

    def out_tensor(buf):
        return arena.read(buf.view().to_c()).reshape(B, A, buf.c)
    return out_tensor(outs["reg"]), out_tensor(outs["cls"])


def teacher_forced_errors(model, x):
    """Returns [(op index, name, max error / allowed rounding slack, max abs err)]; a ratio <= 1 means the op
    differs from the fp32 evaluation by no more than one fp16 rounding step of its pre-activation sum."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        eng = model.engine_for(x)
        g = eng.graph
        ops = g.c_ops()
        wblob, bblob = eng.weights.float(), eng.biases
        q = lambda t: t.half().float()
        xf = x.float()
        out = []
        for i, pop in enumerate(g.ops):
            o = ops[i]
            src = eng.view_tensor(o.src).float() if (o.kind != _capi.OP_S2D and not (o.kind == _capi.OP_CONV and (o.aux & 4))) else None
            res = eng.view_tensor(o.res).float().clone() if (o.kind == _capi.OP_CONV and o.res.c > 0) else None
            band = None
            up = eng.view_tensor(o.up).float() if (o.kind == _capi.OP_CONV and o.up.c > 0) else None
            if o.kind in (_capi.OP_CONV, _capi.OP_DWCONV):         # before the op runs (dst may alias res)
                ref, band = eval_op(o, xf, src, res, wblob, bblob, q, want_band=True, up=up)
            else:
                ref = eval_op(o, xf, src, res, wblob, bblob, q)
            eng.run_ops(x, i, 1)
            torch.cuda.synchronize()
            got = eng.view_tensor(o.dst).float()
            err = (got - ref).abs()
            # allowed: one fp16 step of the pre-activation sum (different fp32 accumulation order) propagated
            # through the activation, plus one fp16 step of the stored result
            # (floor 2^-4: near zero the absolute fp32 summation noise of ~1e4 O(0.1) terms, not the fp16
            # grid, limits agreement)
            tol = ulp16(torch.clamp(ref.abs(), min=2.0 ** -4)) + (band if band is not None else 0.0)
            out.append((i, pop.name, float((err / tol).max()), float(err.max())))
        return out
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
