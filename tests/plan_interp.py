"""Test helper: a CPU interpreter for plan.Graph op lists (torch fp32).  It executes exactly what the
native engine is told to execute — arena offsets, channel slices, packed KRSC weight blobs, merged
convs — so the graph builder / memory planner / weight packer can be verified without a GPU."""
import numpy as np
import torch
import torch.nn.functional as F

import yolox_b200 as yb
from yolox_b200 import _capi

ACT = {0: "none", 1: "silu", 2: "hard_swish", 3: "relu", 4: "lrelu"}


def _act(x, code):
    from oracle.model_ref import activation
    return activation(x, ACT[code])


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
    """image: NCHW fp32.  quantize=True rounds every stored activation / weight to fp16 like the engine."""
    q = (lambda t: t.half().float()) if quantize else (lambda t: t)
    arena = ArenaSim(g.arena_bytes)
    wblob = g.weight_blob.float()
    bblob = g.bias_blob
    ops = g.c_ops()
    for i, pop in enumerate(g.ops):
        o = ops[i]
        if o.kind == _capi.OP_S2D:
            x = image
            tl, tr, bl, br = x[..., ::2, ::2], x[..., ::2, 1::2], x[..., 1::2, ::2], x[..., 1::2, 1::2]
            if o.aux == 1:
                y = torch.stack((tl, tr, bl, br), dim=2).reshape(x.shape[0], 12, x.shape[2] // 2, x.shape[3] // 2)
            else:
                y = torch.cat((tl, bl, tr, br), dim=1)
            y = F.pad(y, (0, 0, 0, 0, 0, 4))
            arena.write(o.dst, q(y.permute(0, 2, 3, 1)))
        elif o.kind == _capi.OP_CONV:
            k = o.ksize
            x = arena.read(o.src).permute(0, 3, 1, 2)
            assert not torch.isnan(x).any(), f"op {i} {pop.name}: reads uninitialised arena memory"
            x = F.pad(x, (0, 0, 0, 0, 0, o.cin_pad - x.shape[1]))
            w = wblob[o.w_offset // 2: o.w_offset // 2 + o.cout_pad * k * k * o.cin_pad]
            w = w.view(o.cout_pad, k, k, o.cin_pad).permute(0, 3, 1, 2)
            b = bblob[o.b_offset // 4: o.b_offset // 4 + o.cout_pad]
            y = F.conv2d(x, w, b, stride=o.stride, padding=k // 2)
            y = _act(q(y), o.act)
            y = y.permute(0, 2, 3, 1)[..., :o.dst.c]
            if o.res.c > 0:
                r = arena.read(o.res)
                assert not torch.isnan(r).any(), f"op {i} {pop.name}: residual reads uninitialised memory"
                y = q(y) + r
            arena.write(o.dst, q(y))
        elif o.kind == _capi.OP_DWCONV:
            k = o.ksize
            x = arena.read(o.src).permute(0, 3, 1, 2)
            c = x.shape[1]
            w = wblob[o.w_offset // 2: o.w_offset // 2 + k * k * c].view(k, k, c).permute(2, 0, 1).unsqueeze(1)
            b = bblob[o.b_offset // 4: o.b_offset // 4 + c]
            y = _act(q(F.conv2d(x, w, b, stride=o.stride, padding=k // 2, groups=c)), o.act)
            arena.write(o.dst, q(y.permute(0, 2, 3, 1)))
        elif o.kind == _capi.OP_SPP:
            x = arena.read(o.src).permute(0, 3, 1, 2)
            ys = [F.max_pool2d(x, ks, 1, ks // 2) for ks in (5, 9, 13)]
            arena.write(o.dst, torch.cat(ys, 1).permute(0, 2, 3, 1))
        elif o.kind == _capi.OP_UPSAMPLE:
            x = arena.read(o.src).permute(0, 3, 1, 2)
            arena.write(o.dst, F.interpolate(x, scale_factor=2, mode="nearest").permute(0, 2, 3, 1))
        else:
            raise AssertionError(f"unknown op kind {o.kind}")
    outs = g.outputs
    A = outs["reg"].h
    B = g.batch

    def out_tensor(buf):
        v = buf.view().to_c()
        return arena.read(v).reshape(B, A, buf.c)
    return out_tensor(outs["reg"]), out_tensor(outs["cls"])
