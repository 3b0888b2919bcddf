"""Module tree mirroring the reference's block vocabulary (same attribute names -> same state_dict
keys), but the modules only HOLD parameters and EMIT ops into a plan.Graph; all arithmetic runs in
the native sm_100a engine.

Mirrors (reference root relative):
  yolox/models/network_blocks.py:44-84 (BaseConv), :107-120 (DWConv), :137-205 (Bottleneck[Custom]),
  :225-246 (SPPBottleneck), :249-320 (CSPLayer[Custom]), :323-361 (Focus[Custom])
  choijhanyangackr/yolox_infer/models/blocks.py (BN-free twins; bn=False here)
"""
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _capi
from .plan import Graph, V


def get_activation(name: str = "silu", inplace=True) -> str:
    """Validates like network_blocks.get_activation (AttributeError on unknown names); returns the name."""
    _capi.act_code(name)
    return name.lower()


def _no_eager(mod):
    raise RuntimeError(
        f"{type(mod).__name__}.forward: sub-modules of the B200 engine are not individually executable; "
        "call the top-level model (YOLOX / YOLOXP6 / YOLOXCustomP6) — there is no eager PyTorch path")


class BaseConv(nn.Module):
    """conv (+ BN when bn=True) + activation.  bn=True  -> keys conv.weight, bn.* (yolox flavour);
    bn=False -> keys conv.weight, conv.bias (inference twin / after fuse_model)."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="silu", bn=True):
        super().__init__()
        pad = (ksize - 1) // 2
        self.conv = nn.Conv2d(in_channels, out_channels, ksize, stride, pad, groups=groups, bias=bias or not bn)
        if bn:
            self.bn = nn.BatchNorm2d(out_channels, eps=1e-3, momentum=0.03)
        self.act_type = get_activation(act)
        self.ksize, self.stride, self.groups = ksize, stride, groups

    def folded(self):
        """(weight, bias) fp32 with BN folded: fuse_conv_and_bn, yolox/utils/model_utils.py:32-63."""
        w = self.conv.weight.detach().float()
        b = self.conv.bias.detach().float() if self.conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        if hasattr(self, "bn"):
            bn = self.bn
            scale = bn.weight.detach().float() / torch.sqrt(bn.eps + bn.running_var.detach().float())
            w = w * scale.reshape(-1, 1, 1, 1)
            b = scale * b + (bn.bias.detach().float() - bn.weight.detach().float() * bn.running_mean.detach().float()
                             / torch.sqrt(bn.running_var.detach().float() + bn.eps))
        return w.cpu(), b.cpu()

    def fuse_(self):
        """In-place fuse_model step for this block (model_utils.py:66-75): conv absorbs bn."""
        if hasattr(self, "bn"):
            w, b = self.folded()
            dev, dt = self.conv.weight.device, self.conv.weight.dtype
            conv = nn.Conv2d(self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size, self.conv.stride,
                             self.conv.padding, groups=self.conv.groups, bias=True)
            conv.weight.data.copy_(w); conv.bias.data.copy_(b)
            self.conv = conv.to(device=dev, dtype=dt).requires_grad_(False)
            del self.bn

    def emit(self, g: Graph, name: str, x: V, out: Optional[V] = None, res: Optional[V] = None) -> V:
        w, b = self.folded()
        cout = w.shape[0]
        pad = (self.ksize - 1) // 2
        ho = (x.H + 2 * pad - self.ksize) // self.stride + 1
        wo = (x.W + 2 * pad - self.ksize) // self.stride + 1
        if out is None:
            out = g.new_buf(name, ho, wo, cout).view()
        if self.groups == 1:
            return g.conv(name, x, out, w, b, self.stride, self.act_type, res)
        assert self.groups == w.shape[0] and res is None, "only dense and depthwise convs are supported"
        return g.dwconv(name, x, out, w, b, self.stride, self.act_type)

    def forward(self, x):
        _no_eager(self)


class DWConv(nn.Module):
    def __init__(self, in_channels, out_channels, ksize, stride=1, act="silu", bn=True):
        super().__init__()
        self.dconv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act, bn=bn)
        self.pconv = BaseConv(in_channels, out_channels, 1, 1, act=act, bn=bn)

    def emit(self, g, name, x, out=None, res=None):
        y = self.dconv.emit(g, name + ".dconv", x)
        return self.pconv.emit(g, name + ".pconv", y, out, res)

    def forward(self, x):
        _no_eager(self)


class DWConvNoP(nn.Module):
    """Depthwise conv only (yolox_infer/models/blocks.py:66-78)."""

    def __init__(self, in_channels, out_channels, ksize, stride=1, act="silu", bn=True):
        super().__init__()
        assert out_channels == in_channels
        self.dconv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act, bn=bn)

    def emit(self, g, name, x, out=None, res=None):
        assert res is None
        return self.dconv.emit(g, name + ".dconv", x, out)

    def forward(self, x):
        _no_eager(self)


class Bottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, shortcut=True, expansion=0.5, depthwise=False, act="silu", bn=True,
                 kernel_size=3, custom=False, is_last=False):
        """custom=True: BottleneckCustom (yolox_infer/models/blocks.py:113-147): a depthwise bottleneck that is neither
        the last of its CSP nor a residual one has no pointwise conv (DWConvNoP)."""
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=act, bn=bn)
        self.use_add = shortcut and in_channels == out_channels
        if depthwise and custom and not is_last and not self.use_add:
            self.conv2 = DWConvNoP(hidden, out_channels, kernel_size, 1, act=act, bn=bn)
        else:
            self.conv2 = (DWConv if depthwise else BaseConv)(hidden, out_channels, kernel_size, 1, act=act, bn=bn)

    def emit(self, g, name, x, out=None):
        y = self.conv1.emit(g, name + ".conv1", x)
        if self.use_add and os.environ.get("YX_INPLACE_RESIDUAL", "1") != "0":
            # y = x + conv2(conv1(x)) computed IN PLACE on x's buffer: the engine stores conv2's tile with a TMA
            # reduce-add (fp16 add in L2), so the residual is never loaded into shared memory.  x has no other reader
            # (network_blocks.py:199-205; inside CSPLayer the chain m.0 -> m.1 -> ... owns its input slice).
            assert out is None or (out.buf is x.buf and out.c_off == x.c_off and out.c == x.c), \
                "in-place Bottleneck: the requested output must be the input slice"
            return self.conv2.emit(g, name + ".conv2", y, x, res=x)
        return self.conv2.emit(g, name + ".conv2", y, out, res=x if self.use_add else None)

    def forward(self, x):
        _no_eager(self)


BottleneckCustom = Bottleneck  # identical arithmetic for the non-depthwise configs (network_blocks.py:171-205)


class SPPBottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_sizes=(5, 9, 13), activation="silu", bn=True):
        super().__init__()
        assert tuple(kernel_sizes) == (5, 9, 13), "the SPP kernel implements the reference's (5, 9, 13) pyramid"
        hidden = in_channels // 2
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=activation, bn=bn)
        self.m = nn.ModuleList([nn.MaxPool2d(ks, 1, ks // 2) for ks in kernel_sizes])  # parameter-free
        self.conv2 = BaseConv(hidden * 4, out_channels, 1, 1, act=activation, bn=bn)
        self.hidden = hidden

    def emit(self, g, name, x, out=None):
        cat = g.new_buf(name + ".cat", x.H, x.W, 4 * self.hidden)
        self.conv1.emit(g, name + ".conv1", x, cat.view(0, self.hidden))
        g.spp(cat.view(0, self.hidden), cat.view(self.hidden, 3 * self.hidden))
        return self.conv2.emit(g, name + ".conv2", cat.view(), out)

    def forward(self, x):
        _no_eager(self)


class CSPLayer(nn.Module):
    """stock CSP (custom=False, network_blocks.py:249-283) or CSPLayerCustom (custom=True, :286-320)."""

    def __init__(self, in_channels, out_channels, n=1, shortcut=True, expansion=0.5, depthwise=False, act="silu",
                 bn=True, custom=False, kernel_size=3):
        super().__init__()
        hidden = int(out_channels * expansion)
        c2 = (in_channels - hidden) if custom else hidden
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=act, bn=bn)
        self.conv2 = BaseConv(in_channels, c2, 1, 1, act=act, bn=bn)
        self.m = nn.Sequential(*[Bottleneck(hidden, hidden, shortcut, 1.0, depthwise, act=act, bn=bn, kernel_size=kernel_size,
                                            custom=custom, is_last=(i == n - 1)) for i in range(n)])
        self.conv3 = BaseConv(hidden + c2, out_channels, 1, 1, act=act, bn=bn)
        self.hidden, self.c2 = hidden, c2

    def emit(self, g, name, x, out=None, up=None):
        """up: the layer's input is torch.cat([upsample2x(up), x], 1) (PAFPN top-down path); neither the upsampled
        tensor nor the concatenation is materialised."""
        h, c2 = self.hidden, self.c2
        cat = g.new_buf(name + ".cat", x.H, x.W, h + c2)  # [x_1 | x_2]
        # conv1 and conv2 read the same tensor: one GEMM with Cout = h + c2 writes [x_0 | x_2]
        w1, b1 = self.conv1.folded()
        w2, b2 = self.conv2.folded()
        assert self.conv1.act_type == self.conv2.act_type
        g.conv(name + ".conv1+2", x, cat.view(), torch.cat([w1, w2], 0), torch.cat([b1, b2], 0), 1, self.conv1.act_type,
               up=up)
        cur = cat.view(0, h)
        n = len(self.m)
        for i, m in enumerate(self.m):
            cur = m.emit(g, f"{name}.m.{i}", cur, out=cat.view(0, h) if i == n - 1 else None)
        return self.conv3.emit(g, name + ".conv3", cat.view(), out)

    def forward(self, x):
        _no_eager(self)


def CSPLayerCustom(*a, **k):
    return CSPLayer(*a, custom=True, **k)


class Focus(nn.Module):
    order = "focus"  # [TL, BL, TR, BR] patch-major (network_blocks.py:333-345)

    def __init__(self, in_channels, out_channels, ksize=1, stride=1, act="silu", bn=True):
        super().__init__()
        assert in_channels == 3
        self.conv = BaseConv(in_channels * 4, out_channels, ksize, stride, act=act, bn=bn)

    def emit(self, g, name, out=None):
        c = self.conv
        if (c.ksize == 3 and c.stride == 1 and c.groups == 1 and os.environ.get("YX_FUSE_S2D", "1") != "0"
                and c.conv.out_channels <= 96      # nine resident weight taps + halo stages + raw patches must fit in smem
                and g.in_w % 16 == 0 and g.in_h // 2 >= 16 and g.in_w // 2 >= 8
                and -(-(g.in_h // 2) // 16) * 16 * (-(-(g.in_w // 2) // 8) * 8) <= 1.25 * (g.in_h // 2) * (g.in_w // 2)):
            # image-fed stem: the conv kernel's producer warps build the halo operand tiles straight from the NCHW image
            # (space-to-depth + input affine fused into the stem's operand build): no s2d launch, no s2d tensor
            w, b = c.folded()
            if out is None:
                out = g.new_buf(name + ".conv", g.in_h // 2, g.in_w // 2, w.shape[0]).view()
            return g.conv_stem_image(name + ".conv", out, w, b, c.act_type, self.order)
        if c.ksize == 3 and c.stride == 1 and c.groups == 1:
            # row-packed stem: padded s2d rows + overlapping-view conv (3 taps of K=48 instead of 9 of K=16)
            s2d = g.new_buf(name + ".s2d", g.in_h // 2, g.in_w // 2 + 4, 16)
            g.s2d(s2d.view(), self.order, padded=True)
            w, b = c.folded()
            if out is None:
                out = g.new_buf(name + ".conv", g.in_h // 2, g.in_w // 2, w.shape[0]).view()
            return g.conv_rowpack(name + ".conv", s2d.view(), out, w, b, c.act_type)
        s2d = g.new_buf(name + ".s2d", g.in_h // 2, g.in_w // 2, 16)
        g.s2d(s2d.view(), self.order)
        return c.emit(g, name + ".conv", s2d.view(), out)

    def forward(self, x):
        _no_eager(self)


class FocusCustom(Focus):
    order = "unshuffle"  # F.pixel_unshuffle channel order (network_blocks.py:357-361, blocks.py:286-304)
