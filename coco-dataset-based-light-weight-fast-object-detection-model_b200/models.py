"""Drop-in model classes for the hot path.  Two flavours share one implementation:

  * "infer" flavour  = choijhanyangackr/yolox_infer/models/{yolox,yolox_p6}.py: BN-free, forward returns
    raw logits (reg[B,A,4], obj[B,A,1], cls[B,A,C])          -> infer.YOLOX, infer.YOLOXP6
  * "yolox" flavour  = yolox/models/{yolox,yolox_p6}.py (+ yolo_head.py:131-225): conv+BN blocks,
    forward returns [B,A,5+C] = [box, sigmoid(obj), sigmoid(cls)], decoded when
    head.decode_in_inference                                 -> models.YOLOX, models.YOLOXCustomP6

state_dict keys equal the reference's (e.g. backbone.backbone.dark2.0.conv.weight,
backbone.C3_p5.m.0.conv2.conv.bias, head.cls_preds.3.weight); load_state_dict(strict=True) works with
reference checkpoints.  forward() builds (once per input shape / weight version) a native engine and
runs it; nothing executes in eager PyTorch.
"""
from typing import Dict, List, Optional, Tuple

import os

import torch
import torch.nn as nn

from . import _capi
from .blocks import BaseConv, CSPLayer, Focus, FocusCustom, SPPBottleneck, _no_eager
from .plan import Engine, Graph, V


# ==========================================================================================
# backbones
# ==========================================================================================
class CSPDarknet(nn.Module):
    """yolox/models/darknet.py:89-171 / yolox_infer/models/darknet.py (stock, P3-P5)."""

    def __init__(self, dep_mul, wid_mul, out_features=("dark3", "dark4", "dark5"), depthwise=False, act="silu", bn=True):
        super().__init__()
        assert out_features, "please provide output features of Darknet"
        assert not depthwise, "depthwise backbone is not used by any configured model"
        self.out_features = out_features
        bc, bd = int(wid_mul * 64), max(round(dep_mul * 3), 1)
        kw = dict(act=act, bn=bn)
        self.stem = Focus(3, bc, ksize=3, **kw)
        self.dark2 = nn.Sequential(BaseConv(bc, bc * 2, 3, 2, **kw), CSPLayer(bc * 2, bc * 2, n=bd, **kw))
        self.dark3 = nn.Sequential(BaseConv(bc * 2, bc * 4, 3, 2, **kw), CSPLayer(bc * 4, bc * 4, n=bd * 3, **kw))
        self.dark4 = nn.Sequential(BaseConv(bc * 4, bc * 8, 3, 2, **kw), CSPLayer(bc * 8, bc * 8, n=bd * 3, **kw))
        self.dark5 = nn.Sequential(BaseConv(bc * 8, bc * 16, 3, 2, **kw),
                                   SPPBottleneck(bc * 16, bc * 16, activation=act, bn=bn),
                                   CSPLayer(bc * 16, bc * 16, n=bd, shortcut=False, **kw))
        self.stages = ("dark2", "dark3", "dark4", "dark5")

    def emit(self, g: Graph, outs: Dict[str, Optional[V]]) -> Dict[str, V]:
        """outs: where each out_feature must be written (a concat slice) or None."""
        p = "backbone.backbone."
        x = self.stem.emit(g, p + "stem")
        feats = {}
        for st in self.stages:
            seq = getattr(self, st)
            x = seq[0].emit(g, f"{p}{st}.0", x)
            for i in range(1, len(seq)):
                last = i == len(seq) - 1
                x = seq[i].emit(g, f"{p}{st}.{i}", x, outs.get(st) if last else None)
            feats[st] = x
        return {k: v for k, v in feats.items() if k in self.out_features}

    def forward(self, x):
        _no_eager(self)


class CSPDarknetCustomP6(CSPDarknet):
    """yolox/models/darknet_p6.py:10-137 / yolox_infer/models/darknet_p6.py (P3-P6, CSPLayerCustom)."""

    def __init__(self, dep_mul, wid_mul, out_features=("dark3", "dark4", "dark5", "dark6"), act="hard_swish", bn=True,
                 v2=False):
        """v2: CSPDarknetP6v2 (yolox_infer/models/darknet_p6_v2.py): 4x4 stride-2 convs, dark5 = 3x bottlenecks with
        shortcuts."""
        nn.Module.__init__(self)
        assert out_features, "please provide output features of Darknet"
        self.out_features = out_features
        bc, bd = int(wid_mul * 64), max(round(dep_mul * 3), 1)
        kw = dict(act=act, bn=bn)
        ckw = dict(act=act, bn=bn, custom=True)
        dk = 4 if v2 else 3
        self.stem = FocusCustom(3, bc, ksize=3, **kw)
        self.dark2 = nn.Sequential(BaseConv(bc, bc * 2, dk, 2, **kw), CSPLayer(bc * 2, bc * 2, n=bd, **ckw))
        self.dark3 = nn.Sequential(BaseConv(bc * 2, bc * 4, dk, 2, **kw), CSPLayer(bc * 4, bc * 4, n=bd * 3, **ckw))
        self.dark4 = nn.Sequential(BaseConv(bc * 4, bc * 8, dk, 2, **kw), CSPLayer(bc * 8, bc * 8, n=bd * 3, **ckw))
        self.dark5 = nn.Sequential(BaseConv(bc * 8, bc * 12, dk, 2, **kw),
                                   CSPLayer(bc * 12, bc * 12, n=bd * 3 if v2 else bd, shortcut=bool(v2), **ckw))
        self.dark6 = nn.Sequential(BaseConv(bc * 12, bc * 16, dk, 2, **kw),
                                   SPPBottleneck(bc * 16, bc * 16, activation=act, bn=bn),
                                   CSPLayer(bc * 16, bc * 16, n=bd, shortcut=False, **ckw))
        self.stages = ("dark2", "dark3", "dark4", "dark5", "dark6")


class CSPDarknetDepthwise(CSPDarknet):
    """choijhanyangackr/yolox_infer/models/darknet_dw.py:6-101: P3-P5 with (4b, 8b, 12b) channels, 4x4 stride-2 convs,
    CSPLayerCustom whose dark3..dark5 bottlenecks are depthwise 5x5."""

    def __init__(self, dep_mul, wid_mul, out_features=("dark3", "dark4", "dark5"), act="hard_swish", bn=True):
        nn.Module.__init__(self)
        self.out_features = out_features
        bc, bd = int(wid_mul * 64), max(round(dep_mul * 3), 1)
        kw = dict(act=act, bn=bn)
        ckw = dict(act=act, bn=bn, custom=True)
        dkw = dict(depthwise=True, kernel_size=5, **ckw)
        self.stem = FocusCustom(3, bc, ksize=3, **kw)
        self.dark2 = nn.Sequential(BaseConv(bc, bc * 2, 4, 2, **kw), CSPLayer(bc * 2, bc * 2, n=bd, **ckw))
        self.dark3 = nn.Sequential(BaseConv(bc * 2, bc * 4, 4, 2, **kw), CSPLayer(bc * 4, bc * 4, n=bd * 3, **dkw))
        self.dark4 = nn.Sequential(BaseConv(bc * 4, bc * 8, 4, 2, **kw), CSPLayer(bc * 8, bc * 8, n=bd * 3, **dkw))
        self.dark5 = nn.Sequential(BaseConv(bc * 8, bc * 12, 4, 2, **kw),
                                   SPPBottleneck(bc * 12, bc * 12, activation=act, bn=bn),
                                   CSPLayer(bc * 12, bc * 12, n=bd, shortcut=False, **dkw))
        self.stages = ("dark2", "dark3", "dark4", "dark5")


# ==========================================================================================
# necks
# ==========================================================================================
def _topdown(g: Graph, csp, name: str, lateral: V, cat: "Buf", c_up: int) -> V:
    """PAFPN top-down merge (yolo_pafpn.py:86-94, yolo_pafpn_p6.py:153-165): CSP(cat([upsample(lateral), feature])).
    `cat` = [c_up channels reserved for the upsampled tensor | backbone feature, already written by its producer].
    When the engine can fold the upsample into the consumer's loads the first half of `cat` is never written."""
    feat = cat.view(c_up, cat.c - c_up)
    if os.environ.get("YX_FUSE_UPSAMPLE", "1") != "0" and Graph.can_fuse_upsample(lateral):
        return csp.emit(g, name, feat, up=lateral)
    g.upsample(lateral, cat.view(0, c_up))
    return csp.emit(g, name, cat.view())


class YOLOPAFPN(nn.Module):
    """yolox/models/yolo_pafpn.py:15-106 / yolox_infer/models/yolo_pafpn.py."""

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024),
                 depthwise=False, act="silu", bn=True):
        super().__init__()
        self.backbone = CSPDarknet(depth, width, act=act, bn=bn)
        self.in_features, self.in_channels = in_features, in_channels
        assert len(in_channels) == 3
        c = [int(ch * width) for ch in in_channels]
        n = round(3 * depth)
        kw = dict(act=act, bn=bn)
        ckw = dict(shortcut=False, depthwise=depthwise, act=act, bn=bn)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = BaseConv(c[2], c[1], 1, 1, **kw)
        self.C3_p4 = CSPLayer(2 * c[1], c[1], n, **ckw)
        self.reduce_conv1 = BaseConv(c[1], c[0], 1, 1, **kw)
        self.C3_p3 = CSPLayer(2 * c[0], c[0], n, **ckw)
        self.bu_conv2 = BaseConv(c[0], c[0], 3, 2, **kw)
        self.C3_n3 = CSPLayer(2 * c[0], c[1], n, **ckw)
        self.bu_conv1 = BaseConv(c[1], c[1], 3, 2, **kw)
        self.C3_n4 = CSPLayer(2 * c[1], c[2], n, **ckw)
        self.c = c

    def emit(self, g: Graph) -> List[V]:
        c = self.c
        H8, W8 = g.in_h // 8, g.in_w // 8
        # concat buffers first, so producers can write into their slices
        cat_p4 = g.new_buf("cat_p4", H8 // 2, W8 // 2, 2 * c[1])   # [up(fpn_out0) | dark4]
        cat_p3 = g.new_buf("cat_p3", H8, W8, 2 * c[0])             # [up(fpn_out1) | dark3]
        cat_n3 = g.new_buf("cat_n3", H8 // 2, W8 // 2, 2 * c[0])   # [bu_conv2 | fpn_out1]
        cat_n4 = g.new_buf("cat_n4", H8 // 4, W8 // 4, 2 * c[1])   # [bu_conv1 | fpn_out0]
        f = self.backbone.emit(g, {"dark3": cat_p3.view(c[0], c[0]), "dark4": cat_p4.view(c[1], c[1]), "dark5": None})
        x0 = f["dark5"]
        fpn_out0 = self.lateral_conv0.emit(g, "backbone.lateral_conv0", x0, cat_n4.view(c[1], c[1]))
        f_out0 = _topdown(g, self.C3_p4, "backbone.C3_p4", fpn_out0, cat_p4, c[1])
        fpn_out1 = self.reduce_conv1.emit(g, "backbone.reduce_conv1", f_out0, cat_n3.view(c[0], c[0]))
        pan_out2 = _topdown(g, self.C3_p3, "backbone.C3_p3", fpn_out1, cat_p3, c[0])
        self.bu_conv2.emit(g, "backbone.bu_conv2", pan_out2, cat_n3.view(0, c[0]))
        pan_out1 = self.C3_n3.emit(g, "backbone.C3_n3", cat_n3.view())
        self.bu_conv1.emit(g, "backbone.bu_conv1", pan_out1, cat_n4.view(0, c[1]))
        pan_out0 = self.C3_n4.emit(g, "backbone.C3_n4", cat_n4.view())
        return [pan_out2, pan_out1, pan_out0]

    def forward(self, x):
        _no_eager(self)


class YOLOPAFPNDepthwise(YOLOPAFPN):
    """choijhanyangackr/yolox_infer/models/yolo_pafpn_dw.py:8-140: the 3-level PAFPN over CSPDarknetDepthwise, depthwise
    5x5 CSPLayerCustom blocks, 4x4 stride-2 bottom-up convs (same dataflow as YOLOPAFPN, so emit() is inherited)."""

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 768),
                 act="hard_swish", bn=True):
        nn.Module.__init__(self)
        self.backbone = CSPDarknetDepthwise(depth, width, act=act, bn=bn)
        self.in_features, self.in_channels = in_features, in_channels
        c = [int(ch * width) for ch in in_channels]
        n = round(3 * depth)
        kw = dict(act=act, bn=bn)
        ckw = dict(shortcut=False, depthwise=True, kernel_size=5, custom=True, act=act, bn=bn)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = BaseConv(c[2], c[1], 1, 1, **kw)
        self.C3_p4 = CSPLayer(2 * c[1], c[1], n, **ckw)
        self.reduce_conv1 = BaseConv(c[1], c[0], 1, 1, **kw)
        self.C3_p3 = CSPLayer(2 * c[0], c[0], n, **ckw)
        self.bu_conv2 = BaseConv(c[0], c[0], 4, 2, **kw)
        self.C3_n3 = CSPLayer(2 * c[0], c[1], n, **ckw)
        self.bu_conv1 = BaseConv(c[1], c[1], 4, 2, **kw)
        self.C3_n4 = CSPLayer(2 * c[1], c[2], n, **ckw)
        self.c = c


class YOLOPAFPNCustomP6(nn.Module):
    """yolox/models/yolo_pafpn_p6.py:14-178 / yolox_infer/models/yolo_pafpn_p6.py."""

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5", "dark6"),
                 in_channels=(256, 512, 768, 1024), act="hard_swish", bn=True, v2=False):
        """v2: YOLOPAFPNP6v2 (yolox_infer/models/yolo_pafpn_p6_v2.py): 4x4 stride-2 bottom-up convs, v2 backbone."""
        super().__init__()
        self.backbone = CSPDarknetCustomP6(depth, width, act=act, bn=bn, v2=v2)
        dk = 4 if v2 else 3
        self.in_features, self.in_channels = in_features, in_channels
        assert len(in_channels) == 4
        c = [int(ch * width) for ch in in_channels]
        n = round(3 * depth)
        kw = dict(act=act, bn=bn)
        ckw = dict(shortcut=False, act=act, bn=bn, custom=True)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = BaseConv(c[3], c[2], 1, 1, **kw)
        self.C3_p5 = CSPLayer(2 * c[2], c[2], n, **ckw)
        self.lateral_conv1 = BaseConv(c[2], c[1], 1, 1, **kw)
        self.C3_p4 = CSPLayer(2 * c[1], c[1], n, **ckw)
        self.reduce_conv1 = BaseConv(c[1], c[0], 1, 1, **kw)
        self.C3_p3 = CSPLayer(2 * c[0], c[0], n, **ckw)
        self.bu_conv2 = BaseConv(c[0], c[0], dk, 2, **kw)
        self.C3_n3 = CSPLayer(2 * c[0], c[1], n, **ckw)
        self.bu_conv1 = BaseConv(c[1], c[1], dk, 2, **kw)
        self.C3_n4 = CSPLayer(2 * c[1], c[2], n, **ckw)
        self.bu_conv0 = BaseConv(c[2], c[2], dk, 2, **kw)
        self.C3_n5 = CSPLayer(2 * c[2], c[3], n, **ckw)
        self.c = c

    def emit(self, g: Graph) -> List[V]:
        c = self.c
        H8, W8 = g.in_h // 8, g.in_w // 8
        cat_p5 = g.new_buf("cat_p5", H8 // 4, W8 // 4, 2 * c[2])   # [up(fpn_out0) | dark5]
        cat_p4 = g.new_buf("cat_p4", H8 // 2, W8 // 2, 2 * c[1])   # [up(fpn_out1) | dark4]
        cat_p3 = g.new_buf("cat_p3", H8, W8, 2 * c[0])             # [up(fpn_out2) | dark3]
        cat_n3 = g.new_buf("cat_n3", H8 // 2, W8 // 2, 2 * c[0])   # [bu_conv2 | fpn_out2]
        cat_n4 = g.new_buf("cat_n4", H8 // 4, W8 // 4, 2 * c[1])   # [bu_conv1 | fpn_out1]
        cat_n5 = g.new_buf("cat_n5", H8 // 8, W8 // 8, 2 * c[2])   # [bu_conv0 | fpn_out0]
        f = self.backbone.emit(g, {"dark3": cat_p3.view(c[0], c[0]), "dark4": cat_p4.view(c[1], c[1]),
                                   "dark5": cat_p5.view(c[2], c[2]), "dark6": None})
        x0 = f["dark6"]
        fpn_out0 = self.lateral_conv0.emit(g, "backbone.lateral_conv0", x0, cat_n5.view(c[2], c[2]))
        f_out0 = _topdown(g, self.C3_p5, "backbone.C3_p5", fpn_out0, cat_p5, c[2])
        fpn_out1 = self.lateral_conv1.emit(g, "backbone.lateral_conv1", f_out0, cat_n4.view(c[1], c[1]))
        f_out1 = _topdown(g, self.C3_p4, "backbone.C3_p4", fpn_out1, cat_p4, c[1])
        fpn_out2 = self.reduce_conv1.emit(g, "backbone.reduce_conv1", f_out1, cat_n3.view(c[0], c[0]))
        pan_out3 = _topdown(g, self.C3_p3, "backbone.C3_p3", fpn_out2, cat_p3, c[0])
        self.bu_conv2.emit(g, "backbone.bu_conv2", pan_out3, cat_n3.view(0, c[0]))
        pan_out2 = self.C3_n3.emit(g, "backbone.C3_n3", cat_n3.view())
        self.bu_conv1.emit(g, "backbone.bu_conv1", pan_out2, cat_n4.view(0, c[1]))
        pan_out1 = self.C3_n4.emit(g, "backbone.C3_n4", cat_n4.view())
        self.bu_conv0.emit(g, "backbone.bu_conv0", pan_out1, cat_n5.view(0, c[2]))
        pan_out0 = self.C3_n5.emit(g, "backbone.C3_n5", cat_n5.view())
        return [pan_out3, pan_out2, pan_out1, pan_out0]

    def forward(self, x):
        _no_eager(self)


# ==========================================================================================
# head
# ==========================================================================================
class YOLOXHead(nn.Module):
    """yolox/models/yolo_head.py:17-110 (eval half) / yolox_infer/models/yolo_head.py:7-101."""

    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024), act="silu", bn=True):
        super().__init__()
        self.n_anchors = 1
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.strides = strides
        self.hw = None
        hc = int(256 * width)
        self.stems, self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        kw = dict(act=act, bn=bn)
        for ch in in_channels:
            self.stems.append(BaseConv(int(ch * width), hc, 1, 1, **kw))
            self.cls_convs.append(nn.Sequential(BaseConv(hc, hc, 3, 1, **kw), BaseConv(hc, hc, 3, 1, **kw)))
            self.reg_convs.append(nn.Sequential(BaseConv(hc, hc, 3, 1, **kw), BaseConv(hc, hc, 3, 1, **kw)))
            self.cls_preds.append(nn.Conv2d(hc, self.n_anchors * num_classes, 1, 1, 0))
            self.reg_preds.append(nn.Conv2d(hc, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hc, self.n_anchors * 1, 1, 1, 0))
        self.hc = hc

    def initialize_biases(self, prior_prob):
        """yolo_head.py:120-129."""
        import math
        with torch.no_grad():   # in-place on the parameter itself: bumps its version, so cached engines are rebuilt
            for conv in list(self.cls_preds) + list(self.obj_preds):
                conv.bias.fill_(-math.log((1 - prior_prob) / prior_prob))

    def emit(self, g: Graph, feats: List[V]) -> Tuple["Buf", "Buf", List[Tuple[int, int]]]:
        assert len(feats) == len(self.strides) == len(self.stems), \
            "number of FPN levels, strides and head branches must agree (SURVEY C5)"
        hc, C = self.hc, self.num_classes
        cpitch = (C + 7) // 8 * 8
        level_hw = [(f.H, f.W) for f in feats]
        A = sum(h * w for h, w in level_hw)
        cls_out = g.new_output("head.cls", A, cpitch)   # [B, A, C]  raw class logits
        reg_out = g.new_output("head.regobj", A, 8)     # [B, A, 8]  = [reg(4), obj(1), 0, 0, 0]
        off = 0
        for k, x in enumerate(feats):
            h, w = level_hw[k]
            x = self.stems[k].emit(g, f"head.stems.{k}", x)
            # cls_convs[k][0] and reg_convs[k][0] share their input: one GEMM with Cout = 2*hc
            wc, bc = self.cls_convs[k][0].folded()
            wr, br = self.reg_convs[k][0].folded()
            both = g.new_buf(f"head.l{k}.cls0+reg0", h, w, 2 * hc)
            g.conv(f"head.cls_convs.{k}.0+reg_convs.{k}.0", x, both.view(), torch.cat([wc, wr], 0),
                   torch.cat([bc, br], 0), 1, self.cls_convs[k][0].act_type)
            cls_feat = self.cls_convs[k][1].emit(g, f"head.cls_convs.{k}.1", both.view(0, hc))
            reg_feat = self.reg_convs[k][1].emit(g, f"head.reg_convs.{k}.1", both.view(hc, hc))
            cp = self.cls_preds[k]
            g.conv(f"head.cls_preds.{k}", cls_feat, V(cls_out, 0, cpitch, lvl_off=off, h=h, w=w, nstride=A * cpitch),
                   cp.weight.detach().float().cpu(), cp.bias.detach().float().cpu(), 1, "none")
            rp, op = self.reg_preds[k], self.obj_preds[k]
            g.conv(f"head.reg_preds.{k}+obj_preds.{k}", reg_feat, V(reg_out, 0, 8, lvl_off=off, h=h, w=w, nstride=A * 8),
                   torch.cat([rp.weight.detach().float().cpu(), op.weight.detach().float().cpu()], 0),
                   torch.cat([rp.bias.detach().float().cpu(), op.bias.detach().float().cpu()], 0), 1, "none")
            off += h * w
        return reg_out, cls_out, level_hw

    # --- reference API: decode_outputs(outputs, dtype) in place, yolo_head.py:210-225 -----------
    def decode_outputs(self, outputs, dtype=None):
        from .postprocess import decode_outputs
        return decode_outputs(outputs, self.hw, self.strides)

    def forward(self, xin, labels=None, imgs=None):
        _no_eager(self)


def YOLOXHeadCustom(num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 768), act="hard_swish", bn=True):
    """yolox/models/yolo_head_custom.py:17-24: YOLOXHead with other defaults."""
    return YOLOXHead(num_classes, width, strides, in_channels, act, bn)


# ==========================================================================================
# top-level models
# ==========================================================================================
class _EngineModel(nn.Module):
    """Shared engine cache + execution.  Subclasses define self.backbone, self.head, flavour."""

    flavour = "infer"   # "infer": (reg, obj, cls) raw logits ; "yolox": [B,A,5+C]
    max_engines = 4

    def _weights_version(self):
        """Identity of the weights an engine was packed from.  In-place edits of a parameter bump `_version`; writes
        through `p.data` do NOT (`.data` carries its own version counter) -- after such an edit call
        invalidate_engines()."""
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def invalidate_engines(self):
        """Drop every cached engine: the next forward repacks the current parameter values.  Needed only after weight
        edits the version counters cannot see (`p.data.copy_(...)`, `p.data.fill_(...)`)."""
        self.__dict__.pop("_engines", None)
        return self

    def build_graph(self, batch: int, in_h: int, in_w: int) -> Graph:
        stride_max = max(self.head.strides)
        if in_h % stride_max or in_w % stride_max:
            raise RuntimeError(f"input {in_h}x{in_w} must be a multiple of the largest stride {stride_max}")
        g = Graph(batch, in_h, in_w)
        feats = self.backbone.emit(g)
        reg_out, cls_out, level_hw = self.head.emit(g, feats)
        g.outputs = dict(reg=reg_out, cls=cls_out, level_hw=level_hw)
        return g.finalize()

    def engine_for(self, x) -> Engine:
        if self.training:
            raise RuntimeError("yolox_b200 implements the inference path only: call model.eval()")
        cache = self.__dict__.setdefault("_engines", {})
        key = (tuple(x.shape), str(x.device))
        ver = self._weights_version()
        ent = cache.get(key)
        if ent is None or ent[0] != ver:
            g = self.build_graph(x.shape[0], x.shape[2], x.shape[3])
            ent = (ver, Engine(g, x.device))
            cache[key] = ent
            while len(cache) > self.max_engines:
                cache.pop(next(iter(cache)))
        return ent[1]

    def run_engine(self, x, in_scale=1.0, in_shift=0.0, use_graph=False):
        """Runs the network; returns (engine, reg8[B,A,8], cls[B,A,Cp]) as VIEWS of the engine arena
        (valid until the next run of the same engine)."""
        _capi.require_cuda(x, "input image")
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError("input must be [B,3,H,W]")
        eng = self.engine_for(x)
        eng.run(x, in_scale, in_shift, use_graph)
        outs = eng.graph.outputs
        A = outs["reg"].h
        reg8 = eng.tensor_of(outs["reg"]).view(x.shape[0], A, 8)
        cls = eng.tensor_of(outs["cls"]).view(x.shape[0], A, outs["cls"].c)
        self.head.hw = [tuple(hw) for hw in outs["level_hw"]]
        return eng, reg8, cls

    def forward(self, x, targets=None):
        eng, reg8, cls = self.run_engine(x)
        C = self.head.num_classes
        out_dtype = torch.float16 if x.dtype == torch.uint8 else x.dtype   # uint8 pixels are evaluated as the same values in fp16
        if self.flavour == "infer":
            # fresh tensors like the reference (yolo_head.py:130-133 returns permuted views of new tensors)
            reg8 = reg8.clone()
            cls = cls.clone().to(out_dtype) if out_dtype != torch.float16 else cls.clone()
            reg8 = reg8.to(out_dtype)
            return reg8[..., :4], reg8[..., 4:5], cls[..., :C]
        from .postprocess import head_assemble
        return head_assemble(reg8, cls, C, self.head.hw, self.head.strides, self.head.decode_in_inference, out_dtype)


class _InferYOLOX(_EngineModel):
    flavour = "infer"

    def __init__(self, depth=1.0, width=1.0, act="silu", num_classes: int = 80):
        super().__init__()
        self.backbone = YOLOPAFPN(depth, width, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024),
                                  act=act, bn=False)
        self.head = YOLOXHead(num_classes, width, strides=(8, 16, 32), in_channels=(256, 512, 1024), act=act, bn=False)
        self.eval()


class _InferYOLOXP6(_EngineModel):
    flavour = "infer"

    def __init__(self, depth=1.0, width=1.0, act="hard_swish", num_classes: int = 80):
        super().__init__()
        self.backbone = YOLOPAFPNCustomP6(depth, width, in_channels=(256, 512, 768, 1024), act=act, bn=False)
        self.head = YOLOXHead(num_classes, width, strides=(8, 16, 32, 64), in_channels=(256, 512, 768, 1024), act=act,
                              bn=False)
        self.eval()


class _InferYOLOXP6v2(_EngineModel):
    """choijhanyangackr/yolox_infer/models/yolox_p6_v2.py (main.py:39-41 builds it with act="silu")."""
    flavour = "infer"

    def __init__(self, depth=1.0, width=1.0, act="hard_swish", num_classes: int = 80):
        super().__init__()
        self.backbone = YOLOPAFPNCustomP6(depth, width, in_channels=(256, 512, 768, 1024), act=act, bn=False, v2=True)
        self.head = YOLOXHead(num_classes, width, strides=(8, 16, 32, 64), in_channels=(256, 512, 768, 1024), act=act,
                              bn=False)
        self.eval()


class _InferYOLOXDepthwise(_EngineModel):
    """choijhanyangackr/yolox_infer/models/yolox_dw.py (main.py:36-38)."""
    flavour = "infer"

    def __init__(self, depth=1.0, width=1.0, act="hard_swish", num_classes: int = 80):
        super().__init__()
        self.backbone = YOLOPAFPNDepthwise(depth, width, in_channels=(256, 512, 768), act=act, bn=False)
        self.head = YOLOXHead(num_classes, width, strides=(8, 16, 32), in_channels=(256, 512, 768), act=act, bn=False)
        self.eval()


class YOLOX(_EngineModel):
    """yolox/models/yolox.py:10-47 (inference branch)."""
    flavour = "yolox"

    def __init__(self, backbone=None, head=None):
        super().__init__()
        self.backbone = backbone if backbone is not None else YOLOPAFPN()
        self.head = head if head is not None else YOLOXHead(80)


class YOLOXCustomP6(_EngineModel):
    """yolox/models/yolox_p6.py:10-49 (inference branch).  NOTE the reference's no-arg default head uses
    in_channels (256,512,1024,1024), which mismatches its own PAFPN; like every real caller,
    the default here is (256,512,768,1024)."""
    flavour = "yolox"

    def __init__(self, backbone=None, head=None):
        super().__init__()
        self.backbone = backbone if backbone is not None else YOLOPAFPNCustomP6(act="hard_swish")
        self.head = head if head is not None else YOLOXHead(80, strides=(8, 16, 32, 64),
                                                            in_channels=(256, 512, 768, 1024), act="hard_swish")


def fuse_model(model):
    """yolox/utils/model_utils.py:66-75: fold every BaseConv's BN into its conv, in place."""
    with torch.no_grad():
        for m in model.modules():
            if type(m) is BaseConv and hasattr(m, "bn"):
                m.fuse_()
    return model
