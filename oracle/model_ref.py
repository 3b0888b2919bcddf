"""TEST INFRASTRUCTURE ONLY — never imported by the product package.

CPU fp32 restatement (plain torch functional ops, no nn.Module) of the reference's
YOLOX / YOLOX-P6 inference forward, BN folding and pruning-mask rule.  It exists so that the
CUDA engine can be checked on boxes where /root/reference is absent.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF imported in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`; checked by `tests/test_oracle_model.py`).

Reference files restated here (all paths relative to /root/reference):
  network ops       yolox/models/network_blocks.py:12-24,44-84,137-168,225-361
                    choijhanyangackr/yolox_infer/models/blocks.py:6-304
  backbones         yolox/models/darknet.py:89-171, yolox/models/darknet_p6.py:10-137
                    choijhanyangackr/yolox_infer/models/darknet.py, darknet_p6.py
  necks             yolox/models/yolo_pafpn.py:80-106, yolox/models/yolo_pafpn_p6.py:143-178
  head              yolox/models/yolo_head.py:131-225,
                    choijhanyangackr/yolox_infer/models/yolo_head.py:103-133
  BN fold           yolox/utils/model_utils.py:32-75, merge_save_p6.py:11-27
  pruning masks     01_mask_generator.py:18-46, 03_jh_merge.py:32-87 (intended semantics)
"""
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# model description
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ModelCfg:
    """kind: "yolox" = stock CSPDarknet/YOLOPAFPN (3 levels, Focus patch-major order, CSPLayer)
             "p6"    = CSPDarknetCustomP6/YOLOPAFPNCustomP6 (4 levels, pixel_unshuffle order,
                       CSPLayerCustom)
             "dw"    = YOLOXDepthwise (yolox_infer/models/{darknet,yolo_pafpn,yolox}_dw.py): 3 levels with channels
                       (256, 512, 768), pixel_unshuffle stem, 4x4 stride-2 convs, CSPLayerCustom whose bottlenecks use
                       depthwise 5x5 convs (BottleneckCustom, blocks.py:113-147)."""
    kind: str = "p6"
    depth: float = 0.67
    width: float = 0.75
    act: str = "hard_swish"
    num_classes: int = 80
    depthwise_neck: bool = False  # YOLOPAFPN(depthwise=True): DWConv in the neck bottlenecks
    v2: bool = False              # P6-v2 (yolox_infer/models/*_p6_v2.py): every stride-2 conv is 4x4, dark5's CSP has
                                  # 3x the bottlenecks WITH shortcuts (darknet_p6_v2.py:65-75), activation SiLU (main.py:40)

    @property
    def strides(self) -> Tuple[int, ...]:
        return (8, 16, 32, 64) if self.kind == "p6" else (8, 16, 32)

    @property
    def head_in_channels(self) -> Tuple[int, ...]:
        if self.kind == "dw":
            return (256, 512, 768)
        return (256, 512, 768, 1024) if self.kind == "p6" else (256, 512, 1024)


CONFIGS = {
    # BASELINE.json configs
    "nano": ModelCfg("yolox", 0.33, 0.25, "silu", 80, True),
    "yolox_m": ModelCfg("yolox", 0.67, 0.75, "silu", 80, False),
    "yolox_l": ModelCfg("yolox", 1.0, 1.0, "silu", 80, False),
    "yolox_m_p6": ModelCfg("p6", 0.67, 0.75, "hard_swish", 80, False),
    # tiny variants used for fast CPU parity tests / committed golden vectors
    "tiny": ModelCfg("yolox", 0.33, 0.25, "silu", 80, False),
    "tiny_p6": ModelCfg("p6", 0.33, 0.25, "hard_swish", 80, False),
    "yolox_m_p6_v2": ModelCfg("p6", 0.67, 0.75, "silu", 80, False, True),
    "tiny_p6_v2": ModelCfg("p6", 0.33, 0.25, "silu", 80, False, True),
    "yolox_l_dw": ModelCfg("dw", 1.0, 1.0, "hard_swish", 80, False),     # choijhanyangackr/config/yolox_l_dw.json
    "tiny_dw": ModelCfg("dw", 0.33, 0.25, "hard_swish", 80, False),
}


def activation(x: torch.Tensor, name: str) -> torch.Tensor:
    """network_blocks.py:12-24 / blocks.py:6-18."""
    name = name.lower()
    if name in ("silu", "swish"):
        return F.silu(x)
    if name in ("hsilu", "hswish", "hard_silu", "hard_swish"):
        return F.hardswish(x)
    if name == "relu":
        return F.relu(x)
    if name in ("lrelu", "leaky_relu"):
        return F.leaky_relu(x, 0.1)
    if name in ("none", "identity"):
        return x
    raise AttributeError("Unsupported act type: {}".format(name))


# --------------------------------------------------------------------------------------
# layer enumeration: (key prefix, cin, cout, k, stride, groups) for every conv of a model,
# in state-dict order.  Used to synthesise weights and to fold BN.
# --------------------------------------------------------------------------------------
def _csp_specs(p, cin, cout, n, custom, depthwise, dw_k=3, nop_rule=False, shortcut=True):
    """nop_rule: BottleneckCustom (blocks.py:129-137): a depthwise bottleneck that is neither the last of its CSP nor a
    residual one has NO pointwise conv (DWConvNoP)."""
    h = int(cout * 0.5)
    out = [(p + ".conv1", cin, h, 1, 1, 1),
           (p + ".conv2", cin, (cin - h) if custom else h, 1, 1, 1)]
    # nn.Module registration order in the reference: conv1, conv2, m, conv3
    for i in range(n):
        out.append((f"{p}.m.{i}.conv1", h, h, 1, 1, 1))
        if depthwise:
            out.append((f"{p}.m.{i}.conv2.dconv", h, h, dw_k, 1, h))
            if not (nop_rule and i != n - 1 and not shortcut):
                out.append((f"{p}.m.{i}.conv2.pconv", h, h, 1, 1, 1))
        else:
            out.append((f"{p}.m.{i}.conv2", h, h, 3, 1, 1))
    out.append((p + ".conv3", cin if custom else 2 * h, cout, 1, 1, 1))
    return out


def _dw_conv_specs(cfg: ModelCfg):
    """YOLOXDepthwise: darknet_dw.py:21-85, yolo_pafpn_dw.py:24-98."""
    base = int(cfg.width * 64)
    bd = max(round(cfg.depth * 3), 1)
    n = round(3 * cfg.depth)
    bb, nb = "backbone.backbone.", "backbone."
    kw = dict(dw_k=5, nop_rule=True)
    s = [(bb + "stem.conv", 12, base, 3, 1, 1), (bb + "dark2.0", base, base * 2, 4, 2, 1)]
    s += _csp_specs(bb + "dark2.1", base * 2, base * 2, bd, True, False)
    s.append((bb + "dark3.0", base * 2, base * 4, 4, 2, 1))
    s += _csp_specs(bb + "dark3.1", base * 4, base * 4, bd * 3, True, True, shortcut=True, **kw)
    s.append((bb + "dark4.0", base * 4, base * 8, 4, 2, 1))
    s += _csp_specs(bb + "dark4.1", base * 8, base * 8, bd * 3, True, True, shortcut=True, **kw)
    s.append((bb + "dark5.0", base * 8, base * 12, 4, 2, 1))
    s.append((bb + "dark5.1.conv1", base * 12, base * 6, 1, 1, 1))
    s.append((bb + "dark5.1.conv2", base * 24, base * 12, 1, 1, 1))
    s += _csp_specs(bb + "dark5.2", base * 12, base * 12, bd, True, True, shortcut=False, **kw)
    ic = [int(c * cfg.width) for c in cfg.head_in_channels]
    s.append((nb + "lateral_conv0", ic[2], ic[1], 1, 1, 1))
    s += _csp_specs(nb + "C3_p4", 2 * ic[1], ic[1], n, True, True, shortcut=False, **kw)
    s.append((nb + "reduce_conv1", ic[1], ic[0], 1, 1, 1))
    s += _csp_specs(nb + "C3_p3", 2 * ic[0], ic[0], n, True, True, shortcut=False, **kw)
    s.append((nb + "bu_conv2", ic[0], ic[0], 4, 2, 1))
    s += _csp_specs(nb + "C3_n3", 2 * ic[0], ic[1], n, True, True, shortcut=False, **kw)
    s.append((nb + "bu_conv1", ic[1], ic[1], 4, 2, 1))
    s += _csp_specs(nb + "C3_n4", 2 * ic[1], ic[2], n, True, True, shortcut=False, **kw)
    hc = int(256 * cfg.width)
    for k, c in enumerate(ic):
        s.append((f"head.stems.{k}", c, hc, 1, 1, 1))
        for j in range(2):
            s.append((f"head.cls_convs.{k}.{j}", hc, hc, 3, 1, 1))
        for j in range(2):
            s.append((f"head.reg_convs.{k}.{j}", hc, hc, 3, 1, 1))
    return s


def conv_specs(cfg: ModelCfg) -> List[Tuple[str, int, int, int, int, int]]:
    if cfg.kind == "dw":
        return _dw_conv_specs(cfg)
    base = int(cfg.width * 64)
    bd = max(round(cfg.depth * 3), 1)
    n_neck = round(3 * cfg.depth)
    custom = cfg.kind == "p6"
    dk = 4 if cfg.v2 else 3  # stride-2 ("down") conv kernel size
    bb = "backbone.backbone."
    s: List[Tuple[str, int, int, int, int, int]] = []
    s.append((bb + "stem.conv", 12, base, 3, 1, 1))
    s.append((bb + "dark2.0", base, base * 2, dk, 2, 1))
    s += _csp_specs(bb + "dark2.1", base * 2, base * 2, bd, custom, False)
    s.append((bb + "dark3.0", base * 2, base * 4, dk, 2, 1))
    s += _csp_specs(bb + "dark3.1", base * 4, base * 4, bd * 3, custom, False)
    s.append((bb + "dark4.0", base * 4, base * 8, dk, 2, 1))
    s += _csp_specs(bb + "dark4.1", base * 8, base * 8, bd * 3, custom, False)
    if custom:
        s.append((bb + "dark5.0", base * 8, base * 12, dk, 2, 1))
        s += _csp_specs(bb + "dark5.1", base * 12, base * 12, bd * 3 if cfg.v2 else bd, True, False)
        last, lc = "dark6", base * 12
    else:
        last, lc = "dark5", base * 8
    s.append((bb + last + ".0", lc, base * 16, dk, 2, 1))
    s.append((bb + last + ".1.conv1", base * 16, base * 8, 1, 1, 1))
    s.append((bb + last + ".1.conv2", base * 32, base * 16, 1, 1, 1))
    s += _csp_specs(bb + last + ".2", base * 16, base * 16, bd, custom, False)

    ic = [int(c * cfg.width) for c in cfg.head_in_channels]
    nb = "backbone."
    dw = cfg.depthwise_neck
    if custom:
        s.append((nb + "lateral_conv0", ic[3], ic[2], 1, 1, 1))
        s += _csp_specs(nb + "C3_p5", 2 * ic[2], ic[2], n_neck, True, dw)
        s.append((nb + "lateral_conv1", ic[2], ic[1], 1, 1, 1))
        s += _csp_specs(nb + "C3_p4", 2 * ic[1], ic[1], n_neck, True, dw)
        s.append((nb + "reduce_conv1", ic[1], ic[0], 1, 1, 1))
        s += _csp_specs(nb + "C3_p3", 2 * ic[0], ic[0], n_neck, True, dw)
        s.append((nb + "bu_conv2", ic[0], ic[0], dk, 2, 1))
        s += _csp_specs(nb + "C3_n3", 2 * ic[0], ic[1], n_neck, True, dw)
        s.append((nb + "bu_conv1", ic[1], ic[1], dk, 2, 1))
        s += _csp_specs(nb + "C3_n4", 2 * ic[1], ic[2], n_neck, True, dw)
        s.append((nb + "bu_conv0", ic[2], ic[2], dk, 2, 1))
        s += _csp_specs(nb + "C3_n5", 2 * ic[2], ic[3], n_neck, True, dw)
    else:
        s.append((nb + "lateral_conv0", ic[2], ic[1], 1, 1, 1))
        s += _csp_specs(nb + "C3_p4", 2 * ic[1], ic[1], n_neck, False, dw)
        s.append((nb + "reduce_conv1", ic[1], ic[0], 1, 1, 1))
        s += _csp_specs(nb + "C3_p3", 2 * ic[0], ic[0], n_neck, False, dw)
        s.append((nb + "bu_conv2", ic[0], ic[0], dk, 2, 1))
        s += _csp_specs(nb + "C3_n3", 2 * ic[0], ic[1], n_neck, False, dw)
        s.append((nb + "bu_conv1", ic[1], ic[1], dk, 2, 1))
        s += _csp_specs(nb + "C3_n4", 2 * ic[1], ic[2], n_neck, False, dw)

    hc = int(256 * cfg.width)
    for k, c in enumerate(ic):
        s.append((f"head.stems.{k}", c, hc, 1, 1, 1))
        for j in range(2):
            s.append((f"head.cls_convs.{k}.{j}", hc, hc, 3, 1, 1))
        for j in range(2):
            s.append((f"head.reg_convs.{k}.{j}", hc, hc, 3, 1, 1))
    return s


def pred_specs(cfg: ModelCfg):
    """The bare nn.Conv2d prediction layers (keys end in .weight/.bias directly)."""
    hc = int(256 * cfg.width)
    out = []
    for k in range(len(cfg.strides)):
        out.append((f"head.cls_preds.{k}", hc, cfg.num_classes))
        out.append((f"head.reg_preds.{k}", hc, 4))
        out.append((f"head.obj_preds.{k}", hc, 1))
    return out


# --------------------------------------------------------------------------------------
# deterministic synthetic weights (numpy RandomState: stable across torch versions)
# --------------------------------------------------------------------------------------
def synth_images(seed: int, batch: int, H: int, W: int) -> torch.Tensor:
    """Seeded multi-octave block noise in [0,255], NCHW fp32.  Unlike iid
    noise it keeps signal at every pyramid level, so deep layers see non-degenerate statistics."""
    rs = np.random.RandomState(seed)
    img = np.zeros((batch, 3, H, W), np.float64)
    tot = 0.0
    for s, wgt in ((64, 1.0), (32, 1.0), (16, 0.8), (8, 0.6), (4, 0.5), (2, 0.4), (1, 0.3)):
        n = rs.uniform(-1, 1, (batch, 3, -(-H // s), -(-W // s)))
        img += wgt * np.repeat(np.repeat(n, s, axis=2), s, axis=3)[:, :, :H, :W]
        tot += wgt
    img = (img / tot * 1.8).clip(-1, 1) * 127.5 + 127.5
    return torch.from_numpy(img.astype(np.float32))


def synth_train_state(cfg: ModelCfg, seed: int = 0, pred_bias: float = None,
                      calibrate: bool = True, calib_hw=(384, 384)) -> Dict[str, torch.Tensor]:
    """Unfused ('training-side') state dict: conv.weight (no bias) + bn.{weight,bias,running_mean,
    running_var} for every BaseConv, weight+bias for the preds.
    calibrate=True sets every BN's running statistics to the statistics its input actually has on a
    seeded noise image (what training would have produced), so activations stay O(1) through the
    ~140 layers; a plain random-init net either explodes or collapses to its biases and hides bugs.
    The statistics of the coarsest levels depend on the image size (13x13 SPP pooling over a 6x6 map
    is global), so pass calib_hw = the resolution the weights will be exercised at.
    pred_bias: value for cls/obj pred biases (the reference's initialize_biases(1e-2) gives -4.595;
    None = small random)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    for (p, cin, cout, k, s, g) in conv_specs(cfg):
        fan_in = (cin // g) * k * k
        w = rs.standard_normal((cout, cin // g, k, k)) * np.sqrt(2.0 / fan_in)
        sd[p + ".conv.weight"] = torch.from_numpy(w.astype(np.float32))
        sd[p + ".bn.weight"] = torch.from_numpy(rs.uniform(0.8, 1.2, cout).astype(np.float32))
        sd[p + ".bn.bias"] = torch.from_numpy(rs.uniform(-0.2, 0.2, cout).astype(np.float32))
        sd[p + ".bn.running_mean"] = torch.from_numpy(rs.uniform(-0.2, 0.2, cout).astype(np.float32))
        sd[p + ".bn.running_var"] = torch.from_numpy(rs.uniform(0.5, 1.5, cout).astype(np.float32))
    for (p, cin, cout) in pred_specs(cfg):
        gain = 0.5 if "reg_preds" in p else 3.0
        w = rs.standard_normal((cout, cin, 1, 1)) * (gain / np.sqrt(cin))
        sd[p + ".weight"] = torch.from_numpy(w.astype(np.float32))
        if pred_bias is not None and ("cls_preds" in p or "obj_preds" in p):
            b = np.full(cout, pred_bias, np.float32)
        elif "reg_preds" in p:
            b = rs.uniform(-0.5, 0.5, cout).astype(np.float32)
        else:
            b = rs.uniform(-3.0, -1.0, cout).astype(np.float32)
        sd[p + ".bias"] = torch.from_numpy(b)
    if calibrate:
        # SPP's 13x13 pooling is global on small maps, so the effective sample count of those
        # channels is the number of IMAGES: use many small images, few large ones.
        area = calib_hw[0] * calib_hw[1]
        nb = 32 if area <= 160 * 160 else 16 if area <= 320 * 320 else 8 if area <= 640 * 640 else 4
        x = synth_images(seed + 77, nb, calib_hw[0], calib_hw[1])
        with torch.no_grad():
            forward_raw(sd, cfg, x, _calibrate=True)
    return sd


def fold_bn(train_sd: Dict[str, torch.Tensor], eps: float = 1e-3) -> Dict[str, torch.Tensor]:
    """fuse_conv_and_bn (yolox/utils/model_utils.py:32-63) applied to every BaseConv, producing the
    flat fused state dict the inference twin loads with strict=True (merge_save_p6.py:18-27).
    W' = diag(g/sqrt(var+eps)) W ;  b' = g/sqrt(var+eps) * b_conv + beta - g*mean/sqrt(var+eps)."""
    out: Dict[str, torch.Tensor] = {}
    for key, w in train_sd.items():
        if key.endswith(".conv.weight") and key[:-len("conv.weight")] + "bn.weight" in train_sd:
            p = key[:-len(".conv.weight")]
            g, beta = train_sd[p + ".bn.weight"], train_sd[p + ".bn.bias"]
            mean, var = train_sd[p + ".bn.running_mean"], train_sd[p + ".bn.running_var"]
            scale = g / torch.sqrt(eps + var)
            out[p + ".conv.weight"] = (torch.diag(scale) @ w.reshape(w.shape[0], -1)).reshape(w.shape)
            b_conv = train_sd.get(p + ".conv.bias", torch.zeros_like(g))
            out[p + ".conv.bias"] = (torch.diag(scale) @ b_conv.reshape(-1, 1)).reshape(-1) + \
                (beta - g * mean / torch.sqrt(var + eps))
        elif ".bn." in key:
            continue
        else:
            out[key] = w
    return out


def magnitude_masks(train_sd: Dict[str, torch.Tensor], prune_pct: float = 49.0) -> Dict[str, torch.Tensor]:
    """01_mask_generator.py:18-39: global magnitude threshold over every 4-D tensor whose key lacks
    'head' (|w| clamped to 1.0, sorted ascending, threshold = element at int(N*pct/100));
    mask = |w| > threshold.  Returned under the conv.weight key."""
    elems = [v.flatten() for k, v in train_sd.items() if "head" not in k and v.ndim == 4]
    allw = torch.cat(elems).abs().clamp_max(1.0)
    thr = allw.sort()[0][int(len(allw) * prune_pct / 100)]
    return {k: torch.greater(v.abs(), thr) for k, v in train_sd.items() if "head" not in k and v.ndim == 4}


def two_four_masks(train_sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Synthetic 2:4-compliant masks: keep the 2 largest |w| of every group of
    4 consecutive input channels; tensors whose Cin is not a multiple of 4 are left dense."""
    out = {}
    for k, v in train_sd.items():
        if "head" in k or v.ndim != 4:
            continue
        co, ci, kh, kw = v.shape
        if ci % 4:
            out[k] = torch.ones_like(v, dtype=torch.bool)
            continue
        a = v.abs().permute(0, 2, 3, 1).reshape(-1, 4)
        idx = a.argsort(dim=1, descending=True, stable=True)[:, :2]
        m = torch.zeros_like(a, dtype=torch.bool)
        m.scatter_(1, idx, True)
        out[k] = m.reshape(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
    return out


def apply_masks(fused_sd, masks):
    """03_jh_merge.py:43-56 intended semantics: W_sparse = fold_bn(W) * mask."""
    out = dict(fused_sd)
    for k, m in masks.items():
        out[k] = fused_sd[k] * m.to(fused_sd[k].dtype)
    return out


def to_sparse_ckpt(fused_sd):
    """03_jh_merge.py:66-87: {"model": {key: coalesced sparse COO}}."""
    return {"model": {k: v.to_sparse().coalesce() for k, v in fused_sd.items()}}


# --------------------------------------------------------------------------------------
# forward (fused flat state dict)
# --------------------------------------------------------------------------------------
_CAL = {"on": False}


def _bconv(sd, p, x, stride, act, groups=1):
    """BaseConv: fused (conv+bias -> act, network_blocks.py:80-84 / blocks.py:46-49) when the state
    dict holds `conv.bias`, else unfused conv -> BN(eps=1e-3, running stats) -> act (:73-78)."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    if p + ".bn.weight" in sd:
        y = F.conv2d(x, w, sd.get(p + ".conv.bias"), stride=stride, padding=(k - 1) // 2, groups=groups)
        if _CAL["on"]:
            sd[p + ".bn.running_mean"] = y.mean(dim=(0, 2, 3)).clone()
            v = y.var(dim=(0, 2, 3), unbiased=False)
            sd[p + ".bn.running_var"] = torch.maximum(v, 0.25 * v.median()).clamp_min(1e-3).clone()
        y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"],
                         sd[p + ".bn.weight"], sd[p + ".bn.bias"], False, 0.0, 1e-3)
    else:
        y = F.conv2d(x, w, sd[p + ".conv.bias"], stride=stride, padding=(k - 1) // 2, groups=groups)
    return activation(y, act)


def _bottleneck(sd, p, x, act, use_add, depthwise):
    y = _bconv(sd, p + ".conv1", x, 1, act)
    if depthwise:
        y = _bconv(sd, p + ".conv2.dconv", y, 1, act, groups=y.shape[1])
        if p + ".conv2.pconv.conv.weight" in sd:      # absent for DWConvNoP (BottleneckCustom, blocks.py:129-131)
            y = _bconv(sd, p + ".conv2.pconv", y, 1, act)
    else:
        y = _bconv(sd, p + ".conv2", y, 1, act)
    return y + x if use_add else y


def _csp(sd, p, x, n, act, shortcut, depthwise=False):
    x0 = _bconv(sd, p + ".conv1", x, 1, act)
    x2 = _bconv(sd, p + ".conv2", x, 1, act)
    for i in range(n):
        x0 = _bottleneck(sd, f"{p}.m.{i}", x0, act, shortcut, depthwise)
    return _bconv(sd, p + ".conv3", torch.cat((x0, x2), dim=1), 1, act)


def _spp(sd, p, x, act):
    x = _bconv(sd, p + ".conv1", x, 1, act)
    xs = [x] + [F.max_pool2d(x, ks, stride=1, padding=ks // 2) for ks in (5, 9, 13)]
    return _bconv(sd, p + ".conv2", torch.cat(xs, dim=1), 1, act)


def space_to_depth(x, order):
    """order "focus": [TL,BL,TR,BR] patch-major (network_blocks.py:333-345);
       order "unshuffle": out ch = c*4 + dy*2 + dx == F.pixel_unshuffle (blocks.py:286-304)."""
    tl, tr = x[..., ::2, ::2], x[..., ::2, 1::2]
    bl, br = x[..., 1::2, ::2], x[..., 1::2, 1::2]
    if order == "focus":
        return torch.cat((tl, bl, tr, br), dim=1)
    b, c, h, w = x.shape
    y = torch.stack((tl, tr, bl, br), dim=2)  # [b, c, 4, h/2, w/2]
    return y.reshape(b, 4 * c, h // 2, w // 2)


def backbone_features(sd, cfg: ModelCfg, x):
    act = cfg.act
    bd = max(round(cfg.depth * 3), 1)
    bb = "backbone.backbone."
    if cfg.kind == "dw":                              # CSPDarknetDepthwise.forward, darknet_dw.py:87-101
        x = _bconv(sd, bb + "stem.conv", space_to_depth(x, "unshuffle"), 1, act)
        x = _csp(sd, bb + "dark2.1", _bconv(sd, bb + "dark2.0", x, 2, act), bd, act, True)
        d3 = x = _csp(sd, bb + "dark3.1", _bconv(sd, bb + "dark3.0", x, 2, act), bd * 3, act, True, True)
        d4 = x = _csp(sd, bb + "dark4.1", _bconv(sd, bb + "dark4.0", x, 2, act), bd * 3, act, True, True)
        x = _spp(sd, bb + "dark5.1", _bconv(sd, bb + "dark5.0", x, 2, act), act)
        return [d3, d4, _csp(sd, bb + "dark5.2", x, bd, act, False, True)]
    custom = cfg.kind == "p6"
    x = space_to_depth(x, "unshuffle" if custom else "focus")
    x = _bconv(sd, bb + "stem.conv", x, 1, act)
    x = _bconv(sd, bb + "dark2.0", x, 2, act)
    x = _csp(sd, bb + "dark2.1", x, bd, act, True)
    x = _bconv(sd, bb + "dark3.0", x, 2, act)
    d3 = x = _csp(sd, bb + "dark3.1", x, bd * 3, act, True)
    x = _bconv(sd, bb + "dark4.0", x, 2, act)
    d4 = x = _csp(sd, bb + "dark4.1", x, bd * 3, act, True)
    feats = [d3, d4]
    if custom:
        x = _bconv(sd, bb + "dark5.0", x, 2, act)
        x = _csp(sd, bb + "dark5.1", x, bd * 3 if cfg.v2 else bd, act, cfg.v2)
        feats.append(x)
        last = "dark6"
    else:
        last = "dark5"
    x = _bconv(sd, bb + last + ".0", x, 2, act)
    x = _spp(sd, bb + last + ".1", x, act)
    x = _csp(sd, bb + last + ".2", x, bd, act, False)
    feats.append(x)
    return feats


def _up(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def neck(sd, cfg: ModelCfg, feats):
    act, n, dw = cfg.act, round(3 * cfg.depth), cfg.depthwise_neck or cfg.kind == "dw"
    nb = "backbone."
    if cfg.kind == "p6":
        x3, x2, x1, x0 = feats
        fpn0 = _bconv(sd, nb + "lateral_conv0", x0, 1, act)
        f0 = _csp(sd, nb + "C3_p5", torch.cat([_up(fpn0), x1], 1), n, act, False, dw)
        fpn1 = _bconv(sd, nb + "lateral_conv1", f0, 1, act)
        f1 = _csp(sd, nb + "C3_p4", torch.cat([_up(fpn1), x2], 1), n, act, False, dw)
        fpn2 = _bconv(sd, nb + "reduce_conv1", f1, 1, act)
        pan3 = _csp(sd, nb + "C3_p3", torch.cat([_up(fpn2), x3], 1), n, act, False, dw)
        p2 = _bconv(sd, nb + "bu_conv2", pan3, 2, act)
        pan2 = _csp(sd, nb + "C3_n3", torch.cat([p2, fpn2], 1), n, act, False, dw)
        p1 = _bconv(sd, nb + "bu_conv1", pan2, 2, act)
        pan1 = _csp(sd, nb + "C3_n4", torch.cat([p1, fpn1], 1), n, act, False, dw)
        p0 = _bconv(sd, nb + "bu_conv0", pan1, 2, act)
        pan0 = _csp(sd, nb + "C3_n5", torch.cat([p0, fpn0], 1), n, act, False, dw)
        return [pan3, pan2, pan1, pan0]
    x2, x1, x0 = feats
    fpn0 = _bconv(sd, nb + "lateral_conv0", x0, 1, act)
    f0 = _csp(sd, nb + "C3_p4", torch.cat([_up(fpn0), x1], 1), n, act, False, dw)
    fpn1 = _bconv(sd, nb + "reduce_conv1", f0, 1, act)
    pan2 = _csp(sd, nb + "C3_p3", torch.cat([_up(fpn1), x2], 1), n, act, False, dw)
    p1 = _bconv(sd, nb + "bu_conv2", pan2, 2, act)
    pan1 = _csp(sd, nb + "C3_n3", torch.cat([p1, fpn1], 1), n, act, False, dw)
    p0 = _bconv(sd, nb + "bu_conv1", pan1, 2, act)
    pan0 = _csp(sd, nb + "C3_n4", torch.cat([p0, fpn0], 1), n, act, False, dw)
    return [pan2, pan1, pan0]


def head_raw(sd, cfg: ModelCfg, fpn_outs):
    """Infer flavour (yolox_infer/models/yolo_head.py:103-133): raw logits
    reg [B,A,4], obj [B,A,1], cls [B,A,C]; levels concatenated stride-ascending, row-major."""
    regs, objs, clss = [], [], []
    for k, x in enumerate(fpn_outs):
        b = x.shape[0]
        x = _bconv(sd, f"head.stems.{k}", x, 1, cfg.act)
        c = _bconv(sd, f"head.cls_convs.{k}.1", _bconv(sd, f"head.cls_convs.{k}.0", x, 1, cfg.act), 1, cfg.act)
        r = _bconv(sd, f"head.reg_convs.{k}.1", _bconv(sd, f"head.reg_convs.{k}.0", x, 1, cfg.act), 1, cfg.act)
        cls = F.conv2d(c, sd[f"head.cls_preds.{k}.weight"], sd[f"head.cls_preds.{k}.bias"])
        reg = F.conv2d(r, sd[f"head.reg_preds.{k}.weight"], sd[f"head.reg_preds.{k}.bias"])
        obj = F.conv2d(r, sd[f"head.obj_preds.{k}.weight"], sd[f"head.obj_preds.{k}.bias"])
        regs.append(reg.reshape(b, 4, -1))
        objs.append(obj.reshape(b, 1, -1))
        clss.append(cls.reshape(b, cfg.num_classes, -1))
    return (torch.cat(regs, 2).permute(0, 2, 1), torch.cat(objs, 2).permute(0, 2, 1),
            torch.cat(clss, 2).permute(0, 2, 1))


def forward_raw(sd, cfg: ModelCfg, x, _calibrate: bool = False):
    """YOLOXP6.forward / YOLOX.forward of the inference twin (yolox_p6.py:31-34)."""
    _CAL["on"] = _calibrate
    try:
        return head_raw(sd, cfg, neck(sd, cfg, backbone_features(sd, cfg, x)))
    finally:
        _CAL["on"] = False


def level_hw(cfg: ModelCfg, H: int, W: int):
    return [(H // s, W // s) for s in cfg.strides]


def grids_and_strides(hw, strides, dtype=torch.float32):
    """yolo_head.py:211-222 / postprocess_utils.py:6-24: x fastest, levels stride-ascending."""
    g, s = [], []
    for (h, w), st in zip(hw, strides):
        yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
        g.append(torch.stack((xv, yv), 2).reshape(1, -1, 2))
        s.append(torch.full((1, h * w, 1), st))
    return torch.cat(g, 1).to(dtype), torch.cat(s, 1).to(dtype)


def forward_yolox(sd, cfg: ModelCfg, x, decode: bool = True):
    """yolox-package flavour (yolo_head.py:167-190): [B,A,5+C] = [reg, sigmoid(obj), sigmoid(cls)],
    decoded in the tensor's dtype when decode_in_inference."""
    reg, obj, cls = forward_raw(sd, cfg, x)
    out = torch.cat([reg, obj.sigmoid(), cls.sigmoid()], dim=2)
    if decode:
        grids, strides = grids_and_strides(level_hw(cfg, x.shape[2], x.shape[3]), cfg.strides, out.dtype)
        out[..., :2] = (out[..., :2] + grids) * strides
        out[..., 2:4] = torch.exp(out[..., 2:4]) * strides
    return out
