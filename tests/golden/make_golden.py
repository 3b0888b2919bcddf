"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference) on
seeded synthetic weights/inputs.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/model_*.npz, tests/golden/post_*.npz.
The weights are not stored: `oracle.model_ref.synth_train_state(cfg, seed)` regenerates them from
numpy RandomState, so tests rebuild the exact same state dict.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)

# the reference imports `thop` at yolox/utils/model_utils.py:9 (absent here) — stub it.
thop = types.ModuleType("thop")
thop.profile = lambda *a, **k: (0, 0)
sys.modules["thop"] = thop
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "choijhanyangackr"))

from oracle import model_ref as mr  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_grad_enabled(False)


def ref_train_model(cfg: mr.ModelCfg):
    """Training-side reference model, as the exp files / merge_save*.py build it."""
    import torch.nn as nn
    if cfg.kind == "p6":
        from yolox.models import YOLOXCustomP6, YOLOPAFPNCustomP6, YOLOXHeadCustom
        backbone = YOLOPAFPNCustomP6(cfg.depth, cfg.width, act=cfg.act, in_channels=[256, 512, 768, 1024])
        head = YOLOXHeadCustom(cfg.num_classes, cfg.width, act=cfg.act, strides=(8, 16, 32, 64),
                               in_channels=[256, 512, 768, 1024])
        model = YOLOXCustomP6(backbone, head)
    else:
        from yolox.models import YOLOX, YOLOPAFPN, YOLOXHead
        backbone = YOLOPAFPN(cfg.depth, cfg.width, in_channels=[256, 512, 1024], act=cfg.act,
                             depthwise=cfg.depthwise_neck)
        head = YOLOXHead(cfg.num_classes, cfg.width, in_channels=[256, 512, 1024], act=cfg.act)
        model = YOLOX(backbone, head)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps = 1e-3
    return model.eval()


def ref_infer_model(cfg: mr.ModelCfg):
    from yolox_infer.models import YOLOX, YOLOXP6
    if cfg.kind == "p6":
        return YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes).eval()
    return YOLOX(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes).eval()


def load_train_state(model, sd):
    own = model.state_dict()
    full = dict(sd)
    for k, v in own.items():
        if k.endswith("num_batches_tracked"):
            full[k] = v
    model.load_state_dict(full, strict=True)


def golden_model(name, H, W, batch, seed):
    from yolox.utils.model_utils import fuse_model
    cfg = mr.CONFIGS[name]
    train_sd = mr.synth_train_state(cfg, seed, calib_hw=(H, W))
    x = mr.synth_images(seed + 1000, batch, H, W)

    tm = ref_train_model(cfg)
    load_train_state(tm, train_sd)
    tm.head.decode_in_inference = True
    y_dec_unfused = tm(x.clone())
    tm = fuse_model(tm)
    fused_ref = {k: v.clone() for k, v in tm.state_dict().items()}
    y_dec = tm(x.clone())
    tm.head.decode_in_inference = False
    y_undec = tm(x.clone())

    out = dict(x=x.numpy(), yolox_decoded=y_dec.numpy(), yolox_undecoded=y_undec.numpy(),
               yolox_decoded_unfused=y_dec_unfused.numpy())
    if not cfg.depthwise_neck:  # the inference twin's stock YOLOX has no depthwise neck
        im = ref_infer_model(cfg)
        im.load_state_dict(fused_ref, strict=True)
        reg, obj, cls = im(x.clone())
        out.update(reg=reg.contiguous().numpy(), obj=obj.contiguous().numpy(), cls=cls.contiguous().numpy())
    # a few fused tensors to pin fold_bn
    keys = [k for k in fused_ref if k.endswith("conv.bias")][:3] + [k for k in fused_ref if k.endswith("conv.weight")][:2]
    for i, k in enumerate(keys):
        out[f"fused_key_{i}"] = np.array(k)
        out[f"fused_val_{i}"] = fused_ref[k].numpy()
    np.savez_compressed(os.path.join(OUT, f"model_{name}_{H}x{W}_b{batch}_s{seed}.npz"), **out)
    print(name, "anchors", y_dec.shape, "fused-vs-unfused max|d|",
          float((y_dec - y_dec_unfused).abs().max()))


def golden_infer_v2(name, H, W, batch, seed):
    """P6-v2 inference twin (choijhanyangackr/yolox_infer/models/yolox_p6_v2.py, built like main.py:39-41 with
    act="silu"): the oracle's folded synthetic weights must load STRICTLY (same keys / shapes: 4x4 stride-2 convs, three
    times the dark5 bottlenecks), and its raw logits are the golden output."""
    cfg = mr.CONFIGS[name]
    fused = mr.fold_bn(mr.synth_train_state(cfg, seed, calib_hw=(H, W)))
    x = mr.synth_images(seed + 1000, batch, H, W)
    if cfg.kind == "dw":      # YOLOXDepthwise (yolox_infer/models/yolox_dw.py), main.py:36-38
        from yolox_infer.models.yolox_dw import YOLOXDepthwise
        im = YOLOXDepthwise(cfg.depth, cfg.width, act=cfg.act).eval()
    else:
        from yolox_infer.models.yolox_p6_v2 import YOLOXP6v2
        im = YOLOXP6v2(cfg.depth, cfg.width, act=cfg.act).eval()
    im.load_state_dict(fused, strict=True)
    reg, obj, cls = im(x.clone())
    np.savez_compressed(os.path.join(OUT, f"infer_{name}_{H}x{W}_b{batch}_s{seed}.npz"), x=x.numpy(),
                        reg=reg.contiguous().numpy(), obj=obj.contiguous().numpy(), cls=cls.contiguous().numpy())
    print(name, "reg", tuple(reg.shape), "n tensors", len(fused))


# ------------------------------------------------------------------------------------------
# post-processing goldens
# ------------------------------------------------------------------------------------------
def synth_head_logits(seed, B, hw, strides, C, dist):
    """SURVEY §8d config 4 distributions. Returns fp32 reg[B,A,4], obj[B,A,1], cls[B,A,C]."""
    rs = np.random.RandomState(seed)
    A = sum(h * w for h, w in hw)
    if dist == "maxcand":
        reg = rs.standard_normal((B, A, 4)).astype(np.float32)
        obj = (rs.standard_normal((B, A, 1)) * 2 - 2).astype(np.float32)
        cls = (rs.standard_normal((B, A, C)) * 2 - 2).astype(np.float32)
        return reg, obj, cls
    # clustered: n_gt boxes per image; anchors inside a box regress to it with jitter
    n_gt = 12 if A < 4000 else 200
    H, W = hw[0][0] * strides[0], hw[0][1] * strides[0]
    reg = np.zeros((B, A, 4), np.float32)
    obj = np.full((B, A, 1), -6.0, np.float32)
    cls = np.full((B, A, C), -6.0, np.float32)
    gx, gy, gs = [], [], []
    for (h, w), s in zip(hw, strides):
        yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        gx.append(xx.reshape(-1)); gy.append(yy.reshape(-1)); gs.append(np.full(h * w, s))
    gx, gy, gs = np.concatenate(gx), np.concatenate(gy), np.concatenate(gs).astype(np.float32)
    for b in range(B):
        cx = rs.uniform(0.1 * W, 0.9 * W, n_gt); cy = rs.uniform(0.1 * H, 0.9 * H, n_gt)
        bw = rs.uniform(0.05 * W, 0.4 * W, n_gt); bh = rs.uniform(0.05 * H, 0.4 * H, n_gt)
        lab = rs.randint(0, C, n_gt)
        for g in range(n_gt):
            ax, ay = (gx + 0.5) * gs, (gy + 0.5) * gs
            inside = (np.abs(ax - cx[g]) < bw[g] / 4) & (np.abs(ay - cy[g]) < bh[g] / 4)
            idx = np.nonzero(inside)[0]
            if len(idx) == 0:
                continue
            j = rs.standard_normal((len(idx), 4)).astype(np.float32) * 0.05
            reg[b, idx, 0] = cx[g] / gs[idx] - gx[idx] + j[:, 0]
            reg[b, idx, 1] = cy[g] / gs[idx] - gy[idx] + j[:, 1]
            reg[b, idx, 2] = np.log(bw[g] / gs[idx]) + j[:, 2]
            reg[b, idx, 3] = np.log(bh[g] / gs[idx]) + j[:, 3]
            obj[b, idx, 0] = rs.uniform(0, 4, len(idx))
            cls[b, idx, lab[g]] = rs.uniform(0, 6, len(idx))
    return reg, obj, cls


def golden_post(tag, seed, B, img, strides, C, dist, conf, nms_thr, half_yolox=False):
    from yolox_infer.postprocess_utils import (yolox_generate_grid, yolox_nms_torch_batch,
                                               yolox_postprocess_output_torch_batch)
    from yolox.utils import postprocess
    hw = [(img // s, img // s) for s in strides]
    reg, obj, cls = synth_head_logits(seed, B, hw, strides, C, dist)
    # main.py flavour ------------------------------------------------------------
    treg, tobj, tcls = (torch.from_numpy(a.astype(np.float16)) for a in (reg, obj, cls))  # fp16 logits, as .half() model
    grids, scales = yolox_generate_grid(img, strides=strides, dtype=torch.float16)
    boxes, objc, clsc = yolox_postprocess_output_torch_batch(treg, tobj, tcls, grids, scales)
    dets = yolox_nms_torch_batch(boxes, objc, clsc, nms_threshold=nms_thr, conf_threshold=conf)
    out = dict(reg=reg.astype(np.float16), obj=obj.astype(np.float16), cls=cls.astype(np.float16),
               boxes=boxes.numpy(), obj_conf=objc.numpy(), cls_conf=clsc.numpy(),
               cls_conf_max=clsc.max(-1)[0].numpy(), cls_arg=clsc.max(-1)[1].numpy().astype(np.int32),
               img=np.array(img), strides=np.array(strides), conf=np.array(conf), nms_thr=np.array(nms_thr))
    for i, d in enumerate(dets):
        out[f"main_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
    # uncapped main flavour (max_num_nms=0, max_num_det huge)
    dets_u = yolox_nms_torch_batch(boxes, objc, clsc, nms_threshold=nms_thr, conf_threshold=conf,
                                   max_num_nms=0, max_num_det=10 ** 9)
    for i, d in enumerate(dets_u):
        out[f"mainu_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
    # class-agnostic
    dets_a = yolox_nms_torch_batch(boxes, objc, clsc, nms_threshold=nms_thr, conf_threshold=conf,
                                   class_agnostic=True)
    for i, d in enumerate(dets_a):
        out[f"maina_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
    # yolox.utils.postprocess flavour: prediction = decoded [B,A,5+C] (cxcywh, sigmoid scores) -------
    dt = torch.float16 if half_yolox else torch.float32
    pred = torch.cat([torch.from_numpy(reg), torch.from_numpy(obj).sigmoid(), torch.from_numpy(cls).sigmoid()], 2)
    g32, s32 = yolox_generate_grid(img, strides=strides, dtype=torch.float32)
    pred[..., :2] = (pred[..., :2] + g32) * s32
    pred[..., 2:4] = torch.exp(pred[..., 2:4]) * s32
    pred = pred.to(dt)
    out["yolox_pred"] = pred.numpy().copy()
    res = postprocess(pred.clone().float() if half_yolox else pred.clone(), C, conf, nms_thr)
    if half_yolox:
        # CPU half lacks some ops in torchvision; run the reference on the fp16-rounded values in fp32
        out["yolox_pred_note"] = np.array("fp16-rounded inputs evaluated in fp32")
    for i, d in enumerate(res):
        out[f"yolox_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
    res_a = postprocess(pred.clone().float(), C, conf, nms_thr, class_agnostic=True)
    for i, d in enumerate(res_a):
        out[f"yoloxa_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
    np.savez_compressed(os.path.join(OUT, f"post_{tag}.npz"), **out)
    print(tag, "A", reg.shape[1], "main kept", [len(out[f'main_det_{i}']) for i in range(B)],
          "yolox kept", [len(out[f'yolox_det_{i}']) for i in range(B)])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "v2":       # only the P6-v2 / depthwise variant vectors
        golden_infer_v2("tiny_p6_v2", 128, 128, 1, 4)
        golden_infer_v2("tiny_dw", 96, 128, 1, 5)
        sys.exit(0)
    golden_infer_v2("tiny_p6_v2", 128, 128, 1, 4)
    golden_infer_v2("tiny_dw", 96, 128, 1, 5)
    golden_model("tiny_p6", 128, 128, 2, 0)
    golden_model("tiny", 96, 160, 1, 1)
    golden_model("nano", 64, 64, 1, 2)
    golden_post("clustered_p6_256", 10, 2, 256, (8, 16, 32, 64), 80, "clustered", 0.001, 0.65)
    golden_post("maxcand_p6_256", 11, 2, 256, (8, 16, 32, 64), 80, "maxcand", 0.001, 0.55)
    golden_post("clustered_p5_320_c20", 12, 1, 320, (8, 16, 32), 20, "clustered", 0.3, 0.45)
    golden_post("clustered_p6_640", 13, 1, 640, (8, 16, 32, 64), 80, "clustered", 0.001, 0.65)
