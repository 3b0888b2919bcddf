"""Pins oracle/model_ref.py (the CPU restatement) against outputs of the unmodified reference
(tests/golden/model_*.npz, produced by tests/golden/make_golden.py in the build container)."""
import glob
import os
import re

import numpy as np
import pytest
import torch

from oracle import model_ref as mr

torch.set_grad_enabled(False)
FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "model_*.npz")))


def _parse(path):
    m = re.match(r"model_(.+)_(\d+)x(\d+)_b(\d+)_s(\d+)\.npz", os.path.basename(path))
    return m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))


def _close(a, b, rtol=2e-4, atol=2e-4):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    err = np.abs(a[fin] - b[fin]) / (atol + rtol * np.abs(b[fin]))
    assert err.max() <= 1.0, f"max normalised err {err.max():.3g}"


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_forward_matches_reference(path):
    name, H, W, B, seed = _parse(path)
    g = np.load(path)
    cfg = mr.CONFIGS[name]
    train_sd = mr.synth_train_state(cfg, seed, calib_hw=(H, W))
    x = mr.synth_images(seed + 1000, B, H, W)
    np.testing.assert_array_equal(x.numpy(), g["x"])
    fused = mr.fold_bn(train_sd)
    # BN fold vs reference fuse_model
    i = 0
    while f"fused_key_{i}" in g:
        _close(fused[str(g[f"fused_key_{i}"])].numpy(), g[f"fused_val_{i}"], 1e-5, 1e-6)
        i += 1
    assert i >= 4
    # infer flavour raw logits
    reg, obj, cls = mr.forward_raw(fused, cfg, x)
    if "reg" in g:
        _close(reg.numpy(), g["reg"]); _close(obj.numpy(), g["obj"]); _close(cls.numpy(), g["cls"])
    # yolox flavour, undecoded and decoded
    _close(mr.forward_yolox(fused, cfg, x, decode=False).numpy(), g["yolox_undecoded"])
    _close(mr.forward_yolox(fused, cfg, x, decode=True).numpy(), g["yolox_decoded"], 5e-4, 5e-4)
    # unfused (conv -> BN -> act) path
    _close(mr.forward_yolox(train_sd, cfg, x, decode=True).numpy(), g["yolox_decoded_unfused"], 5e-4, 5e-4)


def test_p6_v2_forward_matches_reference():
    """P6-v2 (4x4 stride-2 convs, 3x dark5 bottlenecks with shortcuts, SiLU): raw logits of the reference's inference twin
    loaded strictly with these weights (tests/golden/make_golden.py::golden_infer_v2)."""
    path = os.path.join(os.path.dirname(__file__), "golden", "infer_tiny_p6_v2_128x128_b1_s4.npz")
    g = np.load(path)
    cfg = mr.CONFIGS["tiny_p6_v2"]
    fused = mr.fold_bn(mr.synth_train_state(cfg, 4, calib_hw=(128, 128)))
    x = mr.synth_images(1004, 1, 128, 128)
    np.testing.assert_array_equal(x.numpy(), g["x"])
    assert fused["backbone.backbone.dark2.0.conv.weight"].shape[-2:] == (4, 4)
    assert fused["backbone.bu_conv0.conv.weight"].shape[-2:] == (4, 4)
    reg, obj, cls = mr.forward_raw(fused, cfg, x)
    _close(reg.numpy(), g["reg"]); _close(obj.numpy(), g["obj"]); _close(cls.numpy(), g["cls"])


def test_depthwise_variant_forward_matches_reference():
    """YOLOXDepthwise (depthwise 5x5 BottleneckCustom with the DWConvNoP rule, 4x4 stride-2 convs): raw logits of the
    reference's inference twin loaded strictly with these weights."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "infer_tiny_dw_96x128_b1_s5.npz"))
    cfg = mr.CONFIGS["tiny_dw"]
    fused = mr.fold_bn(mr.synth_train_state(cfg, 5, calib_hw=(96, 128)))
    x = mr.synth_images(1005, 1, 96, 128)
    np.testing.assert_array_equal(x.numpy(), g["x"])
    assert fused["backbone.backbone.dark3.1.m.0.conv2.dconv.conv.weight"].shape[1:] == (1, 5, 5)
    assert "backbone.C3_p4.m.0.conv2.pconv.conv.weight" in fused       # n = 1: the only bottleneck is the last one
    reg, obj, cls = mr.forward_raw(fused, cfg, x)
    _close(reg.numpy(), g["reg"]); _close(obj.numpy(), g["obj"]); _close(cls.numpy(), g["cls"])


def test_state_dict_key_count_m_p6():
    """the M-P6 inference twin has 278 tensors."""
    cfg = mr.CONFIGS["yolox_m_p6"]
    n = 2 * len(mr.conv_specs(cfg)) + 2 * len(mr.pred_specs(cfg))
    assert n == 278
    assert len(mr.conv_specs(cfg)) + len(mr.pred_specs(cfg)) == 139


def test_space_to_depth_orders():
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 3, 4, 6)
    assert torch.equal(mr.space_to_depth(x, "unshuffle"), torch.nn.functional.pixel_unshuffle(x, 2))
    f = mr.space_to_depth(x, "focus")
    assert torch.equal(f[:, 0:3], x[..., ::2, ::2]) and torch.equal(f[:, 3:6], x[..., 1::2, ::2])
    assert torch.equal(f[:, 6:9], x[..., ::2, 1::2]) and torch.equal(f[:, 9:12], x[..., 1::2, 1::2])


def test_masks():
    cfg = mr.CONFIGS["tiny_p6"]
    sd = mr.synth_train_state(cfg, 3, calibrate=False)
    m = mr.magnitude_masks(sd, 49.0)
    assert all("head" not in k for k in m)
    tot = sum(v.numel() for v in m.values()); kept = sum(int(v.sum()) for v in m.values())
    assert abs(kept / tot - 0.51) < 0.01
    m24 = mr.two_four_masks(sd)
    for k, v in m24.items():
        if v.shape[1] % 4 == 0:
            g = v.permute(0, 2, 3, 1).reshape(-1, 4).sum(1)
            assert int(g.min()) == 2 and int(g.max()) == 2
    fused = mr.fold_bn(sd)
    sp = mr.to_sparse_ckpt(mr.apply_masks(fused, m))["model"]
    k0 = next(iter(m))
    assert torch.equal(sp[k0].to_dense(), fused[k0] * m[k0])
