"""CPU checks of the host logic: module tree / state_dict surface, graph builder (concat elimination,
merged convs), memory planner, weight packer — by interpreting the planned op list with torch and
comparing with the oracle (which is pinned against the reference)."""
import numpy as np
import pytest
import torch

import yolox_b200 as yb
from oracle import model_ref as mr
from tests.plan_interp import run_graph_cpu

torch.set_grad_enabled(False)


def _q16(sd):
    return {k: (v.half().float() if k.endswith("weight") else v) for k, v in sd.items()}


def _infer_model(name):
    cfg = mr.CONFIGS[name]
    cls = {"p6": yb.infer.YOLOXP6v2 if cfg.v2 else yb.infer.YOLOXP6, "dw": yb.infer.YOLOXDepthwise}.get(cfg.kind, yb.infer.YOLOX)
    return cfg, cls(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)


@pytest.mark.parametrize("name,H,W,B", [("tiny_p6", 128, 192, 2), ("tiny", 96, 160, 1), ("tiny_p6_v2", 128, 128, 1),
                                        ("tiny_dw", 96, 128, 2)])
def test_infer_graph_matches_oracle(name, H, W, B):
    cfg, model = _infer_model(name)
    fused = mr.fold_bn(mr.synth_train_state(cfg, 3, calib_hw=(H, W)))
    model.load_state_dict(fused, strict=True)       # reference key names, strict
    assert set(model.state_dict().keys()) == set(fused.keys())
    g = model.build_graph(B, H, W)
    x = mr.synth_images(11, B, H, W)
    reg8, cls = run_graph_cpu(g, x)
    reg, obj, clso = mr.forward_raw(_q16(fused), cfg, x)   # the engine stores weights in fp16
    np.testing.assert_allclose(reg8[..., :4].numpy(), reg.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(reg8[..., 4:5].numpy(), obj.numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(cls[..., :cfg.num_classes].numpy(), clso.numpy(), rtol=2e-4, atol=2e-4)
    assert float(reg8[..., 5:].abs().max()) == 0.0


def test_yolox_flavour_bn_fold_and_nano_depthwise():
    cfg = mr.CONFIGS["nano"]
    H = W = 64
    train = mr.synth_train_state(cfg, 5, calib_hw=(H, W))
    backbone = yb.models.YOLOPAFPN(cfg.depth, cfg.width, in_channels=[256, 512, 1024], act=cfg.act, depthwise=True)
    head = yb.models.YOLOXHead(cfg.num_classes, cfg.width, in_channels=[256, 512, 1024], act=cfg.act)
    model = yb.models.YOLOX(backbone, head).eval()
    sd = dict(train)
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = v
    model.load_state_dict(sd, strict=True)
    x = mr.synth_images(12, 1, H, W)
    ref = mr.forward_raw(_q16(mr.fold_bn(train)), cfg, x)
    reg8, cls = run_graph_cpu(model.build_graph(1, H, W), x)
    np.testing.assert_allclose(reg8[..., :4].numpy(), ref[0].numpy(), rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(cls.numpy(), ref[2].numpy(), rtol=2e-4, atol=2e-4)
    # fuse_model: same keys as the reference's fused checkpoint, same graph result
    yb.models.fuse_model(model)
    fused = mr.fold_bn(train)
    assert set(model.state_dict().keys()) == set(fused.keys())
    for k, v in model.state_dict().items():
        np.testing.assert_allclose(v.numpy(), fused[k].numpy(), rtol=1e-5, atol=1e-6)
    reg8b, clsb = run_graph_cpu(model.build_graph(1, H, W), x)
    np.testing.assert_allclose(clsb.numpy(), cls.numpy(), rtol=1e-4, atol=1e-4)


def test_m_p6_surface_and_work():
    cfg, model = _infer_model("yolox_m_p6")
    assert len(model.state_dict()) == 278
    assert model.head.strides == (8, 16, 32, 64)
    g = model.build_graph(2, 1280, 1280)
    assert abs(g.conv_flops() / 2 / 1e9 - 315.28) < 0.05       # 315.28 GFLOP / image
    assert sum(h * w for h, w in g.outputs["level_hw"]) == 34000
    # small batches (the multi-stream "lanes" regime): no buffer shares bytes with any other
    assert g.arena_bytes == sum(b.nbytes for b in g.bufs)
    # planner at the bench batch: live ranges of overlapping buffers never share bytes, dead buffers' memory is reused
    g = model.build_graph(16, 1280, 1280)
    for i, a in enumerate(g.bufs):
        for b in g.bufs[i + 1:]:
            if not (a.last < b.first or b.last < a.first):
                assert a.offset + a.nbytes <= b.offset or b.offset + b.nbytes <= a.offset, (a.name, b.name)
    assert g.arena_bytes < sum(b.nbytes for b in g.bufs) / 4


def test_sparse_checkpoint_ingest():
    """main.py:52-55: state_dict()[key].copy_(param.to_dense()) for a {"model": sparse_coo} checkpoint."""
    cfg, model = _infer_model("tiny_p6")
    train = mr.synth_train_state(cfg, 6, calibrate=False)
    fused = mr.apply_masks(mr.fold_bn(train), mr.magnitude_masks(train, 49.0))
    ckpt = mr.to_sparse_ckpt(fused)["model"]
    for key, param in ckpt.items():
        model.state_dict()[key].copy_(param.to_dense().data)
    for k, v in model.state_dict().items():
        assert torch.equal(v, fused[k])
    g = model.build_graph(1, 64, 64)
    nz = float((g.weight_blob != 0).float().mean())
    assert 0.3 < nz < 0.8


def test_errors():
    with pytest.raises(AttributeError):
        yb.infer.YOLOXP6(0.33, 0.25, act="gelu")
    cfg, model = _infer_model("tiny_p6")
    with pytest.raises(RuntimeError):
        model.build_graph(1, 100, 128)                          # not a multiple of 64
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 64, 64))                        # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        yb.postprocess.yolox_nms_torch_batch(None, None, None, soft=True)


def test_peer_out_struct_matches_header():
    """ctypes mirror of yx_peer_out: 4 int32 + 3 pointer tables of YX_MAX_PEERS + 3 pointers; yx_conv_tune: 12 int32."""
    import ctypes
    import os
    import re
    from yolox_b200 import _capi
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "yolox_b200.h")).read()
    assert int(re.search(r"#define YX_MAX_PEERS (\d+)", hdr).group(1)) == _capi.MAX_PEERS
    assert int(re.search(r"#define YX_IPC_HANDLE_BYTES (\d+)", hdr).group(1)) == _capi.IPC_HANDLE_BYTES
    assert ctypes.sizeof(_capi.PeerOut) == 16 + 3 * 8 * _capi.MAX_PEERS + 24
    body = re.search(r"typedef struct yx_peer_out \{(.*?)\} yx_peer_out;", hdr, re.S).group(1)
    names = re.findall(r"(?:int32_t|void\*)\s+(\w+)(?:\[YX_MAX_PEERS\])?;", body)
    assert names == [f[0] for f in _capi.PeerOut._fields_], names
    body = re.search(r"typedef struct yx_conv_tune \{(.*?)\} yx_conv_tune;", hdr, re.S).group(1)
    names = re.findall(r"int32_t\s+(\w+)(?:\[\d+\])?;", body)
    assert names == [f[0] for f in _capi.ConvTune._fields_], names
    assert ctypes.sizeof(_capi.ConvTune) == 48


def test_tune_cache_roundtrip(tmp_path, monkeypatch):
    """plan.TuneCache: keys are per layer geometry and batch, files are rewritten atomically, other kernel revisions are
    dropped; the device name is only needed for the section, so a fake one is injected (no GPU here)."""
    from yolox_b200 import plan
    monkeypatch.setenv("YX_TUNE_CACHE", str(tmp_path / "tc.json"))
    monkeypatch.setattr(plan.TuneCache, "section", lambda self, device: "fake B200|rev1")
    cfg, model = _infer_model("tiny_p6")
    g1, g2 = model.build_graph(1, 64, 64), model.build_graph(2, 64, 64)
    convs = [op for op in g1.ops if op.kind == 0]
    # a layer with 2:4-compliant weights is tuned over more candidates: its persisted choice has its own key
    assert plan.TuneCache.op_key(convs[3], 1, sparse_ok=True) != plan.TuneCache.op_key(convs[3], 1)
    k1 = [plan.TuneCache.op_key(op, 1) for op in convs]
    k2 = [plan.TuneCache.op_key(op, 2) for op in g2.ops if op.kind == 0]
    assert len(set(k1)) <= len(k1) and not set(k1) & set(k2)          # batch is part of the key
    assert any("inplace" in k for k in k1) and any("up" in k and "up-" not in k for k in k1)
    c = plan.TuneCache()
    assert c.load(None) == {}
    c.store(None, {k1[0]: [1, 64, 1, 1, 1, 2, 1, 0, 0, 0, 0]})
    c.store(None, {k1[1]: [2, 96, 1, 2, 1, 2, 1, 0, 0, 0, 1]})
    got = plan.TuneCache().load(None)
    assert got == {k1[0]: [1, 64, 1, 1, 1, 2, 1, 0, 0, 0, 0], k1[1]: [2, 96, 1, 2, 1, 2, 1, 0, 0, 0, 1]}
    monkeypatch.setattr(plan.TuneCache, "section", lambda self, device: "fake B200|rev2")
    assert plan.TuneCache().load(None) == {}
    plan.TuneCache().store(None, {k1[0]: [1, 128, 1, 1, 1, 2, 1, 0, 0, 0, 0]})
    import json
    assert list(json.load(open(tmp_path / "tc.json"))) == ["fake B200|rev2"]
    monkeypatch.setenv("YX_TUNE_CACHE", "0")
    assert plan.TuneCache().load(None) == {}
    from yolox_b200 import _capi
    t = _capi.ConvTune.from_list([2, 96, 1, 2, 1, 2, 1, 0, 1, 0, 1])
    assert t.as_list() == [2, 96, 1, 2, 1, 2, 1, 0, 1, 0, 1] and t.cta_pair == 1 and t.epilogue_alternate == 1


def test_shipped_tune_cache_matches_kernel_revision():
    """The launch-shape cache shipped next to the package is only ever USED for the kernel revision it was measured on
    (plan.TuneCache keys it by device name and a hash of the conv kernel sources); a stale file is dead weight that makes
    every first run tune and rewrite it, so shipping one is a mistake this test catches."""
    import importlib
    import json
    import os
    from yolox_b200 import plan
    b = importlib.import_module(plan.__package__ + "._build")
    path = os.path.join(os.path.dirname(plan.__file__), "tune_cache.json")
    if not os.path.exists(path):
        pytest.skip("no launch-shape cache shipped")
    sections = list(json.load(open(path)))
    assert sections and all(s.endswith("|" + b.kernel_rev()) for s in sections), (sections, b.kernel_rev())
    entries = json.load(open(path))[sections[0]]
    assert all(len(v) in (11, 12) for v in entries.values())
