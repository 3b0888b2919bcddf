"""GPU parity of the whole engine (all kernels, through the public model classes) against the CPU oracle.
Tolerance (stated per BASELINE north_star): pre-NMS head logits of the fp16 engine vs the fp32 oracle
evaluated with the same fp16-rounded weights: relative L2 error <= 0.10 and mean |d| <= 0.10 per output
tensor (logits of O(1..30); fp16 activations through ~60 sequential layers — a CPU emulation of the fp16
roundings alone gives rel-L2 ~1e-2 on the synthetic M-P6).  The strict gate is test_every_op_teacher_forced:
each launch equals fp32 on the same inputs up to one fp16 rounding step."""
import glob
import os
import re

import numpy as np
import pytest
import torch

import yolox_b200 as yb
from oracle import model_ref as mr

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
REL_L2, MEAN_ABS = 0.10, 0.10


def _build(name, H, W, seed, flavour="infer"):
    cfg = mr.CONFIGS[name]
    train = mr.synth_train_state(cfg, seed, calib_hw=(H, W))
    fused = mr.fold_bn(train)
    cls_ = {"p6": yb.infer.YOLOXP6v2 if cfg.v2 else yb.infer.YOLOXP6, "dw": yb.infer.YOLOXDepthwise}.get(cfg.kind, yb.infer.YOLOX)
    model = cls_(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(fused, strict=True)
    return cfg, fused, model.cuda().half()


def _check(a, b, what):
    a = a.float().cpu()
    err = (a - b).abs()
    assert not torch.isnan(a).any(), what
    rel = float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
    assert rel <= REL_L2, \
        f"{what}: rel-L2 {rel:.4g} mean {float(err.mean()):.4g} max {float(err.max()):.4g}"


def _q16(sd):
    return {k: (v.half().float() if k.endswith("weight") else v) for k, v in sd.items()}


def _rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


ANCHOR_RATIO, ANCHOR_FLOOR = 1.5, 2e-3


def _torch_fp16_cuda(fused, cfg, x):
    """The yardstick: the reference's op sequence in torch fp16 on CUDA (cuDNN convs with fp32 accumulation, one fp16
    rounding after every conv and every activation) -- what the reference itself computes with `.cuda().half()`."""
    sd = {k: v.cuda().half() for k, v in fused.items()}
    return mr.forward_raw(sd, cfg, x.cuda().half())


def _check_anchored(got, want32, torch16, what):
    """engine_err <= ANCHOR_RATIO * torch_fp16_err (+ a floor for tensors both sides reproduce almost exactly): the engine
    may not be further from the fp32 oracle than 1.5x what torch's own fp16 CUDA path is on the same weights and input."""
    e, t = _rel_l2(got, want32), _rel_l2(torch16, want32)
    assert not torch.isnan(got.float()).any(), what
    assert e <= ANCHOR_RATIO * t + ANCHOR_FLOOR, f"{what}: engine rel-L2 {e:.4g} vs torch fp16 CUDA {t:.4g} (ratio {e / max(t, 1e-12):.2f})"
    return e, t


@pytest.mark.parametrize("name,H,W,B", [("tiny_p6", 128, 192, 2), ("tiny", 96, 160, 3), ("yolox_m_p6", 320, 320, 2),
                                        ("yolox_m", 384, 384, 1), ("tiny_p6_v2", 128, 192, 2), ("tiny_dw", 96, 160, 2)])
def test_infer_logits_match_oracle(name, H, W, B):
    cfg, fused, model = _build(name, H, W, 3)
    x = mr.synth_images(11, B, H, W)
    reg, obj, cls = model(x.cuda().half())
    rr, ro, rc = mr.forward_raw(_q16(fused), cfg, x.half().float())
    _check(reg, rr, "reg"); _check(obj, ro, "obj"); _check(cls, rc, "cls")
    tr, to, tc = _torch_fp16_cuda(fused, cfg, x)          # the anchor of the whole-network tolerance
    _check_anchored(reg, rr, tr, "reg"); _check_anchored(obj, ro, to, "obj"); _check_anchored(cls, rc, tc, "cls")
    # graph replay gives identical bits
    eng, reg8, cls2 = model.run_engine(x.cuda().half(), use_graph=True)
    assert torch.equal(reg8[..., :4], reg) and torch.equal(cls2[..., :cfg.num_classes], cls)


@pytest.mark.parametrize("name,H,W,B", [("tiny_p6", 128, 192, 2), ("tiny", 96, 160, 3), ("yolox_m_p6", 320, 320, 1)])
def test_stand_alone_s2d_path(name, H, W, B, monkeypatch):
    """YX_FUSE_S2D=0: the stand-alone space-to-depth kernel + the row-packed stem conv (what runs when the image-fed stem does
    not apply, e.g. an image width that is not a multiple of 16) -- logits vs the oracle with the usual anchor, every op
    teacher-forced, and bit-identity of the network's FIRST activation with the image-fed stem for fp16 and uint8 input
    (both build the same fp16 s2d values and sum the same products; the launch shapes differ, so only the op-level bound
    applies further down)."""
    from tests.plan_interp import teacher_forced_errors
    cfg, fused, model = _build(name, H, W, 3)
    x = mr.synth_images(11, B, H, W)
    rr, ro, rc = mr.forward_raw(_q16(fused), cfg, x.half().float())
    tr, to, tc = _torch_fp16_cuda(fused, cfg, x)
    for fuse in ("0", "1"):
        monkeypatch.setenv("YX_FUSE_S2D", fuse)
        monkeypatch.setenv("YX_TUNE", "0")
        model.invalidate_engines()
        reg, obj, cls = model(x.cuda().half())
        kinds = [op.kind for op in model.engine_for(x.cuda().half()).graph.ops]
        assert (kinds[0] == 1) == (fuse == "0"), "op 0 must be the s2d kernel exactly when the fusion is off"
        _check_anchored(reg, rr, tr, f"reg fuse={fuse}"); _check_anchored(cls, rc, tc, f"cls fuse={fuse}")
        bad = [e for e in teacher_forced_errors(model, x.cuda().half()) if e[2] > 2.0]
        assert not bad, (fuse, bad[:5])
        u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device="cuda")          # uint8 batches: same values as fp16
        a = model(u8)
        b = model(u8.half())
        assert all(torch.equal(p, q) for p, q in zip(a, b)), f"uint8 and fp16 input disagree (fuse={fuse})"


@pytest.mark.parametrize("name,H,W,B", [("yolox_m_p6", 320, 320, 1), ("tiny_p6", 128, 192, 2), ("tiny", 96, 160, 3), ("tiny_dw", 128, 96, 1)])
def test_lanes_equal_single_stream(name, H, W, B, monkeypatch):
    """Small batches run the op list over several streams (the head's pyramid levels and cls / reg branches side by side),
    ordered by events derived from the arena ranges each op reads and writes.  The same kernels run either way, so the
    logits must be bit-identical to the single-stream run -- eagerly, from the engine's CUDA graph, and over repeated runs
    (a missing dependency would show as an intermittent difference)."""
    cfg, fused, model = _build(name, H, W, 7)
    x = mr.synth_images(13, B, H, W).cuda().half()
    monkeypatch.setenv("YX_TUNE", "0")
    monkeypatch.setenv("YX_LANES", "0")
    model.invalidate_engines()
    _, reg8, cls = model.run_engine(x)
    ref = (reg8.clone(), cls.clone())
    monkeypatch.setenv("YX_LANES", "1")
    model.invalidate_engines()
    for use_graph in (False, True):                                # eagerly, then from the engine's own CUDA graph
        for it in range(6):
            _, reg8, cls = model.run_engine(x, 1.0, 0.0, use_graph)
            torch.cuda.synchronize()
            assert torch.equal(reg8, ref[0]) and torch.equal(cls, ref[1]), f"run {it} (graph={use_graph}) differs from the single-stream result"


def test_depthwise_kernels_agree(monkeypatch):
    """The three stride-1 depthwise kernels (shared-memory tile, register-blocked strip, one thread per output) sum the
    same products in the same order: the whole depthwise model's logits must be bit-identical whichever runs."""
    cfg, fused, model = _build("tiny_dw", 128, 96, 5)
    x = mr.synth_images(12, 2, 128, 96).cuda().half()
    monkeypatch.setenv("YX_TUNE", "0")
    outs = []
    for tile, strip in (("1", "1"), ("0", "1"), ("0", "0")):
        monkeypatch.setenv("YX_DW_TILE", tile)
        monkeypatch.setenv("YX_DW_STRIP", strip)
        model.invalidate_engines()
        outs.append([t.clone() for t in model(x)])
    kinds = [op.kind for op in model.engine_for(x).graph.ops]
    assert 4 in kinds, "the depthwise model must contain depthwise ops"
    for other in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(outs[0], other))


@pytest.mark.parametrize("name,H,W,seed", [("tiny_p6_v2", 128, 128, 4), ("tiny_dw", 96, 128, 5)])
def test_variant_reference_golden(name, H, W, seed):
    """P6-v2 / depthwise inference twins: the engine vs the reference's raw logits (tests/golden/infer_*.npz)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", f"infer_{name}_{H}x{W}_b1_s{seed}.npz"))
    cfg, fused, model = _build(name, H, W, seed)
    reg, obj, cls = model(torch.from_numpy(g["x"]).cuda().half())
    _check(reg, torch.from_numpy(g["reg"]), "reg"); _check(obj, torch.from_numpy(g["obj"]), "obj")
    _check(cls, torch.from_numpy(g["cls"]), "cls")


@pytest.mark.parametrize("name,H,W,B", [("yolox_m_p6", 320, 320, 2), ("yolox_m_p6", 640, 384, 3), ("yolox_m", 416, 416, 2),
                                        ("yolox_m_p6_v2", 256, 256, 2), ("tiny_dw", 160, 128, 5), ("yolox_m_p6", 1280, 1280, 4)])
def test_every_tuning_candidate_agrees(name, H, W, B, monkeypatch):
    """yx_engine_tune self-check: on every conv of a real network, EVERY candidate launch shape (generic / halo / pair /
    256-pixel tiles / resident or streamed weights / ...) reproduces the default shape's output."""
    monkeypatch.setenv("YX_TUNE_CHECK", "1")
    cfg, fused, model = _build(name, H, W, 3)
    x = mr.synth_images(11, B, H, W).cuda().half()
    model(x)
    bad = model.engine_for(x).tune_mismatches()
    assert not bad, bad[:8]


GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "model_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_reference_golden_vectors(path):
    """Outputs of the UNMODIFIED reference (tests/golden/make_golden.py) vs the engine."""
    m = re.match(r"model_(.+)_(\d+)x(\d+)_b(\d+)_s(\d+)\.npz", os.path.basename(path))
    name, H, W, B, seed = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))
    g = np.load(path)
    cfg = mr.CONFIGS[name]
    train = mr.synth_train_state(cfg, seed, calib_hw=(H, W))
    x = torch.from_numpy(g["x"])
    # yolox flavour: conv+BN modules, [B,A,85] output, decode_in_inference on/off
    if cfg.kind == "p6":
        backbone = yb.models.YOLOPAFPNCustomP6(cfg.depth, cfg.width, act=cfg.act, in_channels=[256, 512, 768, 1024])
        head = yb.models.YOLOXHeadCustom(cfg.num_classes, cfg.width, act=cfg.act, strides=(8, 16, 32, 64),
                                         in_channels=[256, 512, 768, 1024])
        model = yb.models.YOLOXCustomP6(backbone, head)
    else:
        backbone = yb.models.YOLOPAFPN(cfg.depth, cfg.width, in_channels=[256, 512, 1024], act=cfg.act,
                                       depthwise=cfg.depthwise_neck)
        head = yb.models.YOLOXHead(cfg.num_classes, cfg.width, in_channels=[256, 512, 1024], act=cfg.act)
        model = yb.models.YOLOX(backbone, head)
    sd = dict(train)
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = v
    model.load_state_dict(sd, strict=True)
    model = model.eval().cuda()
    model.head.decode_in_inference = False
    und = model(x.cuda())                       # fp32 in -> fp32 out (engine computes in fp16)
    assert und.dtype == torch.float32 and tuple(und.shape) == g["yolox_undecoded"].shape
    ref = torch.from_numpy(g["yolox_undecoded"])
    _check(und[..., :4], ref[..., :4], "undecoded box logits")
    assert float((und[..., 4:].cpu() - ref[..., 4:]).abs().mean()) < 0.02   # sigmoid outputs (fp16 weights + activations)
    model.head.decode_in_inference = True
    dec = model(x.cuda())
    refd = torch.from_numpy(g["yolox_decoded"])
    rel = (dec[..., :4].cpu() - refd[..., :4]).abs() / (refd[..., :4].abs() + 8.0)
    assert float(rel.mean()) < 0.02 and float(rel.median()) < 0.01, (float(rel.mean()), float(rel.max()))
    assert model.head.hw == [tuple(hw) for hw in mr.level_hw(cfg, H, W)]
    if "reg" in g:                              # inference twin, raw logits
        cfg2, fused, im = _build(name, H, W, seed)
        reg, obj, cls = im(x.cuda().half())
        _check(reg, torch.from_numpy(g["reg"]), "reg"); _check(cls, torch.from_numpy(g["cls"]), "cls")


def test_nano_416_config1():
    """BASELINE config 1 geometry (416x416, bs1, depthwise neck, odd 13x13 / 26x26 / 52x52 maps)."""
    cfg = mr.CONFIGS["nano"]
    H = W = 416
    train = mr.synth_train_state(cfg, 2, calib_hw=(H, W))
    backbone = yb.models.YOLOPAFPN(cfg.depth, cfg.width, in_channels=[256, 512, 1024], act=cfg.act, depthwise=True)
    head = yb.models.YOLOXHead(cfg.num_classes, cfg.width, in_channels=[256, 512, 1024], act=cfg.act)
    model = yb.models.YOLOX(backbone, head)
    sd = dict(train)
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = v
    model.load_state_dict(sd, strict=True)
    model = model.eval().cuda().half()
    model.head.decode_in_inference = False
    x = mr.synth_images(4, 1, H, W)
    out = model(x.cuda().half())
    ref = mr.forward_yolox(_q16(mr.fold_bn(train)), cfg, x.half().float(), decode=False)
    _check(out[..., :4], ref[..., :4], "nano box logits")
    assert float((out[..., 4:].float().cpu() - ref[..., 4:]).abs().mean()) < 0.02


def test_predict_loop_end_to_end():
    """main.py:160-188 sequence through the public functions; detections vs oracle on the same logits."""
    from oracle import post_ref as pr
    cfg, fused, model = _build("tiny_p6", 256, 256, 7)
    x = mr.synth_images(21, 2, 256, 256)
    img = x.cuda().half()
    img.mul_(0.9).add_(11.4)
    reg, obj, cls = model(img)
    grids, scales = yb.postprocess.yolox_generate_grid((256, 256), model.head.strides, torch.float16)
    boxes, oc, cc = yb.postprocess.yolox_postprocess_output_torch_batch(reg, obj, cls, grids.cuda(), scales.cuda())
    outs = yb.postprocess.yolox_nms_torch_batch(boxes, oc, cc, nms_threshold=0.55, conf_threshold=0.05)
    for i, d in enumerate(outs):
        ref, _ = pr.nms_image_main(boxes[i].cpu().numpy(), oc[i].cpu().numpy(), cc[i].cpu().numpy(), 0.05, 0.55)
        got = d.cpu().numpy() if d is not None else np.zeros((0, 7), np.float32)
        np.testing.assert_array_equal(got, ref)
    # fused input affine == explicit mul_/add_ in half
    eng, reg8, cls2 = model.run_engine(x.cuda().half(), in_scale=0.9, in_shift=11.4)
    assert torch.equal(reg8[..., :4], reg)


@pytest.mark.parametrize("name,H,W,B", [("yolox_m_p6", 640, 640, 2), ("yolox_m", 320, 320, 1), ("tiny", 160, 96, 2),
                                        ("yolox_m", 640, 640, 2),      # BASELINE config 2 geometry (YOLOX-M 640x640)
                                        ("yolox_l", 640, 640, 1),      # BASELINE config 5 geometry (YOLOX-L 640x640)
                                        ("tiny_p6_v2", 128, 128, 1), ("yolox_m_p6_v2", 256, 256, 1), ("tiny_dw", 128, 96, 1),
                                        ("yolox_l_dw", 256, 256, 1)])
def test_every_op_teacher_forced(name, H, W, B):
    """Each of the ~125 launches of a real network, run one at a time, equals a torch fp32 evaluation of the
    SAME device inputs to within one fp16 rounding step of the pre-activation sum, x2 slack: the deepest reductions
    (K = 8192, the 4x4 stride-2 512->512 conv of the depthwise-L model) sit at 1.6 -- fp32 accumulation order over 8192
    terms plus the two fp16 roundings; every other launch of every configuration is at or below 1.0."""
    from tests.plan_interp import teacher_forced_errors
    cfg, fused, model = _build(name, H, W, 3)
    x = mr.synth_images(11, B, H, W).cuda().half()
    errs = teacher_forced_errors(model, x)
    bad = [e for e in errs if e[2] > 2.0]
    assert not bad, bad[:5]


@pytest.mark.parametrize("H,W,C", [(20, 20, 384), (5, 3, 8), (40, 40, 72), (64, 48, 16)])
def test_spp_pools_exact(H, W, C):
    """SPP (network_blocks.py:239-246): slices 1..3 of the concat buffer = max_pool2d 5/9/13 of slice 0, exactly (max is
    exact in fp16).  Covers the tiled kernel with 4 and 1 channel vectors per CTA and the direct kernel for large maps."""
    import torch.nn.functional as F
    from yolox_b200 import plan
    B = 2
    g = plan.Graph(B, 2 * H, 2 * W)
    img = g.new_buf("img16", H, W, 16)
    cat = g.new_buf("cat", H, W, 4 * C)
    g.s2d(img.view(), "unshuffle")
    g.spp(cat.view(0, C), cat.view(C, 3 * C))
    g.place_buffers()
    g.weight_blob = torch.zeros(128, dtype=torch.float16)
    g.bias_blob = torch.zeros(64, dtype=torch.float32)
    eng = plan.Engine(g, "cuda")
    x = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(H * W + C)).half().cuda()
    t = eng.tensor_of(cat)
    t.zero_()
    t[..., :C] = x
    eng.run_ops(torch.zeros(B, 3, 2 * H, 2 * W, dtype=torch.float16, device="cuda"), 1, 1)
    xn = x.permute(0, 3, 1, 2).float()
    for i, k in enumerate((5, 9, 13)):
        want = F.max_pool2d(xn, k, 1, k // 2).permute(0, 2, 3, 1).half()
        assert torch.equal(t[..., (i + 1) * C:(i + 2) * C], want), f"pool {k}"
    assert torch.equal(t[..., :C], x)


def test_predictor_whole_step_graph_equals_eager():
    """Predictor(whole_graph=True): network + decode + NMS replayed as one CUDA graph gives the eager result, for
    successive different inputs and two shapes."""
    cfg, fused, model = _build("tiny_p6", 128, 128, 9)
    eager = yb.predict.Predictor(model, conf_threshold=0.05, nms_threshold=0.55)
    graphed = yb.predict.Predictor(model, conf_threshold=0.05, nms_threshold=0.55, whole_graph=True)
    for seed, (B, H, W) in enumerate([(1, 128, 128), (1, 128, 128), (2, 128, 192), (1, 128, 128)]):
        x = mr.synth_images(30 + seed, B, H, W).cuda().half()
        d0, c0 = eager(x)
        d1, c1 = graphed(x)
        assert torch.equal(c0, c1) and torch.equal(d0, d1), (seed, c0.tolist(), c1.tolist())
    assert len(graphed._step_graphs) == 2


def test_full_size_batch_independence(monkeypatch):
    """BASELINE.json's full configuration (YOLOX-M-P6, 1280x1280, 64 images per step): a size-independent property in place
    of an oracle run that would take minutes -- no operator on the path mixes images, so every image's logits and
    detections in the bs64 step are BIT-identical to the same image run alone (shape-only launch shapes: YX_TUNE=0;
    tuned shapes may order the K loop differently between batch sizes)."""
    monkeypatch.setenv("YX_TUNE", "0")
    cfg = mr.CONFIGS["yolox_m_p6"]
    model = yb.infer.YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    g = torch.Generator().manual_seed(5)
    for p in model.parameters():          # random weights, default-initialised prediction biases
        with torch.no_grad():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 / max(p[0].numel(), 1) ** 0.5) if p.dim() == 4
                    else torch.randn(p.shape, generator=g) * 0.1)
    model = model.cuda().half().eval()
    B, S = 64, 1280
    x = (torch.rand(B, 3, S, S, device="cuda", generator=torch.Generator("cuda").manual_seed(6)) * 255).half()
    pred = yb.predict.Predictor(model, conf_threshold=0.001, nms_threshold=0.65)
    eng, reg8, cls = model.run_engine(x, 0.9, 11.4)
    reg8, cls = reg8.clone(), cls.clone()
    assert torch.isfinite(reg8.float()).all() and torch.isfinite(cls.float()).all()
    det, cnt = pred(x)
    assert int(cnt.min()) > 0, "random-init predictions should produce candidates (sigmoid ~ 0.5)"
    for i in (0, 17, 63):
        e1, r1, c1 = model.run_engine(x[i:i + 1].contiguous(), 0.9, 11.4)
        assert torch.equal(r1[0], reg8[i]) and torch.equal(c1[0], cls[i]), f"image {i}: logits differ from the bs1 run"
        d1, n1 = pred(x[i:i + 1].contiguous())
        assert int(n1[0]) == int(cnt[i]) and torch.equal(d1[0], det[i]), f"image {i}: detections differ from the bs1 run"


def test_headline_config_against_oracle():
    """BASELINE.json's headline workload pinned against the oracle: pruned YOLOX-M-P6 (hard-swish, 49 % global-magnitude
    masks over the non-head convs, dense-with-zeros) at 1280x1280.  Three images' raw head logits vs the fp32 oracle on
    the same fp16-rounded weights, with the tolerance ANCHORED on what torch fp16 CUDA (the reference's own `.half()` path,
    merge_save_p6.py:29-45 is the reference's template for this comparison) scores against the same oracle; the
    detections vs oracle/post_ref.c on the engine's logits, bit-exact.  Together with test_full_size_batch_independence
    (bs64 == bs1 bit for bit) this pins the bench step."""
    from oracle import post_ref as pr
    cfg = mr.CONFIGS["yolox_m_p6"]
    S, B = 1280, 3
    train = mr.synth_train_state(cfg, 0, calib_hw=(S, S))
    fused = mr.apply_masks(mr.fold_bn(train), mr.magnitude_masks(train, 49.0))
    model = yb.infer.YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(fused, strict=True)
    model = model.cuda().half()
    dens = yb.weights.density({k: v for k, v in model.state_dict().items() if "head" not in k})
    assert 0.49 < dens < 0.53, dens
    x = mr.synth_images(41, B, S, S)
    reg, obj, cls = model(x.cuda().half())
    hw = mr.level_hw(cfg, S, S)
    assert reg.shape == (B, 34000, 4) and cls.shape == (B, 34000, 80) and [h * w for h, w in hw] == [25600, 6400, 1600, 400]
    rr, ro, rc = mr.forward_raw(_q16(fused), cfg, x.half().float())
    tr, to, tc = _torch_fp16_cuda(fused, cfg, x)
    report = []
    for name, a, b, t in (("reg", reg, rr, tr), ("obj", obj, ro, to), ("cls", cls, rc, tc)):
        _check(a, b, name)
        for i in range(B):                      # per image, so one good image cannot hide a bad one
            report.append((name, i) + _check_anchored(a[i], b[i], t[i], f"{name}[{i}]"))
    print("headline parity (tensor, image, engine rel-L2, torch-fp16 rel-L2):", report)
    # per pyramid level (the 160x160 / 80x80 maps of the real workload are covered by nothing smaller)
    off = 0
    for (h, w) in hw:
        sl = slice(off, off + h * w)
        _check_anchored(cls[:, sl], rc[:, sl], tc[:, sl], f"cls level {h}x{w}")
        _check_anchored(reg[:, sl], rr[:, sl], tr[:, sl], f"reg level {h}x{w}")
        off += h * w
    det, cnt, anc = yb.postprocess.detect_main(reg, obj, cls, hw, cfg.strides, 0.001, 0.65, 5000, 300)
    for i in range(B):
        boxes, oc, cc = pr.decode_infer(reg[i].cpu().numpy(), obj[i].cpu().numpy(), cls[i].cpu().numpy(), hw, cfg.strides)
        gb, go, gc = yb.postprocess.decode_infer(reg[i:i + 1], obj[i:i + 1], cls[i:i + 1], hw, cfg.strides)
        np.testing.assert_allclose(gb[0].cpu().numpy(), boxes, rtol=3e-6, atol=2e-3)
        n_cand = int((gc[0].max(-1).values >= 0.001).sum())
        d_ref, _ = pr.nms_image_main(gb[0].cpu().numpy(), go[0].cpu().numpy(), gc[0].cpu().numpy(), 0.001, 0.65, 5000, 300,
                                     pr.torchvision_mode(min(n_cand, 5000), "cuda"))
        n = int(cnt[i])
        assert n == len(d_ref) and n > 0, (n, len(d_ref))
        np.testing.assert_array_equal(det[i, :n].cpu().numpy(), d_ref)


@pytest.mark.parametrize("masks", ["magnitude49", "two_four"])
def test_pruned_checkpoint_through_build_yolox(masks, tmp_path):
    """BASELINE config 3 (pruned model): main.py:31-59's path -- a {"model": sparse COO} checkpoint written to disk, ingested
    by build_yolox(sparse=True), run dense-with-zeros -- against the oracle evaluated on the same masked weights; both the
    49 % global-magnitude masks of the shipped recipe (01_mask_generator.py) and the synthetic 2:4-compliant set."""
    cfg = mr.CONFIGS["tiny_p6"]
    H, W, B = 128, 192, 2
    train = mr.synth_train_state(cfg, 8, calib_hw=(H, W))
    m = mr.magnitude_masks(train, 49.0) if masks == "magnitude49" else mr.two_four_masks(train)
    fused = mr.apply_masks(mr.fold_bn(train), m)
    path = str(tmp_path / "pruned.pth")
    torch.save(mr.to_sparse_ckpt(fused), path)
    model = yb.predict.build_yolox(dict(model=dict(type="yolox-p6", depth=cfg.depth, width=cfg.width), ckpt=path, sparse=True,
                                        half=True))
    for k, v in model.state_dict().items():               # the ingest reproduced the dense-with-zeros weights
        assert torch.equal(v.float().cpu(), fused[k].half().float()), k
    dens = yb.weights.density({k: v for k, v in model.state_dict().items() if "head" not in k})
    assert 0.4 < dens < 0.6
    x = mr.synth_images(12, B, H, W)
    reg, obj, cls = model(x.cuda().half())
    rr, ro, rc = mr.forward_raw(_q16(fused), cfg, x.half().float())
    _check(reg, rr, "reg"); _check(obj, ro, "obj"); _check(cls, rc, "cls")
    eng = model.engine_for(x.cuda().half())
    n_sparse = sum("sparse24" in eng.op_desc(i) for i in range(len(eng.graph.ops)))
    if masks == "magnitude49":
        assert n_sparse == 0, "unstructured 49 % masks are never 2:4-compliant: dense-with-zeros is the only path"


@pytest.mark.parametrize("name,H,W,B", [("tiny_p6", 128, 192, 2), ("yolox_m_p6", 320, 320, 2), ("yolox_m_p6", 640, 640, 1)])
def test_two_four_masks_run_on_sparse_tensor_cores(name, H, W, B, monkeypatch):
    """north_star (1) / SURVEY 8d config 3: a 2:4-compliant mask set runs on the sparse tensor-core path.  With
    YX_SPARSE=force every conv whose packed weights are compliant and whose geometry the variant supports launches the
    tcgen05.mma.sp kernel (asserted from the engine's own launch-shape descriptions); logits vs the oracle on the same
    masked weights, anchored on torch fp16 CUDA like every whole-network comparison; and the dense-with-zeros run of
    the same weights (YX_SPARSE=0) must agree with the sparse run to within the same anchor."""
    cfg = mr.CONFIGS[name]
    train = mr.synth_train_state(cfg, 8, calib_hw=(H, W))
    fused = mr.apply_masks(mr.fold_bn(train), mr.two_four_masks(train))
    x = mr.synth_images(12, B, H, W)
    rr, ro, rc = mr.forward_raw(_q16(fused), cfg, x.half().float())
    tr, to, tc = _torch_fp16_cuda(fused, cfg, x)
    outs = {}
    for mode in ("force", "0"):
        monkeypatch.setenv("YX_SPARSE", mode)
        monkeypatch.setenv("YX_TUNE", "0")
        model = yb.infer.YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
        model.load_state_dict(fused, strict=True)
        model = model.cuda().half()
        reg, obj, cls = model(x.cuda().half())
        eng = model.engine_for(x.cuda().half())
        descs = [eng.op_desc(i) for i in range(len(eng.graph.ops))]
        n_sparse = sum("sparse24" in d for d in descs)
        if mode == "force":
            # every non-head conv with cin % 32 == 0 (and metadata that fits) is eligible; the head is dense (unmasked)
            eligible = [op for op in eng.graph.ops if op.kind == 0 and not op.name.startswith("head") and op.cin_pad % 32 == 0
                        and op.aux == 0 and op.up is None
                        and -(-op.cout_pad // 128) * op.ksize * op.ksize * (op.cin_pad // 32) <= 256]
            assert n_sparse == len(eligible) and n_sparse >= 10, (n_sparse, len(eligible))
            assert not any("sparse24" in d for d, op in zip(descs, eng.graph.ops) if op.name.startswith("head"))
        else:
            assert n_sparse == 0
        _check_anchored(reg, rr, tr, f"reg (YX_SPARSE={mode})")
        _check_anchored(obj, ro, to, f"obj (YX_SPARSE={mode})")
        _check_anchored(cls, rc, tc, f"cls (YX_SPARSE={mode})")
        outs[mode] = (reg.float().cpu(), cls.float().cpu())
    for a, b, t, want in zip(outs["force"], outs["0"], (tr, tc), (rr, rc)):
        assert _rel_l2(a, b) <= ANCHOR_RATIO * _rel_l2(t, want) + ANCHOR_FLOOR
