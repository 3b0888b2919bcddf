"""NOT a pytest module (no test_ prefix): measures the torch + cuDNN fp16 throughput of the SAME op sequence
the reference's nn.Modules execute (the oracle restatement run on CUDA, NCHW, cudnn.benchmark=True, default
stream) — the SURVEY §8d 'reference GPU baseline' — to be read next to bench.py's numbers.  The reference
tree itself does not travel to the GPU box; the oracle is pinned against it by tests/golden.  Usage:
    python tests/perf_reference_gpu.py [batch] [size] [iters]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchvision

from oracle import model_ref as mr

torch.set_grad_enabled(False)
torch.backends.cudnn.benchmark = True
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = mr.CONFIGS["yolox_m_p6"]
sd = {k: v.cuda().half() for k, v in mr.fold_bn(mr.synth_train_state(cfg, 0, calibrate=False)).items()}
x = (torch.rand(B, 3, S, S, device="cuda") * 255).half()
out = {}
for fmt in ("nchw", "channels_last"):
    xi = x.contiguous(memory_format=torch.channels_last) if fmt == "channels_last" else x
    sdi = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 and fmt == "channels_last" else v)
           for k, v in sd.items()}
    for _ in range(5):
        mr.forward_raw(sdi, cfg, xi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        reg, obj, cls = mr.forward_raw(sdi, cfg, xi)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out[fmt] = dict(forward_ms_per_step=ms, forward_images_per_s=B / ms * 1e3)


def ref_post(reg, obj, cls):
    """decode + per-image torchvision NMS loop, as postprocess_utils.py:27-129 does it."""
    hw = mr.level_hw(cfg, S, S)
    grids, strides = mr.grids_and_strides(hw, cfg.strides, torch.float16)
    grids, strides = grids.cuda(), strides.cuda()
    reg, obj, cls = reg.float(), obj.float(), cls.float()
    reg[..., :2].add_(grids).mul_(strides)
    reg[..., 2:].exp_().mul_(strides / 2)
    boxes = torch.stack([reg[..., 0] - reg[..., 2], reg[..., 1] - reg[..., 3], reg[..., 0] + reg[..., 2],
                         reg[..., 1] + reg[..., 3]], -1)
    oc = obj.sigmoid_()
    cc = cls.sigmoid_() * oc
    res = []
    for i in range(B):
        s, l = torch.max(cc[i], -1, keepdim=True)
        m = s.squeeze(-1) >= 0.001
        det = torch.cat((boxes[i], oc[i], s, l.float()), 1)[m]
        if det.size(0) > 5000:
            det = det[torch.argsort(det[:, 5], descending=True)[:5000]]
        keep = torchvision.ops.batched_nms(det[:, :4], det[:, 5], det[:, 6], 0.65)[:300]
        res.append(det[keep])
    return res


reg, obj, cls = mr.forward_raw(sd, cfg, x)
for _ in range(2):
    ref_post(reg.clone(), obj.clone(), cls.clone())
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    ref_post(reg.clone(), obj.clone(), cls.clone())
torch.cuda.synchronize()
out["postprocess_ms_per_step"] = (time.perf_counter() - t0) / 3 * 1e3
out["config"] = dict(batch=B, size=S, dtype="fp16", cudnn_benchmark=True, torch=torch.__version__,
                     note="random-init (uncalibrated) weights, ~all anchors are candidates")
print(json.dumps(out))
