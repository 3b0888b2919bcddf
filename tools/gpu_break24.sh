#!/bin/bash
mkdir -p gpurun_out
python tools/profile_ops.py 64 1280 gpurun_out/prof_ops64.json 3 2>&1 | tail -2
python - <<'PY' 2>&1 | tail -12
import sys, torch, time
sys.path.insert(0, '.')
import bench
from yolox_b200 import postprocess as pp
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
x = (torch.rand(64, 3, 1280, 1280, device=dev) * 255).half()
for _ in range(3):
    eng, reg8, cls = model.run_engine(x, 0.9, 11.4)
def t(fn, n=10):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
det = lambda: pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :80], model.head.hw, bench.MODEL["strides"], bench.CONF_THR, bench.NMS_THR, bench.MAX_NMS, bench.MAX_DET)
print("forward only ms", t(lambda: model.run_engine(x, 0.9, 11.4)))
print("forward graph ms", t(lambda: model.run_engine(x, 0.9, 11.4, use_graph=True)))
print("detect only ms", t(det))
print("forward+detect ms", t(lambda: (model.run_engine(x, 0.9, 11.4), det())))
PY
