#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | grep -E "Error|FAILED|passed|failed|assert" | head -8
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
peaks = bench.measured_peaks()
r = bench.post_stress_lines(64, 20, torch.device("cuda", 0), peaks)
print({k: (round(v["ms_per_step"], 4), round(v["roofline"]["frac"], 3)) for k, v in r.items() if isinstance(v, dict)})
# the selection kernel alone
from yolox_b200 import _capi
import yolox_b200.postprocess as pp
B, A, C = 64, 34000, 80
g = torch.Generator(device="cuda").manual_seed(1)
reg = torch.randn(B, A, 4, device="cuda", generator=g).half(); obj = (torch.randn(B, A, 1, device="cuda", generator=g) * 2 - 2).half()
cls = (torch.randn(B, A, C, device="cuda", generator=g) * 2 - 2).half()
hw = [(160, 160), (80, 80), (40, 40), (20, 20)]
for _ in range(3): pp.detect_main(reg, obj, cls, hw, (8, 16, 32, 64), 0.001, 0.65)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): pp.detect_main(reg, obj, cls, hw, (8, 16, 32, 64), 0.001, 0.65)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "kernel" in e.key: print(e.key[:60], round(e.device_time_total / e.count, 1), "us")
PY
