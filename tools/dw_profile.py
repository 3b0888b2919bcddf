"""Depthwise 5x5 / 3x3 kernels on the depthwise-L model (choijhanyangackr/config/yolox_l_dw.json geometry, 640x640): per-op
CUDA-event time and achieved HBM GB/s (algorithmic bytes = input + output once) against the measured peak.
    YX_DW_STRIP=0|1 python tools/dw_profile.py [batch] [size]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import yolox_b200 as yb

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 640
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = yb.infer.YOLOXDepthwise(1.0, 1.0).to(dev).half().eval()
x = (torch.rand(B, 3, S, S, device=dev) * 255).half()
eng = model.engine_for(x)
for _ in range(3):
    eng.run(x)
torch.cuda.synchronize()
prof = eng.profile(x, iters=10)
peaks = bench.measured_peaks()
dw = [p for p in prof if p["kind"] == 4]
tot_ms, tot_b = sum(p["ms"] for p in dw), sum(p["bytes"] for p in dw)
print(f"YX_DW_STRIP={os.environ.get('YX_DW_STRIP', '1')}  depthwise-L {S}x{S} bs{B}: {len(dw)} depthwise launches, {tot_ms:.3f} ms, "
      f"{tot_b / tot_ms / 1e6:.0f} GB/s aggregate = {tot_b / tot_ms / 1e6 / peaks['hbm_gbs']:.2f} of {peaks['hbm_gbs']:.0f} GB/s; "
      f"whole network {sum(p['ms'] for p in prof):.3f} ms")
for p in dw:
    print(f"  {p['name']:44s} {p['ms']:.4f} ms  {p['bytes'] / 1e6:8.1f} MB  {p['bytes'] / p['ms'] / 1e6:6.0f} GB/s")
