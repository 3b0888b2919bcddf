"""Short target for ncu: the bench step (forward + fused decode/NMS) on the bench workload, 3 steps."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

torch.set_grad_enabled(False)
B = int(os.environ.get("YX_B", "64"))
S = int(os.environ.get("YX_S", "1280"))
dev = torch.device("cuda", 0)
# YX_MASKS=two_four (+ YX_SPARSE=force): the 2:4 mask set on the sparse tensor-core path; YX_MODEL=dw: the depthwise-L model
import yolox_b200 as yb
if os.environ.get("YX_MODEL") == "dw":
    torch.manual_seed(0)
    model = yb.infer.YOLOXDepthwise(1.0, 1.0).to(dev).half().eval()
    strides = (8, 16, 32)
else:
    model = bench.build_model(dev, masks=os.environ.get("YX_MASKS", "magnitude49"))
    strides = bench.MODEL["strides"]
from yolox_b200 import postprocess as pp
x = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, device=dev)     # the bench's input: uint8 pixels
def step():
    eng, reg8, cls = model.run_engine(x, 0.9, 11.4)
    return pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :80], model.head.hw, strides,
                          bench.CONF_THR, bench.NMS_THR, bench.MAX_NMS, bench.MAX_DET)


for _ in range(int(os.environ.get("YX_STEPS", "3"))):   # first step tunes the launch shapes
    det, cnt, _ = step()
torch.cuda.synchronize()
# profile exactly ONE steady-state step (run ncu with --profile-from-start off)
torch.cuda.cudart().cudaProfilerStart()
det, cnt, _ = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", int(cnt.sum()))
