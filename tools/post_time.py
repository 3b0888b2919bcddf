"""Decode + NMS only (bench.post_stress_lines) at a given batch: whole-step CUDA-event times of both distributions.
Under `ncu --metrics gpu__time_duration.sum` it gives the per-kernel split (select / sort / nms).
    python tools/post_time.py <batch> [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
r = bench.post_stress_lines(B, steps, dev, bench.measured_peaks())
for k in ("max_candidate", "clustered"):
    print(f"B={B} {k}: {r[k]['ms_per_step'] * 1e3:.1f} us/step, {r[k]['candidates_per_image']:.0f} candidates, "
          f"{r[k]['detections_per_image']:.1f} detections per image", flush=True)
