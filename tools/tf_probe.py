"""Spread of the teacher-forced error ratio for one configuration over repeated engine builds (tuned shapes vary run to run)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_model import _build
from tests.plan_interp import teacher_forced_errors
from oracle import model_ref as mr
torch.set_grad_enabled(False)
name, H, W, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
for rep in range(int(sys.argv[5])):
    os.environ["YX_TUNE"] = "0" if rep == 0 else "1"
    cfg, fused, model = _build(name, H, W, 3)
    x = mr.synth_images(11, B, H, W).cuda().half()
    errs = sorted(teacher_forced_errors(model, x), key=lambda e: -e[2])[:3]
    eng = model.engine_for(x)
    print("rep", rep, "tune", os.environ["YX_TUNE"], [(e[0], e[1], round(e[2], 3), e[3], eng.op_desc(e[0])[:70]) for e in errs], flush=True)
