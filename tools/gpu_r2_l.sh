#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_evaluator.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
python tools/post_time.py 1; python tools/post_time.py 64
for t in 0 1; do YX_DW_TILE=$t python tools/dw_profile.py 32 640 2>&1 | head -12; done) > gpurun_out/l.log 2>&1
cat gpurun_out/l.log
