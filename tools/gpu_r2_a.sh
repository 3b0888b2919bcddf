#!/bin/bash
# round 2, call A: sparse-MMA probe (metadata layout + rates), then the whole GPU suite with the new parity tests
mkdir -p gpurun_out
timeout 120 ./tools/sp_probe > gpurun_out/sp_probe.txt 2>&1; echo "sp_probe rc=$?"
tail -12 gpurun_out/sp_probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_gpu_r2a.log; echo "pytest rc=${PIPESTATUS[0]}"
tail -15 gpurun_out/pytest_gpu_r2a.log
cp coco-dataset-based-light-weight-fast-object-detection-model_b200/tune_cache.json gpurun_out/tune_cache_r2a.json 2>/dev/null
