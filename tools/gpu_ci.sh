#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/ci.log
: > $LOG
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv exit=$?" >> $LOG
tail -5 gpurun_out/pytest_conv.log >> $LOG
if grep -q "failed\|error" gpurun_out/pytest_conv.log; then cat $LOG; tail -40 gpurun_out/pytest_conv.log; exit 1; fi
YX_TUNE_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_tuned.json > gpurun_out/bench_ci.json 2> gpurun_out/tune_ci.log; echo "bench exit=$?" >> $LOG
cat gpurun_out/bench_ci.json >> $LOG
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit=$?" >> $LOG
tail -5 gpurun_out/pytest_gpu.log >> $LOG
cat $LOG | cut -c1-600
