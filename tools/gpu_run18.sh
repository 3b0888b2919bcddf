#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run18.log
: > $LOG
timeout 900 python -m pytest tests/test_gpu_conv.py -q -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv exit=$?" >> $LOG
tail -15 gpurun_out/pytest_conv.log >> $LOG
if grep -q "failed\|error" gpurun_out/pytest_conv.log; then cat $LOG; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit=$?" >> $LOG
tail -8 gpurun_out/pytest_gpu.log >> $LOG
YX_TUNE_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_tuned.json > gpurun_out/bench18.json 2> gpurun_out/tune18.log; echo "bench exit=$?" >> $LOG
cat gpurun_out/bench18.json >> $LOG
YX_TUNE=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_untuned.json >> $LOG 2>&1; echo "bench untuned exit=$?" >> $LOG
cat $LOG | cut -c1-1500
grep "^tune op" gpurun_out/tune18.log | head -130
