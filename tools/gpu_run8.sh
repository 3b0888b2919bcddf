#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run8.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -8 gpurun_out/pytest_gpu.log >> $LOG
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_r8.json >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
grep -E "exit=|passed|failed|FAILED|value" $LOG | cut -c1-1500
