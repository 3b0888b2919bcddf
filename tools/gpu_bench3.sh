#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/bench3.log
: > $LOG
for ai in 0 300 1000000000; do
  echo "== YX_MEM_AI=$ai" >> $LOG
  YX_MEM_AI=$ai timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_ai$ai.json >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
done
timeout 600 python tools/gpu_debug.py drill yolox_m_p6 640 640 2 >> $LOG 2>&1 || echo "drill exit=$?" >> $LOG
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -12 gpurun_out/pytest_gpu.log >> $LOG
grep -E "==|value|drill|ratio|exit|passed|failed" $LOG | cut -c1-600
