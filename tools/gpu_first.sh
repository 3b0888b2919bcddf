#!/bin/bash
# First GPU bring-up: every stage in its own process with a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/bringup.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
NC=$(python -c "import sys; sys.path.insert(0,'.'); from tests.conv_util import CASES; print(len(CASES))")
for i in $(seq 0 $((NC-1))); do
  timeout 180 python tools/gpu_debug.py conv $i >> $LOG 2>&1 || echo "conv[$i] exit=$?" >> $LOG
done
timeout 300 python tools/gpu_debug.py post >> $LOG 2>&1 || echo "post exit=$?" >> $LOG
timeout 300 python tools/gpu_debug.py model tiny_p6 128 128 2 >> $LOG 2>&1 || echo "model tiny_p6 exit=$?" >> $LOG
timeout 300 python tools/gpu_debug.py model tiny 96 160 1 >> $LOG 2>&1 || echo "model tiny exit=$?" >> $LOG
timeout 300 python tools/gpu_debug.py model nano 416 416 1 >> $LOG 2>&1 || echo "model nano exit=$?" >> $LOG
timeout 600 python tools/gpu_debug.py model yolox_m_p6 640 640 2 >> $LOG 2>&1 || echo "model m_p6 exit=$?" >> $LOG
grep -E "OK|FAIL|exit=|post |model |max\|d\|" $LOG | tail -60
