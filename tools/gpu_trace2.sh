#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_trace.sh > /dev/null 2>&1
grep -A6 "^CASE\|^trace" gpurun_out/trace.log | grep -E "CASE|trace:|^ +[0-9]+ " | awk '/CASE/{print} /trace:/{print} /^ +(0|1|2|3|9|10) /{print}'
LOG=gpurun_out/bench4.log
: > $LOG
for ai in 300 1000000000; do
  echo "== YX_MEM_AI=$ai" >> $LOG
  YX_MEM_AI=$ai timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_b_ai$ai.json >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
done
grep -E "==|value|exit" $LOG | cut -c1-220
