#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run13.log
: > $LOG
for cfg in "YX_HALO=0" "YX_HALO=1" "YX_HALO=1 YX_HALO_MH=1"; do
  echo "== $cfg" >> $LOG
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile_$tag.json >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -6 gpurun_out/pytest_gpu.log >> $LOG
grep -E "==|exit=|passed|failed|FAILED" $LOG | cut -c1-300
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' $LOG
grep -o '"latency_bs1_ms_p50": [0-9.]*' $LOG
