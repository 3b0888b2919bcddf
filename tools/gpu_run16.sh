#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run16.log
: > $LOG
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $LOG
tail -6 gpurun_out/pytest_gpu.log >> $LOG
for cfg in "YX_HALO=0" "YX_HALO=1" "YX_HALO=0 YX_MEM_AI=1000000000"; do
  echo "== $cfg" >> $LOG
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/profile2_$tag.json >> $LOG 2>&1 || echo "bench exit=$?" >> $LOG
done
echo "=== trace" >> $LOG
YX_HALO=0 YX_CONV_TRACE=1 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case
for c in [dict(cin=192, cout=192, k=3, stride=1, H=80, W=80, B=32, act="hard_swish")]:
    print("CASE", c, flush=True)
    r = run_conv_case(**c)
PY
YX_HALO=1 YX_CONV_TRACE=1 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case
for c in [dict(cin=192, cout=192, k=3, stride=1, H=80, W=80, B=32, act="hard_swish"), dict(cin=192, cout=384, k=3, stride=1, H=160, W=160, B=8, act="hard_swish")]:
    print("CASE", c, flush=True)
    r = run_conv_case(**c)
PY
grep -E "==|exit=|passed|failed|FAILED|CASE|trace:|^ +(2|3|4) " $LOG | cut -c1-200
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus"' $LOG
