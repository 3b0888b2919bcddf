#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/trace3.log
: > $LOG
for cfg in "YX_HALO=1" "YX_HALO=1 YX_HALO_MH=1" "YX_HALO=0" "YX_HALO=0 YX_MEM_AI=1000000000"; do
echo "=== $cfg" >> $LOG
env $cfg YX_CONV_TRACE=1 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case
for c in [dict(cin=192, cout=384, k=3, stride=1, H=160, W=160, B=8, act="hard_swish"),
          dict(cin=192, cout=192, k=3, stride=1, H=80, W=80, B=32, act="hard_swish")]:
    print("CASE", c, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
done
grep -E "===|CASE|trace:|^ +(1|2|3|4|5) " $LOG
