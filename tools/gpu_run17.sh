#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/run17.log
: > $LOG
echo "=== tma_probe" >> $LOG
timeout 300 tools/tma_probe > gpurun_out/tma_probe.txt 2>&1; echo "probe exit=$?" >> $LOG
echo "=== traces" >> $LOG
for cfg in "YX_HALO=1" "YX_HALO=0"; do
echo "--- $cfg" >> $LOG
env $cfg YX_CONV_TRACE=1 timeout 300 python - >> $LOG 2>&1 <<'PY'
import sys, os
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case
cases = [dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish"),
         dict(cin=96, cout=192, k=3, stride=2, H=320, W=320, B=8, act="hard_swish"),
         dict(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=8, act="hard_swish", res=True),
         dict(cin=96, cout=96, k=1, stride=1, H=320, W=320, B=16, act="hard_swish"),
         dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish"),
         dict(cin=192, cout=192, k=1, stride=1, H=160, W=160, B=16, act="hard_swish"),
         dict(cin=384, cout=384, k=1, stride=1, H=160, W=160, B=8, act="hard_swish")]
if os.environ.get("YX_HALO") == "0":
    cases = cases[2:3]
for c in cases:
    print("CASE", c, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
done
echo "=== ncu launch list" >> $LOG
YX_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|s2d|spp|upsample|dwconv|select|sort|nms|radix|decode|assemble|topk|hist" -c 400 --csv --log-file gpurun_out/launches_step.csv python tools/ncu_target.py >> $LOG 2>&1
echo "ncu exit=$?" >> $LOG
grep -E "===|---|exit=|CASE|trace:|max_err|^ +(2|3|4) " $LOG | cut -c1-220
cat gpurun_out/tma_probe.txt
