#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/trace.log
: > $LOG
python - >> $LOG 2>&1 <<'PY'
import os, sys
sys.path.insert(0, '.')
os.environ["YX_CONV_TRACE"] = "1"
from tests.conv_util import run_conv_case
cases = [dict(cin=96, cout=96, k=1, stride=1, H=320, W=320, B=8, act="hard_swish"),
         dict(cin=64, cout=64, k=1, stride=1, H=320, W=320, B=8, act="hard_swish"),
         dict(cin=96, cout=96, k=3, stride=1, H=160, W=160, B=8, act="hard_swish"),
         dict(cin=16, cout=48, k=3, stride=1, H=640, W=640, B=2, act="hard_swish"),
         dict(cin=192, cout=384, k=3, stride=1, H=160, W=160, B=4, act="hard_swish")]
for c in cases:
    print("CASE", c, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
cat $LOG | head -150
