#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/halo.log
: > $LOG
run() { # env..., then python code
  echo "== $1" >> $LOG
  env $1 timeout 300 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case
cases = [dict(cin=64, cout=64, k=3, stride=1, H=16, W=16, act="none"),
         dict(cin=64, cout=64, k=3, stride=1, H=32, W=24, act="none", B=1),
         dict(cin=96, cout=96, k=3, stride=1, H=48, W=40, act="hard_swish", res=True),
         dict(cin=48, cout=48, k=3, stride=1, H=64, W=64, act="hard_swish", res=True),
         dict(cin=192, cout=192, k=3, stride=1, H=160, W=160, act="hard_swish", B=2),
         dict(cin=192, cout=384, k=3, stride=1, H=80, W=80, act="silu", B=2),
         dict(cin=288, cout=288, k=3, stride=1, H=40, W=40, act="hard_swish", B=2)]
for c in cases:
    try:
        r = run_conv_case(**c)
        print("case", c, "max_err", r["max_err"], "mean_err", r["mean_err"], "clobbered", r["clobbered"], flush=True)
    except Exception as e:
        print("case", c, "EXC", repr(e)[:300], flush=True)
        break
PY
}
run "YX_HALO=1"
run "YX_HALO=2"
run "YX_HALO=1 YX_HALO_MH=1"
grep -E "==|max_err|EXC" $LOG | cut -c1-260
