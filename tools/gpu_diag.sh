#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/diag22.log
: > $LOG
for d in 0 1 2 3 4 5 7; do
echo "=== DIAG $d" >> $LOG
YX_CONV_DIAG=$d YX_CONV_TRACE=1 timeout 300 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
cases = [
  ("3x3 48 halo mh1 epi2", dict(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=16, act="hard_swish", res=True, tune=_t(2, 48, halves=1, eg=2))),
  ("3x3 48 halo mh1 epi1 nores", dict(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=16, act="hard_swish", res=False, tune=_t(2, 48, halves=1, eg=1))),
  ("s2 48->96 resident", dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96))),
  ("1x1 96->96 ctas1", dict(cin=96, cout=96, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 96, ctas=1))),
  ("3x3 192 halo mh1 BN192", dict(cin=192, cout=192, k=3, stride=1, H=160, W=160, B=16, act="hard_swish", tune=_t(2, 192, halves=1, sb=1))),
]
for name, c in cases:
    print("CASE", name, flush=True)
    r = run_conv_case(**c)
PY
done
grep -E "===|CASE|blocked" $LOG | cut -c1-250
