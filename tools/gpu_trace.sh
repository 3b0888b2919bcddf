#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/trace19.log
: > $LOG
YX_CONV_TRACE=1 timeout 600 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
H3 = dict(cin=192, cout=192, k=3, stride=1, H=160, W=160, B=16, act="hard_swish")
cases = [
  ("3x3 192 generic BN192", dict(H3, tune=_t(1, 192, sb=1))),
  ("3x3 192 generic BN192 w3=0", dict(H3, tune=_t(1, 192, sb=1, w3=0))),
  ("3x3 192 halo mh1 BN192", dict(H3, tune=_t(2, 192, halves=1, sb=1))),
  ("3x3 192 halo mh2 BN128", dict(H3, tune=_t(2, 128, halves=2, sb=1))),
  ("3x3 192->384 halo mh2 BN128", dict(H3, cout=384, tune=_t(2, 128, halves=2, sb=1))),
  ("3x3 96 halo mh2", dict(cin=96, cout=96, k=3, stride=1, H=160, W=160, B=32, act="hard_swish", res=True, tune=_t(2, 96, halves=2))),
  ("3x3 48 halo mh1 epi2", dict(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=16, act="hard_swish", res=True, tune=_t(2, 48, halves=1, eg=2))),
  ("3x3 48 halo mh2 epi2", dict(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=16, act="hard_swish", res=True, tune=_t(2, 48, halves=2, eg=2))),
  ("s2 48->96 resident", dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96))),
  ("s2 96->192", dict(cin=96, cout=192, k=3, stride=2, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 192, sb=1))),
  ("1x1 48->48 ctas2", dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 48, ctas=2))),
  ("1x1 96->96 ctas2", dict(cin=96, cout=96, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 96, ctas=2, sb=1))),
  ("1x1 384->384 BN192", dict(cin=384, cout=384, k=1, stride=1, H=160, W=160, B=16, act="hard_swish", tune=_t(1, 192))),
  ("1x1 768->768 BN256", dict(cin=768, cout=768, k=1, stride=1, H=80, W=80, B=32, act="hard_swish", tune=_t(1, 256, sb=1))),
]
for name, c in cases:
    print("CASE", name, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
grep -E "CASE|trace:|blocked|^ +(3|4|5) " $LOG | cut -c1-260
