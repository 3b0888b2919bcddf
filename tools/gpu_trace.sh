#!/bin/bash
# Per-tile timeline + per-role blocked-cycle shares (YX_CONV_TRACE) of chosen launch shapes.
mkdir -p gpurun_out
LOG=gpurun_out/trace.log
: > $LOG
YX_CONV_TRACE=1 timeout 600 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
C96 = dict(cin=96, cout=96, k=3, stride=1, H=160, W=160, B=32, act="hard_swish", res="inplace")
cases = [
  ("3x3 96 pair-halo resident eg2 sb2", dict(C96, tune=_t(2, 96, pair=1, eg=2))),
  ("3x3 96 pair-halo resident eg1 sb2", dict(C96, tune=_t(2, 96, pair=1, eg=1))),
  ("3x3 96 pair-halo resident eg2 sb1", dict(C96, tune=_t(2, 96, pair=1, eg=2, sb=1))),
  ("3x3 96 pair-halo resident eg2 sb2 no-res-add", dict(C96, res=False, tune=_t(2, 96, pair=1, eg=2))),
  ("3x3 96 halo mh2 (old best)", dict(C96, tune=_t(2, 96, halves=2, eg=2, w3=2))),
  ("1x1 96->96 @160 generic", dict(cin=96, cout=96, k=1, stride=1, H=160, W=160, B=32, act="hard_swish", tune=_t(1, 96, eg=2))),
  ("1x1 48->48 @320 mh2", dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 48, halves=2, eg=2))),
]
for name, c in cases:
    print("CASE", name, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
grep -E "CASE|trace:|blocked|^ +(3|4|5|6) " $LOG | cut -c1-260
