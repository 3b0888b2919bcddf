#!/bin/bash
# Per-tile timeline + per-role blocked-cycle shares (YX_CONV_TRACE) of the layers furthest above their floor (r02).
mkdir -p gpurun_out
LOG=gpurun_out/trace_r2.log
: > $LOG
YX_CONV_TRACE=1 timeout 600 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
cases = [
  ("dark2.0 48->96 3x3 s2 @640->320 generic BN96 eg1 sb1", dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96, eg=1, sb=1))),
  ("dark2.0 same, eg2 sb2", dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96, eg=2, sb=2))),
  ("dark2.0 same, mh2", dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96, halves=2, eg=2))),
  ("m.conv1 48->48 1x1 @320 (slice of 96) BN48 eg2", dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", src_pitch=96, tune=_t(1, 48, eg=2))),
  ("m.conv1 48->48 1x1 @320 (slice of 96) mh2 eg2", dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", src_pitch=96, tune=_t(1, 48, halves=2, eg=2))),
  ("m.conv1 48->48 1x1 @320 ctas2", dict(cin=48, cout=48, k=1, stride=1, H=320, W=320, B=16, act="hard_swish", src_pitch=96, tune=_t(1, 48, ctas=2))),
  ("dark3.0 96->192 3x3 s2 @320->160 pair BN192 sb1", dict(cin=96, cout=192, k=3, stride=2, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 192, pair=1, sb=1))),
  ("dark3.0 generic BN192", dict(cin=96, cout=192, k=3, stride=2, H=320, W=320, B=16, act="hard_swish", tune=_t(1, 192, sb=1))),
  ("C3_p3.conv1+2 up192+192->384 1x1 @160 BN128 mh2", dict(cin=192, cout=384, k=1, stride=1, H=160, W=160, B=16, act="hard_swish", up_c=192, tune=_t(1, 128, halves=2, sb=1))),
  ("C3_p3.conv1+2 pair BN192", dict(cin=192, cout=384, k=1, stride=1, H=160, W=160, B=16, act="hard_swish", up_c=192, tune=_t(1, 192, pair=1, sb=1))),
  ("stem 16->48 rowpack? (plain 3x3 K16 here) mh2", dict(cin=16, cout=48, k=3, stride=1, H=640, W=640, B=4, act="hard_swish", tune=_t(2, 48, halves=2, eg=2))),
]
for name, c in cases:
    print("CASE", name, flush=True)
    r = run_conv_case(**c)
    print("max_err", r["max_err"], flush=True)
PY
grep -E "CASE|trace:|blocked|^ +(8|9|10|11) " $LOG | cut -c1-300
