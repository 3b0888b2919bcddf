#!/bin/bash
# Epilogue duration with and without concurrent MMAs (YX_CONV_DIAG bit 4 = no MMAs issued), pair vs single-CTA.
mkdir -p gpurun_out
LOG=gpurun_out/diag.log
: > $LOG
for diag in 0 8 4 12; do
echo "=== YX_CONV_DIAG=$diag" >> $LOG
YX_CONV_DIAG=$diag YX_CONV_TRACE=1 timeout 300 python - >> $LOG 2>&1 <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
C96 = dict(cin=96, cout=96, k=3, stride=1, H=160, W=160, B=32, act="hard_swish", res=False)
cases = [
  ("pair-halo 96 resident eg2", dict(C96, tune=_t(2, 96, pair=1, eg=2))),
  ("pair-halo 96 resident eg1", dict(C96, tune=_t(2, 96, pair=1, eg=1))),
  ("halo mh1 96 eg2", dict(C96, tune=_t(2, 96, halves=1, eg=2, w3=2))),
  ("halo mh1 96 eg1", dict(C96, tune=_t(2, 96, halves=1, eg=1, w3=2))),
  ("pair-halo 192 eg1", dict(C96, cin=192, cout=192, tune=_t(2, 192, pair=1, eg=1, sb=1, w3=2))),
  ("halo mh1 192 eg1", dict(C96, cin=192, cout=192, tune=_t(2, 192, halves=1, eg=1, sb=1, w3=2))),
]
for name, c in cases:
    print("CASE", name, flush=True)
    try:
        r = run_conv_case(**c)
    except Exception as e:
        print("ERR", repr(e)[:200])
PY
done
grep -E "===|CASE|ERR|trace:|blocked|^ +(4|5|6) " $LOG | cut -c1-230
