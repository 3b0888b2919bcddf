#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/one_conv.py <<'PY'
import sys
sys.path.insert(0, '.')
from tests.conv_util import run_conv_case, _t
which = sys.argv[1]
if which == "c48":
    run_conv_case(cin=48, cout=48, k=3, stride=1, H=320, W=320, B=16, act="hard_swish", res=True, tune=_t(2, 48, halves=1, eg=2))
elif which == "c192":
    run_conv_case(cin=192, cout=192, k=3, stride=1, H=160, W=160, B=16, act="hard_swish", tune=_t(2, 192, halves=1, sb=1))
elif which == "s2":
    run_conv_case(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=8, act="hard_swish", tune=_t(1, 96))
PY
for w in c48 c192 s2; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm -c 1 -f -o gpurun_out/ncu20_$w python /tmp/one_conv.py $w > gpurun_out/ncu20_$w.log 2>&1
  echo "ncu $w exit=$?"
done
ls -la gpurun_out/ncu20_*
