"""Per-tile timeline and per-role blocked-cycle shares (YX_CONV_TRACE=1) of one case of tools/conv_time.py.
    YX_CONV_TRACE=1 python tools/conv_trace.py <case> [shape index ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("YX_CONV_TRACE", "1")
from tests.conv_util import run_conv_case
from tools.conv_time import CASES

case, shapes = CASES[sys.argv[1]]
for i in [int(a) for a in sys.argv[2:]] or range(len(shapes)):
    name, tune = shapes[i]
    print("CASE", sys.argv[1], name, flush=True)
    sys.stderr.flush()
    run_conv_case(**case, tune=tune)
