"""CUDA-event time of single conv layers under forced launch shapes (the same yx_conv2d_ex path the parity tests use; the
result is also checked against torch).    python tools/conv_time.py <case> [<case> ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.conv_util import run_conv_case, _t

CASES = {
    # the 288-channel 3x3 layers of the 40x40 level (dark5.1.m.*, C3_p5.m.*, C3_n4.m.*)
    "c288": (dict(cin=288, cout=288, k=3, stride=1, H=40, W=40, B=64, act="hard_swish"), [
        ("halo BN192 sb1", _t(2, 192, sb=1, w3=2)), ("halo BN192 eg2 sb1", _t(2, 192, eg=2, sb=1, w3=2)),
        ("pair-halo BN192 sb1", _t(2, 192, sb=1, w3=2, pair=1)), ("pair-halo BN192 sb2", _t(2, 192, sb=2, w3=2, pair=1)),
        ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)),
        ("pair-halo BN256 sb1", _t(2, 256, sb=1, w3=2, pair=1)), ("pair-halo BN128 sb1", _t(2, 128, sb=1, w3=2, pair=1)),
        ("halo BN128 sb1", _t(2, 128, sb=1, w3=2)), ("halo BN256 sb1", _t(2, 256, sb=1, w3=2)),
    ]),
    # the stem's MMA + epilogue structure without the image builder (A = a 16-channel NHWC tensor through TMA halo loads)
    "stem16": (dict(cin=16, cout=48, k=3, stride=1, H=640, W=640, B=16, act="hard_swish"), [
        ("halo BN48 mh2 eg2 alt", _t(2, 48, halves=2, eg=2, alt=1)), ("halo BN48 mh2 eg2", _t(2, 48, halves=2, eg=2)),
        ("halo BN48 mh2 eg1", _t(2, 48, halves=2, eg=1)), ("halo BN48 mh1 eg2 alt", _t(2, 48, halves=1, eg=2, alt=1)),
        ("halo BN48 mh1 eg1", _t(2, 48, halves=1, eg=1)),
    ]),
    "d20": (dict(cin=48, cout=96, k=3, stride=2, H=640, W=640, B=16, act="hard_swish"), [
        ("generic BN96 eg2 sb2", _t(1, 96, eg=2, sb=2)), ("generic BN96 eg1 sb1", _t(1, 96, eg=1, sb=1)),
        ("generic BN96 mh2 eg2", _t(1, 96, halves=2, eg=2)), ("pair-generic BN96 sb2", _t(1, 96, pair=1)),
    ]),
    "d30": (dict(cin=96, cout=192, k=3, stride=2, H=320, W=320, B=16, act="hard_swish"), [
        ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)), ("generic BN192 sb1", _t(1, 192, sb=1)),
        ("generic BN128 mh2 eg2", _t(1, 128, halves=2, eg=2)),
    ]),
    "c96x1": (dict(cin=96, cout=96, k=1, stride=1, H=160, W=160, B=64, act="hard_swish", src_pitch=192), [
        ("generic BN96 eg2 sb1", _t(1, 96, eg=2, sb=1)), ("pair-generic BN96 sb2", _t(1, 96, pair=1)),
        ("generic BN96 mh2 eg2", _t(1, 96, halves=2, eg=2)), ("generic BN96 ctas2", _t(1, 96, ctas=2)),
    ]),
    "h192": (dict(cin=192, cout=192, k=3, stride=1, H=160, W=160, B=16, act="hard_swish"), [
        ("pair-halo BN192 sb1", _t(2, 192, sb=1, w3=2, pair=1)), ("halo BN192 sb1", _t(2, 192, sb=1, w3=2)),
        ("halo BN192 eg2 sb2 alt", _t(2, 192, eg=2, sb=2, w3=2, alt=1)),
    ]),
    "h384": (dict(cin=192, cout=384, k=3, stride=1, H=160, W=160, B=16, act="hard_swish"), [
        ("pair-halo BN192 sb1", _t(2, 192, sb=1, w3=2, pair=1)), ("pair-halo BN256 sb1", _t(2, 256, sb=1, w3=2, pair=1)),
        ("halo BN192 sb1", _t(2, 192, sb=1, w3=2)),
    ]),
    "p4c3": (dict(cin=768, cout=384, k=1, stride=1, H=80, W=80, B=64, act="hard_swish"), [
        ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)), ("pair-generic BN256 sb1", _t(1, 256, sb=1, pair=1)),
        ("generic BN128 mh2 eg2", _t(1, 128, halves=2, eg=2)), ("generic BN192 sb1", _t(1, 192, sb=1)),
    ]),
    "d5c12": (dict(cin=576, cout=576, k=1, stride=1, H=40, W=40, B=64, act="hard_swish"), [
        ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)), ("generic BN192 sb1", _t(1, 192, sb=1)),
        ("generic BN128 mh2 eg2", _t(1, 128, halves=2, eg=2)),
    ]),
    "bu1": (dict(cin=384, cout=384, k=3, stride=2, H=80, W=80, B=64, act="hard_swish"), [
        ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)), ("generic BN192 sb1", _t(1, 192, sb=1)),
        ("pair-generic BN128 sb1", _t(1, 128, sb=1, pair=1)),
    ]),
    "c384x1": (dict(cin=384, cout=384, k=1, stride=1, H=80, W=80, B=64, act="hard_swish"), [
        ("generic BN128 mh2", _t(1, 128, halves=2, eg=2)), ("pair-generic BN192 sb1", _t(1, 192, sb=1, pair=1)),
        ("pair-generic BN256 sb1", _t(1, 256, sb=1, pair=1)), ("generic BN192 sb1", _t(1, 192, sb=1)),
    ]),
}
for key in (sys.argv[1:] if __name__ == "__main__" else []):
    case, shapes = CASES[key]
    for name, tune in shapes:
        try:
            r = run_conv_case(**case, tune=tune, time_iters=20)
            print(f"{key:8s} {name:28s} {r['ms']:.4f} ms  max_err {r['max_err']:.3g}", flush=True)
        except Exception as e:  # a shape the planner rejects
            print(f"{key:8s} {name:28s} rejected: {str(e)[:150]}", flush=True)
