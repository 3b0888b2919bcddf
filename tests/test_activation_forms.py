"""The conv epilogue evaluates Hardswish as x * sat(x/6 + 0.5) (FFMA.SAT + FMUL, csrc/yx_conv.cu apply_act) instead of the
reference's x * relu6(x + 3) / 6 (network_blocks.py:12-24 -> nn.Hardswish).  Exhaustive check over every finite fp16 value
(as an fp32 pre-activation): after the single rounding to fp16 the two forms agree except for isolated 1-ulp cases."""
import numpy as np
import torch


def _ulp16(v):
    a = np.abs(v.astype(np.float32))
    e = np.floor(np.log2(np.maximum(a, 2.0 ** -14)))
    return (2.0 ** (e - 10)).astype(np.float32)


def test_hardswish_forms_agree_over_all_fp16_inputs():
    bits = np.arange(0, 1 << 16, dtype=np.uint16)
    x16 = bits.view(np.float16)
    x16 = x16[np.isfinite(x16)]
    x = x16.astype(np.float32)
    ref = torch.nn.functional.hardswish(torch.from_numpy(x)).numpy()                 # fp32, the reference's formula
    # FFMA (single rounding of x*(1/6)+0.5) emulated in float64, saturate, then an fp32 multiply
    t = (x.astype(np.float64) * np.float64(np.float32(1.0 / 6.0)) + 0.5).astype(np.float32)
    ours = (x * np.clip(t, 0.0, 1.0)).astype(np.float32)
    r16, o16 = ref.astype(np.float16), ours.astype(np.float16)
    diff = np.abs(r16.astype(np.float32) - o16.astype(np.float32))
    assert np.all(diff <= _ulp16(r16)), "more than one fp16 ulp apart"
    assert float((diff > 0).mean()) < 0.01, float((diff > 0).mean())
    # exact where the function is linear or zero
    big, neg = x >= 3.0, x <= -3.0
    assert np.array_equal(o16[big], x16[big]) and np.all(o16[neg] == 0)
