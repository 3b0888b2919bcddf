"""GPU parity of the tcgen05 implicit-GEMM conv kernel (through the C ABI, yx_conv2d) against torch's
fp32 conv on identical fp16-rounded inputs.  Tolerance: outputs are fp16 (relative 2^-11 ~ 4.9e-4 of
O(1..8) values) after fp32 accumulation in a different order => abs 1.5e-2."""
import pytest
import torch

from tests.conv_util import CASES, SPARSE_CASES, TUNED_CASES, run_conv_case, tolerance

pytestmark = pytest.mark.gpu


def _id(c):
    return f"{c['cin']}->{c['cout']}_k{c['k']}s{c['stride']}_{c['H']}x{c['W']}" + ("_res" + ("inplace" if c.get("res") == "inplace" else "") if c.get("res") else "") + \
        ("_slice" if c.get("src_pitch") else "")


@pytest.mark.parametrize("case", CASES, ids=[_id(c) for c in CASES])
def test_conv_matches_torch(case):
    r = run_conv_case(**case)
    assert r["max_err"] <= tolerance(case), r["max_err"]
    assert not r["clobbered"], "conv wrote outside its channel slice"
    assert not r["pad_nonzero"], "padded output channels must be zero"


def _tid(c):
    t = c["tune"]
    return _id(c) + f"_v{t['variant']}bn{t['n_tile']}c{t['ctas_per_sm']}h{t['halves']}e{t['epilogue_groups']}s{t['staging_buffers']}" + \
        ("_nores" if t["no_resident_weights"] else "") + ("_pair" if t["cta_pair"] else "") + ("_alt" if t.get("epilogue_alternate") else "") + \
        (f"_up{c['up_c']}" if c.get("up_c") else "")


@pytest.mark.parametrize("case", TUNED_CASES, ids=[_tid(c) for c in TUNED_CASES])
def test_conv_forced_launch_shapes(case):
    """Every launch shape yx_engine_tune may choose (generic / halo, N tile, CTAs per SM, epilogue groups, staging
    buffers, resident or streamed weights) computes the same function."""
    r = run_conv_case(B=3, **case)
    assert r["max_err"] <= tolerance(case), r["max_err"]
    assert not r["clobbered"]


@pytest.mark.parametrize("case", SPARSE_CASES, ids=[_id(c) + f"_sparse_v{c['tune']['variant']}" for c in SPARSE_CASES])
def test_conv_sparse_24(case):
    """The 2:4 sparse tensor-core variant (tcgen05.mma.sp: weights = sparse A operand, pixels = B, transposed accumulator)
    against torch's dense fp32 conv on the same 2:4-masked weights (main.py:52-55 densifies them; 01_mask_generator.py)."""
    if case["cin"] == 384 and case["k"] == 3:     # 3 M tiles x 9 taps x 12 steps = 324 metadata columns > the 256 behind the accumulators
        with pytest.raises(RuntimeError, match="sparse"):
            run_conv_case(B=2, mask24=True, **case)
        return
    r = run_conv_case(B=3, mask24=True, **case)
    assert r["max_err"] <= tolerance(case), r["max_err"]
    assert not r["clobbered"] and not r["pad_nonzero"]


def test_conv_sparse_rejects_dense_weights():
    from tests.conv_util import _sp
    with pytest.raises(RuntimeError, match="2:4"):
        run_conv_case(cin=64, cout=128, k=1, stride=1, H=16, W=16, tune=_sp(1))          # unmasked weights
    with pytest.raises(RuntimeError, match="sparse"):
        run_conv_case(cin=48, cout=96, k=1, stride=1, H=16, W=16, mask24=True, tune=_sp(1))   # cin not a multiple of 32


def test_conv_rejects_unfit_shape():
    import ctypes
    from tests.conv_util import _t
    with pytest.raises(RuntimeError, match="halo"):
        run_conv_case(cin=64, cout=64, k=1, stride=1, H=16, W=16, tune=_t(2, 64))


def test_conv_rejects_bad_geometry():
    import ctypes
    from yolox_b200 import _capi
    lib = _capi.load()
    op = _capi.Op()
    op.kind, op.ksize, op.stride = _capi.OP_CONV, 5, 1
    buf = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    rc = lib.yx_conv2d(ctypes.byref(op), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 0)
    assert rc != 0 and b"ksize" in lib.yx_last_error()
