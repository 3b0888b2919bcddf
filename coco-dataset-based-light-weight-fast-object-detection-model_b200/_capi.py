"""ctypes binding of lib/libyolox_b200.so (C ABI declared in include/yolox_b200.h).

The library is the ONLY compute path: importing the package without it raises, and every entry
point raises RuntimeError with the library's message on a non-zero status.  There is no CPU or
PyTorch fallback."""
import ctypes
import os

from . import _build

c_i32, c_i64, c_f32, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

YX_F16, YX_F32, YX_U8 = 0, 1, 2


def image_dtype(t) -> int:
    """yx_dtype of an image batch (fp16 / fp32 / uint8 pixel values)."""
    import torch
    try:
        return {torch.float16: YX_F16, torch.float32: YX_F32, torch.uint8: YX_U8}[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported image dtype {t.dtype}") from None
ACT = {"none": 0, "identity": 0, "silu": 1, "swish": 1, "hsilu": 2, "hswish": 2, "hard_silu": 2, "hard_swish": 2,
       "relu": 3, "lrelu": 4, "leaky_relu": 4}
OP_CONV, OP_S2D, OP_SPP, OP_UPSAMPLE, OP_DWCONV = 0, 1, 2, 3, 4
NMS_TRICK, NMS_VANILLA, NMS_AGNOSTIC = 0, 1, 2


def act_code(name: str) -> int:
    try:
        return ACT[name.lower()]
    except KeyError:
        raise AttributeError("Unsupported act type: {}".format(name))  # same error as blocks.py:17


class View(ctypes.Structure):
    _fields_ = [("offset", c_i64), ("nstride", c_i64), ("n", c_i32), ("h", c_i32), ("w", c_i32), ("c", c_i32),
                ("pitch", c_i32), ("_pad", c_i32)]


class Op(ctypes.Structure):
    _fields_ = [("kind", c_i32), ("ksize", c_i32), ("stride", c_i32), ("act", c_i32),
                ("src", View), ("dst", View), ("res", View), ("up", View),
                ("w_offset", c_i64), ("b_offset", c_i64), ("cin_pad", c_i32), ("cout_pad", c_i32),
                ("aux", c_i32), ("_pad", c_i32)]


class ConvTune(ctypes.Structure):
    """Mirror of yx_conv_tune (include/yolox_b200.h)."""
    _fields_ = [("variant", c_i32), ("n_tile", c_i32), ("ctas_per_sm", c_i32), ("halves", c_i32),
                ("epilogue_groups", c_i32), ("staging_buffers", c_i32), ("second_producer", c_i32),
                ("no_resident_weights", c_i32), ("cta_pair", c_i32), ("sparse", c_i32), ("epilogue_alternate", c_i32), ("reserved", c_i32)]

    FIELDS = ("variant", "n_tile", "ctas_per_sm", "halves", "epilogue_groups", "staging_buffers", "second_producer",
              "no_resident_weights", "cta_pair", "sparse", "epilogue_alternate")

    def as_list(self):
        return [int(getattr(self, f)) for f in self.FIELDS]

    @classmethod
    def from_list(cls, vals):
        t = cls()
        for f, v in zip(cls.FIELDS, vals):
            setattr(t, f, int(v))
        return t


class Levels(ctypes.Structure):
    _fields_ = [("n_levels", c_i32), ("h", c_i32 * 8), ("w", c_i32 * 8), ("stride", c_i32 * 8)]


def make_levels(level_hw, strides) -> Levels:
    lv = Levels()
    lv.n_levels = len(strides)
    for i, ((h, w), s) in enumerate(zip(level_hw, strides)):
        lv.h[i], lv.w[i], lv.stride[i] = int(h), int(w), int(s)
    return lv


MAX_PEERS, IPC_HANDLE_BYTES = 8, 64


PEER_WAIT_NONE, PEER_WAIT_AFTER, PEER_WAIT_BEFORE = 0, 1, 2


class PeerOut(ctypes.Structure):
    """yx_peer_out (include/yolox_b200.h)."""
    _fields_ = [("world", ctypes.c_int32), ("wait_target", ctypes.c_int32), ("timeout_ms", ctypes.c_int32),
                ("wait_mode", ctypes.c_int32), ("det", ctypes.c_void_p * MAX_PEERS), ("cnt", ctypes.c_void_p * MAX_PEERS),
                ("arrive", ctypes.c_void_p * MAX_PEERS), ("local_arrive", ctypes.c_void_p), ("status", ctypes.c_void_p),
                ("wait_cnt", ctypes.c_void_p)]


SYMBOLS = ["yx_last_error", "yx_abi_version", "yx_engine_create", "yx_engine_destroy", "yx_engine_run",
           "yx_engine_profile", "yx_engine_run_ops", "yx_engine_num_launches", "yx_engine_tune", "yx_engine_get_tune", "yx_engine_set_tune", "yx_engine_op_desc", "yx_engine_op_sparse_ok", "yx_engine_tune_mismatches",
           "yx_conv2d", "yx_conv2d_ex", "yx_decode_infer", "yx_decode_infer_grids", "yx_detect_workspace_bytes",
           "yx_nms_main", "yx_nms_main_ex", "yx_nms_workspace_bytes", "yx_detect_main",
           "yx_detect_main_gather", "yx_peer_wait", "yx_ipc_export", "yx_ipc_open", "yx_ipc_close", "yx_head_assemble", "yx_decode_outputs", "yx_postprocess_yolox",
           "yx_preprocess_batch", "yx_coco_records", "yx_cocoeval_bbox"]

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if the sources are newer) the native library.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if not os.path.exists(path) or _build._stale():
        path = _build.build()
    lib = ctypes.CDLL(path)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise RuntimeError(f"libyolox_b200.so does not export {s}")
    lib.yx_last_error.restype = ctypes.c_char_p
    lib.yx_abi_version.restype = c_i32
    lib.yx_detect_workspace_bytes.restype = c_sz
    lib.yx_detect_workspace_bytes.argtypes = [c_i32, c_i32]
    lib.yx_engine_create.argtypes = [ctypes.POINTER(Op), c_i32, c_vp, c_sz, c_vp, c_sz, c_vp, c_sz, c_i32, c_i32, c_i32,
                                     ctypes.POINTER(c_vp)]
    lib.yx_engine_destroy.argtypes = [c_vp]
    lib.yx_engine_destroy.restype = None
    lib.yx_engine_run.argtypes = [c_vp, c_vp, c_i32, c_f32, c_f32, c_i32, c_vp]
    lib.yx_engine_run_ops.argtypes = [c_vp, c_vp, c_i32, c_f32, c_f32, c_i32, c_i32, c_vp]
    lib.yx_engine_profile.argtypes = [c_vp, c_vp, c_i32, c_i32, c_vp, ctypes.POINTER(c_f32),
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), c_i32]
    lib.yx_engine_num_launches.argtypes = [c_vp]
    lib.yx_conv2d.argtypes = [ctypes.POINTER(Op), c_vp, c_vp, c_vp, c_vp]
    lib.yx_conv2d_ex.argtypes = [ctypes.POINTER(Op), c_vp, c_vp, c_vp, ctypes.POINTER(ConvTune), c_vp]
    lib.yx_engine_tune.argtypes = [c_vp, c_vp, c_i32, c_f32, c_f32, c_i32, c_vp]
    lib.yx_engine_get_tune.argtypes = [c_vp, c_i32, ctypes.POINTER(ConvTune)]
    lib.yx_engine_set_tune.argtypes = [c_vp, c_i32, ctypes.POINTER(ConvTune)]
    lib.yx_engine_op_desc.argtypes = [c_vp, c_i32, ctypes.c_char_p, c_i32]
    lib.yx_engine_op_sparse_ok.argtypes = [c_vp, c_i32]
    lib.yx_engine_tune_mismatches.argtypes = [c_vp, ctypes.c_char_p, c_i32]
    logits = [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64]
    lib.yx_decode_infer.argtypes = logits + [c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Levels), c_vp, c_vp, c_vp, c_vp]
    lib.yx_decode_infer_grids.argtypes = logits + [c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]
    lib.yx_nms_main.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32, c_i32, c_i32, c_vp, c_sz,
                                c_vp, c_vp, c_vp, c_vp]
    lib.yx_nms_main_ex.argtypes = [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32, c_i32, c_i32, c_i32, c_f32,
                                   c_f32, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp]
    lib.yx_nms_workspace_bytes.restype = c_sz
    lib.yx_nms_workspace_bytes.argtypes = [c_i32, c_i32, c_i32, c_i32, c_i32]
    lib.yx_detect_main.argtypes = logits + [c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Levels), c_f32, c_f32, c_i32,
                                            c_i32, c_i32, c_vp, c_sz, c_vp, c_vp, c_vp, c_vp]
    lib.yx_detect_main_gather.argtypes = logits + [c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Levels), c_f32, c_f32, c_i32,
                                                   c_i32, c_i32, c_vp, c_sz, c_vp, c_vp, c_vp, ctypes.POINTER(PeerOut), c_vp]
    lib.yx_peer_wait.argtypes = [c_vp, c_i32, c_i32, c_vp, c_i32, c_vp, c_i32, c_vp]
    lib.yx_ipc_export.argtypes = [c_vp, c_vp, ctypes.POINTER(c_i64)]
    lib.yx_ipc_open.argtypes = [c_vp, ctypes.POINTER(c_vp)]
    lib.yx_ipc_close.argtypes = [c_vp]
    lib.yx_head_assemble.argtypes = logits + [c_i32, c_i32, c_i32, ctypes.POINTER(Levels), c_i32, c_vp, c_i32, c_vp]
    lib.yx_decode_outputs.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, ctypes.POINTER(Levels), c_vp]
    lib.yx_postprocess_yolox.argtypes = [c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_f32, c_i32, c_vp, c_sz, c_vp, c_vp,
                                         c_vp, c_vp]
    lib.yx_preprocess_batch.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_vp]
    lib.yx_coco_records.argtypes = [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp, c_vp]
    lib.yx_cocoeval_bbox.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp,
                                     c_i32, c_vp, c_vp, c_vp]
    if lib.yx_abi_version() != 4:
        raise RuntimeError("libyolox_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().yx_last_error()
        raise RuntimeError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"yolox_b200: {name} must be a CUDA tensor — the engine has no CPU path")
