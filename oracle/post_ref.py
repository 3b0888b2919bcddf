"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/post_ref.c (see its header for the
reference file:line each function restates).  Imported only by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpost_ref.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "post_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.yxref_nms.restype = ctypes.c_int
        _lib.yxref_batched_nms_trick.restype = ctypes.c_int
        _lib.yxref_batched_nms_vanilla.restype = ctypes.c_int
        _lib.yxref_nms_image_main.restype = ctypes.c_int
        _lib.yxref_postprocess_image.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ia(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.int32))


def decode_infer(reg, obj, cls, level_hw, strides):
    """reg [A,4], obj [A] or [A,1], cls [A,C]; fp16 or fp32 numpy. -> boxes[A,4], obj_conf[A], cls_conf[A,C]"""
    is_half = reg.dtype == np.float16
    dt = np.float16 if is_half else np.float32
    reg, obj, cls = (np.ascontiguousarray(a, dtype=dt) for a in (reg, obj, cls))
    A, C = cls.shape
    lh, lw, ls = _ia([h for h, _ in level_hw]), _ia([w for _, w in level_hw]), _ia(strides)
    assert int((lh * lw).sum()) == A
    boxes = np.empty((A, 4), np.float32); oc = np.empty(A, np.float32); cc = np.empty((A, C), np.float32)
    lib().yxref_decode_infer(_p(reg), _p(obj), _p(cls), ctypes.c_int(int(is_half)), A, C, len(ls),
                             _p(lh), _p(lw), _p(ls), _p(boxes), _p(oc), _p(cc))
    return boxes, oc, cc


def decode_yolox(pred, level_hw, strides):
    pred = np.ascontiguousarray(pred, dtype=np.float32).copy()
    A, D = pred.shape
    lh, lw, ls = _ia([h for h, _ in level_hw]), _ia([w for _, w in level_hw]), _ia(strides)
    lib().yxref_decode_yolox(_p(pred), A, D - 5, len(ls), _p(lh), _p(lw), _p(ls))
    return pred


def nms(boxes, scores, thr):
    boxes = np.ascontiguousarray(boxes, np.float32); scores = np.ascontiguousarray(scores, np.float32)
    n = len(scores); keep = np.empty(max(n, 1), np.int32)
    k = lib().yxref_nms(_p(boxes), _p(scores), n, ctypes.c_float(thr), _p(keep))
    return keep[:k].copy()


def batched_nms(boxes, scores, labels, thr, mode="trick"):
    boxes = np.ascontiguousarray(boxes, np.float32); scores = np.ascontiguousarray(scores, np.float32)
    labels = np.ascontiguousarray(labels, np.float32)
    n = len(scores); keep = np.empty(max(n, 1), np.int32)
    f = lib().yxref_batched_nms_trick if mode == "trick" else lib().yxref_batched_nms_vanilla
    k = f(_p(boxes), _p(scores), _p(labels), n, ctypes.c_float(thr), _p(keep))
    return keep[:k].copy()


_MODES = {"trick": 0, "vanilla": 1, "agnostic": 2}


def nms_image_main(boxes, obj_conf, cls_conf, conf_thr, nms_thr, max_nms=5000, max_det=300, mode="trick",
                   multi_class=False, rmmop=None):
    """postprocess_utils.py:73-127 for one image; multi_class / rmmop select the candidate rule (:74-95)."""
    boxes = np.ascontiguousarray(boxes, np.float32); obj_conf = np.ascontiguousarray(obj_conf, np.float32).reshape(-1)
    cls_conf = np.ascontiguousarray(cls_conf, np.float32)
    A, C = cls_conf.shape
    cand = 2 if rmmop is not None else (1 if multi_class else 0)
    r1, r2 = (np.float32(rmmop[0]), np.float32(rmmop[1])) if rmmop is not None else (0.0, 0.0)
    md = min(max_det, A * C if cand == 1 else A)
    det = np.empty((max(md, 1), 7), np.float32); anc = np.empty(max(md, 1), np.int32)
    k = lib().yxref_nms_image_main_ex(_p(boxes), _p(obj_conf), _p(cls_conf), A, C, ctypes.c_float(conf_thr),
                                      ctypes.c_float(nms_thr), int(max_nms), int(md), _MODES[mode], cand,
                                      ctypes.c_float(r1), ctypes.c_float(r2), _p(det), _p(anc))
    return det[:k].copy(), anc[:k].copy()


def postprocess_image(pred, conf_thr, nms_thr, mode="trick"):
    """pred [A,5+C] fp32 cxcywh-decoded. Returns (det[n,7], anchors[n], pred_xyxy)."""
    pred = np.ascontiguousarray(pred, np.float32).copy()
    A, D = pred.shape
    det = np.empty((max(A, 1), 7), np.float32); anc = np.empty(max(A, 1), np.int32)
    k = lib().yxref_postprocess_image(_p(pred), A, D - 5, ctypes.c_float(conf_thr), ctypes.c_float(nms_thr),
                                      _MODES[mode], _p(det), _p(anc))
    return det[:k].copy(), anc[:k].copy(), pred


def torchvision_mode(n_candidates: int, device: str = "cuda") -> str:
    """torchvision 0.26 batched_nms dispatch: vanilla when boxes.numel() > (4000 cpu / 100000 cuda)."""
    return "vanilla" if n_candidates * 4 > (4000 if device == "cpu" else 100000) else "trick"
