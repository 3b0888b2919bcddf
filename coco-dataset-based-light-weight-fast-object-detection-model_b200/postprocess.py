"""Reference-signature post-processing functions on top of the native decode / select / NMS kernels.

  yolox_generate_grid, yolox_postprocess_output_torch_batch, yolox_nms_torch_batch
        == choijhanyangackr/yolox_infer/postprocess_utils.py:6-129 (main.py:174,180,188 call them)
  postprocess == yolox/utils/boxes.py:32-82 (mutates prediction[:, :, :4] to xyxy like the reference)
  decode_outputs == YOLOXHead.decode_outputs, yolox/models/yolo_head.py:210-225 (in place)
  detect_main — the fused production path (decode + threshold + NMS from raw logits, no host sync)

All tensors must live on a CUDA device; there is no CPU path (the oracle under oracle/ is the checker).
"""
import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _capi

_NMS_MODES = {"trick": _capi.NMS_TRICK, "vanilla": _capi.NMS_VANILLA, "agnostic": _capi.NMS_AGNOSTIC, "auto": 3}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float16:
        return _capi.YX_F16
    if t.dtype == torch.float32:
        return _capi.YX_F32
    raise RuntimeError(f"unsupported dtype {t.dtype} (fp16 / fp32 only)")


def _rows(t: torch.Tensor) -> torch.Tensor:
    """[B,A,K] with a contiguous last dim (arbitrary batch / anchor strides are passed through)."""
    _capi.require_cuda(t)
    if t.dim() != 3:
        raise RuntimeError("expected a [B, A, K] tensor")
    return t if t.stride(2) == 1 or t.shape[2] == 1 else t.contiguous()


def _workspace(B: int, A: int, device, nbytes: Optional[int] = None) -> torch.Tensor:
    n = _capi.load().yx_detect_workspace_bytes(B, A) if nbytes is None else nbytes
    ws = torch.empty(n + 256, dtype=torch.uint8, device=device)
    off = (-ws.data_ptr()) % 256
    return ws[off:off + n]


def level_hw_of(img_size, strides) -> List[Tuple[int, int]]:
    if isinstance(img_size, int):
        img_size = (img_size, img_size)
    return [(img_size[0] // s, img_size[1] // s) for s in strides]


def yolox_generate_grid(img_size, strides=(8, 16, 32), dtype=torch.float32):
    """postprocess_utils.py:6-24 — same values/shapes: grids (1,A,2) xy, scales (1,A,1)."""
    hw = level_hw_of(img_size, strides)
    grids, scales = [], []
    for (h, w), s in zip(hw, strides):
        yv, xv = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
        grids.append(torch.stack((xv, yv), 2).view(1, -1, 2))
        scales.append(torch.full((1, h * w, 1), s))
    grids = torch.cat(grids, dim=1).to(dtype)
    scales = torch.cat(scales, dim=1).to(dtype)
    return grids, scales


def yolox_postprocess_output_torch_batch(reg_output, obj_output, cls_output, grids, scales):
    """postprocess_utils.py:27-52: fp32 xyxy boxes [B,A,4], obj_conf [B,A,1], cls_conf [B,A,C]=sig(cls)*sig(obj).

    The kernel READS `grids` (1,A,2) and `scales` (1,A,1) per anchor, as the reference's add_/mul_ do: nothing about
    them is cached or inferred (main.py:171-178 rebinds both whenever the batch's (h, w) changes, and equal-sized
    replacements tend to land on the same device addresses)."""
    lib = _capi.load()
    reg, obj, cls = _rows(reg_output), _rows(obj_output), _rows(cls_output)
    if not (reg.dtype == obj.dtype == cls.dtype):
        raise RuntimeError("reg/obj/cls must share a dtype")
    B, A, C = cls.shape
    _capi.require_cuda(grids, "grids")
    _capi.require_cuda(scales, "scales")
    if grids.numel() != 2 * A or scales.numel() != A:
        raise RuntimeError(f"grids / scales must hold (1,{A},2) / (1,{A},1) values, got {tuple(grids.shape)} / {tuple(scales.shape)}")
    if grids.dtype != scales.dtype:
        scales = scales.to(grids.dtype)
    if grids.dtype not in (torch.float16, torch.float32):
        grids, scales = grids.float(), scales.float()
    grids, scales = grids.contiguous(), scales.contiguous()
    boxes = torch.empty(B, A, 4, dtype=torch.float32, device=cls.device)
    obj_conf = torch.empty(B, A, 1, dtype=torch.float32, device=cls.device)
    cls_conf = torch.empty(B, A, C, dtype=torch.float32, device=cls.device)
    with torch.cuda.device(cls.device):
        _capi.check(lib.yx_decode_infer_grids(reg.data_ptr(), reg.stride(0), reg.stride(1), obj.data_ptr(), obj.stride(0),
                                              obj.stride(1), cls.data_ptr(), cls.stride(0), cls.stride(1), _dt(cls), B, A, C,
                                              grids.data_ptr(), scales.data_ptr(), _dt(grids), boxes.data_ptr(),
                                              obj_conf.data_ptr(), cls_conf.data_ptr(), _capi.current_stream_ptr()),
                    "yx_decode_infer_grids")
    return boxes, obj_conf, cls_conf


def decode_infer(reg_output, obj_output, cls_output, level_hw, strides):
    """The same decode from a pyramid description (level_hw, strides) instead of grids / scales tensors."""
    lib = _capi.load()
    reg, obj, cls = _rows(reg_output), _rows(obj_output), _rows(cls_output)
    B, A, C = cls.shape
    lv = _capi.make_levels(level_hw, strides)
    boxes = torch.empty(B, A, 4, dtype=torch.float32, device=cls.device)
    obj_conf = torch.empty(B, A, 1, dtype=torch.float32, device=cls.device)
    cls_conf = torch.empty(B, A, C, dtype=torch.float32, device=cls.device)
    with torch.cuda.device(cls.device):
        _capi.check(lib.yx_decode_infer(reg.data_ptr(), reg.stride(0), reg.stride(1), obj.data_ptr(), obj.stride(0),
                                        obj.stride(1), cls.data_ptr(), cls.stride(0), cls.stride(1), _dt(cls), B, A, C,
                                        lv, boxes.data_ptr(), obj_conf.data_ptr(), cls_conf.data_ptr(),
                                        _capi.current_stream_ptr()), "yx_decode_infer")
    return boxes, obj_conf, cls_conf


def _split(det: torch.Tensor, count: torch.Tensor, dtype=None) -> List[Optional[torch.Tensor]]:
    counts = count.cpu().tolist()  # the one host sync, as in the reference's boolean indexing
    out = []
    for i, n in enumerate(counts):
        if n == 0:
            out.append(None)
        else:
            d = det[i, :n].clone()
            out.append(d if dtype is None else d.to(dtype))
    return out


CAND_MAX, CAND_MULTI_CLASS, CAND_RMMOP = 0, 1, 2  # yx_cand_mode


def nms_main_raw(reg_boxes, obj_conf, cls_conf, nms_threshold, conf_threshold, max_num_nms, max_num_det, mode="auto",
                 multi_class=False, rmmop=None):
    """Device-side result of yolox_nms_torch_batch: det [B,R,7], count [B], anchor [B,R] (no sync)."""
    lib = _capi.load()
    for t in (reg_boxes, obj_conf, cls_conf):
        _capi.require_cuda(t)
    boxes = reg_boxes.float().contiguous()
    objc = obj_conf.float().contiguous()
    clsc = cls_conf.float().contiguous()
    B, A, C = clsc.shape
    cand = CAND_RMMOP if rmmop is not None else (CAND_MULTI_CLASS if multi_class else CAND_MAX)  # rmmop wins, :74
    r1, r2 = (float(rmmop[0]), float(rmmop[1])) if rmmop is not None else (0.0, 0.0)
    per_image = A * C if cand == CAND_MULTI_CLASS else A
    if max_num_det >= per_image:
        max_num_det = 0                                   # no cap can bind: size det by the candidate count
    rows = max_num_det if max_num_det > 0 else per_image
    det = torch.empty(B, rows, 7, dtype=torch.float32, device=clsc.device)
    cnt = torch.empty(B, dtype=torch.int32, device=clsc.device)
    anc = torch.empty(B, rows, dtype=torch.int32, device=clsc.device)
    if cand == CAND_MULTI_CLASS:
        nbytes = lib.yx_nms_workspace_bytes(B, A, C, cand, int(max_num_nms))
        if nbytes == 0:
            raise RuntimeError("multi_class candidate set too large")
        ws = _workspace(B, A, clsc.device, nbytes)
    else:
        ws = _workspace(B, A, clsc.device)
    with torch.cuda.device(clsc.device):
        _capi.check(lib.yx_nms_main_ex(boxes.data_ptr(), objc.data_ptr(), clsc.data_ptr(), B, A, C, float(conf_threshold),
                                       float(nms_threshold), int(max_num_nms), int(max_num_det), _NMS_MODES[mode], cand,
                                       r1, r2, ws.data_ptr(), ws.numel(), det.data_ptr(), cnt.data_ptr(),
                                       anc.data_ptr(), _capi.current_stream_ptr()), "yx_nms_main_ex")
    return det, cnt, anc


def yolox_nms_torch_batch(reg_boxes, obj_conf, cls_conf, nms_threshold: float = 0.65, conf_threshold: float = 0.001,
                          soft: bool = False, max_num_nms: int = 5000, max_num_det: int = 300,
                          multi_class: bool = False, rmmop=None, class_agnostic: bool = False):
    """postprocess_utils.py:55-129, every candidate rule (default, multi_class, rmmop) and class_agnostic.
    Returns list[B] of [n,7] or None."""
    if soft:
        raise ValueError("Soft-NMS is not installed, but using soft_nms.")  # nms.py:21,36
    det, cnt, _ = nms_main_raw(reg_boxes, obj_conf, cls_conf, nms_threshold, conf_threshold, max_num_nms,
                               max(1, min(int(max_num_det), 2 ** 31 - 1)), "agnostic" if class_agnostic else "auto",
                               multi_class=multi_class, rmmop=rmmop)
    out = _split(det, cnt)
    if max_num_det <= 0:  # :123-124 slices to zero rows; images without candidates stay None
        out = [None if d is None else d[:0] for d in out]
    return out


def detect_main(reg, obj, cls, level_hw, strides, conf_threshold=0.001, nms_threshold=0.65, max_num_nms=5000,
                max_num_det=300, mode="auto", gather=None):
    """Fused decode + threshold + top-k + NMS from raw logits [B,A,*] (fp16/fp32).  No host sync.
    Returns det [B,max_num_det,7] fp32 (rows beyond count are zero), count [B] int32, anchor [B,max_num_det].
    gather: a dist.PeerGather — the NMS kernel also stores the rows into every rank's window (fused all-gather);
    read the global result from gather.result() afterwards."""
    lib = _capi.load()
    reg, obj, cls = _rows(reg), _rows(obj), _rows(cls)
    B, A, C = cls.shape
    lv = _capi.make_levels(level_hw, strides)
    rows = max_num_det if max_num_det > 0 else A
    det = torch.empty(B, rows, 7, dtype=torch.float32, device=cls.device)
    cnt = torch.empty(B, dtype=torch.int32, device=cls.device)
    anc = torch.empty(B, rows, dtype=torch.int32, device=cls.device)
    ws = _workspace(B, A, cls.device)
    if gather is not None:
        po = gather.next_step(B, rows)
        with torch.cuda.device(cls.device):
            _capi.check(lib.yx_detect_main_gather(
                reg.data_ptr(), reg.stride(0), reg.stride(1), obj.data_ptr(), obj.stride(0), obj.stride(1), cls.data_ptr(),
                cls.stride(0), cls.stride(1), _dt(cls), B, A, C, lv, float(conf_threshold), float(nms_threshold),
                int(max_num_nms), int(max_num_det), _NMS_MODES[mode], ws.data_ptr(), ws.numel(), det.data_ptr(),
                cnt.data_ptr(), anc.data_ptr(), ctypes.byref(po), _capi.current_stream_ptr()), "yx_detect_main_gather")
        return det, cnt, anc
    with torch.cuda.device(cls.device):
        _capi.check(lib.yx_detect_main(reg.data_ptr(), reg.stride(0), reg.stride(1), obj.data_ptr(), obj.stride(0),
                                       obj.stride(1), cls.data_ptr(), cls.stride(0), cls.stride(1), _dt(cls), B, A, C, lv,
                                       float(conf_threshold), float(nms_threshold), int(max_num_nms), int(max_num_det),
                                       _NMS_MODES[mode], ws.data_ptr(), ws.numel(), det.data_ptr(), cnt.data_ptr(),
                                       anc.data_ptr(), _capi.current_stream_ptr()), "yx_detect_main")
    return det, cnt, anc


def head_assemble(reg8, cls, num_classes, level_hw, strides, decode: bool, out_dtype):
    """[B,A,5+C] = [reg, sigmoid(obj), sigmoid(cls)] (+ decode) from the engine's packed fp16 logits."""
    lib = _capi.load()
    B, A, _ = cls.shape
    C = num_classes
    out = torch.empty(B, A, 5 + C, dtype=out_dtype, device=cls.device)
    lv = _capi.make_levels(level_hw, strides)
    obj = reg8[..., 4:5]
    with torch.cuda.device(cls.device):
        _capi.check(lib.yx_head_assemble(reg8.data_ptr(), reg8.stride(0), reg8.stride(1), obj.data_ptr(), obj.stride(0),
                                         obj.stride(1), cls.data_ptr(), cls.stride(0), cls.stride(1), B, A, C, lv,
                                         int(bool(decode)), out.data_ptr(), _dt(out), _capi.current_stream_ptr()),
                    "yx_head_assemble")
    return out


def decode_outputs(outputs, level_hw, strides):
    """In-place decode of outputs[..., :4] (yolo_head.py:210-225); returns the same tensor."""
    lib = _capi.load()
    _capi.require_cuda(outputs)
    if not outputs.is_contiguous():
        raise RuntimeError("decode_outputs works in place and needs a contiguous [B,A,5+C] tensor")
    B, A, D = outputs.shape
    lv = _capi.make_levels(level_hw, strides)
    with torch.cuda.device(outputs.device):
        _capi.check(lib.yx_decode_outputs(outputs.data_ptr(), _dt(outputs), B, A, D - 5, lv,
                                          _capi.current_stream_ptr()), "yx_decode_outputs")
    return outputs


def postprocess_raw(prediction, num_classes, conf_threshold, nms_threshold, class_agnostic=False, mode=None):
    lib = _capi.load()
    _capi.require_cuda(prediction)
    if not prediction.is_contiguous():
        raise RuntimeError("postprocess mutates prediction in place and needs a contiguous [B,A,5+C] tensor")
    B, A, D = prediction.shape
    if D != 5 + num_classes:
        raise RuntimeError(f"prediction has {D} columns, expected 5 + num_classes = {5 + num_classes}")
    det = torch.empty(B, A, 7, dtype=torch.float32, device=prediction.device)
    cnt = torch.empty(B, dtype=torch.int32, device=prediction.device)
    anc = torch.empty(B, A, dtype=torch.int32, device=prediction.device)
    ws = _workspace(B, A, prediction.device)
    m = _NMS_MODES[mode] if mode else (_capi.NMS_AGNOSTIC if class_agnostic else 3)
    with torch.cuda.device(prediction.device):
        _capi.check(lib.yx_postprocess_yolox(prediction.data_ptr(), _dt(prediction), B, A, num_classes,
                                             float(conf_threshold), float(nms_threshold), m, ws.data_ptr(), ws.numel(),
                                             det.data_ptr(), cnt.data_ptr(), anc.data_ptr(),
                                             _capi.current_stream_ptr()), "yx_postprocess_yolox")
    return det, cnt, anc


def postprocess(prediction, num_classes: int, conf_threshold: float = 0.7, nms_threshold: float = 0.45,
                class_agnostic: bool = False):
    """yolox/utils/boxes.py:32-82.  prediction[:, :, :4] is converted to xyxy IN PLACE (as the reference
    does); returns list[B] of [n,7] = [x1,y1,x2,y2,obj,class_conf,class_pred] in prediction's dtype, or None."""
    det, cnt, _ = postprocess_raw(prediction, num_classes, conf_threshold, nms_threshold, class_agnostic)
    return _split(det, cnt, prediction.dtype)
