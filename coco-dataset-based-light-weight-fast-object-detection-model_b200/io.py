"""The rows either side of the hot path with the reference's call surface, on the device.

  preprocess_batch(images, img_size)            == yolox_load_one_image_pil (minus the file decode) + yolox_collate_batch,
                                                   choijhanyangackr/yolox_infer/preprocess_utils.py:9-55
  convert_to_coco_format(outputs, img_info, …)  == choijhanyangackr/common/utils.py:27-73 (same records, same values)
  coco_records(det, count, img_hw, img_size)     the device tensor underneath it

Pixel / box arithmetic runs in csrc/yx_io.cu; the host side only builds Pillow's per-axis coefficient tables (O(W + H) per
image, double precision like Resample.c) and Python dicts.  No CPU fallback."""
import functools
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi

PRECISION_BITS = 32 - 8 - 2

COCO_CLASS_ID = [
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32,
    33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59,
    60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90
]


def resized_shape(h: int, w: int, img_size: int) -> Tuple[int, int]:
    """(new_h, new_w) exactly as preprocess_utils.py:17-22 computes them."""
    if w > h:
        new_w = img_size
        new_h = int(h * new_w / w)
    else:
        new_h = img_size
        new_w = int(w * new_h / h)
    return new_h, new_w


@functools.lru_cache(maxsize=512)
def _coeffs(in_size: int, out_size: int):
    """(cached per (in, out) pair: datasets repeat a handful of image sizes; the arrays are treated as read-only)
    Pillow's precompute_coeffs + normalize_coeffs_8bpc for the triangle (BILINEAR) filter, vectorised over the output
    index in float64 (same operation order as the C code).  -> bounds int32 [out, 2], kk int32 [out, ksize]."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # (int) truncation of a positive value
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    t = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * (1.0 / filterscale))
    k = np.where(t < 1.0, 1.0 - t, 0.0)
    k = np.where(x < xmax[:, None], k, 0.0)
    ww = np.zeros(out_size)
    for i in range(ksize):                                                     # sequential sum like the C loop
        ww = ww + k[:, i]
    k = np.where(ww[:, None] != 0.0, k / np.where(ww[:, None] != 0.0, ww[:, None], 1.0), k)
    kk = (0.5 + k * (1 << PRECISION_BITS)).astype(np.int64).astype(np.int32)   # weights are >= 0 for this filter
    return np.stack([xmin, xmax], 1).astype(np.int32), kk


_STAGING = {"buf": None, "event": None}


def _staging(nbytes: int) -> torch.Tensor:
    """Grow-only pinned host buffer for the raw pixels of a batch (pinning fresh memory per call costs milliseconds)."""
    if _STAGING["event"] is not None:
        _STAGING["event"].synchronize()
    if _STAGING["buf"] is None or _STAGING["buf"].numel() < nbytes:
        _STAGING["buf"] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8).pin_memory()
    return _STAGING["buf"]


def preprocess_batch(images: Sequence, img_size: int, device="cuda", dtype=torch.float32):
    """images: decoded RGB uint8 arrays / tensors [h, w, 3] (what PIL's Image.open(...).convert("RGB") holds).
    Returns (batch [B,3,Hp,Wp] on `device`, BGR 0..255, pad 114, Hp/Wp multiples of 64 (32 if img_size % 64),
             img_info [(h, w)]) — the tensor yolox_collate_batch builds, computed on the GPU."""
    lib = _capi.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("yolox_b200.io.preprocess_batch runs on CUDA only; there is no CPU path")
    arrs = [np.ascontiguousarray(im.cpu().numpy() if isinstance(im, torch.Tensor) else im) for im in images]
    for a in arrs:
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise RuntimeError("images must be uint8 [h, w, 3] RGB")
    B = len(arrs)
    geom = np.zeros((B, 4), np.int32)
    tabs = []
    for i, a in enumerate(arrs):
        h, w = a.shape[:2]
        nh, nw = resized_shape(h, w, img_size)
        geom[i] = (h, w, nh, nw)
        tabs.append((_coeffs(w, nw), _coeffs(h, nh)))
    mult = 64 if img_size % 64 == 0 else 32
    Hp = int(math.ceil(geom[:, 2].max() / mult) * mult)
    Wp = int(math.ceil(geom[:, 3].max() / mult) * mult)
    ks_h = max(t[0][1].shape[1] for t in tabs)
    ks_v = max(t[1][1].shape[1] for t in tabs)
    bh, kh = np.zeros((B, Wp, 2), np.int32), np.zeros((B, Wp, ks_h), np.int32)
    bv, kv = np.zeros((B, Hp, 2), np.int32), np.zeros((B, Hp, ks_v), np.int32)
    for i, ((b0, k0), (b1, k1)) in enumerate(tabs):
        bh[i, :b0.shape[0]] = b0; kh[i, :k0.shape[0], :k0.shape[1]] = k0
        bv[i, :b1.shape[0]] = b1; kv[i, :k1.shape[0], :k1.shape[1]] = k1
    offs = np.zeros(B, np.int64)
    total = 0
    for i, a in enumerate(arrs):
        offs[i] = total
        total += a.size
    packed = _staging(total)
    pk = packed.numpy()
    for i, a in enumerate(arrs):
        np.copyto(pk[offs[i]:offs[i] + a.size], a.reshape(-1))
    with torch.cuda.device(dev):
        d = lambda x: torch.from_numpy(x).to(dev, non_blocking=True)
        src, d_off, d_geom = packed[:total].to(dev, non_blocking=True), d(offs), d(geom)
        _STAGING["event"] = torch.cuda.Event()
        _STAGING["event"].record()                      # the staging buffer may be rewritten once this copy has run
        d_bh, d_kh, d_bv, d_kv = d(bh), d(kh), d(bv), d(kv)
        out = torch.empty(B, 3, Hp, Wp, dtype=dtype, device=dev)
        if dtype not in (torch.float16, torch.float32, torch.uint8):
            raise RuntimeError("dtype must be float16, float32 or uint8")
        dt = {torch.float16: _capi.YX_F16, torch.float32: _capi.YX_F32, torch.uint8: _capi.YX_U8}[dtype]
        _capi.check(lib.yx_preprocess_batch(src.data_ptr(), d_off.data_ptr(), d_geom.data_ptr(), d_bh.data_ptr(),
                                            d_kh.data_ptr(), d_bv.data_ptr(), d_kv.data_ptr(), ks_h, ks_v, B, Hp, Wp,
                                            out.data_ptr(), dt, _capi.current_stream_ptr()), "yx_preprocess_batch")
        for t in (src, d_off, d_geom, d_bh, d_kh, d_bv, d_kv):
            t.record_stream(torch.cuda.current_stream())
    return out, [(int(g[0]), int(g[1])) for g in geom]


def coco_records(det: torch.Tensor, count: torch.Tensor, img_hw: Sequence[Tuple[int, int]], img_size,
                 class_ids: Optional[Sequence[int]] = None) -> torch.Tensor:
    """det [B,max_det,7] fp32 (detect_main / Predictor output), count [B] -> records [B,max_det,6] fp32 on the device:
    [x, y, w, h, score, category_id], rows beyond count zero."""
    _capi.require_cuda(det, "det")
    lib = _capi.load()
    B, M, seven = det.shape
    if seven != 7 or det.dtype != torch.float32:
        raise RuntimeError("det must be float32 [B, max_det, 7]")
    ids = torch.tensor(COCO_CLASS_ID if class_ids is None else list(class_ids), dtype=torch.int32, device=det.device)
    sh, sw = (img_size, img_size) if isinstance(img_size, (int, float)) else img_size   # evaluator: (h, w)
    scale = torch.tensor([min(sh / float(h), sw / float(w)) for h, w in img_hw], dtype=torch.float32,
                         device=det.device)   # Python double -> float32, like `boxes /= scale` on a float32 tensor
    rec = torch.empty(B, M, 6, dtype=torch.float32, device=det.device)
    det = det.contiguous()
    cnt = count.to(torch.int32).contiguous()
    with torch.cuda.device(det.device):
        _capi.check(lib.yx_coco_records(det.data_ptr(), cnt.data_ptr(), B, M, scale.data_ptr(), ids.data_ptr(), ids.numel(),
                                        rec.data_ptr(), _capi.current_stream_ptr()), "yx_coco_records")
    return rec


def convert_to_coco_format(outputs, img_info, img_size, class_ids=None) -> List[dict]:
    """Same signature and same records as common/utils.py:27-73.  outputs: the reference's list of [n,7] tensors / None,
    or the engine's (det [B,max_det,7], count [B]) pair.  img_info: [(h, w, file_name)]."""
    if isinstance(outputs, tuple) and len(outputs) == 2 and isinstance(outputs[0], torch.Tensor) and outputs[0].dim() == 3:
        det, count = outputs
    else:
        B = len(outputs)
        M = max([o.shape[0] for o in outputs if o is not None] + [1])
        dev = next((o.device for o in outputs if o is not None), torch.device("cuda"))
        det = torch.zeros(B, M, 7, dtype=torch.float32, device=dev)
        count = torch.zeros(B, dtype=torch.int32, device=dev)
        for i, o in enumerate(outputs):
            if o is not None:
                det[i, :o.shape[0]] = o.float()
                count[i] = o.shape[0]
    rec = coco_records(det, count, [(h, w) for h, w, _ in img_info], img_size, class_ids).cpu()
    cnt = count.cpu().tolist()
    data_list = []
    for b, (img_h, img_w, img_path) in enumerate(img_info):
        image_id = int(img_path.split("_")[-1].split(".")[0])
        if cnt[b] == 0:                                     # utils.py:43-51 (the reference's `output is None` case)
            data_list.append({"image_id": image_id, "category_id": 0, "bbox": [0, 0, 0, 0], "score": 0.0})
            continue
        rows = rec[b, :cnt[b]].tolist()
        for r in rows:
            data_list.append({"image_id": image_id, "category_id": int(r[5]), "bbox": r[:4], "score": r[4]})
    return data_list
