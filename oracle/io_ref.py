"""CPU oracle (TEST INFRASTRUCTURE ONLY — never imported by the product path) for the rows either side of the hot path
(SURVEY §8f N1 / N2):

  N1  pre-processing: aspect-preserving Pillow BILINEAR resize + top-left paste into a 114-filled batch padded to a
      multiple of 64, RGB -> BGR, NCHW, no normalisation
      reference: choijhanyangackr/yolox_infer/preprocess_utils.py:9-55 (yolox_load_one_image_pil, yolox_collate_batch)
      third-party arithmetic: Pillow's ImagingResample (src/libImaging/Resample.c; Pillow is unpinned in
      choijhanyangackr/requirements.txt, 12.2.0 installed) — restated here from its published algorithm: per-axis
      coefficient tables over a support scaled by the down-sampling ratio, normalised in double precision, converted
      to 22-bit fixed point, horizontal pass then vertical pass, each rounded and clipped to uint8.
  N2  detections -> COCO records: unscale by min(S/h, S/w), xyxy -> xywh, score = obj * cls (det columns 4 and 5),
      category id from the 80-class COCO table
      reference: choijhanyangackr/common/utils.py:5-73 (convert_to_coco_format)

Pinned by tests/test_oracle_io.py against Pillow itself and against golden vectors produced by the reference functions
(tests/golden/make_golden_io.py).
"""
import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c

COCO_CLASS_ID = [
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32,
    33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59,
    60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90
]  # common/utils.py:5-10


# ------------------------------------------------------------------------------------------------ N1
def resized_shape(h: int, w: int, img_size: int) -> Tuple[int, int]:
    """preprocess_utils.py:17-22 -> (new_h, new_w)."""
    if w > h:
        new_w = img_size
        new_h = int(h * new_w / w)
    else:
        new_h = img_size
        new_w = int(w * new_h / h)
    return new_h, new_w


def bilinear_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0).
    Returns bounds int32 [out, 2] = (first input index, tap count) and kk int32 [out, ksize] (22-bit fixed point)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w = 1.0 - t if t < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            k[:xmax] /= ww
        for x in range(ksize):
            v = k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_bilinear(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """img uint8 [h, w, 3] -> uint8 [new_h, new_w, 3], bit-identical to PIL.Image.resize((new_w, new_h), BILINEAR):
    horizontal pass, then vertical pass (ImagingResample)."""
    h, w, _ = img.shape
    bh, kh = bilinear_coeffs(w, new_w)
    bv, kv = bilinear_coeffs(h, new_h)
    half = 1 << (PRECISION_BITS - 1)
    tmp = np.zeros((h, new_w, 3), np.uint8)
    src = img.astype(np.int64)
    for xx in range(new_w):
        x0, n = bh[xx]
        acc = np.full((h, 3), half, np.int64)
        for i in range(n):
            acc += src[:, x0 + i, :] * int(kh[xx, i])
        tmp[:, xx, :] = _clip8(acc)
    out = np.zeros((new_h, new_w, 3), np.uint8)
    t64 = tmp.astype(np.int64)
    for yy in range(new_h):
        y0, n = bv[yy]
        acc = np.full((new_w, 3), half, np.int64)
        for j in range(n):
            acc += t64[y0 + j] * int(kv[yy, j])
        out[yy] = _clip8(acc)
    return out


def collate(images: Sequence[np.ndarray], img_size: int) -> Tuple[np.ndarray, List[Tuple[int, int]]]:
    """yolox_load_one_image_pil + yolox_collate_batch on decoded RGB uint8 images.
    -> float32 [B, 3, max_h, max_w] (BGR, 0-255, pad 114) and [(h, w)] of the originals."""
    resized, info = [], []
    for im in images:
        h, w, _ = im.shape
        nh, nw = resized_shape(h, w, img_size)
        resized.append(pil_resize_bilinear(im, nw, nh))
        info.append((h, w))
    mult = 64 if img_size % 64 == 0 else 32
    max_h = int(math.ceil(max(r.shape[0] for r in resized) / mult) * mult)
    max_w = int(math.ceil(max(r.shape[1] for r in resized) / mult) * mult)
    batch = np.full((len(images), max_h, max_w, 3), 114, np.uint8)
    for i, r in enumerate(resized):
        batch[i, :r.shape[0], :r.shape[1], :] = r[..., ::-1]
    return np.ascontiguousarray(batch.transpose(0, 3, 1, 2), dtype=np.float32), info


# ------------------------------------------------------------------------------------------------ N2
def coco_records(det: np.ndarray, count: np.ndarray, img_hw: Sequence[Tuple[int, int]], img_size: int,
                 class_ids=None) -> np.ndarray:
    """det float32 [B, max_det, 7] = [x1,y1,x2,y2,obj,cls*obj,label], count [B].
    -> float32 [B, max_det, 6] = [x, y, w, h, score, category_id] (rows >= count are zero), with the reference's
    arithmetic: fp32 division of the corners by float32(scale), then w = x2' - x1', score = col4 * col5."""
    class_ids = np.asarray(COCO_CLASS_ID if class_ids is None else class_ids, np.float32)
    B, M, _ = det.shape
    out = np.zeros((B, M, 6), np.float32)
    for b in range(B):
        n = int(count[b])
        h, w = img_hw[b]
        scale = np.float32(min(img_size / float(h), img_size / float(w)))
        d = det[b, :n].astype(np.float32)
        box = d[:, :4] / scale
        out[b, :n, 0] = box[:, 0]
        out[b, :n, 1] = box[:, 1]
        out[b, :n, 2] = box[:, 2] - box[:, 0]
        out[b, :n, 3] = box[:, 3] - box[:, 1]
        out[b, :n, 4] = d[:, 4] * d[:, 5]
        out[b, :n, 5] = class_ids[d[:, 6].astype(np.int64)]
    return out
