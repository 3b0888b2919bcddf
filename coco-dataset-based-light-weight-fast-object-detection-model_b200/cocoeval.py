"""COCO bounding-box evaluation through the native library (yx_cocoeval_bbox): what the reference gets from
pycocotools' COCOeval(cocoGt, cocoDt, "bbox") evaluate() / accumulate() / summarize()
(yolox/evaluators/coco_evaluator.py:198-215, yolox/layers/fast_coco_eval_api.py)."""
import ctypes
from typing import Iterable, List, Optional, Sequence

import numpy as np

from . import _capi

_TITLES = [("Average Precision", "(AP)", "0.50:0.95", "all", 100), ("Average Precision", "(AP)", "0.50", "all", 100),
           ("Average Precision", "(AP)", "0.75", "all", 100), ("Average Precision", "(AP)", "0.50:0.95", "small", 100),
           ("Average Precision", "(AP)", "0.50:0.95", "medium", 100), ("Average Precision", "(AP)", "0.50:0.95", "large", 100),
           ("Average Recall", "(AR)", "0.50:0.95", "all", 1), ("Average Recall", "(AR)", "0.50:0.95", "all", 10),
           ("Average Recall", "(AR)", "0.50:0.95", "all", 100), ("Average Recall", "(AR)", "0.50:0.95", "small", 100),
           ("Average Recall", "(AR)", "0.50:0.95", "medium", 100), ("Average Recall", "(AR)", "0.50:0.95", "large", 100)]


def _ground_truth(gt):
    """Accepts a pycocotools-style COCO object (has .dataset), a COCO json dict, or a list of annotation dicts.
    -> (annotations, image ids, category ids)."""
    ds = getattr(gt, "dataset", gt)
    if isinstance(ds, dict):
        anns = ds.get("annotations", [])
        img_ids = [im["id"] for im in ds.get("images", [])] or sorted({a["image_id"] for a in anns})
        cat_ids = [c["id"] for c in ds.get("categories", [])] or sorted({a["category_id"] for a in anns})
        return anns, img_ids, cat_ids
    anns = list(ds)
    return anns, sorted({a["image_id"] for a in anns}), sorted({a["category_id"] for a in anns})


class COCOevalBBox:
    """evaluate() fills .stats (12 numbers), .precision [T,R,K,A,M], .recall [T,K,A,M]; summarize() returns the text
    pycocotools prints."""

    def __init__(self, gt, detections: Sequence[dict], img_ids: Optional[Iterable[int]] = None,
                 cat_ids: Optional[Iterable[int]] = None):
        self.anns, gi, gc = _ground_truth(gt)
        self.dets = list(detections)
        self.img_ids = list(gi if img_ids is None else img_ids)
        self.cat_ids = list(gc if cat_ids is None else cat_ids)
        self.stats = self.precision = self.recall = None

    def evaluate(self):
        lib = _capi.load()
        A, D = self.anns, self.dets

        def arr(values, dtype, width=None):
            a = np.ascontiguousarray(np.array(list(values), dtype=dtype))
            return a.reshape(-1, width) if width else a

        def area(a):  # COCO.loadRes fills a missing area with w*h; ground truth carries its own (segment area)
            return a["area"] if "area" in a else a["bbox"][2] * a["bbox"][3]

        gi, gc = arr((a["image_id"] for a in A), np.int64), arr((a["category_id"] for a in A), np.int32)
        gb = arr((a["bbox"] for a in A), np.float64, 4) if A else np.zeros((0, 4))
        ga, gw = arr((area(a) for a in A), np.float64), arr((int(a.get("iscrowd", 0)) for a in A), np.int32)
        di, dc = arr((d["image_id"] for d in D), np.int64), arr((d["category_id"] for d in D), np.int32)
        db = arr((d["bbox"] for d in D), np.float64, 4) if D else np.zeros((0, 4))
        dsc = arr((d["score"] for d in D), np.float64)
        ii, cc = arr(self.img_ids, np.int64), arr(self.cat_ids, np.int32)
        K = len(np.unique(cc))
        stats = np.zeros(12)
        prec, rec = np.zeros((10, 101, K, 4, 3)), np.zeros((10, K, 4, 3))
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        _capi.check(lib.yx_cocoeval_bbox(p(gi), p(gc), p(gb), p(ga), p(gw), len(A), p(di), p(dc), p(db), p(dsc), len(D),
                                         p(ii), len(ii), p(cc), len(cc), p(stats), p(prec), p(rec)), "yx_cocoeval_bbox")
        self.stats, self.precision, self.recall = stats, prec, rec
        return self

    def summarize(self) -> str:
        if self.stats is None:
            self.evaluate()
        lines = []
        for (title, typ, iou, area, md), v in zip(_TITLES, self.stats):
            lines.append(" {:<18} {} @[ IoU={:<9} | area={:>6s} | maxDets={:>3d} ] = {:0.3f}".format(title, typ, iou, area, md, v))
        return "\n".join(lines) + "\n"
