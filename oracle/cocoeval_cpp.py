"""The reference's OWN compiled COCO evaluation (TEST INFRASTRUCTURE ONLY): yolox/layers/csrc/cocoeval/cocoeval.cpp
(EvaluateImages :140-199, Accumulate :370-502) + csrc/vision.cpp, compiled where they lie under /root/reference into
oracle/_ref/ by build() below (g++ and pybind11 only; no source is copied into this repo), and driven the way
yolox/layers/fast_coco_eval_api.py:25-147 drives it for iouType="bbox".

What this pins: the per-image greedy matching and the precision / recall accumulation of the restatement
(oracle/cocoeval_ref.py) and of the product (csrc/yx_cocoeval.cu) against the reference's C++.  What stays restated:
pycocotools' COCOeval._prepare and computeIoU (maskApi bbIou), which are third-party Python/C not present in this image;
they only group the annotations and evaluate the box IoU formula.
"""
import glob
import importlib.util
import os
import subprocess
import sys
import sysconfig

import numpy as np

from . import cocoeval_ref as cr

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/yolox/layers/csrc"
MODULE = "yolox_ref_C"


def lib_path():
    return os.path.join(REF_DIR, MODULE + sysconfig.get_config_var("EXT_SUFFIX"))


def available() -> bool:
    return os.path.exists(lib_path())


def build(force: bool = False) -> str:
    """g++ on the reference's two source files, output only under oracle/_ref/ (git-ignored, travels with gpurun)."""
    out = lib_path()
    if os.path.exists(out) and not force:
        return out
    srcs = [os.path.join(REF_SRC, "vision.cpp"), os.path.join(REF_SRC, "cocoeval", "cocoeval.cpp")]
    if not all(os.path.exists(s) for s in srcs):
        raise RuntimeError("the reference tree is not present: oracle/_ref can only be built in the build container")
    import pybind11
    os.makedirs(REF_DIR, exist_ok=True)
    tmp = out + f".{os.getpid()}.tmp"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-DTORCH_EXTENSION_NAME={MODULE}",
                           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-I" + REF_SRC]
                          + srcs + ["-o", tmp])
    os.replace(tmp, out)
    return out


def _module():
    if MODULE in sys.modules:
        return sys.modules[MODULE]
    spec = importlib.util.spec_from_file_location(MODULE, lib_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[MODULE] = mod
    return mod


class _Params:
    """The attributes Accumulate reads from pycocotools' Params (cocoeval.cpp:373-385)."""

    def __init__(self, img_ids, cat_ids):
        self.imgIds, self.catIds = list(img_ids), list(cat_ids)
        self.iouThrs, self.recThrs = cr.IOU_THRS, cr.REC_THRS
        self.maxDets, self.areaRng, self.useCats = list(cr.MAX_DETS), [list(a) for a in cr.AREA_RNG], 1


def evaluate(gts, dts, img_ids, cat_ids):
    """Same inputs / outputs as cocoeval_ref.evaluate, with evaluateImg + accumulate done by the reference's C++."""
    C = _module()
    img_ids = [int(i) for i in np.unique(img_ids)]
    cat_ids = [int(c) for c in np.unique(cat_ids)]
    _gts, _dts = {}, {}
    for n, g in enumerate(gts):      # COCOeval._prepare: ids as COCO.createIndex / loadRes assign them, ignore = iscrowd
        _gts.setdefault((g["image_id"], g["category_id"]), []).append(dict(g, id=n + 1, ignore=int(g.get("iscrowd", 0))))
    for n, d in enumerate(dts):
        _dts.setdefault((d["image_id"], d["category_id"]), []).append(dict(d, id=n + 1, area=d["bbox"][2] * d["bbox"][3]))
    max_det = cr.MAX_DETS[-1]

    def ious(i, k):                  # COCOeval.computeIoU for iouType == "bbox"
        gt, dt = _gts.get((i, k), []), _dts.get((i, k), [])
        if len(gt) == 0 and len(dt) == 0:
            return []
        inds = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[j] for j in inds][:max_det]
        return [[cr.bb_iou(d["bbox"], g["bbox"], bool(g.get("iscrowd", 0))) for g in gt] for d in dt]

    def inst(objs, is_det):          # fast_coco_eval_api.py:70-83
        return [C.InstanceAnnotation(int(o["id"]), float(o["score"] if is_det else o.get("score", 0.0)), float(o["area"]),
                                     bool(o.get("iscrowd", 0)), bool(o.get("ignore", 0))) for o in objs]

    gt_inst = [[inst(_gts.get((i, k), []), False) for k in cat_ids] for i in img_ids]
    dt_inst = [[inst(_dts.get((i, k), []), True) for k in cat_ids] for i in img_ids]
    iou_all = [[ious(i, k) for k in cat_ids] for i in img_ids]
    ev = C.COCOevalEvaluateImages([list(map(float, a)) for a in cr.AREA_RNG], max_det, [float(t) for t in cr.IOU_THRS],
                                  iou_all, gt_inst, dt_inst)
    acc = C.COCOevalAccumulate(_Params(img_ids, cat_ids), ev)
    counts = list(acc["counts"])
    precision = np.array(acc["precision"]).reshape(counts)
    recall = np.array(acc["recall"]).reshape(counts[:1] + counts[2:])
    return dict(precision=precision, recall=recall)
