"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the COCO bounding-box evaluation that ends the reference's evaluator
(yolox/evaluators/coco_evaluator.py:198-215 -> pycocotools.cocoeval.COCOeval(cocoGt, cocoDt, "bbox") .evaluate() /
.accumulate() / .summarize(), optionally through yolox/layers/fast_coco_eval_api.py).

pycocotools is a third-party dependency (requirements.txt: "pycocotools>=2.0.2", unpinned) and is NOT installed in this
image, so this file restates its published algorithm function by function (COCOeval._prepare, computeIoU, evaluateImg,
accumulate, summarize; maskApi bbIou for the box IoU).  PIN: evaluateImg + accumulate are checked bit for bit against the
reference's OWN C++ COCO evaluation compiled into oracle/_ref (oracle/cocoeval_cpp.py, tests/test_cocoeval.py);
_prepare / computeIoU / summarize remain restated third-party code (grouping, the box IoU formula, twelve means).
It also carries hand-computed cases (tests/test_cocoeval.py) and is the independent checker of the product's native
implementation (csrc/yx_cocoeval.cu), which is structured differently on purpose.
"""
from collections import defaultdict

import numpy as np

IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
MAX_DETS = [1, 10, 100]
AREA_RNG = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]


def bb_iou(d, g, crowd):
    """maskApi.c bbIou for one pair of [x, y, w, h] boxes (double precision)."""
    w = min(d[2] + d[0], g[2] + g[0]) - max(d[0], g[0])
    if w <= 0:
        return 0.0
    h = min(d[3] + d[1], g[3] + g[1]) - max(d[1], g[1])
    if h <= 0:
        return 0.0
    i = w * h
    u = d[2] * d[3] if crowd else d[2] * d[3] + g[2] * g[3] - i
    return i / u


def evaluate(gts, dts, img_ids, cat_ids, cpp_precision=True):
    """gts: list of dict(image_id, category_id, bbox [x,y,w,h], area, iscrowd); dts: list of dict(image_id, category_id,
    bbox, score).  Returns dict(stats [12], precision [T,R,K,A,M], recall [T,K,A,M]).

    cpp_precision: the reference evaluator first tries its C++ accelerator (yolox/evaluators/coco_evaluator.py:204-205),
    whose precision is tp / (tp + fp) exactly (yolox/layers/csrc/cocoeval/cocoeval.cpp:332-335); False gives pycocotools'
    Python form tp / (fp + tp + np.spacing(1)), the fallback when the extension is not built.  The two differ by one ulp
    on some table entries (found by pinning this file against the compiled reference, oracle/cocoeval_cpp.py)."""
    img_ids = list(np.unique(img_ids))
    cat_ids = list(np.unique(cat_ids))
    _gts, _dts = defaultdict(list), defaultdict(list)
    for n, g in enumerate(gts):
        g = dict(g, id=n + 1, ignore=int(g.get("iscrowd", 0)))
        _gts[g["image_id"], g["category_id"]].append(g)
    for n, d in enumerate(dts):
        d = dict(d, id=n + 1, area=d["bbox"][2] * d["bbox"][3])
        _dts[d["image_id"], d["category_id"]].append(d)

    def compute_iou(i, k):
        gt, dt = _gts[i, k], _dts[i, k]
        if len(gt) == 0 and len(dt) == 0:
            return []
        inds = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[j] for j in inds][:MAX_DETS[-1]]
        out = np.zeros((len(dt), len(gt)))
        for a, d in enumerate(dt):
            for b, g in enumerate(gt):
                out[a, b] = bb_iou(d["bbox"], g["bbox"], bool(g.get("iscrowd", 0)))
        return out

    ious = {(i, k): compute_iou(i, k) for i in img_ids for k in cat_ids}

    def evaluate_img(i, k, rng, max_det):
        gt, dt = _gts[i, k], _dts[i, k]
        if len(gt) == 0 and len(dt) == 0:
            return None
        ig = [1 if (g["ignore"] or g["area"] < rng[0] or g["area"] > rng[1]) else 0 for g in gt]
        gtind = np.argsort(ig, kind="mergesort")
        gt = [gt[j] for j in gtind]
        dtind = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt = [dt[j] for j in dtind[:max_det]]
        iscrowd = [int(g.get("iscrowd", 0)) for g in gt]
        io = ious[i, k][:, gtind] if len(ious[i, k]) > 0 else ious[i, k]
        T, G, D = len(IOU_THRS), len(gt), len(dt)
        gtm, dtm = np.zeros((T, G)), np.zeros((T, D))
        gt_ig = np.array([ig[j] for j in gtind])
        dt_ig = np.zeros((T, D))
        if not len(io) == 0:
            for tind, t in enumerate(IOU_THRS):
                for dind, d in enumerate(dt):
                    iou = min([t, 1 - 1e-10])
                    m = -1
                    for gind, g in enumerate(gt):
                        if gtm[tind, gind] > 0 and not iscrowd[gind]:
                            continue
                        if m > -1 and gt_ig[m] == 0 and gt_ig[gind] == 1:
                            break
                        if io[dind, gind] < iou:
                            continue
                        iou = io[dind, gind]
                        m = gind
                    if m == -1:
                        continue
                    dt_ig[tind, dind] = gt_ig[m]
                    dtm[tind, dind] = gt[m]["id"]
                    gtm[tind, m] = d["id"]
        a = np.array([d["area"] < rng[0] or d["area"] > rng[1] for d in dt]).reshape((1, len(dt)))
        dt_ig = np.logical_or(dt_ig, np.logical_and(dtm == 0, np.repeat(a, T, 0)))
        return dict(dtm=dtm, scores=[d["score"] for d in dt], gt_ig=gt_ig, dt_ig=dt_ig)

    max_det = MAX_DETS[-1]
    eval_imgs = [evaluate_img(i, k, rng, max_det) for k in cat_ids for rng in AREA_RNG for i in img_ids]

    T, R, K, A, M = len(IOU_THRS), len(REC_THRS), len(cat_ids), len(AREA_RNG), len(MAX_DETS)
    precision = -np.ones((T, R, K, A, M))
    recall = -np.ones((T, K, A, M))
    I0 = len(img_ids)
    for k in range(K):
        for a in range(A):
            for m, md in enumerate(MAX_DETS):
                E = [eval_imgs[k * A * I0 + a * I0 + i] for i in range(I0)]
                E = [e for e in E if e is not None]
                if len(E) == 0:
                    continue
                scores = np.concatenate([e["scores"][0:md] for e in E])
                inds = np.argsort(-scores, kind="mergesort")
                dtm = np.concatenate([e["dtm"][:, 0:md] for e in E], axis=1)[:, inds]
                dt_ig = np.concatenate([e["dt_ig"][:, 0:md] for e in E], axis=1)[:, inds]
                gt_ig = np.concatenate([e["gt_ig"] for e in E])
                npig = np.count_nonzero(gt_ig == 0)
                if npig == 0:
                    continue
                tps = np.logical_and(dtm, np.logical_not(dt_ig))
                fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
                fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
                for t, (tp, fp) in enumerate(zip(tp_sum, fp_sum)):
                    tp, fp = np.array(tp), np.array(fp)
                    nd = len(tp)
                    rc = tp / npig
                    if cpp_precision:
                        with np.errstate(invalid="ignore", divide="ignore"):
                            pr = np.where(tp + fp > 0, tp / (tp + fp), 0.0)
                    else:
                        pr = tp / (fp + tp + np.spacing(1))
                    q = np.zeros((R,))
                    recall[t, k, a, m] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    for j in range(nd - 1, 0, -1):
                        if pr[j] > pr[j - 1]:
                            pr[j - 1] = pr[j]
                    pos = np.searchsorted(rc, REC_THRS, side="left")
                    try:
                        for ri, pi in enumerate(pos):
                            q[ri] = pr[pi]
                    except IndexError:
                        pass
                    precision[t, :, k, a, m] = np.array(q)

    def summ(ap, iou_thr=None, area=0, md=2):
        if ap:
            s = precision
            if iou_thr is not None:
                s = s[np.where(iou_thr == IOU_THRS)[0]]
            s = s[:, :, :, area, md]
        else:
            s = recall[:, :, area, md]
        return -1.0 if len(s[s > -1]) == 0 else float(np.mean(s[s > -1]))

    stats = [summ(1), summ(1, .5), summ(1, .75), summ(1, area=1), summ(1, area=2), summ(1, area=3),
             summ(0, md=0), summ(0, md=1), summ(0, md=2), summ(0, area=1), summ(0, area=2), summ(0, area=3)]
    return dict(stats=np.array(stats), precision=precision, recall=recall)
