"""COCO bbox evaluation: hand-computed known answers for the oracle (pycocotools is not installed: parity unpinned,
see oracle/cocoeval_ref.py), then the native implementation (csrc/yx_cocoeval.cu through the C ABI, host code — runs
without a GPU) against the oracle on randomised datasets with crowds, all area ranges, ties and > 100 detections."""
import os

import numpy as np
import pytest

import yolox_b200 as yb
from oracle import cocoeval_ref as cr


def _gt(img, cat, box, crowd=0, area=None):
    return dict(image_id=img, category_id=cat, bbox=list(map(float, box)), iscrowd=crowd,
                area=float(box[2] * box[3] if area is None else area))


def _dt(img, cat, box, score):
    return dict(image_id=img, category_id=cat, bbox=list(map(float, box)), score=float(score))


def _both(gts, dts, imgs, cats):
    ref = cr.evaluate(gts, dts, imgs, cats)
    ev = yb.cocoeval.COCOevalBBox(gts, dts, imgs, cats).evaluate()
    np.testing.assert_array_equal(ev.precision, ref["precision"])   # the tables are bit-identical
    np.testing.assert_array_equal(ev.recall, ref["recall"])
    return ref, ev


def test_known_answers():
    # 1) perfect detections: AP = AR = 1 in the populated area range, -1 where no ground truth falls
    gts = [_gt(1, 1, (10, 10, 50, 50)), _gt(1, 2, (100, 100, 20, 20)), _gt(2, 1, (0, 0, 200, 200))]
    dts = [_dt(g["image_id"], g["category_id"], g["bbox"], 0.9) for g in gts]
    ref, ev = _both(gts, dts, [1, 2], [1, 2])
    np.testing.assert_allclose(ref["stats"][[0, 1, 2, 6, 7, 8]], 1.0)
    np.testing.assert_allclose(ref["stats"][[3, 4, 5]], 1.0)   # 400 px^2 small, 2500 medium, 40000 large
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only
    # 2) no detections at all: recall 0, precision 0 everywhere a ground truth exists
    ref, ev = _both(gts, [], [1, 2], [1, 2])
    assert ref["stats"][0] == 0.0 and ref["stats"][8] == 0.0
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only
    # 3) one gt, detections [FP (score .9), TP (score .8)]: precision after the envelope is 0.5 at every recall point
    gts = [_gt(1, 1, (0, 0, 100, 100))]
    dts = [_dt(1, 1, (300, 300, 100, 100), 0.9), _dt(1, 1, (0, 0, 100, 100), 0.8)]
    ref, ev = _both(gts, dts, [1], [1])
    assert abs(ref["stats"][0] - 0.5) < 1e-12 and ref["stats"][6] == 0.0 and ref["stats"][7] == 1.0
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only
    # 4) IoU exactly between thresholds: box shifted so that IoU = 0.6 -> matched at .50/.55/.60, missed above
    #    (intersection 75 x 100 of two 100 x 100 boxes -> 7500 / 12500 = 0.6)
    dts = [_dt(1, 1, (25, 0, 100, 100), 0.9)]
    ref, ev = _both(gts, dts, [1], [1])
    assert abs(ref["stats"][0] - 0.3) < 1e-12 and abs(ref["stats"][1] - 1.0) < 1e-12 and ref["stats"][2] == 0.0
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only
    # 5) a crowd region absorbs any number of detections without penalty and is never counted as a miss
    gts = [_gt(1, 1, (0, 0, 100, 100)), _gt(1, 1, (200, 200, 100, 100), crowd=1)]
    dts = [_dt(1, 1, (0, 0, 100, 100), 0.9), _dt(1, 1, (210, 210, 50, 50), 0.8), _dt(1, 1, (220, 220, 50, 50), 0.7)]
    ref, ev = _both(gts, dts, [1], [1])
    assert abs(ref["stats"][0] - 1.0) < 1e-12 and ref["stats"][8] == 1.0
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only


def _random_dataset(seed, n_img, n_cat, max_gt, max_dt, tie_scores=False):
    rs = np.random.RandomState(seed)
    gts, dts = [], []
    for i in range(1, n_img + 1):
        for _ in range(rs.randint(0, max_gt + 1)):
            w, h = rs.choice([8, 20, 60, 150, 300]) * rs.uniform(0.5, 1.5, 2)
            x, y = rs.uniform(0, 500, 2)
            cat = int(rs.randint(1, n_cat + 1))
            crowd = int(rs.rand() < 0.1)
            gts.append(_gt(i, cat, (x, y, w, h), crowd, area=w * h * rs.uniform(0.5, 1.0)))
            if rs.rand() < 0.8:   # a detection near this ground truth, sometimes of the wrong class
                j = rs.normal(0, 0.08, 4) * (w, h, w, h)
                c2 = cat if rs.rand() < 0.85 else int(rs.randint(1, n_cat + 1))
                s = round(rs.rand(), 2) if tie_scores else rs.rand()
                dts.append(_dt(i, c2, (x + j[0], y + j[1], max(w + j[2], 1), max(h + j[3], 1)), s))
        for _ in range(rs.randint(0, max_dt + 1)):
            w, h = rs.uniform(5, 300, 2)
            s = round(rs.rand(), 2) if tie_scores else rs.rand()
            dts.append(_dt(i, int(rs.randint(1, n_cat + 1)), (*rs.uniform(0, 500, 2), w, h), s))
    rs.shuffle(dts)
    return gts, dts


@pytest.mark.parametrize("seed,n_img,n_cat,max_gt,max_dt,ties", [(0, 12, 3, 6, 10, False), (1, 5, 2, 3, 160, False),
                                                                 (2, 20, 5, 8, 12, True), (3, 3, 1, 0, 4, False)])
def test_native_matches_oracle(seed, n_img, n_cat, max_gt, max_dt, ties):
    gts, dts = _random_dataset(seed, n_img, n_cat, max_gt, max_dt, ties)
    imgs, cats = list(range(1, n_img + 1)), list(range(1, n_cat + 1))
    ref, ev = _both(gts, dts, imgs, cats)
    np.testing.assert_array_equal(ev.precision, ref["precision"])
    np.testing.assert_array_equal(ev.recall, ref["recall"])
    np.testing.assert_allclose(ev.stats, ref["stats"], rtol=0, atol=1e-12)  # mean: summation order only
    assert "Average Precision  (AP) @[ IoU=0.50:0.95 | area=   all | maxDets=100 ] = " in ev.summarize()


def test_accepts_coco_json_and_objects():
    gts, dts = _random_dataset(5, 4, 2, 3, 3)
    ds = dict(images=[dict(id=i) for i in range(1, 5)], categories=[dict(id=1), dict(id=2)], annotations=gts)

    class FakeCOCO:          # what pycocotools.coco.COCO exposes to the evaluator
        dataset = ds

    a = yb.cocoeval.COCOevalBBox(ds, dts).evaluate().stats
    b = yb.cocoeval.COCOevalBBox(FakeCOCO(), dts).evaluate().stats
    c = cr.evaluate(gts, dts, [1, 2, 3, 4], [1, 2])["stats"]
    np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(a, c, rtol=0, atol=1e-12)


def _cpp():
    """The reference's compiled COCOeval (oracle/_ref): built on demand where /root/reference exists."""
    from oracle import cocoeval_cpp as cc
    if not cc.available():
        if not os.path.isdir(cc.REF_SRC):
            pytest.skip("oracle/_ref is not built and the reference tree is not here")
        cc.build()
    return cc


@pytest.mark.parametrize("seed,n_img,n_cat,max_gt,max_dt,ties", [(0, 12, 3, 6, 10, False), (1, 5, 2, 3, 160, False),
                                                                 (2, 20, 5, 8, 12, True), (3, 3, 1, 0, 4, False),
                                                                 (7, 40, 6, 10, 30, True), (8, 6, 2, 12, 220, False)])
def test_reference_cpp_pins_matching_and_accumulation(seed, n_img, n_cat, max_gt, max_dt, ties):
    """Pins row N3: the reference's own cocoeval.cpp (EvaluateImages :140, Accumulate :370), compiled into oracle/_ref,
    gives the same precision / recall tables -- bit for bit -- as the restatement and as the native yx_cocoeval_bbox."""
    cc = _cpp()
    gts, dts = _random_dataset(seed, n_img, n_cat, max_gt, max_dt, ties)
    imgs, cats = list(range(1, n_img + 1)), list(range(1, n_cat + 1))
    want = cc.evaluate(gts, dts, imgs, cats)
    ref = cr.evaluate(gts, dts, imgs, cats)
    np.testing.assert_array_equal(ref["precision"], want["precision"])
    np.testing.assert_array_equal(ref["recall"], want["recall"])
    ev = yb.cocoeval.COCOevalBBox(gts, dts, imgs, cats).evaluate()
    np.testing.assert_array_equal(ev.precision, want["precision"])
    np.testing.assert_array_equal(ev.recall, want["recall"])


def test_reference_cpp_known_answers():
    cc = _cpp()
    gts = [_gt(1, 1, (0, 0, 100, 100)), _gt(1, 1, (200, 200, 100, 100), crowd=1)]
    dts = [_dt(1, 1, (0, 0, 100, 100), 0.9), _dt(1, 1, (210, 210, 50, 50), 0.8), _dt(1, 1, (25, 0, 100, 100), 0.7)]
    want = cc.evaluate(gts, dts, [1], [1])
    ev = yb.cocoeval.COCOevalBBox(gts, dts, [1], [1]).evaluate()
    np.testing.assert_array_equal(ev.precision, want["precision"])
    np.testing.assert_array_equal(ev.recall, want["recall"])
    assert want["recall"][0, 0, 0, 2] == 1.0
