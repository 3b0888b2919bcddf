"""Pins oracle/post_ref.c against (a) torchvision's own CPU ops (the third-party arithmetic the
reference calls at yolox/utils/boxes.py:62,68 and yolox_infer/nms.py:19,40) and (b) detections the
unmodified reference produced (tests/golden/post_*.npz)."""
import glob
import os

import numpy as np
import pytest
import torch
import torchvision
from torchvision.ops import boxes as tvb

from oracle import post_ref as pr

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "post_*.npz")))


def _rand_boxes(rs, n, extent=640.0, cluster=False):
    if cluster:
        c = rs.uniform(0, extent, (max(n // 20, 1), 2))
        ctr = c[rs.randint(0, len(c), n)] + rs.normal(0, 6, (n, 2))
    else:
        ctr = rs.uniform(0, extent, (n, 2))
    wh = rs.uniform(4, 120, (n, 2))
    return np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)


@pytest.mark.parametrize("n,cluster,thr", [(0, False, 0.5), (1, False, 0.5), (257, True, 0.65), (900, True, 0.45),
                                          (2000, True, 0.55), (1500, False, 0.3)])
def test_nms_matches_torchvision(n, cluster, thr):
    rs = np.random.RandomState(n + 7)
    b = _rand_boxes(rs, n, cluster=cluster)
    s = rs.uniform(0, 1, n).astype(np.float32)
    if n > 10:  # force score ties and duplicate boxes
        s[5:9] = s[4]; b[6] = b[5]
    ref = torchvision.ops.nms(torch.from_numpy(b).reshape(-1, 4), torch.from_numpy(s), thr).numpy()
    np.testing.assert_array_equal(pr.nms(b, s, thr), ref)


@pytest.mark.parametrize("n", [0, 3, 700, 3000])
def test_batched_nms_modes_match_torchvision(n):
    rs = np.random.RandomState(n + 11)
    b = _rand_boxes(rs, n, cluster=True)
    s = rs.uniform(0, 1, n).astype(np.float32)
    lab = rs.randint(0, 7, n).astype(np.float32)
    tb, ts, tl = torch.from_numpy(b).reshape(-1, 4), torch.from_numpy(s), torch.from_numpy(lab)
    if n == 0:
        assert len(pr.batched_nms(b, s, lab, 0.5, "trick")) == 0
        return
    np.testing.assert_array_equal(pr.batched_nms(b, s, lab, 0.5, "trick"),
                                  tvb._batched_nms_coordinate_trick(tb, ts, tl, 0.5).numpy())
    np.testing.assert_array_equal(pr.batched_nms(b, s, lab, 0.5, "vanilla"),
                                  tvb._batched_nms_vanilla(tb, ts, tl, 0.5).numpy())


def test_iou_edge_semantics():
    # IoU == thr is kept (strict >); zero-area pairs give NaN -> kept   
    b = np.array([[0, 0, 2, 2], [0, 0, 2, 1], [5, 5, 5, 5], [5, 5, 5, 5]], np.float32)
    s = np.array([0.9, 0.8, 0.7, 0.6], np.float32)
    np.testing.assert_array_equal(pr.nms(b, s, 0.5), [0, 1, 2, 3])
    np.testing.assert_array_equal(torchvision.ops.nms(torch.from_numpy(b), torch.from_numpy(s), 0.5).numpy(), [0, 1, 2, 3])
    np.testing.assert_array_equal(pr.nms(b, s, 0.49), [0, 2, 3])


def _levels(img, strides):
    return [(img // s, img // s) for s in strides]


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_golden_main_flavour(path):
    g = np.load(path)
    img, strides = int(g["img"]), [int(s) for s in g["strides"]]
    conf, thr = float(g["conf"]), float(g["nms_thr"])
    B = g["reg"].shape[0]
    for i in range(B):
        boxes, oc, cc = pr.decode_infer(g["reg"][i], g["obj"][i], g["cls"][i], _levels(img, strides), strides)
        np.testing.assert_allclose(boxes, g["boxes"][i], rtol=2e-6, atol=1e-4)
        np.testing.assert_allclose(oc, g["obj_conf"][i].reshape(-1), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(cc.max(-1), g["cls_conf_max"][i], rtol=4e-6, atol=1e-7)
        # NMS parity must be judged on IDENTICAL decoded inputs: feed the reference's own decode.
        rb, ro = g["boxes"][i], g["obj_conf"][i].reshape(-1)
        rc = g["cls_conf"][i]
        np.testing.assert_allclose(cc, rc, rtol=4e-6, atol=1e-7)
        n_cand = int((rc.max(-1) >= conf).sum())
        for key, kw in (("main", dict(mode=pr.torchvision_mode(min(n_cand, 5000), "cpu"))),
                        ("mainu", dict(max_nms=0, max_det=10 ** 9, mode=pr.torchvision_mode(n_cand, "cpu"))),
                        ("maina", dict(mode="agnostic"))):
            det, anc = pr.nms_image_main(rb, ro, rc, conf, thr, **kw)
            np.testing.assert_array_equal(det, g[f"{key}_det_{i}"], err_msg=key)


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_golden_yolox_postprocess(path):
    g = np.load(path)
    conf, thr = float(g["conf"]), float(g["nms_thr"])
    pred = g["yolox_pred"].astype(np.float32)
    for i in range(pred.shape[0]):
        C = pred.shape[2] - 5
        score = pred[i, :, 4] * pred[i, :, 5:].max(-1)
        n_cand = int((score >= conf).sum())
        det, anc, _ = pr.postprocess_image(pred[i], conf, thr, pr.torchvision_mode(n_cand, "cpu"))
        np.testing.assert_array_equal(det, g[f"yolox_det_{i}"])
        det, anc, _ = pr.postprocess_image(pred[i], conf, thr, "agnostic")
        np.testing.assert_array_equal(det, g[f"yoloxa_det_{i}"])


# ---- alternate candidate rules (multi_class / rmmop), goldens from the reference function itself -------------
MODE_CASES = {  # key -> oracle kwargs (mirrors tests/golden/make_golden_modes.py CASES)
    "mc": dict(multi_class=True),
    "mcu": dict(multi_class=True, max_nms=0, max_det=10 ** 9),
    "mca": dict(multi_class=True, mode="agnostic"),
    "rm": dict(rmmop=(2.0, 0.5)),
    "rmu": dict(rmmop=(1.0, 1.5), max_nms=0, max_det=10 ** 9),
    "rml": dict(rmmop=(1.05, 0.3)),
}


def mode_candidates(rc, ro, conf, kw):
    """Candidate count per image for the torchvision trick/vanilla dispatch."""
    if kw.get("multi_class"):
        return int((rc >= conf).sum())
    if kw.get("rmmop") is not None:
        r1, r2 = (np.float32(v) for v in kw["rmmop"])
        srt = np.sort(rc, -1)[:, ::-1]
        return int(((srt[:, 0] >= srt[:, 1] * r1) & (ro * ro >= srt[:, 0] * r2)).sum())
    return int((rc.max(-1) >= conf).sum())


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_golden_candidate_modes(path):
    mpath = os.path.join(os.path.dirname(path), "postmodes_" + os.path.basename(path)[len("post_"):])
    if not os.path.exists(mpath):
        pytest.skip("no mode vectors for this case")
    g, gm = np.load(path), np.load(mpath)
    conf, thr = float(g["conf"]), float(g["nms_thr"])
    for i in range(g["reg"].shape[0]):
        rb, ro, rc = g["boxes"][i], g["obj_conf"][i].reshape(-1), gm["cls_conf"][i]  # tie-free scores
        for key, kw in MODE_CASES.items():
            if f"{key}_det_{i}" not in gm:
                continue
            kw = dict(kw)
            n = mode_candidates(rc, ro, conf, kw)
            if kw.get("max_nms", 5000) > 0:
                n = min(n, kw.get("max_nms", 5000))
            kw.setdefault("mode", pr.torchvision_mode(n, "cpu"))
            det, anc = pr.nms_image_main(rb, ro, rc, conf, thr, **kw)
            np.testing.assert_array_equal(det, gm[f"{key}_det_{i}"], err_msg=key)
