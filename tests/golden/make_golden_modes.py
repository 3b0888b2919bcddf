"""Golden vectors for the alternate candidate rules of yolox_nms_torch_batch (multi_class, rmmop;
choijhanyangackr/yolox_infer/postprocess_utils.py:74-95), produced by the UNMODIFIED reference function on the
decoded tensors already stored in tests/golden/post_*.npz (those were produced by the reference's own decode), with the
class scores perturbed until every value of an image is distinct: these rules rank up to A*C candidates with
torch.argsort(descending=True), which is not a stable sort, so the reference's own result on tied scores is an accident
of the sort implementation (the synthetic distributions tie heavily: fp16 logits, constant -6 off-class logit).

    python tests/golden/make_golden_modes.py        # needs /root/reference; writes tests/golden/postmodes_*.npz
"""
import glob
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/choijhanyangackr")
from yolox_infer.postprocess_utils import yolox_nms_torch_batch  # noqa: E402

# (key, kwargs) — rmmop ratios chosen so that both masks cut a visible share of the anchors
CASES = [
    ("mc", dict(multi_class=True)),
    ("mcu", dict(multi_class=True, max_num_nms=0, max_num_det=10 ** 9)),
    ("mca", dict(multi_class=True, class_agnostic=True)),
    ("rm", dict(rmmop=(2.0, 0.5))),
    ("rmu", dict(rmmop=(1.0, 1.5), max_num_nms=0, max_num_det=10 ** 9)),
    ("rml", dict(rmmop=(1.05, 0.3))),
]

if __name__ == "__main__":
    torch.set_grad_enabled(False)
    for path in sorted(glob.glob(os.path.join(HERE, "post_*.npz"))):
        g = np.load(path)
        conf, thr = float(g["conf"]), float(g["nms_thr"])
        rs = np.random.RandomState(1234)
        cc = g["cls_conf"].copy()
        cc *= (1.0 + 1e-3 * rs.random_sample(cc.shape)).astype(np.float32)
        for i in range(cc.shape[0]):
            while True:
                flat = cc[i].reshape(-1)
                _, first, counts = np.unique(flat, return_index=True, return_counts=True)
                if (counts == 1).all():
                    break
                dup = np.setdiff1d(np.arange(flat.size), first)
                flat[dup] *= (1.0 + 1e-3 * rs.random_sample(dup.size)).astype(np.float32)
        boxes, objc, clsc = torch.from_numpy(g["boxes"]), torch.from_numpy(g["obj_conf"]), torch.from_numpy(cc)
        out = {"cls_conf": cc}
        for key, kw in CASES:
            if key == "mcu" and boxes.shape[1] * clsc.shape[2] > 200000:
                continue  # uncapped multi-class on the 640 case is minutes of CPU NMS; covered by the 256 cases
            dets = yolox_nms_torch_batch(boxes, objc, clsc, nms_threshold=thr, conf_threshold=conf, **kw)
            for i, d in enumerate(dets):
                out[f"{key}_det_{i}"] = d.numpy() if d is not None else np.zeros((0, 7), np.float32)
            print(os.path.basename(path), key, [len(out[f"{key}_det_{i}"]) for i in range(len(dets))])
        tag = os.path.basename(path)[len("post_"):]
        np.savez_compressed(os.path.join(HERE, "postmodes_" + tag), **out)
