"""COCOEvaluator drop-in (yolox/evaluators/coco_evaluator.py) on a synthetic dataset: the device-side
records equal the reference's per-detection CPU arithmetic bit for bit, and the AP numbers equal the oracle COCO
evaluation of those records."""
import numpy as np
import pytest
import torch

import yolox_b200 as yb
from oracle import cocoeval_ref as cr
from oracle import model_ref as mr

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
H = W = 128
CLASS_IDS = yb.io.COCO_CLASS_ID


class _Dataset:
    class_ids = CLASS_IDS
    coco = None


class _Loader:
    """What the evaluator reads from a torch DataLoader: iteration, len(), batch_size, .dataset."""

    def __init__(self, batches, batch_size):
        self.batches, self.batch_size, self.dataset = batches, batch_size, _Dataset()

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def _model():
    cfg = mr.CONFIGS["tiny_p6"]
    train = mr.synth_train_state(cfg, 3, calib_hw=(H, W))
    backbone = yb.models.YOLOPAFPNCustomP6(cfg.depth, cfg.width, act=cfg.act, in_channels=[256, 512, 768, 1024])
    head = yb.models.YOLOXHeadCustom(cfg.num_classes, cfg.width, act=cfg.act, strides=(8, 16, 32, 64),
                                     in_channels=[256, 512, 768, 1024])
    model = yb.models.YOLOXCustomP6(backbone, head)
    sd = dict(train)
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = v
    model.load_state_dict(sd, strict=True)
    return cfg, model.eval().cuda()


def test_evaluator_matches_reference_arithmetic_and_oracle_ap():
    cfg, model = _model()
    conf, nms = 0.3, 0.65
    sizes = [(480, 640), (375, 500), (427, 640), (600, 400), (333, 500)]
    batches, img_id = [], 11
    for b0 in (0, 2, 4):                         # batches of 2, 2, 1 (short last batch, excluded from the timing)
        hw = sizes[b0:b0 + 2]
        x = mr.synth_images(50 + b0, len(hw), H, W)
        ids = torch.arange(img_id, img_id + len(hw))
        img_id += len(hw)
        info = (torch.tensor([h for h, _ in hw]), torch.tensor([w for _, w in hw]))
        batches.append((x, None, info, ids))
    loader = _Loader(batches, 2)
    ev = yb.evaluator.COCOEvaluator(loader, (H, W), conf, nms, cfg.num_classes)

    # reference-style pass: list outputs -> convert_to_coco_format (per-detection CPU arithmetic of :135-165)
    ref_records = []
    for x, _, info, ids in batches:
        outs = yb.postprocess.postprocess(model(x.cuda()), cfg.num_classes, conf, nms)
        ref_records.extend(ev.convert_to_coco_format(outs, info, ids))
    assert len(ref_records) > 20, "the synthetic model should produce detections at this threshold"

    # ground truth: the strongest detections of every image (so AP is neither 0 nor 1) plus one decoy and one crowd
    by_img = {}
    for r in ref_records:
        by_img.setdefault(r["image_id"], []).append(r)
    anns = []
    for i, rs in by_img.items():
        for r in sorted(rs, key=lambda r: -r["score"])[:4]:
            x, y, w, h = r["bbox"]
            anns.append(dict(image_id=i, category_id=r["category_id"], bbox=[x + 0.07 * w, y, w, 0.9 * h], area=w * h * 0.8,
                             iscrowd=0))
        anns.append(dict(image_id=i, category_id=rs[0]["category_id"], bbox=[5000.0, 5000.0, 40.0, 40.0], area=1600.0, iscrowd=0))
        anns.append(dict(image_id=i, category_id=rs[-1]["category_id"], bbox=[0.0, 0.0, 300.0, 300.0], area=90000.0, iscrowd=1))
    img_ids = list(range(11, 16))
    loader.dataset.coco = dict(images=[dict(id=i) for i in img_ids], categories=[dict(id=c) for c in CLASS_IDS],
                               annotations=anns)

    ap, ap50, summary = ev.evaluate(model)
    got_records = []
    for x, _, info, ids in batches:                   # the evaluator's own (device-side) records, same order
        det, cnt, _ = yb.postprocess.postprocess_raw(model(x.cuda()), cfg.num_classes, conf, nms)
        got_records.extend(ev._records_dense(det, cnt, info, ids))
    assert len(got_records) == len(ref_records)
    for a, b in zip(got_records, ref_records):
        assert a["image_id"] == b["image_id"] and a["category_id"] == b["category_id"]
        assert a["bbox"] == b["bbox"] and a["score"] == b["score"], (a, b)     # bit-exact fp32 values

    want = cr.evaluate(anns, ref_records, img_ids, CLASS_IDS)["stats"]
    assert abs(ap - want[0]) < 1e-12 and abs(ap50 - want[1]) < 1e-12
    assert 0.0 < ap50 <= 1.0
    assert "Average forward time" in summary and "Average Precision  (AP) @[ IoU=0.50:0.95 | area=   all | maxDets=100 ]" in summary
    assert "= {:0.3f}".format(want[8]) in summary.splitlines()[9]


def test_evaluator_without_detections_and_half():
    cfg, model = _model()
    x = mr.synth_images(1, 2, H, W)
    info = (torch.tensor([480, 480]), torch.tensor([640, 640]))
    loader = _Loader([(x, None, info, torch.tensor([1, 2])), (x, None, info, torch.tensor([3, 4]))], 2)
    loader.dataset.coco = dict(images=[dict(id=i) for i in (1, 2, 3, 4)], categories=[dict(id=c) for c in CLASS_IDS],
                               annotations=[])
    ev = yb.evaluator.COCOEvaluator(loader, (H, W), 1.1, 0.65, cfg.num_classes)     # nothing passes a threshold > 1
    ap, ap50, summary = ev.evaluate(model, half=True)
    assert (ap, ap50) == (0, 0) and summary.startswith("Average forward time")
    with pytest.raises(NotImplementedError):
        ev.evaluate(model, trt_file="model_trt.pth")
