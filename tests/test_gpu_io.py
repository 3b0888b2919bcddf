"""GPU parity of the rows next to the hot path (csrc/yx_io.cu through yolox_b200.io) against the CPU oracle, the
reference-generated golden vectors and Pillow: pixels and records bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import io_ref

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_preprocess_matches_reference_golden():
    from yolox_b200 import io as yio
    z = np.load(os.path.join(G, "io_preprocess.npz"))
    images = [z[f"img{i}"] for i in range(len(z["sizes"]))]
    for dtype in (torch.float32, torch.float16):
        out, info = yio.preprocess_batch(images, int(z["img_size"]), dtype=dtype)
        assert out.dtype == dtype and tuple(out.shape) == z["batch"].shape
        assert np.array_equal(out.float().cpu().numpy().astype(np.uint8), z["batch"])
        assert info == [im.shape[:2] for im in images]


@pytest.mark.parametrize("sizes,img_size", [([(480, 640), (640, 480), (333, 500)], 640), ([(1080, 1920), (720, 1280)], 1280),
                                            ([(200, 3000)], 1280), ([(50, 40), (64, 64)], 416)])
def test_preprocess_matches_oracle_and_pillow(sizes, img_size):
    from yolox_b200 import io as yio
    rng = np.random.RandomState(len(sizes) * 31 + img_size)
    images = [(rng.rand(h, w, 3) * 255).astype(np.uint8) for h, w in sizes]
    out, _ = yio.preprocess_batch(images, img_size)
    got = out.cpu().numpy().astype(np.uint8)
    try:
        from PIL import Image
    except ImportError:
        Image = None
    for i, im in enumerate(images):
        nh, nw = io_ref.resized_shape(im.shape[0], im.shape[1], img_size)
        ref = np.asarray(Image.fromarray(im).resize((nw, nh), resample=Image.BILINEAR)) if Image is not None \
            else io_ref.pil_resize_bilinear(im, nw, nh)
        assert np.array_equal(got[i, :, :nh, :nw], ref[..., ::-1].transpose(2, 0, 1)), f"image {i}: resized pixels differ"
        pad = got[i].copy()
        pad[:, :nh, :nw] = 114
        assert (pad == 114).all(), "padding must be 114"
    assert got.shape[2] % (64 if img_size % 64 == 0 else 32) == 0 and got.shape[3] % (64 if img_size % 64 == 0 else 32) == 0


def test_coco_records_match_golden_and_oracle():
    from yolox_b200 import io as yio
    z = np.load(os.path.join(G, "io_coco.npz"))
    det, cnt = torch.from_numpy(z["det"]).cuda(), torch.from_numpy(z["count"]).cuda()
    hw = [tuple(int(v) for v in x) for x in z["hw"]]
    rec = yio.coco_records(det, cnt, hw, int(z["img_size"]))
    assert np.array_equal(rec.cpu().numpy(), io_ref.coco_records(z["det"], z["count"], hw, int(z["img_size"])))
    info = [(h, w, str(n)) for (h, w), n in zip(hw, z["names"])]
    recs = yio.convert_to_coco_format((det, cnt), info, int(z["img_size"]))
    flat = np.array([[r["image_id"], r["category_id"], *r["bbox"], r["score"]] for r in recs], np.float64)
    assert np.array_equal(flat, z["records"]), "records differ from the reference's convert_to_coco_format"
    # reference-style list input (tensors / None)
    outs = [det[b, :int(cnt[b])] if int(cnt[b]) else None for b in range(det.shape[0])]
    assert yio.convert_to_coco_format(outs, info, int(z["img_size"])) == recs


def test_predict_loop_with_device_io():
    """image arrays -> device pre-processing -> Predictor -> COCO records: the whole main.py loop body on the device."""
    import yolox_b200 as yb
    from oracle import model_ref as mr
    cfg = mr.CONFIGS["tiny_p6"]
    model = yb.infer.YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(mr.fold_bn(mr.synth_train_state(cfg, 3, calib_hw=(128, 128))), strict=True)
    model = model.cuda().half().eval()
    rng = np.random.RandomState(5)
    images = [(rng.rand(90, 120, 3) * 255).astype(np.uint8), (rng.rand(128, 100, 3) * 255).astype(np.uint8)]
    batch, info = yb.io.preprocess_batch(images, 128, dtype=torch.float16)
    det, cnt = yb.predict.Predictor(model, conf_threshold=0.3, nms_threshold=0.5)(batch)
    recs = yb.io.convert_to_coco_format((det, cnt), [(h, w, f"x_{i}.jpg") for i, (h, w) in enumerate(info)], 128)
    assert len(recs) == sum(max(int(c), 1) for c in cnt.tolist())
    assert all(set(r) == {"image_id", "category_id", "bbox", "score"} for r in recs)


def test_uint8_images_end_to_end():
    """uint8 image batches (yx_dtype YX_U8): preprocess_batch can emit them and the engine consumes them, bit-identical to the
    same pixel values passed as fp16 -- 4x fewer bytes than the reference's float32 batch on the host->device link."""
    import yolox_b200 as yb
    from oracle import model_ref as mr
    rs = np.random.RandomState(3)
    images = [rs.randint(0, 256, (h, w, 3), dtype=np.uint8) for h, w in ((120, 160), (97, 131))]
    b8, info8 = yb.io.preprocess_batch(images, 128, "cuda", torch.uint8)
    b16, info16 = yb.io.preprocess_batch(images, 128, "cuda", torch.float16)
    assert b8.dtype == torch.uint8 and info8 == info16 and torch.equal(b8.half(), b16)
    cfg = mr.CONFIGS["tiny_p6"]
    model = yb.infer.YOLOXP6(cfg.depth, cfg.width, act=cfg.act, num_classes=cfg.num_classes)
    model.load_state_dict(mr.fold_bn(mr.synth_train_state(cfg, 5, calib_hw=(128, 128))), strict=True)
    model = model.cuda().half().eval()
    pred = yb.predict.Predictor(model, conf_threshold=0.05)
    d8, c8 = pred(b8)
    d16, c16 = pred(b16)
    assert torch.equal(c8, c16) and torch.equal(d8, d16)
    r8, o8, k8 = model(b8)
    r16, o16, k16 = model(b16)
    assert r8.dtype == torch.float16 and torch.equal(r8, r16) and torch.equal(o8, o16) and torch.equal(k8, k16)
