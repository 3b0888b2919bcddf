"""Golden vectors for the rows next to the hot path, produced by the UNMODIFIED reference functions
imported from /root/reference (run in the build container only):

    python tests/golden/make_golden_io.py   ->  tests/golden/io_preprocess.npz, tests/golden/io_coco.npz

  N1: yolox_collate_batch(img_size, [(PIL image resized like yolox_load_one_image_pil, info), ...])
      choijhanyangackr/yolox_infer/preprocess_utils.py:9-55 (the file read is replaced by Image.fromarray)
  N2: convert_to_coco_format(outputs, img_info, img_size)   choijhanyangackr/common/utils.py:27-73
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

REF = "/root/reference/choijhanyangackr"
sys.path.insert(0, REF)
OUT = os.path.dirname(os.path.abspath(__file__))


def synth_image(rng, h, w):
    """smooth structure + noise, so the resize has real gradients to interpolate"""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    base = np.stack([127 + 120 * np.sin(xx / (7 + 3 * c) + c) * np.cos(yy / (11 + 2 * c)) for c in range(3)], -1)
    return np.clip(base + rng.randn(h, w, 3) * 25, 0, 255).astype(np.uint8)


def main():
    from yolox_infer.preprocess_utils import yolox_collate_batch
    from common.utils import convert_to_coco_format
    rng = np.random.RandomState(7)
    img_size = 128
    sizes = [(60, 90), (150, 100), (49, 167), (128, 128), (240, 320), (32, 25)]   # (h, w): down/up-scaling, both orientations
    images, batch = [], []
    for h, w in sizes:
        im = synth_image(rng, h, w)
        images.append(im)
        pil = Image.fromarray(im)                       # == Image.open(...).convert("RGB")
        if w > h:                                       # preprocess_utils.py:17-22
            new_w = img_size; new_h = int(h * new_w / w)
        else:
            new_h = img_size; new_w = int(w * new_h / h)
        batch.append((pil.resize((new_w, new_h), resample=Image.BILINEAR), (h, w, "img_%d.jpg" % len(images), new_h, new_w)))
    ref_batch, info = yolox_collate_batch(img_size, batch)
    np.savez_compressed(os.path.join(OUT, "io_preprocess.npz"), img_size=img_size, sizes=np.array(sizes),
                        batch=ref_batch.numpy().astype(np.uint8), **{f"img{i}": im for i, im in enumerate(images)})

    # N2
    B, M, S = 5, 12, 640
    hw = [(480, 640), (1080, 1920), (333, 500), (640, 427), (100, 100)]
    counts = [12, 5, 0, 1, 9]
    det = np.zeros((B, M, 7), np.float32)
    outputs, img_info = [], []
    for b in range(B):
        n = counts[b]
        x1y1 = rng.rand(n, 2).astype(np.float32) * 500
        wh = rng.rand(n, 2).astype(np.float32) * 300 + 1
        d = np.concatenate([x1y1, x1y1 + wh, rng.rand(n, 2).astype(np.float32), rng.randint(0, 80, (n, 1)).astype(np.float32)], 1)
        det[b, :n] = d
        outputs.append(torch.from_numpy(d.copy()) if n else None)
        img_info.append((hw[b][0], hw[b][1], "val_%012d.jpg" % (1000 + b)))
    recs = convert_to_coco_format(outputs, img_info, S)
    flat = np.array([[r["image_id"], r["category_id"], *r["bbox"], r["score"]] for r in recs], np.float64)
    np.savez_compressed(os.path.join(OUT, "io_coco.npz"), det=det, count=np.array(counts, np.int32), hw=np.array(hw), img_size=S,
                        records=flat, names=np.array([i[2] for i in img_info]))
    print("wrote io_preprocess.npz", ref_batch.shape, "io_coco.npz", flat.shape)


if __name__ == "__main__":
    main()
