"""CPU: the oracle for the rows next to the hot path (oracle/io_ref.py) is pinned against golden vectors produced by the
reference functions (tests/golden/make_golden_io.py) and, where Pillow is importable, against Pillow itself; the host side
of yolox_b200.io (coefficient tables) is checked against the oracle's scalar restatement of Resample.c."""
import os

import numpy as np
import pytest

from oracle import io_ref

G = os.path.join(os.path.dirname(__file__), "golden")


def _golden_images():
    z = np.load(os.path.join(G, "io_preprocess.npz"))
    n = len(z["sizes"])
    return int(z["img_size"]), [z[f"img{i}"] for i in range(n)], z["batch"]


def test_preprocess_oracle_matches_reference_golden():
    img_size, images, ref = _golden_images()
    got, info = io_ref.collate(images, img_size)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.array_equal(got.astype(np.uint8), ref), "oracle collate differs from yolox_collate_batch"
    assert info == [im.shape[:2] for im in images]


@pytest.mark.parametrize("h,w,nh,nw", [(37, 53, 14, 20), (96, 128, 48, 64), (75, 100, 96, 128), (120, 333, 46, 128),
                                        (64, 64, 64, 64), (200, 150, 64, 48), (31, 17, 128, 70)])
def test_resize_oracle_bit_identical_to_pillow(h, w, nh, nw):
    Image = pytest.importorskip("PIL.Image")
    img = (np.random.RandomState(h * 1000 + w).rand(h, w, 3) * 255).astype(np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((nw, nh), resample=Image.BILINEAR))
    assert np.array_equal(io_ref.pil_resize_bilinear(img, nw, nh), ref)


@pytest.mark.parametrize("n_in,n_out", [(53, 20), (640, 427), (500, 1280), (1920, 1280), (100, 100), (4000, 1280), (7, 64)])
def test_host_coefficient_tables_match_oracle(n_in, n_out):
    from yolox_b200 import io as yio
    b0, k0 = io_ref.bilinear_coeffs(n_in, n_out)
    b1, k1 = yio._coeffs(n_in, n_out)
    assert np.array_equal(b0, b1) and np.array_equal(k0, k1)


def test_resized_shape_matches_reference_rule():
    from yolox_b200 import io as yio
    for h, w in [(480, 640), (640, 480), (333, 500), (1080, 1920), (100, 100), (97, 300)]:
        assert yio.resized_shape(h, w, 1280) == io_ref.resized_shape(h, w, 1280)
    assert io_ref.resized_shape(480, 640, 1280) == (960, 1280) and io_ref.resized_shape(640, 480, 1280) == (1280, 960)


def test_coco_oracle_matches_reference_golden():
    z = np.load(os.path.join(G, "io_coco.npz"))
    rec = io_ref.coco_records(z["det"], z["count"], [tuple(x) for x in z["hw"]], int(z["img_size"]))
    rows = []
    for b, name in enumerate(z["names"]):
        image_id = int(str(name).split("_")[-1].split(".")[0])
        n = int(z["count"][b])
        if n == 0:
            rows.append([image_id, 0, 0, 0, 0, 0, 0.0])       # the reference's placeholder record for an empty image
        for r in rec[b, :n]:
            rows.append([image_id, int(r[5]), *[float(v) for v in r[:4]], float(r[4])])
    assert np.array_equal(np.array(rows, np.float64), z["records"]), "oracle records differ from convert_to_coco_format"
