"""Measurement of the rows next to the hot path on one GPU, each beside the reference's own CPU way of
doing the same step on this host: device-side pre-processing vs Pillow + numpy collate, COCO records vs the per-detection
Python loop, the multi_class / rmmop candidate rules vs the oracle port, the native COCO evaluation vs the Python
restatement.  Writes a markdown table (argv[1], default gpurun_out/next_rows.md).  Timing: CUDA events after warm-up
for device work (through the public API, so host-side table building and H2D of the raw images are reported separately),
perf_counter for host work."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import yolox_b200 as yb
from oracle import cocoeval_ref, io_ref, post_ref

torch.set_grad_enabled(False)
OUT = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/next_rows.md"
rows = []


def cuda_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def host_ms(fn, iters=3):
    fn()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t) * 1e3 / iters


# ---- N1: pre-processing, 64 COCO-sized images -> [64,3,1280,1280] ------------------------------------------------
rs = np.random.RandomState(0)
sizes = [(480, 640), (427, 640), (640, 480), (375, 500), (500, 333), (640, 640), (360, 640), (612, 612)] * 8
images = [rs.randint(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
t_api = cuda_ms(lambda: yb.io.preprocess_batch(images, 1280, "cuda", torch.float16), iters=5, warm=2)
batch, _ = yb.io.preprocess_batch(images, 1280, "cuda", torch.float16)
src_bytes, out_bytes = sum(im.size for im in images), batch.numel() * 2
t_cpu = host_ms(lambda: io_ref.collate(images[:8], 1280), iters=1) * 8          # oracle == Pillow bit for bit; 8 of 64 timed
try:
    from PIL import Image

    def pil_path():
        out = []
        for im in images:
            h, w = im.shape[:2]
            nh, nw = io_ref.resized_shape(h, w, 1280)
            out.append(np.asarray(Image.fromarray(im).resize((nw, nh), Image.BILINEAR)))
        return out
    t_pil = host_ms(pil_path, iters=2)
except Exception:  # noqa: BLE001
    t_pil = float("nan")
rows.append(("N1 pre-processing, 64 images (~0.9 MB each) -> fp16 [64,3,1280,1280]",
             f"{t_api:.2f} ms per batch through `io.preprocess_batch` (host coefficient tables + H2D of {src_bytes / 1e6:.0f} MB raw pixels + kernel; "
             f"output {out_bytes / 1e6:.0f} MB)", f"Pillow BILINEAR resize alone, 1 thread: {t_pil:.0f} ms per batch (numpy oracle: {t_cpu:.0f} ms)"))

# ---- N2: COCO records ---------------------------------------------------------------------------------------------
det = torch.rand(64, 300, 7, device="cuda") * 100
det[..., 6] = torch.randint(0, 80, (64, 300), device="cuda").float()
cnt = torch.full((64,), 300, dtype=torch.int32, device="cuda")
hw = [(480, 640)] * 64
t_rec = cuda_ms(lambda: yb.io.coco_records(det, cnt, hw, 1280))
t_loop = host_ms(lambda: io_ref.coco_records(det.cpu().numpy(), cnt.cpu().numpy(), hw, 1280), iters=3)
rows.append(("N2 COCO records, 64 x 300 detections", f"{t_rec * 1e3:.0f} us (`io.coco_records`, incl. the scale / class-id table uploads)",
             f"numpy restatement of the reference arithmetic: {t_loop:.1f} ms (the reference itself loops per detection in Python)"))

# ---- N4: alternate candidate rules at full anchor count ------------------------------------------------------------
B, A, C = 8, 34000, 80
g = torch.Generator("cuda").manual_seed(1)
boxes = torch.rand(B, A, 4, device="cuda", generator=g) * 1000
boxes[..., 2:] += boxes[..., :2]
obj = torch.sigmoid(torch.randn(B, A, 1, device="cuda", generator=g) * 2 - 2)
cls = torch.sigmoid(torch.randn(B, A, C, device="cuda", generator=g) * 2 - 2) * obj
for name, kw in (("default", {}), ("multi_class", dict(multi_class=True)), ("rmmop=(1.05, 0.3)", dict(rmmop=(1.05, 0.3)))):
    t = cuda_ms(lambda: yb.postprocess.nms_main_raw(boxes, obj, cls, 0.65, 0.001, 5000, 300, **kw), iters=5, warm=2)
    bn, on, cn = boxes[0].cpu().numpy(), obj[0].cpu().numpy(), cls[0].cpu().numpy()
    okw = dict(multi_class=kw.get("multi_class", False), rmmop=kw.get("rmmop"))
    t_o = host_ms(lambda: post_ref.nms_image_main(bn, on, cn, 0.001, 0.65, 5000, 300, "trick", **okw), iters=1) * B
    rows.append((f"N4 `yolox_nms_torch_batch` {name}, {B} x {A} anchors x {C} classes, top-5000 -> 300",
                 f"{t:.2f} ms per batch ({t / B * 1e3:.0f} us per image)", f"C oracle, 1 thread: {t_o:.0f} ms per batch"))

# ---- N3: COCO evaluation ----------------------------------------------------------------------------------------------
rs = np.random.RandomState(3)
gts, dts = [], []
for i in range(1, 501):
    for _ in range(rs.randint(1, 12)):
        x, y = rs.uniform(0, 500, 2)
        w, h = rs.uniform(8, 300, 2)
        c = int(rs.randint(1, 81))
        gts.append(dict(image_id=i, category_id=c, bbox=[x, y, w, h], area=w * h, iscrowd=int(rs.rand() < 0.05)))
        for _ in range(rs.randint(0, 4)):
            j = rs.normal(0, 0.1, 4) * (w, h, w, h)
            dts.append(dict(image_id=i, category_id=c, bbox=[x + j[0], y + j[1], max(w + j[2], 1), max(h + j[3], 1)], score=rs.rand()))
    for _ in range(40):
        dts.append(dict(image_id=i, category_id=int(rs.randint(1, 81)), bbox=[*rs.uniform(0, 500, 2), *rs.uniform(8, 300, 2)], score=rs.rand() * 0.5))
imgs, cats = list(range(1, 501)), list(range(1, 81))
t_nat = host_ms(lambda: yb.cocoeval.COCOevalBBox(gts, dts, imgs, cats).evaluate(), iters=2)
t_py = host_ms(lambda: cocoeval_ref.evaluate(gts, dts, imgs, cats), iters=1)
rows.append((f"N3 COCO bbox evaluation, 500 images, {len(gts)} ground truths, {len(dts)} detections, 80 classes",
             f"{t_nat:.0f} ms (`yx_cocoeval_bbox`, host C++, incl. marshalling the dicts)", f"Python / numpy restatement of pycocotools: {t_py / 1e3:.1f} s"))

with open(OUT, "w") as f:
    f.write("# Rows next to the hot path (SURVEY §8f) — measured on one B200 box (`tools/bench_next_rows.py`)\n\n")
    f.write("| row / workload | this package | the reference's way, on this host's CPU |\n|---|---|---|\n")
    for r in rows:
        f.write("| " + " | ".join(r) + " |\n")
print(open(OUT).read())
