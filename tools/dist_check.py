"""2-GPU functional check of Predictor.predict_sharded: sharded + all-gathered detections == single-GPU run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import yolox_b200 as yb

torch.set_grad_enabled(False)
os.environ["YX_TUNE"] = "0"   # bit-equality across DIFFERENT per-rank batch sizes needs the shape-only heuristic launch shapes
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
model = bench.build_model(torch.device("cuda", local))
pred = yb.predict.Predictor(model)
g = torch.Generator().manual_seed(7)
world = dist.get_world_size()
ok = True
for n_img in (3 * world + 1, world - 1):      # ragged shards (last ranks padded); fewer images than ranks (empty shards)
    img = (torch.rand(n_img, 3, 256, 320, generator=g) * 255).half().cuda()
    det_one, cnt_one = pred(img)
    for fused in ("1", "0"):                  # NMS-fused peer gather, then the NCCL all-gather it replaces
        os.environ["YX_PEER_GATHER"] = fused
        for it in range(5):                   # several steps: window parity, arrival-counter targets
            det_all, cnt_all = pred.predict_sharded(img)
            same = torch.equal(det_all, det_one) and torch.equal(cnt_all, cnt_one)
            ok = ok and same
        print(f"dist_check rank {rank}: {n_img} images peer_gather={fused} sharded==single {same} counts {cnt_all.tolist()}", flush=True)
for g in pred._gathers.values():
    st = g.status()
    ok = ok and st == 0
    print(f"dist_check rank {rank}: gather status {st} after {g.step} steps", flush=True)
    g.close()

# ---- COCOEvaluator.evaluate(distributed=True): per-rank loaders, records gathered to rank 0 (coco_evaluator.py:125-129) ----
from oracle import model_ref as mr                                    # noqa: E402  (checker only)
from tests.test_gpu_evaluator import _Loader, _model, CLASS_IDS, H, W  # noqa: E402

cfg, emodel = _model()
sizes = [(480, 640), (375, 500), (427, 640), (600, 400), (333, 500), (500, 500), (640, 480), (300, 400)]
all_batches = []
for b in range(4):                                   # 4 batches of 2 images; rank r evaluates batches r, r + world, ...
    info = (torch.tensor([sizes[2 * b][0], sizes[2 * b + 1][0]]), torch.tensor([sizes[2 * b][1], sizes[2 * b + 1][1]]))
    all_batches.append((mr.synth_images(70 + b, 2, H, W), None, info, torch.tensor([100 + 2 * b, 101 + 2 * b])))
probe = yb.evaluator.COCOEvaluator(_Loader(all_batches, 2), (H, W), 0.3, 0.65, cfg.num_classes)
anns = []                                            # ground truth = the strongest detections, slightly shifted (0 < AP < 1)
for x, _, info, ids in all_batches:
    det, cnt, _ = yb.postprocess.postprocess_raw(emodel(x.cuda()), cfg.num_classes, 0.3, 0.65)
    recs = probe._records_dense(det, cnt, info, ids)
    for i in ids.tolist():
        for r in sorted((r for r in recs if r["image_id"] == i), key=lambda r: -r["score"])[:3]:
            bx, by, bw, bh = r["bbox"]
            anns.append(dict(image_id=i, category_id=r["category_id"], bbox=[bx + 0.05 * bw, by, bw, 0.92 * bh], area=bw * bh,
                             iscrowd=0))
coco = dict(images=[dict(id=100 + i) for i in range(8)], categories=[dict(id=c) for c in CLASS_IDS], annotations=anns)
mine = _Loader(all_batches[rank::world], 2)
mine.dataset.coco = coco
ev = yb.evaluator.COCOEvaluator(mine, (H, W), 0.3, 0.65, cfg.num_classes)
ap, ap50, summary = ev.evaluate(emodel, distributed=True)
# every rank runs the single-process reference too: evaluate() ends with a barrier whenever a process group exists
# (the reference's synchronize(), coco_evaluator.py:132), so it must be called by all ranks together
whole = _Loader(all_batches, 2)
whole.dataset.coco = coco
ap1, ap501, _ = yb.evaluator.COCOEvaluator(whole, (H, W), 0.3, 0.65, cfg.num_classes).evaluate(emodel)
if rank == 0:
    same = abs(ap - ap1) < 1e-12 and abs(ap50 - ap501) < 1e-12 and 0.0 < ap50 <= 1.0
    ok = ok and same
    print(f"dist_check rank 0: distributed evaluator AP {ap:.6f} / AP50 {ap50:.6f} == single-process {same}", flush=True)
else:
    ok = ok and (ap, ap50, summary) == (0, 0, None)
    print(f"dist_check rank {rank}: evaluator returned (0, 0, None) off the main process {(ap, ap50, summary) == (0, 0, None)}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
