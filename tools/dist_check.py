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
dist.destroy_process_group()
sys.exit(0 if ok else 1)
