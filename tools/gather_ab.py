"""A/B of the two ways to end a sharded step, inside ONE process per GPU and on the SAME tuned engine:
(a) NMS-fused peer gather (yx_detect_main_gather), (b) detect + packed NCCL all-gather.  torchrun, 2+ GPUs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import yolox_b200 as yb
from yolox_b200 import postprocess as pp

torch.set_grad_enabled(False)
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = int(os.environ.get("AB_BATCH", 64)), int(os.environ.get("AB_SIZE", 1280))
model = bench.build_model(dev)
M = bench.MODEL
img = (torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(rank)) * 255).half().to(dev)
peer = yb.dist.PeerGather(B, bench.MAX_DET, dev)
gathered = torch.empty(world, B, bench.MAX_DET * 7 + 1, dtype=torch.float32, device=dev)


def step(mode):
    eng, reg8, cls = model.run_engine(img, in_scale=0.9, in_shift=11.4)
    det, cnt, _ = pp.detect_main(reg8[..., :4], reg8[..., 4:5], cls[..., :M["num_classes"]], model.head.hw, M["strides"],
                                 bench.CONF_THR, bench.NMS_THR, bench.MAX_NMS, bench.MAX_DET,
                                 gather=peer if mode == "peer" else None)
    if mode == "nccl":
        packed = torch.cat([det.view(B, -1), cnt.view(B, 1).float()], dim=1)
        dist.all_gather_into_tensor(gathered.view(world * B, -1), packed)


def timed(mode, k):
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        step(mode)
    b.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / k], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for m in ("none", "peer", "nccl"):
    for _ in range(3):
        step(m)
res = {m: [] for m in ("none", "peer", "nccl")}
for rep in range(4):
    for m in ("none", "peer", "nccl"):
        res[m].append(timed(m, 10))
if rank == 0:
    for m, v in res.items():
        print(f"gather_ab {m:5s} ms/step " + " ".join(f"{x:.3f}" for x in v) + f"  min {min(v):.3f}", flush=True)
    print("gather_ab status", peer.status(), flush=True)
dist.destroy_process_group()
